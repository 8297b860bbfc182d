// Device-resident PolynomialBatch (coefficients + coset-major LDE + Poseidon Merkle tree) and the
// process-wide device context.  Mirrors plonky2's fri::oracle::PolynomialBatch as used by the
// reference (prover/src/prover.rs:154-163,514-521,576-587; fields evidenced at prover.rs:180,687).
#pragma once
#include "dev.cuh"
#include "ntt.cuh"
#include "merkle.cuh"
#include <functional>
#include <vector>

namespace zkm {

struct Arena;
struct Ctx {
    int device = -1;
    Arena* arena = nullptr;                // worker contexts own one; the process-wide context uses the default arena
    cudaStream_t stream = 0;
    cudaStream_t copy_stream = 0;          // host->device trace uploads, overlapped with the commitments
    NttTables ntt;
    // Pinned bounce ring for host->device copies from PAGEABLE memory (the real caller's `Vec<F>` columns): the uploader copies
    // chunk k into slot k % BOUNCE_SLOTS with a few host threads while the DMA engine drains the previous slots, instead of the
    // driver's own serial stage-then-copy.  Allocated on first use (upload_bounced, capi.cu), released with the context.
    static constexpr int BOUNCE_SLOTS = 4;
    static constexpr size_t BOUNCE_BYTES = (size_t)16 << 20;
    void* bounce[BOUNCE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t bounce_free[BOUNCE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    int bounce_next = 0;
    ~Ctx();
};
Ctx& ctx();                    // throws if zkm_b200_init has not succeeded
bool ctx_ready();
void ctx_init(int device);
void ctx_shutdown();
// Worker contexts: a second (third, ...) set of streams + tables + arena on the initialised device.  A host thread that has
// bound one runs every library call on it; two bound threads can prove concurrently.
Ctx* worker_create();
void worker_bind(Ctx* w);       // nullptr unbinds
void worker_destroy(Ctx* w);

struct Batch {
    int ncols = 0, log_n = 0, rate_bits = 0, cap_height = 0;
    DevBuf coeffs;             // ncols x n            (column-major, natural order)
    DevBuf lde;                // ncols x (n<<rate)    (column-major, coset-major rows)
    MerkleTreeDev tree;
    bool sharded = false;      // in-segment sharding (shard.cuh): only this rank's cosets of `lde` / quarters of `tree` exist here
    size_t n() const { return (size_t)1 << log_n; }
    size_t lde_n() const { return (size_t)1 << (log_n + rate_bits); }
    int lde_bits() const { return log_n + rate_bits; }
};

// values (device, column-major ncols x n; consumed: may alias coeffs) -> batch
void batch_from_values_dev(Batch& b, DevBuf&& values, int ncols, int log_n, int rate_bits, int cap_height);
void batch_from_coeffs_dev(Batch& b, DevBuf&& coeffs, int ncols, int log_n, int rate_bits, int cap_height);
// Same result as ifft + batch_from_coeffs_dev, but the columns are transformed group by group: group k = columns
// [col_ends[k-1], col_ends[k]) is touched only after wait_group(k) returns (the uploader has delivered it), so the NTTs of the
// first groups overlap the upload of the later ones; the leaf hashing (which needs every column) runs last.
void batch_from_values_grouped_dev(Batch& b, const u64* values, DevBuf&& coeffs, int ncols, int log_n, int rate_bits, int cap_height,
                                   const std::vector<int>& col_ends, const std::function<void(size_t)>& wait_group);

}  // namespace zkm
