// Device runtime plumbing shared by the kernels' host launchers: error handling, RAII device
// buffers on the stream-ordered allocator, and per-size twiddle/power tables.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include "gl.cuh"

namespace zkm {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define ZKM_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            throw ::zkm::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " @" +  \
                                   __FILE__ + ":" + std::to_string(__LINE__));                   \
    } while (0)

#define ZKM_CHECK(cond, msg)                                                                     \
    do { if (!(cond)) throw std::runtime_error(std::string(msg)); } while (0)

// Count of kernel launches issued by this library (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;
#define ZKM_LAUNCHED() do { __atomic_fetch_add(&::zkm::g_launch_count, 1ULL, __ATOMIC_RELAXED); ZKM_CUDA(cudaGetLastError()); } while (0)

// Device memory arena: blocks come from cudaMalloc once and are then recycled through a size-keyed free
// list, so a steady-state proof issues no driver allocation calls at all (the stream-ordered driver pool
// cost 100s of ms per proof in map/unmap work at multi-GB sizes).  All library work runs on one stream, so
// handing a freed block to the next user is ordered after the previous user's kernels.
void* arena_alloc(size_t bytes);
void arena_free(void* p, size_t bytes);
void arena_trim();                 // cudaFree every cached block
size_t arena_cached_bytes();

// Host wait for a stream: cudaStreamSynchronize (the driver spins), or -- ZKM_BLOCKING_SYNC=1 -- a wait on an event created with
// cudaEventBlockingSync, which frees the waiting thread's core.  Measured (profiles/r2q_blocking_sync_ab.txt, 24 host threads
// for 2 GPUs x 3 proofs in flight): spinning is 2.6 % faster on one GPU and 11 % faster end to end on two, so spinning is the
// default; the blocking wait is for hosts with fewer cores than waiting threads.
cudaError_t stream_sync(cudaStream_t s);
bool blocking_sync_enabled();

struct DevBuf {
    u64* p = nullptr;
    size_t n = 0;            // elements (u64)
    cudaStream_t stream = 0;
    DevBuf() {}
    DevBuf(size_t n_, cudaStream_t s = 0) { alloc(n_, s); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), stream(o.stream) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; stream = o.stream; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n_, cudaStream_t s = 0) {
        release();
        n = n_; stream = s;
        if (n) p = (u64*)arena_alloc(n * sizeof(u64));
    }
    void release() {
        if (p) { arena_free(p, n * sizeof(u64)); p = nullptr; n = 0; }
    }
    void zero() { if (p) ZKM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(u64), stream)); }
    void upload(const u64* h, size_t cnt, size_t off = 0) {
        ZKM_CUDA(cudaMemcpyAsync(p + off, h, cnt * sizeof(u64), cudaMemcpyHostToDevice, stream));
    }
    void download(u64* h, size_t cnt, size_t off = 0) const {
        ZKM_CUDA(cudaMemcpyAsync(h, p + off, cnt * sizeof(u64), cudaMemcpyDeviceToHost, stream));
        ZKM_CUDA(stream_sync(stream));
    }
};

// Per-kernel-family device timing (CUDA events on the launching stream), switched on by
// zkm_b200_profile_enable.  Scope names follow the kernel families listed in DESIGN.md; bench.py
// reads them back for the roofline entry.  Disabled: zero overhead beyond one branch.
struct ProfScope {
    const char* name; cudaStream_t s; cudaEvent_t e0 = nullptr, e1 = nullptr; double bytes, aux;
    // algorithmic_bytes: SURVEY section 8(d) accounting (inputs read once + outputs written once); aux: a second per-family
    // counter -- bytes moved by the shared-memory passes for the NTT family, Poseidon permutations for the hashing families
    ProfScope(const char* name_, cudaStream_t s_, double algorithmic_bytes = 0, double aux = 0);
    ~ProfScope();
};
void prof_enable(bool on);
void prof_reset();
// Resolves pending events (synchronises) and returns totals for one family; false if never seen.
bool prof_get(const char* name, double* ms, unsigned long long* launches, double* bytes, double* aux = nullptr);
// Names seen so far, '\n'-separated.
std::string prof_names();

// Timings keyed by the reference's TimingTree scope strings (prover.rs:86,146,152,193,204,250..411,479,513,545,562,578,620):
// the reference times nested wall-clock scopes on the host; here every scope is bracketed by CUDA events on the launching
// stream, so the figure is the device time of that scope.  Off by default (zkm_b200_timing_enable); resolved once per proof.
struct TimedScope {
    cudaStream_t s; int idx = -1;
    TimedScope(const std::string& name, cudaStream_t s_);
    ~TimedScope();
    TimedScope(const TimedScope&) = delete;
};
void scopes_enable(bool on);
void scopes_begin();                       // start of a proof on this thread's context
void scopes_finish();                      // end of the proof: resolves the events into the text returned by scopes_last()
const std::string& scopes_last();          // "<depth>\t<ms>\t<scope name>\n" per scope, in opening order

// Row-major n x ncols block -> column-major ncols x n (transpose.cu).
void transpose_rows_to_cols(const u64* rows, u64* cols, size_t n, int ncols, cudaStream_t s);

// Arithmetic table (column-major, on the device): fills RANGE_COUNTER and adds the shared columns' histogram to RC_FREQUENCIES.
void arith_generate_range_checks(u64* cols, size_t n, int first_shared, int num_shared, int counter_col, int freq_col,
                                 unsigned* d_bad, cudaStream_t s);

// Memory table from the memory-operation log (memtrace.cu): returns the table height, fills 13 x height columns.
size_t memory_generate_trace_dev(const u64* h_ops, size_t n_ops, struct DevBuf& cols, cudaStream_t s);

// Tables generated from operation logs (tracegen.cu): return the table height, fill ncols x height columns.
size_t arithmetic_generate_trace_dev(const u64* h_ops, size_t n_ops, struct DevBuf& cols, cudaStream_t s);
size_t logic_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
size_t keccak_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
size_t poseidon_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
// The six hash-precompile tables (tracegen_hash.cu); the two byte sponges take a variable-width log (word 0 = its word count).
size_t sha_extend_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
size_t sha_extend_sponge_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
size_t sha_compress_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
size_t sha_compress_sponge_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
size_t keccak_sponge_generate_trace_dev(const u64* h_log, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);
size_t poseidon_sponge_generate_trace_dev(const u64* h_log, size_t n_ops, size_t min_rows, struct DevBuf& cols, cudaStream_t s);

// Two-level table of powers of one field element g:  g^e = lo[e & (2^lo_bits-1)] * hi[e >> lo_bits].
struct PowTable {
    const u64* lo = nullptr;
    const u64* hi = nullptr;
    int lo_bits = 10;
};
#ifdef __CUDACC__
__device__ __forceinline__ gl pow_lookup(const PowTable& t, u64 e) {
    gl l(t.lo[e & ((1u << t.lo_bits) - 1)]);
    u64 h = e >> t.lo_bits;
    if (h == 0) return l;
    return l * gl(t.hi[h]);
}
#endif

struct PowTableOwner {
    DevBuf lo, hi;
    PowTable view;
};
// Builds (on the host, uploads) the table for base g covering exponents < 2^max_bits.
std::shared_ptr<PowTableOwner> make_pow_table(gl g, int max_bits, cudaStream_t s);

}  // namespace zkm
