// Poseidon Merkle trees over device-resident leaves (see merkle.cu).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "dev.cuh"

namespace zkm {

// Digest tree: level 0 = leaf digests (num_leaves x 4 u64), level l = num_leaves >> l digests, up to
// the cap level (1 << cap_height digests).  All levels live in one buffer.
struct MerkleTreeDev {
    int log_leaves = 0, cap_height = 0;
    DevBuf digests;
    std::vector<size_t> level_off;          // element (u64) offset of each level
    std::vector<u64> cap;                   // host copy: (1<<cap_height)*4 words
    // in-segment sharding (shard.cuh): this rank fills and builds only the leaf quarters of the cosets it owns; the cap is
    // completed by an all-gather.  Digests of the other quarters are never written or read on this rank.
    bool sharded = false;
    int num_levels() const { return log_leaves - cap_height + 1; }
    size_t num_leaves() const { return (size_t)1 << log_leaves; }
};

void merkle_alloc(MerkleTreeDev& t, int log_leaves, int cap_height, cudaStream_t s);
// Builds levels 1.. from level 0 (already filled) and downloads the cap.
void merkle_build_from_leaf_digests(MerkleTreeDev& t, cudaStream_t s);

// Leaf digests of a coset-major LDE (ntt.cuh lde_coset): leaf index = bitrev(natural LDE index).
// ncols <= 4: the row itself, zero padded (plonky2 hash_or_noop); else overwrite-mode sponge.
// [coset_begin, coset_begin + coset_count): hash only the leaves of those cosets (coset_count < 0: all remaining), and of
// those only the rows i = row_offset mod 2^row_log_stride (a contiguous 2^-row_log_stride of each coset's leaf quarter).
void lde_leaf_hash(const u64* lde, size_t col_stride, int ncols, int log_n, int rate_bits, u64* leaf_digests, cudaStream_t s,
                   int coset_begin = 0, int coset_count = -1, int row_log_stride = 0, int row_offset = 0);

// Leaf digests for rows stored row-major and already in leaf order: rows[leaf*width .. +width).
void rows_leaf_hash(const u64* rows, int width, size_t num_leaves, u64* leaf_digests, cudaStream_t s);

// out[q*path_len*4 ..] = siblings of leaf idx[q], leaf level first.  idx is a device array.
void merkle_gather_paths(const MerkleTreeDev& t, const u32* d_idx, int nq, u64* d_out, cudaStream_t s);

// out[q*ncols + c] = LDE row of leaf idx[q] (leaf order = bit-reversed natural order).
void lde_gather_rows(const u64* lde, size_t col_stride, int ncols, int log_n, int rate_bits, const u32* d_idx, int nq, u64* d_out,
                     cudaStream_t s);

}  // namespace zkm
