// Poseidon Merkle-cap commitment kernels.  Replaces plonky2's MerkleTree::new(leaves, cap_height)
// as reached from PolynomialBatch::from_values/from_coeffs (reference prover/src/prover.rs:154-163,
// 514-521,576-587) and the per-fold trees inside prove_openings (:621).
//   leaf   = hash_or_noop(row): <= 4 elements copied (zero padded), else Poseidon overwrite-mode
//            sponge, rate 8, output state[0..4)            
//   node   = Poseidon compress(left, right)
//   cap[i] = root of the i-th contiguous block of leaves   (Appendix A.5)
// One thread owns one 12-word sponge state in registers.  The LDE stays column-major (coalesced
// across the warp for every column); the row-major leaf matrix of the reference is never built.
#include "merkle.cuh"
#include "poseidon_v2.cuh"
#include "shard.cuh"
#include <cstring>

namespace zkm {

__device__ __forceinline__ size_t lde_pos_of_leaf(u32 leaf, int log_n, int rate_bits) {
    u32 m = bitrev32(leaf, log_n + rate_bits);
    u32 j = m & ((1u << rate_bits) - 1), i = m >> rate_bits;
    return ((size_t)j << log_n) + i;
}

__global__ void __launch_bounds__(128, 8) lde_leaf_hash_kernel(const u64* __restrict__ lde, size_t cs, int ncols, int log_n,
                                                            int rate_bits, u64* __restrict__ dig, size_t pos_begin, size_t pos_end, int row_log_stride,
                                                            int row_offset) {
    size_t pos = pos_begin + ((((size_t)blockIdx.x * blockDim.x + threadIdx.x) << row_log_stride) | row_offset);
    if (pos >= pos_end) return;
    // position -> natural index -> leaf index
    u32 j = (u32)(pos >> log_n), i = (u32)(pos & (((size_t)1 << log_n) - 1));
    u32 m = (i << rate_bits) | j;
    u32 leaf = bitrev32(m, log_n + rate_bits);
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    const u64* col = lde + pos;
    if (ncols <= 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < ncols) s[k] = col[(size_t)k * cs];
    } else {
        // one copy of the permutation code (instruction-cache footprint): the ragged last chunk is handled by
        // predicated loads inside the same loop
#pragma unroll 1
        for (int c = 0; c < ncols; c += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) if (c + k < ncols) s[k] = __ldg(col + (size_t)(c + k) * cs);
            poseidon_permute_dev(s);
        }
    }
    ulonglong2* o = reinterpret_cast<ulonglong2*>(dig + (size_t)leaf * 4);
    o[0] = make_ulonglong2(s[0], s[1]);
    o[1] = make_ulonglong2(s[2], s[3]);
}

__global__ void __launch_bounds__(128, 8) rows_leaf_hash_kernel(const u64* __restrict__ rows, int width, size_t num_leaves,
                                                             u64* __restrict__ dig) {
    size_t leaf = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= num_leaves) return;
    const u64* row = rows + leaf * (size_t)width;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    if (width <= 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < width) s[k] = row[k];
    } else {
#pragma unroll 1
        for (int c = 0; c < width; c += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) if (c + k < width) s[k] = row[c + k];
            poseidon_permute_dev(s);
        }
    }
    ulonglong2* o = reinterpret_cast<ulonglong2*>(dig + leaf * 4);
    o[0] = make_ulonglong2(s[0], s[1]);
    o[1] = make_ulonglong2(s[2], s[3]);
}

// Which quarters of a level this launch covers (in-segment sharding, shard.cuh: a rank builds only the subtrees over the
// leaf quarters it owns).  Thread i works on node q[i >> log_qsize] * qsize + (i & (qsize - 1)); the identity map
// {4, {0,1,2,3}} is the whole level.
struct QuarterMap { int nq; int q[4]; int log_qsize; };      // "quarter" = one of the 4 * parts() leaf segments (shard.cuh)
__device__ __forceinline__ size_t quarter_node(const QuarterMap& m, size_t i) {
    return ((size_t)m.q[i >> m.log_qsize] << m.log_qsize) | (i & (((size_t)1 << m.log_qsize) - 1));
}
__global__ void __launch_bounds__(128, 8) merkle_level_kernel(const u64* __restrict__ child, u64* __restrict__ parent, size_t n_parents,
                                                           QuarterMap qm) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parents) return;
    i = quarter_node(qm, i);
    const ulonglong2* c = reinterpret_cast<const ulonglong2*>(child + i * 8);
    ulonglong2 a = c[0], b = c[1], d = c[2], e = c[3];
    u64 s[12] = {a.x, a.y, b.x, b.y, d.x, d.y, e.x, e.y, 0, 0, 0, 0};
    poseidon_permute_dev(s);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(parent + i * 4);
    o[0] = make_ulonglong2(s[0], s[1]);
    o[1] = make_ulonglong2(s[2], s[3]);
}

// Small levels (<= 2^14 parents) cannot fill the machine with one thread per permutation: a level then costs one
// single-thread permutation latency (~30 us, 11 such levels per 2^22-leaf tree and ~80 trees per proof).  Here 12 lanes
// share one permutation, one state word each: the S-boxes of a full round run side by side and the MDS row of a lane is
// 12 shuffled multiply-adds, so the dependent instruction chain is ~4x shorter.  Two permutations per warp (lanes 0-11
// and 16-27).  Same merged partial-round constants as poseidon_permute_v9; outputs are bit-identical.
__global__ void __launch_bounds__(128) merkle_level_coop_kernel(const u64* __restrict__ child, u64* __restrict__ parent, size_t n_parents,
                                                             QuarterMap qm) {
    const unsigned lane = threadIdx.x & 31, sub = lane & 15, grp = lane >> 4;
    const size_t slot = ((size_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * 2 + grp;
    const bool live = sub < 12 && slot < n_parents;
    const size_t node = slot < n_parents ? quarter_node(qm, slot) : 0;
    const unsigned base = grp << 4;                      // first lane of this permutation's group
    u64 s = (live && sub < 8) ? child[node * 8 + sub] : 0;
    const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const unsigned l = sub < 12 ? sub : 0;
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        if (full) {
            const u64* rc = (r == 26) ? D_POSEIDON_RC26_MERGED : (D_POSEIDON_RC + 12 * r);
            s = p9_sbox7(p2_add_canon(s, rc[l]));
        } else {
            u64 t = p9_sbox7(p2_add_canon(s, D_POSEIDON_PARTIAL_A[r - 4]));
            if (l == 0) s = t;
        }
        const u32 lo = (u32)s, hi = (u32)(s >> 32);
        u64 al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            unsigned src = l + i; src = src >= 12 ? src - 12 : src;
            al = p2_madw(__shfl_sync(0xffffffffu, lo, base + src), C[i], al);
            ah = p2_madw(__shfl_sync(0xffffffffu, hi, base + src), C[i], ah);
        }
        if (l == 0) { al = p2_madw(lo, 8u, al); ah = p2_madw(hi, 8u, ah); }
        // value = al + ah * 2^32 (< 2^75), folded as in p9_recombine
        u32 al0 = (u32)al, al1 = (u32)(al >> 32), ah0 = (u32)ah, ah1 = (u32)(ah >> 32);
        u32 o0, o1;
        asm("{\n\t.reg .u32 l1,h,cy,m,e0,e1;\n\t.reg .u64 t;\n\t"
            "add.cc.u32 l1, %3, %4;\n\taddc.u32 h, %5, 0;\n\t"
            "mul.wide.u32 t, h, 0xffffffff;\n\tmov.b64 {e0, e1}, t;\n\t"
            "add.cc.u32 e0, e0, %2;\n\taddc.cc.u32 e1, e1, l1;\n\taddc.u32 cy, 0, 0;\n\t"
            "sub.u32 m, 0, cy;\n\t"
            "add.cc.u32 %0, e0, m;\n\taddc.u32 %1, e1, 0;\n\t}"
            : "=r"(o0), "=r"(o1) : "r"(al0), "r"(al1), "r"(ah0), "r"(ah1));
        s = (u64)o0 | ((u64)o1 << 32);
    }
    if (live && sub < 4) parent[node * 4 + sub] = lz_canon(s);
}

void merkle_alloc(MerkleTreeDev& t, int log_leaves, int cap_height, cudaStream_t s) {
    ZKM_CHECK(cap_height <= log_leaves, "Merkle cap height exceeds tree height");
    t.log_leaves = log_leaves; t.cap_height = cap_height;
    t.level_off.clear();
    size_t off = 0;
    for (int l = 0; l <= log_leaves - cap_height; l++) {
        t.level_off.push_back(off);
        off += ((size_t)1 << (log_leaves - l)) * 4;
    }
    t.digests.alloc(off, s);
}

void merkle_build_from_leaf_digests(MerkleTreeDev& t, cudaStream_t s) {
    const Shard& sh = shard();
    const bool sharded = t.sharded && sh.active();
    QuarterMap qm = {4, {0, 1, 2, 3}, 0};
    int log_segs = 2;
    if (sharded) {
        log_segs = sh.log_segs();
        ZKM_CHECK(t.cap_height >= log_segs, "sharded trees need whole cap subtrees per leaf segment");
        qm.nq = sh.num_owned_segs();
        for (int k = 0; k < qm.nq; k++) qm.q[k] = shard_seg_of(sh.coset_begin() + k, sh.part(), sh.world);
    }
    const int total_segs = 1 << log_segs;
    {
    const double parents = (double)(t.num_leaves() - ((size_t)1 << t.cap_height)) * qm.nq / total_segs;
    ProfScope ps("merkle_levels", s, 96.0 * parents, parents);      // 64 B in + 32 B out, one permutation per parent
    for (int l = 1; l < t.num_levels(); l++) {
        const int log_np = t.log_leaves - l;
        qm.log_qsize = log_np - log_segs;
        size_t np = ((size_t)1 << qm.log_qsize) * qm.nq;             // parents this rank computes on this level
        if (np <= ((size_t)1 << 14)) {
            merkle_level_coop_kernel<<<(unsigned)((np + 7) / 8), 128, 0, s>>>(t.digests.p + t.level_off[l - 1], t.digests.p + t.level_off[l], np, qm);
        } else {
            unsigned blocks = (unsigned)((np + 127) / 128);
            merkle_level_kernel<<<blocks, 128, 0, s>>>(t.digests.p + t.level_off[l - 1], t.digests.p + t.level_off[l], np, qm);
        }
        ZKM_LAUNCHED();
    }
    }
    const size_t ncap = (size_t)4 << t.cap_height;          // words
    t.cap.resize(ncap);
    if (!sharded) {
        ZKM_CUDA(cudaMemcpyAsync(t.cap.data(), t.digests.p + t.level_off.back(), ncap * sizeof(u64), cudaMemcpyDeviceToHost, s));
        ZKM_CUDA(stream_sync(s));
        return;
    }
    // the only exchange of a sharded commitment: every rank contributes the cap entries of its leaf quarters (ncclAllGather,
    // 512 B per tree in total); afterwards all ranks hold the same cap and run the same transcript
    const size_t qwords = ncap / total_segs, mine = qwords * qm.nq;
    DevBuf send(mine, s), recv(mine * sh.world, s);
    for (int k = 0; k < qm.nq; k++)
        ZKM_CUDA(cudaMemcpyAsync(send.p + k * qwords, t.digests.p + t.level_off.back() + (size_t)qm.q[k] * qwords, qwords * sizeof(u64),
                                 cudaMemcpyDeviceToDevice, s));
    shard_all_gather(send.p, recv.p, mine, s);
    std::vector<u64> all(mine * sh.world);
    recv.download(all.data(), all.size());
    for (int r = 0; r < sh.world; r++)
        for (int k = 0; k < qm.nq; k++) {
            Shard other; other.rank = r; other.world = sh.world;
            const int q = shard_seg_of(other.coset_begin() + k, other.part(), sh.world);
            memcpy(t.cap.data() + (size_t)q * qwords, all.data() + ((size_t)r * qm.nq + k) * qwords, qwords * sizeof(u64));
        }
}

void lde_leaf_hash(const u64* lde, size_t col_stride, int ncols, int log_n, int rate_bits, u64* leaf_digests, cudaStream_t s,
                   int coset_begin, int coset_count, int row_log_stride, int row_offset) {
    if (coset_count < 0) coset_count = (1 << rate_bits) - coset_begin;
    const size_t pos_begin = (size_t)coset_begin << log_n, pos_end = (size_t)(coset_begin + coset_count) << log_n;
    size_t N = (pos_end - pos_begin) >> row_log_stride;      // leaves hashed by this call
    unsigned blocks = (unsigned)((N + 127) / 128);
    ProfScope ps("leaf_hash", s, (double)N * (8.0 * ncols + 32.0), ncols > 4 ? (double)N * ((ncols + 7) / 8) : 0.0);
    lde_leaf_hash_kernel<<<blocks, 128, 0, s>>>(lde, col_stride, ncols, log_n, rate_bits, leaf_digests, pos_begin, pos_end, row_log_stride,
                                                row_offset);
    ZKM_LAUNCHED();
}

void rows_leaf_hash(const u64* rows, int width, size_t num_leaves, u64* leaf_digests, cudaStream_t s) {
    unsigned blocks = (unsigned)((num_leaves + 127) / 128);
    ProfScope ps("leaf_hash_rows", s, (double)num_leaves * (8.0 * width + 32.0), width > 4 ? (double)num_leaves * ((width + 7) / 8) : 0.0);
    rows_leaf_hash_kernel<<<blocks, 128, 0, s>>>(rows, width, num_leaves, leaf_digests);
    ZKM_LAUNCHED();
}

struct LevelOffsets { size_t off[40]; };

__global__ void gather_paths_kernel(const u64* __restrict__ dig, LevelOffsets lo, int path_len, const u32* __restrict__ idx, int nq,
                                    u64* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * path_len * 4) return;
    int w = t & 3, l = (t >> 2) % path_len, q = (t >> 2) / path_len;
    size_t node = ((size_t)idx[q] >> l) ^ 1;
    out[t] = dig[lo.off[l] + node * 4 + w];
}

void merkle_gather_paths(const MerkleTreeDev& t, const u32* d_idx, int nq, u64* d_out, cudaStream_t s) {
    int path_len = t.log_leaves - t.cap_height;
    if (path_len == 0 || nq == 0) return;
    LevelOffsets lo;
    for (int l = 0; l < path_len; l++) lo.off[l] = t.level_off[l];
    int total = nq * path_len * 4;
    gather_paths_kernel<<<(total + 255) / 256, 256, 0, s>>>(t.digests.p, lo, path_len, d_idx, nq, d_out);
    ZKM_LAUNCHED();
}

__global__ void gather_rows_kernel(const u64* __restrict__ lde, size_t cs, int ncols, int log_n, int rate_bits,
                                   const u32* __restrict__ idx, int nq, u64* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * ncols) return;
    int q = t / ncols, c = t - q * ncols;
    size_t pos = lde_pos_of_leaf(idx[q], log_n, rate_bits);
    out[t] = lde[(size_t)c * cs + pos];
}

void lde_gather_rows(const u64* lde, size_t col_stride, int ncols, int log_n, int rate_bits, const u32* d_idx, int nq, u64* d_out,
                     cudaStream_t s) {
    int total = nq * ncols;
    if (!total) return;
    gather_rows_kernel<<<(total + 255) / 256, 256, 0, s>>>(lde, col_stride, ncols, log_n, rate_bits, d_idx, nq, d_out);
    ZKM_LAUNCHED();
}

}  // namespace zkm
