// Device-side generators of the six hash-precompile tables that tracegen.cu does not cover (the rest of
// Traces::into_tables, reference witness/traces.rs:271-301): PoseidonSponge, KeccakSponge, ShaExtend, ShaExtendSponge,
// ShaCompress, ShaCompressSponge.  As in tracegen.cu only the operation log crosses PCIe and the table is built column-major
// in HBM.  tests/hash_gen.py is the Python restatement of the reference generators these kernels are checked against.
//
// Fixed-width logs, one entry per ROW (u64 words):
//   ShaExtend (6)          5: w[i-15], w[i-2], w[i-16], w[i-7] (u32), timestamp          sha_extend_stark.rs:122-237
//   ShaExtendSponge (7)   13: round i (0..47), the same four words, input_virt[4], output_virt, context, segment, timestamp
//                                                                                         sha_extend_sponge_stark.rs:128-227
//   ShaCompress (8)       15: a..h (u32), w_i, k_i, round (0..64), w_i_virt, segment, context, timestamp
//                             -- the reference's ([u8; 41], MemoryAddress, usize) per row   sha_compress_stark.rs:234-391
//   ShaCompressSponge (9) 86: hx[8], w[64] (u32), hx_virt[8], w_start virt / segment / context, context, segment, timestamp
//                                                                                         sha_compress_sponge_stark.rs:120-237
// Variable-width logs (the two byte sponges): word 0 = total number of words of the log, then per operation
//   context, segment, timestamp, len (bytes), n_addr, virt[n_addr], the input bytes packed little endian 8 per word
// (KeccakSpongeOp / PoseidonSpongeOp {base_address, timestamp, input}: context and segment are base_address[0]'s, virt[k] =
// base_address[k].virt).  An operation gives len / rate + 1 rows that chain through the sponge state, so one thread walks one
// operation; these tables have 2^6..2^12 rows in every real segment.
#include "dev.cuh"
#include "poseidon_v2.cuh"
#include "tables/keccak.h"
#include "tables/keccak_sponge.h"
#include "tables/poseidon_sponge.h"
#include "tables/sha_extend.h"
#include "tables/sha_extend_sponge.h"
#include "tables/sha_compress.h"
#include "tables/sha_compress_sponge.h"

namespace zkm {
namespace {

size_t padded_rows(size_t n_rows, size_t min_rows) {
    size_t n = n_rows > min_rows ? n_rows : min_rows, p = 1;
    while (p < n) p <<= 1;
    return p;
}

struct ColWriter {
    u64* cols; size_t n, row;
    __device__ __forceinline__ void put(int c, u64 v) const { cols[(size_t)c * n + row] = v; }
    __device__ __forceinline__ void le4(int c, u32 v) const {
#pragma unroll
        for (int i = 0; i < 4; i++) put(c + i, (v >> (8 * i)) & 0xFF);
    }
};
__device__ __forceinline__ u32 rotr32(u32 v, int r) { return r ? (v >> r) | (v << (32 - r)) : v; }
// RotateRightOp / ShiftRightOp::generate_trace (sha_extend/rotate_right.rs:14-27,108-117, shift_right.rs:15-28):
// value = the rotated / shifted word as bytes, shift = input >> r, carry = the r low bits
__device__ __forceinline__ u32 shift_op(const ColWriter& w, int at, u32 v, int r, bool rotate) {
    const u32 out = rotate ? rotr32(v, r) : v >> r;
    w.le4(at, out);
    w.put(at + 4, v >> r);
    w.put(at + 5, v & ((1u << r) - 1));
    return out;
}
// WrappingAddNOp::generate_trace (wrapping_add_2.rs / _4.rs / _5.rs): value bytes + one-hot carry
__device__ __forceinline__ u32 wrapping_add(const ColWriter& w, int at, u64 total) {
    w.le4(at, (u32)total);
    w.put(at + 4 + (int)(total >> 32), 1);
    return (u32)total;
}

// ---------------------------------------------------------------------------------------------------------- ShaExtend
__global__ void sha_extend_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols, unsigned* bad) {
    namespace se = tables::sha_extend;
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ops) return;
    const u64* o = ops + 5 * r;
    if ((o[0] | o[1] | o[2] | o[3]) >> 32) { atomicExch(bad, 1u); return; }
    const u32 w15 = (u32)o[0], w2 = (u32)o[1], w16 = (u32)o[2], w7 = (u32)o[3];
    ColWriter w{cols, n, r};
    w.put(se::TIMESTAMP, o[4]); w.put(se::IS_REAL_ROUND, 1);
    w.le4(se::W_I_MINUS_15, w15); w.le4(se::W_I_MINUS_2, w2); w.le4(se::W_I_MINUS_16, w16); w.le4(se::W_I_MINUS_7, w7);
    const u32 rr7 = shift_op(w, se::W_I_MINUS_15_RR_7, w15, 7, true), rr18 = shift_op(w, se::W_I_MINUS_15_RR_18, w15, 18, true),
              rs3 = shift_op(w, se::W_I_MINUS_15_RS_3, w15, 3, false);
    const u32 s0i = rr7 ^ rr18, s0 = s0i ^ rs3;
    const u32 rr17 = shift_op(w, se::W_I_MINUS_2_RR_17, w2, 17, true), rr19 = shift_op(w, se::W_I_MINUS_2_RR_19, w2, 19, true),
              rs10 = shift_op(w, se::W_I_MINUS_2_RS_10, w2, 10, false);
    const u32 s1i = rr17 ^ rr19, s1 = s1i ^ rs10;
    w.le4(se::S_0_INTER, s0i); w.le4(se::S_0, s0); w.le4(se::S_1_INTER, s1i); w.le4(se::S_1, s1);
    wrapping_add(w, se::W_I_VALUE, (u64)s1 + w7 + s0 + w16);
}

// ---------------------------------------------------------------------------------------------------- ShaExtendSponge
__global__ void sha_extend_sponge_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols, unsigned* bad) {
    namespace ss = tables::sha_extend_sponge;
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ops) return;
    const u64* o = ops + 13 * r;
    if (o[0] >= (u64)ss::NUM_ROUNDS || (o[1] | o[2] | o[3] | o[4]) >> 32) { atomicExch(bad, 1u); return; }
    const u32 w15 = (u32)o[1], w2 = (u32)o[2], w16 = (u32)o[3], w7 = (u32)o[4];
    ColWriter w{cols, n, r};
    w.put(ss::ROUND + (int)o[0], 1);
    w.le4(ss::W_I_MINUS_15, w15); w.le4(ss::W_I_MINUS_2, w2); w.le4(ss::W_I_MINUS_16, w16); w.le4(ss::W_I_MINUS_7, w7);
    const u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3), s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
    w.le4(ss::W_I, s1 + w16 + s0 + w7);                                           // compute_w_i :208-222
#pragma unroll
    for (int i = 0; i < 4; i++) w.put(ss::INPUT_VIRT + i, o[5 + i]);
    w.put(ss::OUTPUT_VIRT, o[9]); w.put(ss::CONTEXT, o[10]); w.put(ss::SEGMENT, o[11]); w.put(ss::TIMESTAMP, o[12]);
}

// -------------------------------------------------------------------------------------------------------- ShaCompress
__global__ void sha_compress_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols, unsigned* bad) {
    namespace sc = tables::sha_compress;
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ops) return;
    const u64* o = ops + 15 * r;
    u64 any = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) any |= o[i];
    if (any >> 32 || o[10] >= (u64)sc::NUM_COMPRESS_ROWS) { atomicExch(bad, 1u); return; }
    const u32 a = (u32)o[0], b = (u32)o[1], c = (u32)o[2], d = (u32)o[3], e = (u32)o[4], f = (u32)o[5], g = (u32)o[6], h = (u32)o[7],
              w_i = (u32)o[8], k_i = (u32)o[9];
    ColWriter w{cols, n, r};
    w.put(sc::W_I_VIRT, o[11]); w.put(sc::SEGMENT, o[12]); w.put(sc::CONTEXT, o[13]); w.put(sc::TIMESTAMP, o[14]);
    w.put(sc::ROUND + (int)o[10], 1);
#pragma unroll
    for (int i = 0; i < 8; i++) w.le4(sc::state4(i), (u32)o[i]);
    w.le4(sc::W_I, w_i); w.le4(sc::K_I, k_i);
    const u32 rr6 = shift_op(w, sc::E_RR_6, e, 6, true), rr11 = shift_op(w, sc::E_RR_11, e, 11, true), rr25 = shift_op(w, sc::E_RR_25, e, 25, true);
    const u32 s1i = rr6 ^ rr11, s1 = s1i ^ rr25, e_and_f = e & f, e_not = ~e, enag = e_not & g, ch = e_and_f ^ enag;
    w.le4(sc::S_1_INTER, s1i); w.le4(sc::S_1, s1); w.le4(sc::E_AND_F, e_and_f); w.le4(sc::E_NOT, e_not);
    w.le4(sc::E_NOT_AND_G, enag); w.le4(sc::CH, ch);
    const u32 temp1 = wrapping_add(w, sc::TEMP1, (u64)h + s1 + ch + k_i + w_i);
    const u32 rr2 = shift_op(w, sc::A_RR_2, a, 2, true), rr13 = shift_op(w, sc::A_RR_13, a, 13, true), rr22 = shift_op(w, sc::A_RR_22, a, 22, true);
    const u32 s0i = rr2 ^ rr13, s0 = s0i ^ rr22, ab = a & b, ac = a & c, bc = b & c, mi = ab ^ ac, maj = mi ^ bc;
    w.le4(sc::S_0_INTER, s0i); w.le4(sc::S_0, s0); w.le4(sc::A_AND_B, ab); w.le4(sc::A_AND_C, ac); w.le4(sc::B_AND_C, bc);
    w.le4(sc::MAJ_INTER, mi); w.le4(sc::MAJ, maj);
    const u32 temp2 = wrapping_add(w, sc::TEMP2, (u64)s0 + maj);
    wrapping_add(w, sc::D_ADD_TEMP1, (u64)d + temp1);
    wrapping_add(w, sc::TEMP1_ADD_TEMP2, (u64)temp1 + temp2);
}

// -------------------------------------------------------------------------------------------------- ShaCompressSponge
static __device__ __constant__ const u32 D_SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__global__ void sha_compress_sponge_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols, unsigned* bad) {
    namespace sp = tables::sha_compress_sponge;
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ops) return;
    const u64* o = ops + 86 * r;
    u64 any = 0;
    for (int i = 0; i < 72; i++) any |= o[i];
    if (any >> 32) { atomicExch(bad, 1u); return; }
    ColWriter w{cols, n, r};
    u32 st[8], hx[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { hx[i] = st[i] = (u32)o[i]; w.le4(sp::HX + 4 * i, hx[i]); w.put(sp::HX_VIRT + i, o[72 + i]); }
    w.put(sp::W_START_VIRT, o[80]); w.put(sp::W_START_SEGMENT, o[81]); w.put(sp::W_START_CONTEXT, o[82]);
    w.put(sp::CONTEXT, o[83]); w.put(sp::SEGMENT, o[84]); w.put(sp::TIMESTAMP, o[85]); w.put(sp::IS_REAL_ROUND, 1);
#pragma unroll 1
    for (int i = 0; i < 64; i++) {                                                  // compress :195-223
        const u32 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
        const u32 s1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25), ch = (e & f) ^ (~e & g);
        const u32 temp1 = h + s1 + ch + D_SHA_K[i] + (u32)o[8 + i];
        const u32 s0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22), maj = (a & b) ^ (a & c) ^ (b & c);
        st[7] = g; st[6] = f; st[5] = e; st[4] = d + temp1; st[3] = c; st[2] = b; st[1] = a; st[0] = temp1 + s0 + maj;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        w.le4(sp::OUTPUT_STATE + 4 * i, st[i]);
        wrapping_add(w, sp::OUTPUT_HX + 6 * i, (u64)hx[i] + st[i]);
    }
}

// ------------------------------------------------------------------------------------------------------ byte sponges
// Index of one operation of a variable-width log: word offset of its record, first row of the table it fills.
struct SpongeIndex { std::vector<u64> idx; size_t rows = 0; };
SpongeIndex index_sponge_log(const u64* log, size_t n_ops, size_t rate_bytes, const char* what) {
    SpongeIndex si;
    ZKM_CHECK(log || n_ops == 0, "null operation log");
    const size_t total = n_ops ? (size_t)log[0] : 1;
    size_t at = 1;
    for (size_t k = 0; k < n_ops; k++) {
        ZKM_CHECK(at + 5 <= total, std::string(what) + " log is truncated");
        const size_t len = (size_t)log[at + 3], n_addr = (size_t)log[at + 4];
        ZKM_CHECK(len < ((size_t)1 << 32) && n_addr >= 1 && n_addr < ((size_t)1 << 24), std::string(what) + " log: bad length or address count");
        const size_t words = 5 + n_addr + (len + 7) / 8;
        ZKM_CHECK(at + words <= total, std::string(what) + " log is truncated");
        si.idx.push_back(at); si.idx.push_back(si.rows);
        si.rows += len / rate_bytes + 1;
        at += words;
    }
    ZKM_CHECK(at == total, std::string(what) + " log: the word count in word 0 does not match its records");
    return si;
}
__device__ __forceinline__ u32 log_byte(const u64* data, size_t i) { return (u32)(data[i >> 3] >> (8 * (i & 7))) & 0xFF; }

// Fields common to both sponges' rows (generate_common_fields: keccak_sponge_stark.rs:351-375, poseidon_sponge_stark.rs:302-326)
// and the block bytes with the pad10*1 rule of the final row (:319-349 / :270-300).  Returns the padded block as u32 words.
template <int RATE_BYTES, int C_FULL, int C_CONTEXT, int C_VIRT, int C_TIMESTAMP, int C_LEN, int C_ALREADY, int C_FINAL_LEN, int C_BLOCK>
__device__ __forceinline__ void sponge_row_head(const ColWriter& w, const u64* rec, size_t block, u32* words) {
    constexpr int RATE_U32S = RATE_BYTES / 4;
    const size_t len = (size_t)rec[3], n_addr = (size_t)rec[4];
    const u64* virt = rec + 5;
    const u64* data = virt + n_addr;
    const size_t already = block * RATE_BYTES, nfull = len / RATE_BYTES;
    w.put(C_CONTEXT, rec[0]); w.put(C_CONTEXT + 1, rec[1]);                  // context, segment are adjacent columns
    w.put(C_TIMESTAMP, rec[2]); w.put(C_LEN, len); w.put(C_ALREADY, already);
    const size_t idx = already / 4;
    size_t end = (already + RATE_BYTES) / 4;
    if (end > n_addr) end = n_addr;
    for (size_t i = idx; i < end; i++) w.put(C_VIRT + (int)(i - idx), virt[i]);
    const bool full = block < nfull;
    const int have = full ? RATE_BYTES : (int)(len - already);
    if (full) w.put(C_FULL, 1); else w.put(C_FINAL_LEN + have, 1);
#pragma unroll 1
    for (int i = 0; i < RATE_U32S; i++) {
        u32 word = 0;
        for (int j = 0; j < 4; j++) {
            const int at = 4 * i + j;
            u32 byte = at < have ? log_byte(data, already + at) : 0;
            if (!full) {
                if (have == RATE_BYTES - 1) { if (at == have) byte = 0x81; }
                else { if (at == have) byte = 1; if (at == RATE_BYTES - 1) byte = 0x80; }
            }
            if (byte) w.put(C_BLOCK + at, byte);
            word |= byte << (8 * j);
        }
        words[i] = word;
    }
}

__device__ void keccakf_u32s(u32* s) {                                             // cpu/kernel/keccak_util.rs:6-18
    using namespace tables::keccak;                                                 // ZKM_K(KECCAK_R), ZKM_K(KECCAK_RC)
    u64 A[25];                                                                      // A[x + 5 y]
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = (u64)s[2 * i] | ((u64)s[2 * i + 1] << 32);
    auto rotl = [](u64 v, int r) { r &= 63; return r ? (v << r) | (v >> (64 - r)) : v; };
#pragma unroll 1
    for (int rnd = 0; rnd < 24; rnd++) {
        u64 C[5], D[5], B[25];
#pragma unroll
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl(C[(x + 1) % 5], 1);
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(A[x + 5 * y] ^ D[x], (int)ZKM_K(KECCAK_R)[x * 5 + y]);
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= ZKM_K(KECCAK_RC)[rnd];
    }
#pragma unroll
    for (int i = 0; i < 25; i++) { s[2 * i] = (u32)A[i]; s[2 * i + 1] = (u32)(A[i] >> 32); }
}

__global__ void __launch_bounds__(64) keccak_sponge_rows_kernel(const u64* __restrict__ log, const u64* __restrict__ index, size_t n_ops, size_t n,
                                                                u64* __restrict__ cols) {
    namespace ks = tables::keccak_sponge;
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ops) return;
    const u64* rec = log + index[2 * k];
    const size_t row0 = (size_t)index[2 * k + 1], nblocks = (size_t)rec[3] / ks::KECCAK_RATE_BYTES + 1;
    u32 state[50];
#pragma unroll 1
    for (int i = 0; i < 50; i++) state[i] = 0;
#pragma unroll 1
    for (size_t b = 0; b < nblocks; b++) {
        ColWriter w{cols, n, row0 + b};
#pragma unroll 1
        for (int i = 0; i < ks::KECCAK_RATE_U32S; i++) if (state[i]) w.put(ks::ORIGINAL_RATE_U32S + i, state[i]);
#pragma unroll 1
        for (int i = 0; i < ks::KECCAK_CAPACITY_U32S; i++) if (state[ks::KECCAK_RATE_U32S + i]) w.put(ks::ORIGINAL_CAPACITY_U32S + i, state[ks::KECCAK_RATE_U32S + i]);
        u32 words[ks::KECCAK_RATE_U32S];
        sponge_row_head<ks::KECCAK_RATE_BYTES, ks::IS_FULL_INPUT_BLOCK, ks::CONTEXT, ks::VIRT, ks::TIMESTAMP, ks::LEN, ks::ALREADY_ABSORBED_BYTES,
                        ks::IS_FINAL_INPUT_LEN, ks::BLOCK_BYTES>(w, rec, b, words);
#pragma unroll 1
        for (int i = 0; i < ks::KECCAK_RATE_U32S; i++) { state[i] ^= words[i]; w.put(ks::XORED_RATE_U32S + i, state[i]); }
        keccakf_u32s(state);
#pragma unroll 1
        for (int i = 0; i < ks::KECCAK_WIDTH_MINUS_DIGEST_U32S; i++) w.put(ks::PARTIAL_UPDATED_STATE_U32S + i, state[ks::KECCAK_DIGEST_U32S + i]);
#pragma unroll 1
        for (int l = 0; l < ks::KECCAK_DIGEST_U32S; l++) w.le4(ks::UPDATED_DIGEST_STATE_BYTES + 4 * l, state[l]);
    }
}

__global__ void __launch_bounds__(64) poseidon_sponge_rows_kernel(const u64* __restrict__ log, const u64* __restrict__ index, size_t n_ops, size_t n,
                                                                  u64* __restrict__ cols) {
    namespace ps = tables::poseidon_sponge;
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ops) return;
    const u64* rec = log + index[2 * k];
    const size_t row0 = (size_t)index[2 * k + 1], nblocks = (size_t)rec[3] / ps::POSEIDON_RATE_BYTES + 1;
    u64 state[12];
#pragma unroll
    for (int i = 0; i < 12; i++) state[i] = 0;
#pragma unroll 1
    for (size_t b = 0; b < nblocks; b++) {
        ColWriter w{cols, n, row0 + b};
#pragma unroll
        for (int i = 0; i < ps::SPONGE_RATE; i++) w.put(ps::ORIGINAL_RATE + i, state[i]);
#pragma unroll
        for (int i = 0; i < ps::SPONGE_CAPACITY; i++) w.put(ps::ORIGINAL_CAPACITY + i, state[ps::SPONGE_RATE + i]);
        u32 words[ps::SPONGE_RATE];
        sponge_row_head<ps::POSEIDON_RATE_BYTES, ps::IS_FULL_INPUT_BLOCK, ps::CONTEXT, ps::VIRT, ps::TIMESTAMP, ps::LEN, ps::ALREADY_ABSORBED_BYTES,
                        ps::IS_FINAL_INPUT_LEN, ps::BLOCK_BYTES>(w, rec, b, words);
#pragma unroll
        for (int i = 0; i < ps::SPONGE_RATE; i++) { state[i] = words[i]; w.put(ps::NEW_RATE + i, words[i]); }   // the block OVERWRITES the rate
        poseidon_permute_dev(state);
#pragma unroll
        for (int i = 0; i < ps::WIDTH_MINUS_DIGEST; i++) w.put(ps::PARTIAL_UPDATED_STATE + i, state[ps::POSEIDON_DIGEST + i]);
#pragma unroll
        for (int i = 0; i < ps::POSEIDON_DIGEST; i++) w.put(ps::UPDATED_DIGEST_STATE + i, state[i]);
    }
}

template <class Kernel>
size_t fixed_log_table(Kernel kernel, const char* family, const char* range_message, int words_per_op, int ncols, const u64* h_ops, size_t n_ops,
                       size_t min_rows, DevBuf& cols, cudaStream_t s) {
    ZKM_CHECK(n_ops <= ((size_t)1 << 24), "too many operations");
    const size_t n = padded_rows(n_ops, min_rows);
    DevBuf ops((size_t)words_per_op * n_ops + 1, s), flag(1, s);
    if (n_ops) ops.upload(h_ops, (size_t)words_per_op * n_ops);
    flag.zero();
    cols.alloc((size_t)ncols * n, s);
    cols.zero();                                   // padding rows are all-zero in all four tables; the kernels write non-zero cells only where needed
    ProfScope ps(family, s, 8.0 * words_per_op * (double)n_ops + 8.0 * ncols * (double)n);
    if (n_ops) {
        kernel<<<(unsigned)((n_ops + 127) / 128), 128, 0, s>>>(ops.p, n_ops, n, cols.p, (unsigned*)flag.p);
        ZKM_LAUNCHED();
    }
    u64 bad = 0;
    flag.download(&bad, 1);
    ZKM_CHECK((unsigned)bad == 0, range_message);
    return n;
}

template <class Kernel>
size_t sponge_log_table(Kernel kernel, const char* family, const char* what, size_t rate_bytes, int ncols, const u64* h_log, size_t n_ops,
                        size_t min_rows, DevBuf& cols, cudaStream_t s) {
    ZKM_CHECK(n_ops <= ((size_t)1 << 22), "too many operations");
    SpongeIndex si = index_sponge_log(h_log, n_ops, rate_bytes, what);
    const size_t n = padded_rows(si.rows, min_rows);
    ZKM_CHECK(n <= ((size_t)1 << 24), "sponge table too large");
    const size_t total = n_ops ? (size_t)h_log[0] : 1;
    DevBuf log(total, s), index(si.idx.size() + 1, s);
    if (n_ops) { log.upload(h_log, total); index.upload(si.idx.data(), si.idx.size()); }
    cols.alloc((size_t)ncols * n, s);
    cols.zero();                                   // generate_padding_row = the default (all-zero) view
    ProfScope ps(family, s, 8.0 * (double)total + 8.0 * ncols * (double)n);
    if (n_ops) {
        kernel<<<(unsigned)((n_ops + 63) / 64), 64, 0, s>>>(log.p, index.p, n_ops, n, cols.p);
        ZKM_LAUNCHED();
        ZKM_CUDA(stream_sync(s));        // si.idx is host memory of this frame
    }
    return n;
}

}  // namespace

size_t sha_extend_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    return fixed_log_table(sha_extend_rows_kernel, "sha_extend_trace", "sha extend input is not a 32-bit word", 5, tables::sha_extend::NUM_COLUMNS,
                           h_ops, n_ops, min_rows, cols, s);
}
size_t sha_extend_sponge_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    return fixed_log_table(sha_extend_sponge_rows_kernel, "sha_extend_sponge_trace", "sha extend sponge operation out of range (round 0..47, 32-bit words)",
                           13, tables::sha_extend_sponge::NUM_COLUMNS, h_ops, n_ops, min_rows, cols, s);
}
size_t sha_compress_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    return fixed_log_table(sha_compress_rows_kernel, "sha_compress_trace", "sha compress row out of range (round 0..64, 32-bit words)", 15,
                           tables::sha_compress::NUM_COLUMNS, h_ops, n_ops, min_rows, cols, s);
}
size_t sha_compress_sponge_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    return fixed_log_table(sha_compress_sponge_rows_kernel, "sha_compress_sponge_trace", "sha compress sponge input is not a 32-bit word", 86,
                           tables::sha_compress_sponge::NUM_COLUMNS, h_ops, n_ops, min_rows, cols, s);
}
size_t keccak_sponge_generate_trace_dev(const u64* h_log, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    return sponge_log_table(keccak_sponge_rows_kernel, "keccak_sponge_trace", "keccak sponge", tables::keccak_sponge::KECCAK_RATE_BYTES,
                            tables::keccak_sponge::NUM_COLUMNS, h_log, n_ops, min_rows, cols, s);
}
size_t poseidon_sponge_generate_trace_dev(const u64* h_log, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    return sponge_log_table(poseidon_sponge_rows_kernel, "poseidon_sponge_trace", "poseidon sponge", tables::poseidon_sponge::POSEIDON_RATE_BYTES,
                            tables::poseidon_sponge::NUM_COLUMNS, h_log, n_ops, min_rows, cols, s);
}

}  // namespace zkm
