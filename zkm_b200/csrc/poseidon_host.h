// Host-side Goldilocks Poseidon permutation for the Fiat-Shamir transcript (HostChallenger, prover.cu).
// Same function as poseidon.cuh poseidon_permute / the device kernels; this copy is tuned for the CPU because the
// transcript is the one serial host step between GPU phases: a U20 proof absorbs ~18 000 opening values, i.e.
// ~2 200 dependent permutations while the device idles.  It follows the fast schedule of the reference's
// in-tree permutation (prover/src/poseidon/poseidon_stark.rs:76-95,389-400,463-501 with the FAST_PARTIAL_* tables of
// poseidon/constants.rs:107-870): 4 full rounds, one dense 11x11 pre-multiplication, 22 partial rounds that cost one
// S-box, one 12-term dot product and 11 multiply-adds each, 4 full rounds.  The full-round MDS is evaluated in the
// frequency domain of the length-12 cyclic convolution on 32-bit halves (adds and shifts only; derivation in
// poseidon_v2.cuh / tools/gen_mds_freq.py).  Values are lazy residues in [0, 2^64) until the end.
#pragma once
#include "poseidon.cuh"

namespace zkm {
namespace phost {

typedef unsigned __int128 u128;
static const u64 RC[360] = POSEIDON_ALL_ROUND_CONSTANTS_INIT;
static const u64 FIRST[12] = POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT_INIT;
static const u64 PRC[22] = POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS_INIT;
static const u64 VS[22 * 11] = POSEIDON_FAST_PARTIAL_ROUND_VS_INIT;
static const u64 WHAT[22 * 11] = POSEIDON_FAST_PARTIAL_ROUND_W_HATS_INIT;
static const u64 INIT[11 * 11] = POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX_INIT;

// x (< 2^128) -> lazy residue.  Branch-free: the wrap of t0 + t1 happens for about half of all inputs, a conditional jump
// there mispredicts constantly (measured 10 ns per multiplication with branches, 2.5 ns without).
static inline u64 red128(u128 x) {
    const u64 lo = (u64)x, hi = (u64)(x >> 64);
    const u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    t0 -= (0 - (u64)(lo < hh)) & GL_EPS;
    const u64 t1 = hl * GL_EPS;
    u64 r = t0 + t1;
    r += (0 - (u64)(r < t1)) & GL_EPS;
    return r;
}
static inline u64 mul(u64 a, u64 b) { return red128((u128)a * b); }
// lazy a + lazy b
static inline u64 add(u64 a, u64 b) {
    u64 s = a + b;
    const u64 c1 = (0 - (u64)(s < a)) & GL_EPS;     // wrapped: 2^64 = 2^32 - 1
    s += c1;
    const u64 c2 = (0 - (u64)(s < c1)) & GL_EPS;    // adding the correction wrapped again (needs a, b both near 2^64)
    return s + c2;
}
static inline u64 sbox(u64 x) {
    u64 x2 = mul(x, x), x3 = mul(x2, x), x4 = mul(x2, x2);
    return mul(x3, x4);
}
// exact circulant MDS on 12 small non-negative integers (< 2^33): out[r] = sum_i C[i] x[(i+r)%12] + (r==0) 8 x[0]
static inline void mds_half(const int64_t* x, int64_t* o) {
    int64_t S1[3], Sm[3], Sr[3], Si[3];
    for (int b = 0; b < 3; b++) {
        int64_t e0 = x[b] + x[6 + b], e1 = x[3 + b] + x[9 + b];
        Sr[b] = x[b] - x[6 + b]; Si[b] = x[3 + b] - x[9 + b];
        S1[b] = e0 + e1; Sm[b] = e0 - e1;
    }
    const int64_t t = S1[0] + S1[1] + S1[2];
    const int64_t A0 = (t + S1[2]) * 16, A1 = (t + S1[0]) * 16, A2 = (t + S1[1]) * 16;
    const int64_t B0 = Sm[2] * 8 - Sm[1] * 2 - Sm[0];
    const int64_t B1 = -Sm[0] * 8 - Sm[2] * 2 - Sm[1];
    const int64_t B2 = -Sm[1] * 8 + Sm[0] * 2 - Sm[2];
    const int64_t P0 = Si[2] * 4 - Si[1] * 16 + Sr[0] * 2 + Sr[1] + Sr[2] - Si[0];
    const int64_t Q0 = -Sr[2] * 4 + Sr[1] * 16 + Si[0] * 2 + Si[1] + Si[2] + Sr[0];
    const int64_t P1 = -Si[2] * 16 + Sr[1] * 2 - Sr[0] * 4 + Sr[2] + Si[0] - Si[1];
    const int64_t Q1 = Sr[2] * 16 + Si[1] * 2 - Si[0] * 4 + Si[2] - Sr[0] + Sr[1];
    const int64_t P2 = Sr[2] * 2 - Sr[1] * 4 + Sr[0] * 16 + Si[0] + Si[1] - Si[2];
    const int64_t Q2 = Si[2] * 2 - Si[1] * 4 + Si[0] * 16 + Sr[2] - Sr[0] - Sr[1];
    const int64_t u0 = A0 + B0, v0 = A0 - B0, u1 = A1 + B1, v1 = A1 - B1, u2 = A2 + B2, v2 = A2 - B2;
    o[0] = u0 + P0 + x[0] * 8; o[3] = v0 + Q0; o[6] = u0 - P0; o[9] = v0 - Q0;
    o[1] = u1 + P1; o[4] = v1 + Q1; o[7] = u1 - P1; o[10] = v1 - Q1;
    o[2] = u2 + P2; o[5] = v2 + Q2; o[8] = u2 - P2; o[11] = v2 - Q2;
}
static inline void mds(u64* s) {
    int64_t lo[12], hi[12], al[12], ah[12];
    for (int i = 0; i < 12; i++) { lo[i] = (int64_t)(s[i] & GL_EPS); hi[i] = (int64_t)(s[i] >> 32); }
    mds_half(lo, al);
    mds_half(hi, ah);
    for (int i = 0; i < 12; i++) s[i] = red128((u128)(u64)al[i] + ((u128)(u64)ah[i] << 32));      // al, ah < 2^41
}
static inline void full_round(u64* s, const u64* rc) {
    for (int i = 0; i < 12; i++) s[i] = sbox(add(s[i], rc[i]));
    mds(s);
}

}  // namespace phost

// In-place permutation on 12 canonical words; output canonical.
static inline void poseidon_permute_host(u64* s) {
    using namespace phost;
    for (int r = 0; r < 4; r++) full_round(s, RC + 12 * r);
    // partial_first_constant_layer + mds_partial_layer_init
    for (int i = 0; i < 12; i++) s[i] = add(s[i], FIRST[i]);
    {
        u64 t[12];
        t[0] = s[0];
        for (int c = 1; c < 12; c++) {
            u128 lo = 0, hi = 0;                        // sum_r s[r] * INIT[r-1][c-1], 11 terms: no overflow in 128 bits each
            for (int r = 1; r < 12; r++) { u128 p = (u128)s[r] * INIT[(r - 1) * 11 + (c - 1)]; lo += (u64)p; hi += (u64)(p >> 64); }
            // total = lo + hi * 2^64 with lo, hi < 2^68
            u64 a = red128(lo);
            u64 b = red128(hi * (u128)GL_EPS);          // hi * 2^64 = hi * (2^32 - 1)  (mod p), < 2^100
            t[c] = add(a, b);
        }
        for (int i = 0; i < 12; i++) s[i] = t[i];
    }
    for (int i = 0; i < 22; i++) {
        const u64 s0 = add(sbox(s[0]), PRC[i]);
        // mds_partial_layer_fast: d = s0 * (CIRC[0] + DIAG[0]) + sum_j s[j] * W_HAT[i][j-1];  s[j] += s0 * VS[i][j-1]
        u128 lo = (u128)s0 * 25u, hi = 0;
        for (int j = 1; j < 12; j++) { u128 p = (u128)s[j] * WHAT[i * 11 + j - 1]; lo += (u64)p; hi += (u64)(p >> 64); }
        const u64 d = add(red128(lo), red128(hi * (u128)GL_EPS));
        for (int j = 1; j < 12; j++) s[j] = add(s[j], mul(s0, VS[i * 11 + j - 1]));
        s[0] = d;
    }
    for (int r = 0; r < 4; r++) full_round(s, RC + 12 * (26 + r));
    for (int i = 0; i < 12; i++) s[i] = lz_canon(s[i]);
}

}  // namespace zkm
