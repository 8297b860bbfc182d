// Goldilocks Poseidon permutation (width 12, rate 8, capacity 4; 4 full + 22 partial + 4 full rounds,
// S-box x^7), host+device.  Same function as the reference's in-tree permutation
// (prover/src/poseidon/poseidon_stark.rs:51-95 with constants.rs) and plonky2's PoseidonHash that the
// reference uses for Merkle trees and the Fiat-Shamir challenger (prover.rs:154-163,182-190).
//
// B200 notes: the 12-word state lives in 24 registers of one thread.  Values are kept as
// NON-canonical u64 residues inside the permutation (one canonicalisation at the very end): this
// drops a compare+subtract from every one of the ~1000 modular multiplications.  The MDS layer
// multiplies by the small circulant constants on 32-bit halves with u64 accumulators (IMAD.WIDE
// on the FMA pipe, no carries) and reduces once per output word.
#pragma once
#include "gl.cuh"

namespace zkm {

#define ZKM_POSEIDON_NO_ARRAYS
#include "poseidon_consts.h"
// Host copy (used by the Fiat-Shamir challenger on the CPU side) and device __constant__ copy.
static const u64 H_POSEIDON_RC[360] = POSEIDON_ALL_ROUND_CONSTANTS_INIT;
#ifdef __CUDACC__
static __device__ __constant__ const u64 D_POSEIDON_RC[360] = POSEIDON_ALL_ROUND_CONSTANTS_INIT;
#endif

// ---- lazy (non-canonical) residue arithmetic on raw u64 ----
ZKM_HD u64 lz_reduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;
    u64 t1 = hl * GL_EPS;
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    return r;
}
ZKM_HD u64 lz_mul(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    return lz_reduce128(a * b, __umul64hi(a, b));
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    return lz_reduce128((u64)p, (u64)(p >> 64));
#endif
}
// a: any u64 residue; c: canonical constant (< p)
ZKM_HD u64 lz_add_canon(u64 a, u64 c) {
    u64 s = a + c;
    if (s < a) s += GL_EPS;
    return s;
}
ZKM_HD u64 lz_canon(u64 a) { return a >= GL_P ? a - GL_P : a; }
ZKM_HD u64 lz_sbox7(u64 x) {
    u64 x2 = lz_mul(x, x);
    u64 x3 = lz_mul(x2, x);
    u64 x4 = lz_mul(x2, x2);
    return lz_mul(x3, x4);
}

#ifdef __CUDA_ARCH__
#define ZKM_RC(i) D_POSEIDON_RC[i]
#else
#define ZKM_RC(i) H_POSEIDON_RC[i]
#endif

// out[r] = sum_i s[(i+r)%12]*CIRC[i] + (r==0 ? 8*s[0] : 0), as lazy residues.
ZKM_HD void poseidon_mds(u64* s) {
    // circulant constants as immediates (constants.rs:104-105)
    const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { lo[i] = (u32)s[i]; hi[i] = (u32)(s[i] >> 32); }
    u64 out[12];
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (u64)lo[(i + r) % 12] * C[i];
            ah += (u64)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) { al += (u64)lo[0] * 8u; ah += (u64)hi[0] * 8u; }
        // value = al + ah*2^32  (< 2^74): 96-bit (hi32 : lo64)
        u64 l = al + (ah << 32);
        u32 h = (u32)(ah >> 32) + (l < al ? 1u : 0u);
        u64 t1 = (u64)h * GL_EPS;
        u64 rr = l + t1;
        if (rr < t1) rr += GL_EPS;
        out[r] = rr;
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = out[i];
}

// In-place permutation on 12 canonical words; output canonical.
ZKM_HD void poseidon_permute(u64* s) {
    int rc = 0;
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = lz_sbox7(lz_add_canon(s[i], ZKM_RC(rc + i)));
        rc += 12;
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = lz_add_canon(s[i], ZKM_RC(rc + i));
        rc += 12;
        s[0] = lz_sbox7(s[0]);
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = lz_sbox7(lz_add_canon(s[i], ZKM_RC(rc + i)));
        rc += 12;
        poseidon_mds(s);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = lz_canon(s[i]);
}

// two_to_one / compress (plonky2 hash/hashing.rs compress): state = (l, r, 0,0,0,0)
ZKM_HD void poseidon_two_to_one(const u64* l, const u64* r, u64* out) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) { s[i] = l[i]; s[4 + i] = r[i]; s[8 + i] = 0; }
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = s[i];
}

}  // namespace zkm
