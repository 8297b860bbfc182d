// Device Goldilocks Poseidon permutation (same function as poseidon.cuh poseidon_permute, which stays as the portable
// reference the micro-benchmark checks against).  What the SASS of the portable version led to, in order of effect
// (profiles/r1_poseidon_variants_microbench.txt has every intermediate variant's throughput; the superseded code was removed):
//  * the 128-bit product from the compiler's own 64x64 multiply (IMAD.WIDE.U32 with carry-out predicate + IMAD.WIDE.U32.X:
//    4 IMAD + 2 IADD3, a form PTX cannot express), the 2^64 = 2^32 - 1 / 2^96 = -1 folding as one hand-written carry chain,
//    values kept as lazy residues in [0, 2^64) until the end;
//  * the MDS layer on the FP64 pipe, in the frequency domain of the length-12 cyclic convolution (below);
//  * partial rounds in pairs, merged partial-round constants (tools/gen_poseidon_merged.py), one rolled round loop.
#pragma once
#include "poseidon.cuh"

namespace zkm {
#ifdef __CUDACC__

// a + c for a lazy residue a and a canonical constant c
__device__ __forceinline__ u64 p2_add_canon(u64 a, u64 c) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), c0 = (u32)c, c1 = (u32)(c >> 32);
    u32 o0, o1;
    asm("{\n\t.reg .u32 t0,t1,cy,m;\n\t"
        "add.cc.u32 t0, %2, %4;\n\taddc.cc.u32 t1, %3, %5;\n\taddc.u32 cy, 0, 0;\n\t"
        "sub.u32 m, 0, cy;\n\t"
        "add.cc.u32 %0, t0, m;\n\taddc.u32 %1, t1, 0;\n\t}"
        : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(c0), "r"(c1));
    return (u64)o0 | ((u64)o1 << 32);
}
__device__ __forceinline__ u64 p2_madw(u32 a, u32 c, u64 acc) {
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(c), "l"(acc));
    return r;
}
// u32 -> double by the 2^52 magic-number packing (one DADD); p9_cvt selects between this and I2F
__device__ __forceinline__ double p3_u32_to_f64(u32 x) { return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0; }
// Merged partial-round constants.  In a partial round only element 0 passes the S-box, so the constants of elements 1..11
// commute with the (linear) MDS layer: writing the state as u + k with k a data-independent vector, k_{r+1} = M * (0, (k_r +
// c_r)[1..11]), only a_r = (k_r + c_r)[0] has to be added (to element 0, before its S-box), and the residue k_26 is folded
// into the constants of full round 26.  22 additions instead of 264 per permutation; identical outputs (derivation + check:
// tools/gen_poseidon_merged.py).
static __device__ __constant__ const u64 D_POSEIDON_PARTIAL_A[22] = {0x3cc3f892184df408ULL, 0x6754826bf0555feaULL, 0x07f136d86fe52ec6ULL, 0xcc31104c136e624cULL, 0xe75f601068e70acaULL, 0x27bb558ab181ed5eULL, 0xe0593bc645a018abULL, 0xa7ef236c2c5f0e1fULL, 0x2f2ed40f0e211d79ULL, 0x65b0ab09bec15af9ULL, 0x28f0f3bb03d3d776ULL, 0xb60fb82205a86176ULL, 0x6685ae6e5db8023dULL, 0x9b2390b07020a27cULL, 0xc3607e7232b11cefULL, 0xf618ae7058499e24ULL, 0xc18226dd334780c2ULL, 0xb57bff1387506176ULL, 0xec2475152d5a08ffULL, 0x04df43ffd0b458ffULL, 0x10f46236adcc3e98ULL, 0x52588ae3575e2ce9ULL};
static __device__ __constant__ const u64 D_POSEIDON_RC26_MERGED[12] = {0x5405cc09b3ff0c06ULL, 0xe14dc071ace29846ULL, 0xbb56729c7877aa9eULL, 0x474fb5726a0068f3ULL, 0x2629b158383529bfULL, 0xe1cabee6fa7a9532ULL, 0xced9e7d28e6a6de9ULL, 0xd0fd98f1e129850fULL, 0x9689ab45a6d09dd7ULL, 0xba9673a4862f9848ULL, 0xa5c3a8c0fcbdbd41ULL, 0x8f4411a1226beb35ULL};
// ---- v9: fewer instructions per permutation (the kernels are issue-slot bound, ncu: profiles/r1_*):
//  (a) the 128-bit product comes from the compiler's own 64x64 multiply (IMAD.WIDE.U32 with carry-out predicate +
//      IMAD.WIDE.U32.X carry-in: 4 IMAD + 2 IADD3, not reachable from PTX mad/add.cc), the 2^64 = 2^32-1 / 2^96 = -1
//      folding stays a hand-written carry chain;
//  (b) the circulant MDS is evaluated in the "frequency domain" of the length-12 cyclic convolution
//      (x^12 - 1 = prod_{y^4=1} (x^3 - y): 4-point DFTs of the three stride-3 subsequences, a 3x3 block product per
//      frequency, inverse DFTs).  The Poseidon MDS was chosen so that the block constants are tiny: {16,16,32},
//      {-1,-2,8}, {2+i, 1+16i, 1-4i} (derived and checked against the direct sum by tools/gen_mds_freq.py).  All values
//      are integers below 2^53, so the FP64 pipe computes them exactly: 88 DADD/DFMA per 32-bit half instead of 144;
//  (c) partial rounds run in pairs: between the two MDS layers of a pair only element 0 (the one that passes the S-box)
//      is brought back to a 64-bit residue; elements 1..11 stay as exact double halves (< 2^40.1 after one layer,
//      < 2^48.2 after two, intermediates < 300 * 2^40.1 < 2^49), saving 11 of 12 recombine+split steps every other round.
__device__ __forceinline__ u64 p9_mul(u64 a, u64 b) {
    unsigned __int128 pr = (unsigned __int128)a * b;
    u64 lo = (u64)pr, hi = (u64)(pr >> 64);
    u32 r0 = (u32)lo, r1 = (u32)(lo >> 32), r2 = (u32)hi, r3 = (u32)(hi >> 32);
    u32 o0, o1;
    asm("{\n\t.reg .u32 s0,s1,t0,tt1,b,c,m;\n\t"
        "add.cc.u32 s0, %4, %5;\n\taddc.u32 s1, 0, 0;\n\t"
        "sub.cc.u32 t0, %2, s0;\n\tsubc.cc.u32 tt1, %3, s1;\n\tsubc.u32 b, 0, 0;\n\t"
        "sub.cc.u32 t0, t0, b;\n\tsubc.u32 tt1, tt1, 0;\n\t"
        "add.cc.u32 tt1, tt1, %4;\n\taddc.u32 c, 0, 0;\n\t"
        "sub.u32 m, 0, c;\n\t"
        "add.cc.u32 %0, t0, m;\n\taddc.u32 %1, tt1, 0;\n\t}"
        : "=r"(o0), "=r"(o1) : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return (u64)o0 | ((u64)o1 << 32);
}
// The same multiply with the 128-bit product written in PTX (4 mul.wide + two add.cc chains): its carry handling lands on the ALU
// pipe (IADD3 / IADD3.X) where the compiler's own product uses IMAD.X / IMAD.MOV on the port DFMA and IMAD.WIDE share (DESIGN.md
// section 3).  ZKM_P9_MULMIX selects, per multiply of the S-box (bit 0: x*x, 1: x2*x, 2: x2*x2, 3: x3*x4), which product is used;
// 0 = the compiler's everywhere (the product build).  A/B knob of tools/micro/poseidon_bench.cu.
#ifndef ZKM_P9_MULMIX
#define ZKM_P9_MULMIX 0
#endif
__device__ __forceinline__ u64 p9_mul_ptx(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 o0, o1;
    asm("{\n\t.reg .u64 p00,p01,p10,p11;\n\t.reg .u32 r0,r1,r2,r3,t1,u1,u2,v1,v2,w2,w3,s0,s1,t0,tt1,b,c,m;\n\t"
        "mul.wide.u32 p00, %2, %4;\n\tmul.wide.u32 p01, %2, %5;\n\tmul.wide.u32 p10, %3, %4;\n\tmul.wide.u32 p11, %3, %5;\n\t"
        "mov.b64 {r0, t1}, p00;\n\tmov.b64 {u1, u2}, p01;\n\tmov.b64 {v1, v2}, p10;\n\tmov.b64 {w2, w3}, p11;\n\t"
        "add.cc.u32 r1, t1, u1;\n\taddc.cc.u32 r2, u2, w2;\n\taddc.u32 r3, w3, 0;\n\t"
        "add.cc.u32 r1, r1, v1;\n\taddc.cc.u32 r2, r2, v2;\n\taddc.u32 r3, r3, 0;\n\t"
        "add.cc.u32 s0, r2, r3;\n\taddc.u32 s1, 0, 0;\n\t"
        "sub.cc.u32 t0, r0, s0;\n\tsubc.cc.u32 tt1, r1, s1;\n\tsubc.u32 b, 0, 0;\n\t"
        "sub.cc.u32 t0, t0, b;\n\tsubc.u32 tt1, tt1, 0;\n\t"
        "add.cc.u32 tt1, tt1, r2;\n\taddc.u32 c, 0, 0;\n\t"
        "sub.u32 m, 0, c;\n\t"
        "add.cc.u32 %0, t0, m;\n\taddc.u32 %1, tt1, 0;\n\t}"
        : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return (u64)o0 | ((u64)o1 << 32);
}
template <int BIT>
__device__ __forceinline__ u64 p9_mul_sel(u64 a, u64 b) { return ((ZKM_P9_MULMIX >> BIT) & 1) ? p9_mul_ptx(a, b) : p9_mul(a, b); }
__device__ __forceinline__ u64 p9_sbox7(u64 x) {
    u64 x2 = p9_mul_sel<0>(x, x);
    u64 x3 = p9_mul_sel<1>(x2, x);
    u64 x4 = p9_mul_sel<2>(x2, x2);
    return p9_mul_sel<3>(x3, x4);
}
// exact integer MDS on 12 doubles (one 32-bit half of every state word, or an unreduced half from the previous layer)
__device__ __forceinline__ void p9_mds_half(const double* x, double* o) {
    double S1[3], Sm[3], Sr[3], Si[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
        double e0 = x[b] + x[6 + b], e1 = x[3 + b] + x[9 + b];
        Sr[b] = x[b] - x[6 + b]; Si[b] = x[3 + b] - x[9 + b];
        S1[b] = e0 + e1; Sm[b] = e0 - e1;
    }
    // y = 1 : rows of {16,16,32}
    const double t16 = ((S1[0] + S1[1]) + S1[2]) * 16.0;
    const double A0 = fma(S1[2], 16.0, t16), A1 = fma(S1[0], 16.0, t16), A2 = fma(S1[1], 16.0, t16);
    // y = -1 : {-1,-2,8}, {-8,-1,-2}, {2,-8,-1}
    const double B0 = fma(Sm[2], 8.0, fma(Sm[1], -2.0, -Sm[0]));
    const double B1 = fma(Sm[0], -8.0, fma(Sm[2], -2.0, -Sm[1]));
    const double B2 = fma(Sm[1], -8.0, fma(Sm[0], 2.0, -Sm[2]));
    // y = i : {2+i, 1+16i, 1-4i}, {-4-i, 2+i, 1+16i}, {16-i, -4-i, 2+i};  P = sum Sr*gr - Si*gi, Q = sum Sr*gi + Si*gr
    const double P0 = fma(Si[2], 4.0, fma(Si[1], -16.0, fma(Sr[0], 2.0, (Sr[1] + Sr[2]) - Si[0])));
    const double Q0 = fma(Sr[2], -4.0, fma(Sr[1], 16.0, fma(Si[0], 2.0, (Si[1] + Si[2]) + Sr[0])));
    const double P1 = fma(Si[2], -16.0, fma(Sr[1], 2.0, fma(Sr[0], -4.0, (Sr[2] + Si[0]) - Si[1])));
    const double Q1 = fma(Sr[2], 16.0, fma(Si[1], 2.0, fma(Si[0], -4.0, (Si[2] - Sr[0]) + Sr[1])));
    const double P2 = fma(Sr[2], 2.0, fma(Sr[1], -4.0, fma(Sr[0], 16.0, (Si[0] + Si[1]) - Si[2])));
    const double Q2 = fma(Si[2], 2.0, fma(Si[1], -4.0, fma(Si[0], 16.0, Sr[2] - (Sr[0] + Sr[1]))));
    const double u0 = A0 + B0, v0 = A0 - B0, u1 = A1 + B1, v1 = A1 - B1, u2 = A2 + B2, v2 = A2 - B2;
    o[0] = fma(x[0], 8.0, u0 + P0); o[3] = v0 + Q0; o[6] = u0 - P0; o[9] = v0 - Q0;
    o[1] = u1 + P1; o[4] = v1 + Q1; o[7] = u1 - P1; o[10] = v1 - Q1;
    o[2] = u2 + P2; o[5] = v2 + Q2; o[8] = u2 - P2; o[11] = v2 - Q2;
}
// u32 -> double.  CV bit 0: I2F.F64.U32 on the XU pipe (one instruction, otherwise idle pipe) instead of the 2^52
// magic-number packing (a DADD plus register moves on the FP64/FMA pipes, the busiest ones after the v9 changes);
// CV bit 2: pack the magic-number double with one IMAD.WIDE (x * 1 + 0x4330000000000000) instead of two moves.
template <int CV>
__device__ __forceinline__ double p9_cvt(u32 x) {
    if (CV & 1) return (double)x;
    if (CV & 4) {
        u64 t;
        asm("mad.wide.u32 %0, %1, 1, 0x4330000000000000;" : "=l"(t) : "r"(x));
        return __longlong_as_double((long long)t) - 4503599627370496.0;
    }
    return p3_u32_to_f64(x);
}
// al + ah * 2^32 (exact non-negative integers < 2^52 held in doubles) -> lazy residue.  CV bit 1: F2I.U64.F64 (XU pipe)
// instead of DADD + mask.
template <int CV>
__device__ __forceinline__ u64 p9_recombine(double al, double ah) {
    u32 al0, al1, ah0, ah1;
    if (CV & 2) {
        u64 a = __double2ull_rz(al), b = __double2ull_rz(ah);
        al0 = (u32)a; al1 = (u32)(a >> 32); ah0 = (u32)b; ah1 = (u32)(b >> 32);
    } else {
        double tl = al + 4503599627370496.0, th = ah + 4503599627370496.0;
        al0 = (u32)__double2loint(tl); al1 = (u32)__double2hiint(tl) & 0xfffffu;
        ah0 = (u32)__double2loint(th); ah1 = (u32)__double2hiint(th) & 0xfffffu;
    }
    u32 o0, o1;
    asm("{\n\t.reg .u32 l1,h,cy,m,e0,e1;\n\t.reg .u64 t;\n\t"
        "add.cc.u32 l1, %3, %4;\n\taddc.u32 h, %5, 0;\n\t"
        "mul.wide.u32 t, h, 0xffffffff;\n\tmov.b64 {e0, e1}, t;\n\t"
        "add.cc.u32 e0, e0, %2;\n\taddc.cc.u32 e1, e1, l1;\n\taddc.u32 cy, 0, 0;\n\t"
        "sub.u32 m, 0, cy;\n\t"
        "add.cc.u32 %0, e0, m;\n\taddc.u32 %1, e1, 0;\n\t}"
        : "=r"(o0), "=r"(o1) : "r"(al0), "r"(al1), "r"(ah0), "r"(ah1));
    return (u64)o0 | ((u64)o1 << 32);
}
// The 12 S-boxes of a full round.  Rolled as 3 x 4 with a register rotation (one copy of four S-boxes in the instruction
// cache: measured faster than straight-line code when the MDS was the integer version); ZKM_P9_SBOX_UNROLL=1 emits the 12
// S-boxes straight-line (A/B knob, tools/micro/poseidon_bench.cu).
#ifndef ZKM_P9_SBOX_UNROLL
#define ZKM_P9_SBOX_UNROLL 0
#endif
__device__ __forceinline__ void p9_sbox_layer(u64* s) {
#if ZKM_P9_SBOX_UNROLL
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p9_sbox7(s[i]);
#else
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        u64 a = p9_sbox7(s[0]), b = p9_sbox7(s[1]), c = p9_sbox7(s[2]), d = p9_sbox7(s[3]);
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = s[i + 4];
        s[8] = a; s[9] = b; s[10] = c; s[11] = d;
    }
#endif
}
// One of the 19 steps of the permutation: 4 full rounds, 11 PAIRS of partial rounds, 4 full rounds (merged partial-round
// constants).  In a pair only element 0 is brought back to a 64-bit residue between the two MDS layers.
template <bool PAIR, int CV>
__device__ __forceinline__ void p9_step(u64* s, int st) {
    const bool full = (st < 4) || (st >= 15);
    double lo[12], hi[12];
    if (full) {
        const int r = st < 4 ? st : st + 11;
        const u64* rc = (r == 26) ? D_POSEIDON_RC26_MERGED : (D_POSEIDON_RC + 12 * r);
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = p2_add_canon(s[i], rc[i]);
        p9_sbox_layer(s);
#pragma unroll
        for (int i = 0; i < 12; i++) { lo[i] = p9_cvt<CV>((u32)s[i]); hi[i] = p9_cvt<CV>((u32)(s[i] >> 32)); }
    } else {
        const int r = 4 + 2 * (st - 4);
        s[0] = p9_sbox7(p2_add_canon(s[0], D_POSEIDON_PARTIAL_A[r - 4]));
        double xl[12], xh[12];
#pragma unroll
        for (int i = 0; i < 12; i++) { xl[i] = p9_cvt<CV>((u32)s[i]); xh[i] = p9_cvt<CV>((u32)(s[i] >> 32)); }
        p9_mds_half(xl, lo);
        p9_mds_half(xh, hi);
        if (PAIR) {
            u64 s0 = p9_sbox7(p2_add_canon(p9_recombine<CV>(lo[0], hi[0]), D_POSEIDON_PARTIAL_A[r - 3]));
            lo[0] = p9_cvt<CV>((u32)s0); hi[0] = p9_cvt<CV>((u32)(s0 >> 32));
        } else {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = p9_recombine<CV>(lo[i], hi[i]);
            s[0] = p9_sbox7(p2_add_canon(s[0], D_POSEIDON_PARTIAL_A[r - 3]));
#pragma unroll
            for (int i = 0; i < 12; i++) { lo[i] = p9_cvt<CV>((u32)s[i]); hi[i] = p9_cvt<CV>((u32)(s[i] >> 32)); }
        }
    }
    double al[12], ah[12];
    p9_mds_half(lo, al);
    p9_mds_half(hi, ah);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p9_recombine<CV>(al[i], ah[i]);
}
template <bool PAIR, int CV = 0>
__device__ __forceinline__ void poseidon_permute_v9_t(u64* s) {
#pragma unroll 1
    for (int st = 0; st < 19; st++) p9_step<PAIR, CV>(s, st);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = lz_canon(s[i]);
}
// Two independent states advanced step by step in the same loop body, so that the scheduler can issue the S-box phase of one
// (ALU / FMA pipes) next to the MDS phase of the other (FP64 / XU pipes).  Needs ~2x the registers (A/B knob).
template <int CV = 3>
__device__ __forceinline__ void poseidon_permute_v9_x2(u64* a, u64* b) {
#pragma unroll 1
    for (int st = 0; st < 19; st++) { p9_step<true, CV>(a, st); p9_step<true, CV>(b, st); }
#pragma unroll
    for (int i = 0; i < 12; i++) { a[i] = lz_canon(a[i]); b[i] = lz_canon(b[i]); }
}
__device__ __forceinline__ void poseidon_permute_v9(u64* s) { poseidon_permute_v9_t<true, 3>(s); }
// the permutation every product kernel calls
#define poseidon_permute_dev poseidon_permute_v9

#endif
}  // namespace zkm
