// Goldilocks field p = 2^64 - 2^32 + 1 and its quadratic extension F_p[X]/(X^2 - 7), host+device.
// Replaces the reference's use of plonky2_field::goldilocks_field::GoldilocksField and
// extension::quadratic::QuadraticExtension (call sites: reference prover/src/prover.rs:141-789).
// Representation: a u64 that is ALWAYS canonical (< p) at rest, so that buffers compare bit-exactly
// with the reference's `to_canonical_u64()` view.  Montgomery form is not used:
// 2^64 = 2^32 - 1 and 2^96 = -1 (mod p) give a shift/add reduction of the 128-bit product.
#pragma once
#include <stdint.h>

#include "tables/hd.h"
#ifdef __CUDACC__
#define ZKM_D __device__ __forceinline__
#else
#define ZKM_D inline
#endif

namespace zkm {

typedef uint64_t u64;
typedef uint32_t u32;

static const u64 GL_P = 0xFFFFFFFF00000001ULL;
static const u64 GL_EPS = 0xFFFFFFFFULL;

struct gl {
    u64 v;
    ZKM_HD gl() : v(0) {}
    ZKM_HD explicit gl(u64 x) : v(x) {}              // caller guarantees x < p
    static ZKM_HD gl from_u64(u64 x) { return gl(x >= GL_P ? x - GL_P : x); }
    static ZKM_HD gl zero() { return gl(0); }
    static ZKM_HD gl one() { return gl(1); }
};

// ZKM_GLADD selects the device add/sub: 0 = compare/select C code, 1 = borrow-chain PTX (a - b: a borrow is fixed by
// subtracting 2^32 - 1, i.e. adding p mod 2^64; a + b = a - (p - b)).
#ifndef ZKM_GLADD
#define ZKM_GLADD 1
#endif
ZKM_HD gl operator-(gl a, gl b) {
#if defined(__CUDA_ARCH__) && ZKM_GLADD == 1
    u32 a0 = (u32)a.v, a1 = (u32)(a.v >> 32), b0 = (u32)b.v, b1 = (u32)(b.v >> 32), o0, o1;
    asm("{\n\t.reg .u32 t0,t1,m;\n\tsub.cc.u32 t0, %2, %4;\n\tsubc.cc.u32 t1, %3, %5;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, t0, m;\n\tsubc.u32 %1, t1, 0;\n\t}" : "=r"(o0), "=r"(o1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return gl((u64)o0 | ((u64)o1 << 32));
#else
    u64 d = a.v - b.v;
    if (a.v < b.v) d += GL_P;
    return gl(d);
#endif
}
ZKM_HD gl operator+(gl a, gl b) {
#if defined(__CUDA_ARCH__) && ZKM_GLADD == 1
    return a - gl(GL_P - b.v);      // p - b in (0, p]: the borrow chain treats p like 0
#else
    u64 s = a.v + b.v;
    // a,b < p: true sum < 2p; wrapped iff s < a.v
    if (s < a.v || s >= GL_P) s -= GL_P;
    return gl(s);
#endif
}
ZKM_HD gl operator-(gl a) { return gl(a.v ? GL_P - a.v : 0); }

// (hi:lo) mod p, canonical.
ZKM_HD u64 gl_reduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;
    u64 t1 = hl * GL_EPS;
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}
// (x2:x1:x0 as 32|64 bits) i.e. lo + hi32*2^64, mod p, canonical
ZKM_HD u64 gl_reduce96(u64 lo, u32 hi32) {
    u64 t1 = (u64)hi32 * GL_EPS;
    u64 r = lo + t1;
    if (r < t1) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}

// ZKM_GLMUL selects the device multiply: 0 = PTX product + PTX folding, 1 = compiler product + 128-bit-sum folding,
// 2 = compiler product + compare/select folding, 3 = compiler product + PTX folding, 4 = 3 with the canonicalisation done by
// two compares + two predicated moves (the product build: NTT -3.8 %, quotient -3.8 %, openings -10 % against 3 on U20,
// profiles/r2d_glmul_ab.txt).  The kernels built on this are bound by
// the ALU pipe (IADD3/LOP3/SEL, ncu: 72-78 % vs 15 % on the FMA pipe), so what counts is how much of the carry handling the
// compiler can place on the FMA pipe (IMAD.WIDE with carry-out, IMAD.X), not the instruction total.
#ifndef ZKM_GLMUL
#define ZKM_GLMUL 4
#endif
ZKM_HD gl operator*(gl a, gl b) {
#if defined(__CUDA_ARCH__) && ZKM_GLMUL == 0
    // 4 IMAD.WIDE + two IADD3 carry chains (product, then 2^64 = 2^32 - 1 / 2^96 = -1 folding); see
    // poseidon_v2.cuh p2_mul for the derivation.  ~20 SASS instructions instead of ~28.
    u32 a0 = (u32)a.v, a1 = (u32)(a.v >> 32), b0 = (u32)b.v, b1 = (u32)(b.v >> 32);
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
    u64 r = (u64)o0 | ((u64)o1 << 32);
    return gl(r >= GL_P ? r - GL_P : r);
#elif defined(__CUDA_ARCH__) && ZKM_GLMUL == 1
    // x = lo + r2 * (2^32 - 1) - r3 computed as the 66-bit sum s = lo + r2 * EPS + (EPS - r3) = x + 2^64 (mod p);
    // s = sl + sh * 2^64 with sh in {0, 1, 2}  =>  x = sl + (sh - 1) * EPS, which neither wraps nor goes negative.
    unsigned __int128 pr = (unsigned __int128)a.v * b.v;
    u64 lo = (u64)pr, hi = (u64)(pr >> 64);
    u32 r2 = (u32)hi, r3 = (u32)(hi >> 32);
    unsigned __int128 sm = (unsigned __int128)lo + (unsigned __int128)((u64)r2 * GL_EPS) + (unsigned __int128)(GL_EPS - r3);
    u64 sl = (u64)sm;
    long long sh = (long long)(u64)(sm >> 64) - 1;
    u64 r = sl + (((u64)sh) << 32) - (u64)sh;
    return gl(r >= GL_P ? r - GL_P : r);
#elif defined(__CUDA_ARCH__) && ZKM_GLMUL == 3
    unsigned __int128 pr = (unsigned __int128)a.v * b.v;
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
    u64 r = (u64)o0 | ((u64)o1 << 32);
    return gl(r >= GL_P ? r - GL_P : r);
#elif defined(__CUDA_ARCH__) && ZKM_GLMUL == 4
    // as 3, with the final canonicalisation folded into the chain: x >= p  <=>  x_hi == 2^32 - 1 and x_lo != 0, and then
    // x - p = (0 : x_lo - 1): two compares + two predicated moves instead of a 64-bit subtract, compare and select.
    // (A first version also took the carry mask with `subc` right after `add.cc`; mixing the add and subtract carry
    // conventions gave wrong results on the device -- profiles/r2c_glmul4_note.txt -- so the mask is built as in 3.)
    unsigned __int128 pr = (unsigned __int128)a.v * b.v;
    u64 lo = (u64)pr, hi = (u64)(pr >> 64);
    u32 r0 = (u32)lo, r1 = (u32)(lo >> 32), r2 = (u32)hi, r3 = (u32)(hi >> 32);
    u32 o0, o1;
    asm("{\n\t.reg .u32 s0,s1,t0,tt1,b,c,m,x0,x1;\n\t.reg .pred p,q;\n\t"
        "add.cc.u32 s0, %4, %5;\n\taddc.u32 s1, 0, 0;\n\t"
        "sub.cc.u32 t0, %2, s0;\n\tsubc.cc.u32 tt1, %3, s1;\n\tsubc.u32 b, 0, 0;\n\t"
        "sub.cc.u32 t0, t0, b;\n\tsubc.u32 tt1, tt1, 0;\n\t"
        "add.cc.u32 tt1, tt1, %4;\n\taddc.u32 c, 0, 0;\n\t"
        "sub.u32 m, 0, c;\n\t"
        "add.cc.u32 x0, t0, m;\n\taddc.u32 x1, tt1, 0;\n\t"
        "setp.ne.u32 p, x0, 0;\n\tsetp.eq.and.u32 q, x1, 0xffffffff, p;\n\t"
        "@q sub.u32 x0, x0, 1;\n\t@q mov.u32 x1, 0;\n\t"
        "mov.u32 %0, x0;\n\tmov.u32 %1, x1;\n\t}"
        : "=r"(o0), "=r"(o1) : "r"(r0), "r"(r1), "r"(r2), "r"(r3));
    return gl((u64)o0 | ((u64)o1 << 32));
#else
    unsigned __int128 p = (unsigned __int128)a.v * b.v;
    u64 lo = (u64)p, hi = (u64)(p >> 64);
    return gl(gl_reduce128(lo, hi));
#endif
}
ZKM_HD gl& operator+=(gl& a, gl b) { a = a + b; return a; }
ZKM_HD gl& operator-=(gl& a, gl b) { a = a - b; return a; }
ZKM_HD gl& operator*=(gl& a, gl b) { a = a * b; return a; }
ZKM_HD bool operator==(gl a, gl b) { return a.v == b.v; }
ZKM_HD bool operator!=(gl a, gl b) { return a.v != b.v; }

ZKM_HD gl gl_pow(gl b, u64 e) {
    gl r = gl::one();
    while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
    return r;
}
ZKM_HD gl gl_inv(gl a) { return gl_pow(a, GL_P - 2); }
ZKM_HD gl gl_exp2(gl a, unsigned k) { while (k--) a = a * a; return a; }

static const u64 GL_GENERATOR = 7;
static const u64 GL_POWER_OF_TWO_GENERATOR = 1753635133440165772ULL;
ZKM_HD gl gl_root_of_unity(unsigned log_n) { return gl_exp2(gl(GL_POWER_OF_TWO_GENERATOR), 32 - log_n); }

// ---- quadratic extension, W = 7 ----
struct gl2 {
    gl a, b;
    ZKM_HD gl2() {}
    ZKM_HD gl2(gl a_, gl b_) : a(a_), b(b_) {}
    ZKM_HD explicit gl2(gl a_) : a(a_), b() {}
    static ZKM_HD gl2 zero() { return gl2(); }
    static ZKM_HD gl2 one() { return gl2(gl::one(), gl()); }
};
ZKM_HD gl2 mk2(u64 a, u64 b) { return gl2(gl(a), gl(b)); }
ZKM_HD gl2 operator+(gl2 x, gl2 y) { return gl2(x.a + y.a, x.b + y.b); }
ZKM_HD gl2 operator-(gl2 x, gl2 y) { return gl2(x.a - y.a, x.b - y.b); }
ZKM_HD gl2 operator-(gl2 x) { return gl2(-x.a, -x.b); }
ZKM_HD gl2 operator*(gl2 x, gl2 y) {
    gl bb = x.b * y.b;
    gl seven_bb = gl(7) * bb;
    return gl2(x.a * y.a + seven_bb, x.a * y.b + x.b * y.a);
}
ZKM_HD gl2 operator*(gl2 x, gl s) { return gl2(x.a * s, x.b * s); }
ZKM_HD gl2 operator+(gl2 x, gl s) { return gl2(x.a + s, x.b); }
ZKM_HD gl2 operator-(gl2 x, gl s) { return gl2(x.a - s, x.b); }
ZKM_HD gl2& operator+=(gl2& x, gl2 y) { x = x + y; return x; }
ZKM_HD gl2& operator-=(gl2& x, gl2 y) { x = x - y; return x; }
ZKM_HD gl2& operator*=(gl2& x, gl2 y) { x = x * y; return x; }
ZKM_HD bool operator==(gl2 x, gl2 y) { return x.a == y.a && x.b == y.b; }
ZKM_HD bool operator!=(gl2 x, gl2 y) { return !(x == y); }
ZKM_HD gl2 gl2_inv(gl2 x) {
    gl norm = x.a * x.a - gl(7) * (x.b * x.b);
    gl ni = gl_inv(norm);
    return gl2(x.a * ni, -(x.b * ni));
}
ZKM_HD gl2 gl2_pow(gl2 b, u64 e) {
    gl2 r = gl2::one();
    while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
    return r;
}
ZKM_HD gl2 gl2_exp2(gl2 a, unsigned k) { while (k--) a = a * a; return a; }

ZKM_HD u32 bitrev32(u32 x, unsigned bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    u32 r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
#endif
}

}  // namespace zkm
