// ORACLE (test infrastructure, not product code): CPU restatement of the Goldilocks field and its
// quadratic extension as used by the reference through the un-vendored dependency
// plonky2_field 0.1.1 @ zkMIPS/plonky2#f1e28a6d (prover/examples/Cargo.lock:3234-3283).
// Restated from the published algorithm; pinned in-tree by
//   * GOLDILOCKS_INVERSE_2EXP32 = 18446744065119617026 (reference prover/src/cpu/jumps.rs:15)
//   * the Poseidon known answers which exercise add/mul end to end.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference may use this.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <cassert>

namespace orc {

typedef uint64_t u64;
typedef unsigned __int128 u128;

static const u64 GL_P = 0xFFFFFFFF00000001ULL;     // 2^64 - 2^32 + 1
static const u64 GL_EPS = 0xFFFFFFFFULL;           // 2^32 - 1 = 2^64 mod p

// x mod p for a 128-bit x, using 2^64 = 2^32-1 and 2^96 = -1 (mod p).
static inline u64 gl_reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & GL_EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= GL_EPS;                      // borrowed 2^64 = eps (mod p)
    u64 t1 = hl * GL_EPS;                           // < 2^64
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;                        // carried 2^64
    if (r >= GL_P) r -= GL_P;
    return r;
}

struct Fp {
    u64 v;                                           // always canonical (< p)
    Fp() : v(0) {}
    explicit Fp(u64 x) : v(x >= GL_P ? x - GL_P : x) {}
    static Fp zero() { return Fp(); }
    static Fp one() { Fp r; r.v = 1; return r; }
    static Fp from_i64(int64_t x) { return x >= 0 ? Fp((u64)x) : -Fp((u64)(-x)); }
    Fp operator+(Fp o) const {
        u64 s = v + o.v;
        if (s < v || s >= GL_P) s -= GL_P;
        Fp r; r.v = s; return r;
    }
    Fp operator-(Fp o) const {
        Fp r; r.v = v >= o.v ? v - o.v : v + (GL_P - o.v); return r;
    }
    Fp operator-() const { Fp r; r.v = v ? GL_P - v : 0; return r; }
    Fp operator*(Fp o) const { Fp r; r.v = gl_reduce128((u128)v * o.v); return r; }
    Fp& operator+=(Fp o) { *this = *this + o; return *this; }
    Fp& operator-=(Fp o) { *this = *this - o; return *this; }
    Fp& operator*=(Fp o) { *this = *this * o; return *this; }
    bool operator==(Fp o) const { return v == o.v; }
    bool operator!=(Fp o) const { return v != o.v; }
    bool is_zero() const { return v == 0; }
    Fp square() const { return *this * *this; }
    Fp pow(u64 e) const {
        Fp b = *this, r = one();
        while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    Fp exp_power_of_2(unsigned k) const { Fp r = *this; while (k--) r = r.square(); return r; }
    Fp inverse() const { assert(v != 0); return pow(GL_P - 2); }
};

// plonky2: MULTIPLICATIVE_GROUP_GENERATOR = coset_shift() = 7, TWO_ADICITY = 32,
// POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32) = 1753635133440165772 (Appendix A.1).
static const u64 GL_GENERATOR = 7;
static const u64 GL_POWER_OF_TWO_GENERATOR = 1753635133440165772ULL;
static inline Fp primitive_root_of_unity(unsigned log_n) {
    assert(log_n <= 32);
    return Fp(GL_POWER_OF_TWO_GENERATOR).exp_power_of_2(32 - log_n);
}

// Montgomery batch inversion; all inputs must be non-zero (plonky2 batch_multiplicative_inverse).
static inline std::vector<Fp> batch_inverse(const std::vector<Fp>& x) {
    size_t n = x.size();
    std::vector<Fp> out(n);
    if (!n) return out;
    std::vector<Fp> pre(n);
    Fp acc = Fp::one();
    for (size_t i = 0; i < n; i++) { pre[i] = acc; acc *= x[i]; }
    Fp inv = acc.inverse();
    for (size_t i = n; i-- > 0;) { out[i] = inv * pre[i]; inv *= x[i]; }
    return out;
}

// F_p[X]/(X^2 - 7)  (plonky2 QuadraticExtension<GoldilocksField>, W = 7).
struct Ext2 {
    Fp a, b;                                         // a + b*X
    Ext2() {}
    explicit Ext2(u64 c) : a(c), b() {}              // from a canonical base-field constant
    Ext2(Fp a_, Fp b_) : a(a_), b(b_) {}
    static Ext2 from_base(Fp x) { return Ext2(x, Fp()); }
    static Ext2 zero() { return Ext2(); }
    static Ext2 one() { return Ext2(Fp::one(), Fp()); }
    Ext2 operator+(Ext2 o) const { return Ext2(a + o.a, b + o.b); }
    Ext2 operator-(Ext2 o) const { return Ext2(a - o.a, b - o.b); }
    Ext2 operator-() const { return Ext2(-a, -b); }
    Ext2 operator*(Ext2 o) const {
        return Ext2(a * o.a + Fp(7) * (b * o.b), a * o.b + b * o.a);
    }
    Ext2 operator*(Fp s) const { return Ext2(a * s, b * s); }      // scalar_mul
    Ext2& operator+=(Ext2 o) { *this = *this + o; return *this; }
    Ext2& operator-=(Ext2 o) { *this = *this - o; return *this; }
    Ext2& operator*=(Ext2 o) { *this = *this * o; return *this; }
    bool operator==(Ext2 o) const { return a == o.a && b == o.b; }
    bool operator!=(Ext2 o) const { return !(*this == o); }
    bool is_zero() const { return a.is_zero() && b.is_zero(); }
    Ext2 square() const { return *this * *this; }
    Ext2 pow(u64 e) const {
        Ext2 base = *this, r = one();
        while (e) { if (e & 1) r *= base; base *= base; e >>= 1; }
        return r;
    }
    Ext2 exp_power_of_2(unsigned k) const { Ext2 r = *this; while (k--) r = r.square(); return r; }
    Ext2 inverse() const {
        Fp norm = a * a - Fp(7) * (b * b);
        Fp ni = norm.inverse();
        return Ext2(a * ni, -(b * ni));
    }
};

static inline unsigned log2_strict(size_t n) {
    unsigned l = 0;
    while (((size_t)1 << l) < n) l++;
    assert(((size_t)1 << l) == n);
    return l;
}
static inline size_t reverse_bits(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

}  // namespace orc
