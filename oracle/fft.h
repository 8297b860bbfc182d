// ORACLE (test infrastructure): radix-2 NTT / coset LDE over Goldilocks with the conventions of
// plonky2_field 0.1.1 fft.rs / polynomial/mod.rs (not in /root/reference; restated from SURVEY
// Appendix A.2): values[i] = p(w_n^i) natural order; ifft inverts; coset_fft(s)[i] = p(s*w^i);
// coset_ifft(s) = ifft then scale coefficient i by s^-i; lde(r) zero-pads coefficients.
// Call sites being mirrored: reference prover/src/prover.rs:154,514 (from_values), :579
// (from_coeffs), :678-681 (lde_onto_coset), :787 (coset_ifft).
#pragma once
#include "field.h"
#include "par.h"
#include <map>
#include <mutex>
#include <memory>

namespace orc {

// Root table for size 2^k: w^0..w^(n/2-1).
static inline const std::vector<Fp>& root_table(unsigned log_n) {
    static std::mutex mu;
    static std::map<unsigned, std::unique_ptr<std::vector<Fp>>> cache;
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(log_n);
    if (it != cache.end()) return *it->second;
    size_t half = log_n ? ((size_t)1 << (log_n - 1)) : 1;
    auto t = std::make_unique<std::vector<Fp>>(half);
    Fp w = primitive_root_of_unity(log_n), cur = Fp::one();
    for (size_t i = 0; i < half; i++) { (*t)[i] = cur; cur *= w; }
    auto& ref = *t;
    cache[log_n] = std::move(t);
    return ref;
}

template <class T>
static inline void bit_reverse_permute(T* a, size_t n) {
    unsigned lg = log2_strict(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = reverse_bits(i, lg);
        if (i < j) std::swap(a[i], a[j]);
    }
}

// In-place forward NTT, natural order in -> natural order out: a[k] <- sum_i a[i] w^(ik).
static inline void fft_inplace(Fp* a, size_t n) {
    if (n <= 1) return;
    unsigned lg = log2_strict(n);
    const std::vector<Fp>& roots = root_table(lg);
    bit_reverse_permute(a, n);
    for (unsigned s = 1; s <= lg; s++) {
        size_t m = (size_t)1 << s, half = m >> 1, stride = n >> s;
        for (size_t k = 0; k < n; k += m)
            for (size_t j = 0; j < half; j++) {
                Fp t = a[k + j + half] * roots[j * stride];
                Fp u = a[k + j];
                a[k + j] = u + t;
                a[k + j + half] = u - t;
            }
    }
}

static inline void ifft_inplace(Fp* a, size_t n) {
    if (n <= 1) return;
    fft_inplace(a, n);
    // inverse = forward, reversed index, scaled by 1/n
    Fp ninv = Fp((u64)n).inverse();
    a[0] *= ninv;
    if (n > 1) a[n / 2] *= ninv;
    for (size_t i = 1; i < n / 2; i++) {
        Fp x = a[i] * ninv, y = a[n - i] * ninv;
        a[i] = y; a[n - i] = x;
    }
}

static inline void coset_fft_inplace(Fp* a, size_t n, Fp shift) {
    Fp cur = Fp::one();
    for (size_t i = 0; i < n; i++) { a[i] *= cur; cur *= shift; }
    fft_inplace(a, n);
}
static inline void coset_ifft_inplace(Fp* a, size_t n, Fp shift) {
    ifft_inplace(a, n);
    Fp sinv = shift.inverse(), cur = Fp::one();
    for (size_t i = 0; i < n; i++) { a[i] *= cur; cur *= sinv; }
}

// coeffs (len n) -> values on the coset 7*H_{n*2^rate_bits}, natural order (lde + coset_fft(7)).
static inline std::vector<Fp> lde_coset_values(const std::vector<Fp>& coeffs, unsigned rate_bits) {
    std::vector<Fp> v(coeffs.size() << rate_bits);
    std::copy(coeffs.begin(), coeffs.end(), v.begin());
    coset_fft_inplace(v.data(), v.size(), Fp(GL_GENERATOR));
    return v;
}

static inline Fp poly_eval(const std::vector<Fp>& c, Fp x) {
    Fp acc;
    for (size_t i = c.size(); i-- > 0;) acc = acc * x + c[i];
    return acc;
}
static inline Ext2 poly_eval_ext(const std::vector<Fp>& c, Ext2 x) {
    Ext2 acc;
    for (size_t i = c.size(); i-- > 0;) acc = acc * x + Ext2::from_base(c[i]);
    return acc;
}
static inline Ext2 poly_eval_ext(const std::vector<Ext2>& c, Ext2 x) {
    Ext2 acc;
    for (size_t i = c.size(); i-- > 0;) acc = acc * x + c[i];
    return acc;
}

// Extension-valued transforms are component-wise (the twiddles are in the base field).
static inline void ext_coset_fft_inplace(std::vector<Ext2>& a, Fp shift) {
    size_t n = a.size();
    std::vector<Fp> x(n), y(n);
    for (size_t i = 0; i < n; i++) { x[i] = a[i].a; y[i] = a[i].b; }
    coset_fft_inplace(x.data(), n, shift);
    coset_fft_inplace(y.data(), n, shift);
    for (size_t i = 0; i < n; i++) a[i] = Ext2(x[i], y[i]);
}

}  // namespace orc
