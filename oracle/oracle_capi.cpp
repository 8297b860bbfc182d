// ORACLE (test infrastructure): C entry points over the CPU restatement, loaded with ctypes by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs ONLY.
// The product library (zkm_b200/libzkm_b200.so) never links or loads this.
#include "field.h"
#include "poseidon.h"
#include "fft.h"
#include "plonky2_restated.h"
#include <cstring>
#include <string>

using namespace orc;

static thread_local std::string g_err;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
void orc_set_threads(int n) { set_threads(n); }
int orc_get_threads() { return get_threads(); }

// ---- field ----
uint64_t orc_mul(uint64_t a, uint64_t b) { return (Fp(a) * Fp(b)).v; }
uint64_t orc_add(uint64_t a, uint64_t b) { return (Fp(a) + Fp(b)).v; }
uint64_t orc_sub(uint64_t a, uint64_t b) { return (Fp(a) - Fp(b)).v; }
uint64_t orc_inv(uint64_t a) { return Fp(a).inverse().v; }
uint64_t orc_root_of_unity(unsigned log_n) { return primitive_root_of_unity(log_n).v; }
void orc_ext_mul(const uint64_t* a, const uint64_t* b, uint64_t* out) {
    Ext2 r = Ext2(Fp(a[0]), Fp(a[1])) * Ext2(Fp(b[0]), Fp(b[1]));
    out[0] = r.a.v; out[1] = r.b.v;
}
void orc_ext_inv(const uint64_t* a, uint64_t* out) {
    Ext2 r = Ext2(Fp(a[0]), Fp(a[1])).inverse();
    out[0] = r.a.v; out[1] = r.b.v;
}

// ---- Poseidon ----
void orc_poseidon_permute(uint64_t* st, int fast) {
    PState s;
    for (int i = 0; i < 12; i++) s[i] = Fp(st[i]);
    if (fast) poseidon_fast(s); else poseidon_naive(s);
    for (int i = 0; i < 12; i++) st[i] = s[i].v;
}
void orc_poseidon_permute_many(uint64_t* st, size_t count) {
    parallel_for(count, [&](size_t k) {
        PState s;
        for (int i = 0; i < 12; i++) s[i] = Fp(st[k * 12 + i]);
        poseidon(s);
        for (int i = 0; i < 12; i++) st[k * 12 + i] = s[i].v;
    });
}
void orc_hash_or_noop(const uint64_t* in, size_t n, uint64_t* out) {
    std::vector<Fp> v(n);
    for (size_t i = 0; i < n; i++) v[i] = Fp(in[i]);
    Digest d = hash_or_noop(v.data(), n);
    for (int i = 0; i < 4; i++) out[i] = d.e[i].v;
}
void orc_two_to_one(const uint64_t* l, const uint64_t* r, uint64_t* out) {
    Digest a, b;
    for (int i = 0; i < 4; i++) { a.e[i] = Fp(l[i]); b.e[i] = Fp(r[i]); }
    Digest d = two_to_one(a, b);
    for (int i = 0; i < 4; i++) out[i] = d.e[i].v;
}

// ---- transforms: kind 0 fft, 1 ifft, 2 coset_ifft(7), 3 coset_fft(7); column-major ncols x n ----
void orc_ntt(uint64_t* data, uint32_t ncols, uint32_t log_n, int kind) {
    size_t n = (size_t)1 << log_n;
    parallel_for(ncols, [&](size_t c) {
        std::vector<Fp> v(n);
        for (size_t i = 0; i < n; i++) v[i] = Fp(data[c * n + i]);
        if (kind == 0) fft_inplace(v.data(), n);
        else if (kind == 1) ifft_inplace(v.data(), n);
        else if (kind == 2) coset_ifft_inplace(v.data(), n, Fp(GL_GENERATOR));
        else coset_fft_inplace(v.data(), n, Fp(GL_GENERATOR));
        for (size_t i = 0; i < n; i++) data[c * n + i] = v[i].v;
    }, 1);
}

// ---- PolynomialBatch ----
struct OrcBatch { PolynomialBatch b; };

void* orc_commit(const uint64_t* const* cols, uint32_t ncols, uint32_t log_n, uint32_t rate_bits, uint32_t cap_height,
                 int from_values, uint64_t* cap_out) {
    try {
        size_t n = (size_t)1 << log_n;
        std::vector<std::vector<Fp>> v(ncols, std::vector<Fp>(n));
        for (uint32_t c = 0; c < ncols; c++)
            for (size_t i = 0; i < n; i++) v[c][i] = Fp(cols[c][i]);
        auto* h = new OrcBatch;
        h->b = from_values ? PolynomialBatch::from_values(std::move(v), rate_bits, cap_height)
                           : PolynomialBatch::from_coeffs(std::move(v), rate_bits, cap_height);
        if (cap_out)
            for (size_t i = 0; i < h->b.merkle_tree.cap.size(); i++)
                for (int k = 0; k < 4; k++) cap_out[i * 4 + k] = h->b.merkle_tree.cap[i].e[k].v;
        return h;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void orc_batch_free(void* h) { delete (OrcBatch*)h; }
void orc_batch_get_coeffs(const void* h, uint32_t col, uint64_t* out) {
    const auto& p = ((const OrcBatch*)h)->b.polynomials[col];
    for (size_t i = 0; i < p.size(); i++) out[i] = p[i].v;
}
// natural-order LDE values of one polynomial
void orc_batch_get_lde(const void* h, uint32_t col, uint64_t* out) {
    const PolynomialBatch& b = ((const OrcBatch*)h)->b;
    size_t N = b.merkle_tree.leaves.size();
    for (size_t m = 0; m < N; m++) out[m] = b.get_lde_values(m, 1)[col].v;
}
void orc_batch_open(const void* h, uint32_t leaf, uint64_t* leaf_out, uint64_t* siblings_out) {
    const PolynomialBatch& b = ((const OrcBatch*)h)->b;
    const auto& row = b.merkle_tree.get(leaf);
    for (size_t i = 0; i < row.size(); i++) leaf_out[i] = row[i].v;
    MerkleProof p = b.merkle_tree.prove(leaf);
    for (size_t i = 0; i < p.siblings.size(); i++)
        for (int k = 0; k < 4; k++) siblings_out[i * 4 + k] = p.siblings[i].e[k].v;
}

}  // extern "C"
