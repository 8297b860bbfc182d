// ORACLE (test infrastructure): (de)serialisation of AllProof in the flat little-endian u64 layout
// documented in include/zkm_b200.h ("Proof buffer layout").  Written independently of the product's
// writer so that byte equality of the two buffers is a real check.
#pragma once
#include "stark.h"

namespace orc {

struct Writer {
    std::vector<u64> w;
    void u(u64 x) { w.push_back(x); }
    void f(Fp x) { w.push_back(x.v); }
    void e(Ext2 x) { w.push_back(x.a.v); w.push_back(x.b.v); }
    void digest(const Digest& d) { for (int i = 0; i < 4; i++) f(d.e[i]); }
    void cap(const MerkleCap& c) { u(c.size()); for (auto& d : c) digest(d); }
    void exts(const std::vector<Ext2>& v) { u(v.size()); for (auto x : v) e(x); }
    void fps(const std::vector<Fp>& v) { u(v.size()); for (auto x : v) f(x); }
    void path(const MerkleProof& p) { u(p.siblings.size()); for (auto& d : p.siblings) digest(d); }
};

static const u64 PROOF_MAGIC = 0x464F4F52504D4B5AULL;      // "ZKMPROOF" little endian

static inline std::vector<u64> serialize(const AllProof& ap) {
    Writer W;
    W.u(PROOF_MAGIC); W.u(1); W.u(ap.stark_proofs.size());
    W.u(ap.ctl_challenges.size());
    for (auto& c : ap.ctl_challenges) { W.f(c.beta); W.f(c.gamma); }
    for (int i = 0; i < 8; i++) W.u(ap.public_values.roots_before[i]);
    for (int i = 0; i < 8; i++) W.u(ap.public_values.roots_after[i]);
    W.u(ap.public_values.userdata.size());
    for (uint8_t b : ap.public_values.userdata) W.u(b);
    for (auto& sp : ap.stark_proofs) {
        for (int i = 0; i < 12; i++) W.f(sp.init_challenger_state[i]);
        const StarkProof& p = sp.proof;
        W.cap(p.trace_cap); W.cap(p.auxiliary_polys_cap); W.cap(p.quotient_polys_cap);
        W.exts(p.openings.local_values); W.exts(p.openings.next_values);
        W.exts(p.openings.auxiliary_polys); W.exts(p.openings.auxiliary_polys_next);
        W.fps(p.openings.ctl_zs_first); W.exts(p.openings.quotient_polys);
        const FriProof& fp = p.opening_proof;
        W.u(fp.commit_phase_merkle_caps.size());
        for (auto& c : fp.commit_phase_merkle_caps) W.cap(c);
        W.u(fp.query_round_proofs.size());
        for (auto& r : fp.query_round_proofs) {
            W.u(r.initial_trees_proof.size());
            for (auto& ep : r.initial_trees_proof) { W.fps(ep.first); W.path(ep.second); }
            W.u(r.steps.size());
            for (auto& st : r.steps) { W.exts(st.evals); W.path(st.merkle_proof); }
        }
        W.exts(fp.final_poly);
        W.f(fp.pow_witness);
    }
    return W.w;
}

struct Reader {
    const u64* p; size_t n, i = 0;
    Reader(const u64* p_, size_t n_) : p(p_), n(n_) {}
    u64 u() { if (i >= n) throw VerifyError("proof buffer truncated"); return p[i++]; }
    size_t len(size_t unit = 1) { u64 l = u(); if (l > (n - i) / unit + 1) throw VerifyError("proof buffer: bad length"); return (size_t)l; }
    Fp f() { u64 x = u(); if (x >= GL_P) throw VerifyError("non-canonical field element"); Fp r; r.v = x; return r; }
    Ext2 e() { Fp a = f(); Fp b = f(); return Ext2(a, b); }
    Digest digest() { Digest d; for (int k = 0; k < 4; k++) d.e[k] = f(); return d; }
    MerkleCap cap() { size_t l = len(4); MerkleCap c; for (size_t k = 0; k < l; k++) c.push_back(digest()); return c; }
    std::vector<Ext2> exts() { size_t l = len(2); std::vector<Ext2> v; for (size_t k = 0; k < l; k++) v.push_back(e()); return v; }
    std::vector<Fp> fps() { size_t l = len(); std::vector<Fp> v; for (size_t k = 0; k < l; k++) v.push_back(f()); return v; }
    MerkleProof path() { size_t l = len(4); MerkleProof m; for (size_t k = 0; k < l; k++) m.siblings.push_back(digest()); return m; }
};

static inline AllProof deserialize(const u64* buf, size_t n) {
    Reader R(buf, n);
    if (R.u() != PROOF_MAGIC) throw VerifyError("bad proof magic");
    if (R.u() != 1) throw VerifyError("bad proof version");
    size_t T = R.len();
    AllProof ap;
    size_t nc = R.len();
    for (size_t k = 0; k < nc; k++) { GrandProductChallenge c; c.beta = R.f(); c.gamma = R.f(); ap.ctl_challenges.push_back(c); }
    for (int k = 0; k < 8; k++) ap.public_values.roots_before[k] = (uint32_t)R.u();
    for (int k = 0; k < 8; k++) ap.public_values.roots_after[k] = (uint32_t)R.u();
    size_t ul = R.len();
    for (size_t k = 0; k < ul; k++) ap.public_values.userdata.push_back((uint8_t)R.u());
    for (size_t t = 0; t < T; t++) {
        StarkProofWithMetadata sp;
        for (int k = 0; k < 12; k++) sp.init_challenger_state[k] = R.f();
        StarkProof& p = sp.proof;
        p.trace_cap = R.cap(); p.auxiliary_polys_cap = R.cap(); p.quotient_polys_cap = R.cap();
        p.openings.local_values = R.exts(); p.openings.next_values = R.exts();
        p.openings.auxiliary_polys = R.exts(); p.openings.auxiliary_polys_next = R.exts();
        p.openings.ctl_zs_first = R.fps(); p.openings.quotient_polys = R.exts();
        FriProof& fp = p.opening_proof;
        size_t ncap = R.len();
        for (size_t k = 0; k < ncap; k++) fp.commit_phase_merkle_caps.push_back(R.cap());
        size_t nq = R.len();
        for (size_t q = 0; q < nq; q++) {
            FriQueryRound r;
            size_t no = R.len();
            for (size_t o = 0; o < no; o++) { std::vector<Fp> leaf = R.fps(); MerkleProof mp = R.path(); r.initial_trees_proof.emplace_back(std::move(leaf), std::move(mp)); }
            size_t ns = R.len();
            for (size_t s = 0; s < ns; s++) { FriQueryStep st; st.evals = R.exts(); st.merkle_proof = R.path(); r.steps.push_back(std::move(st)); }
            fp.query_round_proofs.push_back(std::move(r));
        }
        fp.final_poly = R.exts();
        fp.pow_witness = R.f();
        ap.stark_proofs.push_back(std::move(sp));
    }
    if (R.i != n) throw VerifyError("trailing data after proof");
    return ap;
}

}  // namespace orc
