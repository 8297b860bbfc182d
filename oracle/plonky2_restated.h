// ORACLE (test infrastructure): restatement of the plonky2 0.1.4 pieces the reference's STARK
// prover calls (Merkle tree, Challenger, PolynomialBatch, FRI prover + verifier).  plonky2 is an
// un-vendored git dependency of the reference (zkMIPS/plonky2 @ zkm_dev
// f1e28a6d85422edc8b0cf717b691cf2858339dfd, prover/examples/Cargo.lock:3234-3283) and is NOT on this
// machine, so every rule here is restated from the published algorithm; the item numbers (A.n)
// refer to SURVEY.md Appendix A.  PARITY UNPINNED at this boundary: the reference holds no golden
// caps/challenges/FRI vectors; what is pinned is Poseidon (Appendix D) and the
// self-consistency "restated verifier accepts restated prover, rejects tampering".
// Reference call sites mirrored: prover.rs:154,514,579,621 ; proof.rs:310-333 ;
// get_challenges.rs:128-147,213-231 ; verifier.rs:276.
#pragma once
#include "field.h"
#include "poseidon.h"
#include "fft.h"
#include "par.h"
#include <string>
#include <stdexcept>

namespace orc {

struct VerifyError : std::runtime_error { using std::runtime_error::runtime_error; };
#define ORC_ENSURE(c, msg) do { if (!(c)) throw ::orc::VerifyError(msg); } while (0)

// ---------------------------------------------------------------------------------- A.5 Merkle
typedef std::vector<Digest> MerkleCap;
struct MerkleProof { std::vector<Digest> siblings; };

struct MerkleTree {
    std::vector<std::vector<Fp>> leaves;             // leaf rows
    std::vector<std::vector<Digest>> levels;         // levels[0] = leaf digests, ... up to cap level
    MerkleCap cap;
    unsigned cap_height = 0;

    MerkleTree() {}
    MerkleTree(std::vector<std::vector<Fp>> lv, unsigned cap_h) : leaves(std::move(lv)), cap_height(cap_h) {
        size_t n = leaves.size();
        unsigned lg = log2_strict(n);
        if (cap_h > lg) throw std::runtime_error("cap height exceeds tree height");
        levels.emplace_back(n);
        parallel_for(n, [&](size_t i) { levels[0][i] = hash_or_noop(leaves[i].data(), leaves[i].size()); });
        for (unsigned h = lg; h > cap_h; h--) {
            const std::vector<Digest>& prev = levels.back();
            std::vector<Digest> cur(prev.size() / 2);
            parallel_for(cur.size(), [&](size_t i) { cur[i] = two_to_one(prev[2 * i], prev[2 * i + 1]); });
            levels.push_back(std::move(cur));
        }
        cap = levels.back();
    }
    const std::vector<Fp>& get(size_t i) const { return leaves[i]; }
    MerkleProof prove(size_t idx) const {
        MerkleProof p;
        for (size_t l = 0; l + 1 < levels.size(); l++) { p.siblings.push_back(levels[l][idx ^ 1]); idx >>= 1; }
        return p;
    }
};

static inline void verify_merkle_proof_to_cap(const std::vector<Fp>& leaf, size_t index, const MerkleCap& cap,
                                              const MerkleProof& proof) {
    Digest cur = hash_or_noop(leaf.data(), leaf.size());
    for (const Digest& sib : proof.siblings) {
        cur = (index & 1) ? two_to_one(sib, cur) : two_to_one(cur, sib);
        index >>= 1;
    }
    ORC_ENSURE(index < cap.size() && cur == cap[index], "Invalid Merkle proof.");
}

// ------------------------------------------------------------------------------ A.6 Challenger
struct Challenger {
    PState state;                                    // zero initialised
    std::vector<Fp> input, output;
    void duplexing() {
        for (size_t i = 0; i < input.size(); i++) state[i] = input[i];
        input.clear();
        poseidon(state);
        output.assign(state.begin(), state.begin() + 8);
    }
    void observe_element(Fp x) {
        output.clear();
        input.push_back(x);
        if (input.size() == 8) duplexing();
    }
    void observe_elements(const std::vector<Fp>& v) { for (Fp x : v) observe_element(x); }
    void observe_digest(const Digest& d) { for (int i = 0; i < 4; i++) observe_element(d.e[i]); }
    void observe_cap(const MerkleCap& c) { for (const Digest& d : c) observe_digest(d); }
    void observe_ext(Ext2 x) { observe_element(x.a); observe_element(x.b); }
    void observe_exts(const std::vector<Ext2>& v) { for (Ext2 x : v) observe_ext(x); }
    Fp get_challenge() {
        if (!input.empty() || output.empty()) duplexing();
        Fp r = output.back();
        output.pop_back();
        return r;
    }
    std::vector<Fp> get_n_challenges(size_t n) {
        std::vector<Fp> r;
        for (size_t i = 0; i < n; i++) r.push_back(get_challenge());
        return r;
    }
    Ext2 get_extension_challenge() { Fp a = get_challenge(); Fp b = get_challenge(); return Ext2(a, b); }
    PState compact() {
        if (!input.empty()) duplexing();
        output.clear();
        return state;
    }
};

// ------------------------------------------------------------------------------ A.7 FRI params
struct FriConfig {
    unsigned rate_bits = 2, cap_height = 4, proof_of_work_bits = 16;
    unsigned arity_bits = 4, final_poly_bits = 5;    // ConstantArityBits(4, 5)
    unsigned num_query_rounds = 37;
};
struct FriParams {
    FriConfig config;
    unsigned degree_bits = 0;
    std::vector<unsigned> reduction_arity_bits;
    unsigned total_arities() const { unsigned s = 0; for (unsigned a : reduction_arity_bits) s += a; return s; }
    size_t lde_size() const { return (size_t)1 << (degree_bits + config.rate_bits); }
};
static inline FriParams fri_params(const FriConfig& c, unsigned degree_bits) {
    FriParams p; p.config = c; p.degree_bits = degree_bits;
    unsigned d = degree_bits;
    while (d > c.final_poly_bits && d + c.rate_bits - c.arity_bits >= c.cap_height) {
        p.reduction_arity_bits.push_back(c.arity_bits);
        d -= c.arity_bits;
    }
    return p;
}

// -------------------------------------------------------------------- A.3 PolynomialBatch
struct PolynomialBatch {
    std::vector<std::vector<Fp>> polynomials;        // coefficients, length n each
    MerkleTree merkle_tree;                          // leaf j = LDE row bitrev(j)
    unsigned degree_log = 0, rate_bits = 0;

    static PolynomialBatch from_coeffs(std::vector<std::vector<Fp>> polys, unsigned rate_bits, unsigned cap_height) {
        PolynomialBatch b;
        size_t n = polys[0].size(), ncols = polys.size();
        b.degree_log = log2_strict(n);
        b.rate_bits = rate_bits;
        size_t N = n << rate_bits;
        unsigned lgN = b.degree_log + rate_bits;
        std::vector<std::vector<Fp>> lde(ncols);
        parallel_for(ncols, [&](size_t c) { lde[c] = lde_coset_values(polys[c], rate_bits); }, 1);
        std::vector<std::vector<Fp>> leaves(N);
        parallel_for(N, [&](size_t j) {
            size_t r = reverse_bits(j, lgN);
            std::vector<Fp>& row = leaves[j];
            row.resize(ncols);
            for (size_t c = 0; c < ncols; c++) row[c] = lde[c][r];
        });
        b.polynomials = std::move(polys);
        b.merkle_tree = MerkleTree(std::move(leaves), cap_height);
        return b;
    }
    static PolynomialBatch from_values(std::vector<std::vector<Fp>> values, unsigned rate_bits, unsigned cap_height) {
        parallel_for(values.size(), [&](size_t c) { ifft_inplace(values[c].data(), values[c].size()); }, 1);
        return from_coeffs(std::move(values), rate_bits, cap_height);
    }
    // get_lde_values(i, step) = leaves[bitrev(i*step)]
    const std::vector<Fp>& get_lde_values(size_t index, size_t step) const {
        return merkle_tree.leaves[reverse_bits(index * step, degree_log + rate_bits)];
    }
};

// ------------------------------------------------------------------ FRI instance / proof shapes
struct FriPolynomialInfo { unsigned oracle_index, polynomial_index; };
struct FriBatchInfo { Ext2 point; std::vector<FriPolynomialInfo> polynomials; };
struct FriInstanceInfo { std::vector<unsigned> oracle_num_polys; std::vector<FriBatchInfo> batches; };
struct FriOpenings { std::vector<std::vector<Ext2>> batches; };

struct FriQueryStep { std::vector<Ext2> evals; MerkleProof merkle_proof; };
struct FriQueryRound {
    std::vector<std::pair<std::vector<Fp>, MerkleProof>> initial_trees_proof;
    std::vector<FriQueryStep> steps;
};
struct FriProof {
    std::vector<MerkleCap> commit_phase_merkle_caps;
    std::vector<FriQueryRound> query_round_proofs;
    std::vector<Ext2> final_poly;
    Fp pow_witness;
};
struct FriChallenges {
    Ext2 fri_alpha;
    std::vector<Ext2> fri_betas;
    Fp fri_pow_response;
    std::vector<size_t> fri_query_indices;
};

// A.9: smallest witness w such that, with w written after the buffered inputs and the state
// permuted, state[7] has >= pow_bits leading zeros.  (Upstream: find_any over a parallel range.)
static inline Fp fri_proof_of_work(Challenger& ch, const FriConfig& cfg) {
    unsigned min_lz = cfg.proof_of_work_bits;        // + (64 - 64)
    PState base = ch.state;
    size_t pos = ch.input.size();
    for (size_t i = 0; i < pos; i++) base[i] = ch.input[i];
    std::atomic<u64> best(~(u64)0);
    const u64 block = 1 << 12;
    std::atomic<u64> next(0);
    auto worker = [&] {
        for (;;) {
            u64 b = next.fetch_add(block);
            if (b >= best.load()) break;
            for (u64 w = b; w < b + block; w++) {
                PState s = base;
                s[pos] = Fp(w);
                poseidon(s);
                u64 r = s[7].v;
                unsigned lz = r ? (unsigned)__builtin_clzll(r) : 64;
                if (lz >= min_lz) {
                    u64 cur = best.load();
                    while (w < cur && !best.compare_exchange_weak(cur, w)) {}
                    break;
                }
            }
        }
    };
    int nt = get_threads();
    std::vector<std::thread> ts;
    for (int t = 1; t < nt; t++) ts.emplace_back(worker);
    worker();
    for (auto& t : ts) t.join();
    Fp w(best.load());
    ch.observe_element(w);
    Fp resp = ch.get_challenge();
    unsigned lz = resp.v ? (unsigned)__builtin_clzll(resp.v) : 64;
    if (lz < min_lz) throw std::runtime_error("PoW self-check failed");
    return w;
}

// A.8 commit phase + A.9 + A.10
static inline FriProof fri_proof(const std::vector<const MerkleTree*>& initial_trees, std::vector<Ext2> coeffs,
                                 std::vector<Ext2> values, Challenger& ch, const FriParams& params) {
    size_t n = values.size();
    FriProof proof;
    std::vector<MerkleTree> trees;
    Fp shift(GL_GENERATOR);
    for (unsigned arity_bits : params.reduction_arity_bits) {
        size_t arity = (size_t)1 << arity_bits;
        bit_reverse_permute(values.data(), values.size());
        std::vector<std::vector<Fp>> leaves(values.size() / arity);
        for (size_t i = 0; i < leaves.size(); i++) {
            leaves[i].resize(2 * arity);
            for (size_t j = 0; j < arity; j++) {
                leaves[i][2 * j] = values[i * arity + j].a;
                leaves[i][2 * j + 1] = values[i * arity + j].b;
            }
        }
        trees.emplace_back(std::move(leaves), params.config.cap_height);
        ch.observe_cap(trees.back().cap);
        Ext2 beta = ch.get_extension_challenge();
        std::vector<Ext2> folded(coeffs.size() / arity);
        for (size_t i = 0; i < folded.size(); i++) {
            Ext2 acc;
            for (size_t j = arity; j-- > 0;) acc = acc * beta + coeffs[i * arity + j];   // reduce_with_powers
            folded[i] = acc;
        }
        coeffs = std::move(folded);
        shift = shift.pow(arity);
        values = coeffs;
        ext_coset_fft_inplace(values, shift);
    }
    coeffs.resize(coeffs.size() >> params.config.rate_bits);
    ch.observe_exts(coeffs);
    proof.final_poly = coeffs;
    proof.pow_witness = fri_proof_of_work(ch, params.config);
    for (unsigned q = 0; q < params.config.num_query_rounds; q++) {
        FriQueryRound round;
        size_t x_index = (size_t)(ch.get_challenge().v % n);
        for (const MerkleTree* t : initial_trees) round.initial_trees_proof.emplace_back(t->get(x_index), t->prove(x_index));
        for (size_t i = 0; i < trees.size(); i++) {
            unsigned ab = params.reduction_arity_bits[i];
            const std::vector<Fp>& leaf = trees[i].get(x_index >> ab);
            FriQueryStep st;
            for (size_t j = 0; j < leaf.size(); j += 2) st.evals.emplace_back(leaf[j], leaf[j + 1]);
            st.merkle_proof = trees[i].prove(x_index >> ab);
            round.steps.push_back(std::move(st));
            x_index >>= ab;
        }
        proof.query_round_proofs.push_back(std::move(round));
    }
    for (const MerkleTree& t : trees) proof.commit_phase_merkle_caps.push_back(t.cap);
    return proof;
}

// A.8 PolynomialBatch::prove_openings
static inline FriProof prove_openings(const FriInstanceInfo& instance, const std::vector<const PolynomialBatch*>& oracles,
                                      Challenger& ch, const FriParams& params) {
    Ext2 alpha = ch.get_extension_challenge();
    size_t n = oracles[0]->polynomials[0].size();
    std::vector<Ext2> final_poly;                    // empty = zero
    for (const FriBatchInfo& batch : instance.batches) {
        // composition = sum_k alpha^k p_k  (ReducingFactor::reduce_polys_base)
        std::vector<Ext2> comp(n);
        size_t cnt = batch.polynomials.size();
        std::vector<Ext2> apow(cnt);
        { Ext2 cur = Ext2::one(); for (size_t k = 0; k < cnt; k++) { apow[k] = cur; cur *= alpha; } }
        parallel_for(n, [&](size_t i) {
            Ext2 acc;
            for (size_t k = 0; k < cnt; k++) {
                const FriPolynomialInfo& pi = batch.polynomials[k];
                acc += apow[k] * oracles[pi.oracle_index]->polynomials[pi.polynomial_index][i];
            }
            comp[i] = acc;
        });
        // quotient = (comp - comp(point)) / (X - point): synthetic division, then pad one zero.
        std::vector<Ext2> quot(n);
        Ext2 carry;
        for (size_t i = n; i-- > 0;) {
            Ext2 c = comp[i] + carry * batch.point;  // b_{i} = a_i + z*b_{i+1}
            if (i > 0) quot[i - 1] = c;
            carry = c;
        }
        quot[n - 1] = Ext2();
        // alpha.shift_poly(final): final *= alpha^cnt ; final += quotient
        Ext2 sh = alpha.pow(cnt);
        if (final_poly.empty()) final_poly.assign(n, Ext2());
        for (size_t i = 0; i < n; i++) final_poly[i] = final_poly[i] * sh + quot[i];
    }
    std::vector<Ext2> lde_coeffs(n << params.config.rate_bits);
    std::copy(final_poly.begin(), final_poly.end(), lde_coeffs.begin());
    std::vector<Ext2> lde_values = lde_coeffs;
    ext_coset_fft_inplace(lde_values, Fp(GL_GENERATOR));
    std::vector<const MerkleTree*> trees;
    for (const PolynomialBatch* b : oracles) trees.push_back(&b->merkle_tree);
    return fri_proof(trees, std::move(lde_coeffs), std::move(lde_values), ch, params);
}

// Challenger::fri_challenges (verifier side replay)
static inline FriChallenges fri_challenges(Challenger& ch, const std::vector<MerkleCap>& caps, const std::vector<Ext2>& final_poly,
                                           Fp pow_witness, unsigned degree_bits, const FriConfig& cfg) {
    FriChallenges c;
    size_t lde_size = (size_t)1 << (degree_bits + cfg.rate_bits);
    c.fri_alpha = ch.get_extension_challenge();
    for (const MerkleCap& cap : caps) { ch.observe_cap(cap); c.fri_betas.push_back(ch.get_extension_challenge()); }
    ch.observe_exts(final_poly);
    ch.observe_element(pow_witness);
    c.fri_pow_response = ch.get_challenge();
    for (unsigned i = 0; i < cfg.num_query_rounds; i++) c.fri_query_indices.push_back((size_t)(ch.get_challenge().v % lde_size));
    return c;
}

// Interpolate {(x_i, y_i)} and evaluate at z (Lagrange; plonky2 interpolate/barycentric_weights).
static inline Ext2 interpolate_eval(const std::vector<Ext2>& xs, const std::vector<Ext2>& ys, Ext2 z) {
    size_t n = xs.size();
    Ext2 acc;
    for (size_t i = 0; i < n; i++) {
        Ext2 num = Ext2::one(), den = Ext2::one();
        for (size_t j = 0; j < n; j++) if (j != i) { num *= (z - xs[j]); den *= (xs[i] - xs[j]); }
        acc += ys[i] * num * den.inverse();
    }
    return acc;
}

// plonky2 fri/verifier.rs compute_evaluation
static inline Ext2 compute_evaluation(Fp x, size_t x_index_within_coset, unsigned arity_bits, const std::vector<Ext2>& evals_in,
                                      Ext2 beta) {
    size_t arity = (size_t)1 << arity_bits;
    Fp g = primitive_root_of_unity(arity_bits);
    std::vector<Ext2> evals = evals_in;
    bit_reverse_permute(evals.data(), arity);
    size_t rev = reverse_bits(x_index_within_coset, arity_bits);
    Fp coset_start = x * g.pow(arity - rev);
    std::vector<Ext2> pts(arity);
    Fp cur = Fp::one();
    for (size_t i = 0; i < arity; i++) { pts[i] = Ext2::from_base(coset_start * cur); cur *= g; }
    return interpolate_eval(pts, evals, beta);
}

// plonky2 fri/verifier.rs verify_fri_proof (A.10), incl. validate_fri_proof_shape.
static inline void verify_fri_proof(const FriInstanceInfo& instance, const FriOpenings& openings, const FriChallenges& ch,
                                    const std::vector<MerkleCap>& initial_caps, const FriProof& proof, const FriParams& params) {
    const FriConfig& cfg = params.config;
    // shape
    ORC_ENSURE(proof.commit_phase_merkle_caps.size() == params.reduction_arity_bits.size(), "fri shape: caps");
    for (const MerkleCap& c : proof.commit_phase_merkle_caps) ORC_ENSURE(c.size() == ((size_t)1 << cfg.cap_height), "fri shape: cap height");
    ORC_ENSURE(proof.query_round_proofs.size() == cfg.num_query_rounds, "Number of query rounds does not match config.");
    unsigned lde_bits = params.degree_bits + cfg.rate_bits;
    for (const FriQueryRound& r : proof.query_round_proofs) {
        ORC_ENSURE(r.initial_trees_proof.size() == instance.oracle_num_polys.size(), "fri shape: oracles");
        for (size_t o = 0; o < r.initial_trees_proof.size(); o++) {
            ORC_ENSURE(r.initial_trees_proof[o].first.size() == instance.oracle_num_polys[o], "fri shape: leaf len");
            ORC_ENSURE(r.initial_trees_proof[o].second.siblings.size() == lde_bits - cfg.cap_height, "fri shape: path len");
        }
        ORC_ENSURE(r.steps.size() == params.reduction_arity_bits.size(), "fri shape: steps");
        unsigned codeword_bits = lde_bits;
        for (size_t i = 0; i < r.steps.size(); i++) {
            unsigned ab = params.reduction_arity_bits[i];
            ORC_ENSURE(r.steps[i].evals.size() == ((size_t)1 << ab), "fri shape: evals");
            codeword_bits -= ab;
            ORC_ENSURE(r.steps[i].merkle_proof.siblings.size() == codeword_bits - cfg.cap_height, "fri shape: step path");
        }
    }
    ORC_ENSURE(proof.final_poly.size() == ((size_t)1 << (params.degree_bits - params.total_arities())), "fri shape: final poly");

    size_t n = params.lde_size();
    unsigned lz = ch.fri_pow_response.v ? (unsigned)__builtin_clzll(ch.fri_pow_response.v) : 64;
    ORC_ENSURE(lz >= cfg.proof_of_work_bits, "Invalid proof of work witness.");

    // PrecomputedReducedOpenings
    std::vector<Ext2> reduced_openings;
    for (const std::vector<Ext2>& b : openings.batches) {
        Ext2 acc;
        for (size_t k = b.size(); k-- > 0;) acc = acc * ch.fri_alpha + b[k];
        reduced_openings.push_back(acc);
    }
    ORC_ENSURE(reduced_openings.size() == instance.batches.size(), "openings/batches mismatch");
    unsigned log_n = log2_strict(n);
    for (size_t q = 0; q < proof.query_round_proofs.size(); q++) {
        size_t x_index = ch.fri_query_indices[q];
        const FriQueryRound& round = proof.query_round_proofs[q];
        for (size_t o = 0; o < initial_caps.size(); o++)
            verify_merkle_proof_to_cap(round.initial_trees_proof[o].first, x_index, initial_caps[o], round.initial_trees_proof[o].second);
        Fp subgroup_x = Fp(GL_GENERATOR) * primitive_root_of_unity(log_n).pow(reverse_bits(x_index, log_n));
        // fri_combine_initial
        Ext2 sum;
        for (size_t b = 0; b < instance.batches.size(); b++) {
            const FriBatchInfo& batch = instance.batches[b];
            Ext2 red;
            for (size_t k = batch.polynomials.size(); k-- > 0;) {
                const FriPolynomialInfo& pi = batch.polynomials[k];
                red = red * ch.fri_alpha + Ext2::from_base(round.initial_trees_proof[pi.oracle_index].first[pi.polynomial_index]);
            }
            Ext2 numerator = red - reduced_openings[b];
            Ext2 denominator = Ext2::from_base(subgroup_x) - batch.point;
            sum = sum * ch.fri_alpha.pow(batch.polynomials.size());
            sum += numerator * denominator.inverse();
        }
        Ext2 old_eval = sum;
        for (size_t i = 0; i < params.reduction_arity_bits.size(); i++) {
            unsigned ab = params.reduction_arity_bits[i];
            size_t arity = (size_t)1 << ab;
            const std::vector<Ext2>& evals = round.steps[i].evals;
            size_t coset_index = x_index >> ab, within = x_index & (arity - 1);
            ORC_ENSURE(evals[within] == old_eval, "FRI fold consistency check failed.");
            old_eval = compute_evaluation(subgroup_x, within, ab, evals, ch.fri_betas[i]);
            std::vector<Fp> flat(2 * arity);
            for (size_t j = 0; j < arity; j++) { flat[2 * j] = evals[j].a; flat[2 * j + 1] = evals[j].b; }
            verify_merkle_proof_to_cap(flat, coset_index, proof.commit_phase_merkle_caps[i], round.steps[i].merkle_proof);
            subgroup_x = subgroup_x.exp_power_of_2(ab);
            x_index = coset_index;
        }
        ORC_ENSURE(poly_eval_ext(proof.final_poly, Ext2::from_base(subgroup_x)) == old_eval, "Final polynomial evaluation is invalid.");
    }
}

}  // namespace orc
