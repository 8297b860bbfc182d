// ORACLE (test infrastructure): CPU restatement of the reference's STARK layer over a `System` of
// tables — prove_with_traces / prove_single_table / compute_quotient_polys (prover/src/prover.rs:
// 130-232, 441-641, 645-789), cross_table_lookup_data and the CTL checks (cross_table_lookup.rs:
// 634-872, 1006-1150, 1415-1452), the logUp lookups (lookup.rs:46-198), StarkOpeningSet (proof.rs:
// 299-367), the Fiat-Shamir order (get_challenges.rs:91-148,190-233) and verify_proof (verifier.rs:
// 27-354).  The table descriptions and constraint templates are the shared headers under
// zkm_b200/csrc/tables/ (one source for the CUDA kernels and this oracle, like the reference's
// eval_packed_generic); everything else here is independent of the product code.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference may use this.
#pragma once
#include "plonky2_restated.h"
#include "../zkm_b200/csrc/tables/registry.h"

namespace orc {

using zkm::tables::Column;
using zkm::tables::Filter;
using zkm::tables::Lookup;
using zkm::tables::TableWithColumns;
using zkm::tables::CrossTableLookup;
using zkm::tables::System;
using zkm::tables::TableLayout;
using zkm::tables::CtlZInfo;

typedef std::vector<std::vector<Fp>> Trace;           // column-major: trace[c][row]

struct StarkConfig {
    FriConfig fri;
    unsigned num_challenges = 2;
};
struct GrandProductChallenge { Fp beta, gamma; };

struct PublicValues { uint32_t roots_before[8], roots_after[8]; std::vector<uint8_t> userdata; };

// ------------------------------------------------------------- Column / Filter on trace values
// Column::eval_table (cross_table_lookup.rs:266-285): next-row terms are dropped on the last row.
static inline Fp col_eval_table(const Column& c, const Trace& t, size_t row) {
    Fp res;
    for (auto& p : c.lin) res += t[p.first][row] * Fp(p.second);
    res += Fp(c.constant);
    if (!c.next.empty() && row + 1 < t[0].size())
        for (auto& p : c.next) res += t[p.first][row + 1] * Fp(p.second);
    return res;
}
static inline Fp filter_eval_table(const Filter& f, const Trace& t, size_t row) {
    Fp s;
    for (auto& pr : f.products) s += col_eval_table(pr.first, t, row) * col_eval_table(pr.second, t, row);
    for (auto& c : f.constants) s += col_eval_table(c, t, row);
    return s;
}
// GrandProductChallenge::combine = reduce_with_powers(terms, beta) + gamma (:494-504)
template <class P>
static inline P combine(const std::vector<P>& terms, P beta, P gamma) {
    P acc = P(0);
    for (size_t i = terms.size(); i-- > 0;) acc = acc * beta + terms[i];
    return acc + gamma;
}

// get_helper_cols (cross_table_lookup.rs:709-795)
static inline Trace get_helper_cols(const Trace& trace, const std::vector<std::pair<const std::vector<Column>*, const Filter*>>& cf,
                                    GrandProductChallenge ch) {
    size_t degree = trace[0].size();
    Trace helpers;
    for (size_t g = 0; g < cf.size(); g += 2) {
        std::vector<Fp> acc;
        for (size_t k = g; k < std::min(cf.size(), g + 2); k++) {
            std::vector<Fp> filt(degree), comb(degree);
            parallel_for(degree, [&](size_t d) {
                Fp f = cf[k].second->present ? filter_eval_table(*cf[k].second, trace, d) : Fp::one();
                filt[d] = f;
                if (f == Fp::one()) {
                    std::vector<Fp> ev;
                    for (const Column& c : *cf[k].first) ev.push_back(col_eval_table(c, trace, d));
                    comb[d] = combine(ev, ch.beta, ch.gamma);
                } else {
                    if (!f.is_zero()) throw std::runtime_error("Non-binary filter?");
                    comb[d] = Fp::one();
                }
            });
            for (size_t d = 0; d < degree; d++) if (comb[d].is_zero()) throw std::runtime_error("batch inverse of zero");
            std::vector<Fp> inv = batch_inverse(comb);
            for (size_t d = 0; d < degree; d++) if (filt[d].is_zero()) inv[d] = Fp::zero();
            if (k == g) acc = std::move(inv);
            else for (size_t d = 0; d < degree; d++) acc[d] += inv[d];
        }
        helpers.push_back(std::move(acc));
    }
    return helpers;
}

// partial_sums (:841-872): returns helper columns (only if > 1 column set) followed by Z.
static inline Trace partial_sums(const Trace& trace, const std::vector<TableWithColumns>& parts, GrandProductChallenge ch) {
    std::vector<std::pair<const std::vector<Column>*, const Filter*>> cf;
    for (const TableWithColumns& p : parts) cf.push_back({&p.columns, &p.filter});
    size_t degree = trace[0].size();
    Trace helpers = get_helper_cols(trace, cf, ch);
    std::vector<Fp> z(degree);
    Fp run;
    for (size_t i = degree; i-- > 0;) {
        for (auto& h : helpers) run += h[i];
        z[i] = run;
    }
    if (parts.size() > 1) { helpers.push_back(std::move(z)); return helpers; }
    Trace only; only.push_back(std::move(z));
    return only;
}

// lookup_helper_columns (lookup.rs:46-124)
static inline Trace lookup_helper_columns(const Lookup& lk, const Trace& trace, Fp challenge) {
    size_t degree = trace[0].size();
    std::vector<std::vector<Column>> looking;
    for (const Column& c : lk.columns) looking.push_back({c});
    std::vector<std::pair<const std::vector<Column>*, const Filter*>> cf;
    for (size_t i = 0; i < looking.size(); i++) cf.push_back({&looking[i], &lk.filter_columns[i]});
    GrandProductChallenge gc{Fp::one(), challenge};
    Trace helpers = get_helper_cols(trace, cf, gc);
    std::vector<Fp> table(degree);
    for (size_t i = 0; i < degree; i++) table[i] = challenge + col_eval_table(lk.table_column, trace, i);
    std::vector<Fp> tinv = batch_inverse(table);
    std::vector<Fp> z(degree);
    for (size_t i = 0; i + 1 < degree; i++) {
        Fp x;
        for (auto& h : helpers) x += h[i];
        x -= col_eval_table(lk.frequencies_column, trace, i) * tinv[i];
        z[i + 1] = z[i] + x;
    }
    helpers.push_back(std::move(z));
    return helpers;
}

// All auxiliary polynomials of one table in commitment order (prover.rs:469-509).
static inline Trace auxiliary_columns(const TableLayout& L, const Trace& trace, const std::vector<GrandProductChallenge>& chs) {
    Trace aux;
    for (const Lookup& lk : L.lookups)
        for (const GrandProductChallenge& c : chs)
            for (auto& col : lookup_helper_columns(lk, trace, c.beta)) aux.push_back(std::move(col));
    Trace zs;
    for (const CtlZInfo& z : L.zs) {
        Trace hz = partial_sums(trace, z.parts, chs[z.challenge]);
        for (size_t i = 0; i + 1 < hz.size(); i++) aux.push_back(std::move(hz[i]));
        zs.push_back(std::move(hz.back()));
    }
    for (auto& z : zs) aux.push_back(std::move(z));
    return aux;
}

// ------------------------------------------------------------------ constraint evaluation
// ConstraintConsumer (constraint_consumer.rs:10-75)
template <class P>
struct Consumer {
    std::vector<P> alphas, accs;
    P z_last, l_first, l_last;
    Consumer(const std::vector<P>& a, P zl, P lf, P ll) : alphas(a), accs(a.size(), P(0)), z_last(zl), l_first(lf), l_last(ll) {}
    void constraint(P c) { for (size_t i = 0; i < alphas.size(); i++) accs[i] = accs[i] * alphas[i] + c; }
    void constraint_transition(P c) { constraint(c * z_last); }
    void constraint_first_row(P c) { constraint(c * l_first); }
    void constraint_last_row(P c) { constraint(c * l_last); }
    void checkpoint() {}            // device consumers re-converge the CTA here (instruction-cache sharing); no-op on the CPU
};

template <class P>
struct RowView {
    const P* p;
    P operator[](int i) const { return p[i]; }
};

// Column::eval_with_next / eval (:247-263), Filter::eval_filter (:56-70)
template <class P>
static inline P col_eval_with_next(const Column& c, const P* lv, const P* nv) {
    P r = P(0);
    for (auto& p : c.lin) r = r + lv[p.first] * P(p.second);
    for (auto& p : c.next) r = r + nv[p.first] * P(p.second);
    return r + P(c.constant);
}
template <class P>
static inline P col_eval(const Column& c, const P* lv) {
    P r = P(0);
    for (auto& p : c.lin) r = r + lv[p.first] * P(p.second);
    return r + P(c.constant);
}
template <class P>
static inline P filter_eval(const Filter& f, const P* lv, const P* nv) {
    if (!f.present) return P(1);
    P s = P(0);
    for (auto& pr : f.products) s = s + col_eval_with_next(pr.first, lv, nv) * col_eval_with_next(pr.second, lv, nv);
    for (auto& c : f.constants) s = s + col_eval_with_next(c, lv, nv);
    return s;
}

// eval_helper_columns (cross_table_lookup.rs:1006-1057)
template <class P>
static inline void eval_helper_columns(const std::vector<const Filter*>& filter, const std::vector<std::vector<P>>& columns, const P* lv,
                                       const P* nv, const P* helpers, size_t num_helpers, P beta, P gamma, Consumer<P>& yc) {
    if (num_helpers == 0) return;
    for (size_t j = 0; 2 * j < columns.size(); j++) {
        size_t len = std::min<size_t>(2, columns.size() - 2 * j);
        P h = helpers[j];
        if (len == 2) {
            P c0 = combine(columns[2 * j], beta, gamma), c1 = combine(columns[2 * j + 1], beta, gamma);
            P f0 = filter_eval(*filter[2 * j], lv, nv), f1 = filter_eval(*filter[2 * j + 1], lv, nv);
            yc.constraint(c1 * c0 * h - f0 * c1 - f1 * c0);
        } else {
            P c = combine(columns[2 * j], beta, gamma);
            P f0 = filter_eval(*filter[2 * j], lv, nv);
            yc.constraint(c * h - f0);
        }
    }
}

// eval_packed_lookups_generic (lookup.rs:138-198).  aux_l / aux_n: auxiliary polys at local / next.
template <class P>
static inline void eval_lookups(const std::vector<Lookup>& lookups, const P* lv, const P* nv, const P* aux_l, const P* aux_n,
                                const std::vector<Fp>& challenges, Consumer<P>& yc) {
    size_t start = 0;
    for (const Lookup& lk : lookups) {
        size_t nh = lk.num_helper_columns(zkm::tables::CONSTRAINT_DEGREE);
        for (Fp chf : challenges) {
            std::vector<std::vector<P>> cols;
            for (const Column& c : lk.columns) cols.push_back({col_eval_with_next(c, lv, nv)});
            std::vector<const Filter*> fl;
            for (const Filter& f : lk.filter_columns) fl.push_back(&f);
            eval_helper_columns(fl, cols, lv, nv, aux_l + start, nh - 1, P(1), P(chf.v), yc);
            P challenge = P(chf.v);
            P z = aux_l[start + nh - 1], next_z = aux_n[start + nh - 1];
            P twc = col_eval(lk.table_column, lv) + challenge;
            P hs = P(0);
            for (size_t i = 0; i + 1 < nh; i++) hs = hs + aux_l[start + i];
            P y = hs * twc - col_eval(lk.frequencies_column, lv);
            yc.constraint_first_row(z);
            yc.constraint((next_z - z) * twc - y);
            start += nh;
        }
    }
}

// One CtlCheckVars (cross_table_lookup.rs:876-889) in oracle form.
template <class P>
struct CtlVars {
    std::vector<P> helpers;
    P local_z, next_z;
    GrandProductChallenge ch;
    std::vector<const std::vector<Column>*> columns;
    std::vector<const Filter*> filter;
};

// eval_cross_table_lookup_checks (:1067-1150)
template <class P>
static inline void eval_ctl_checks(const P* lv, const P* nv, const std::vector<CtlVars<P>>& vars, Consumer<P>& yc) {
    for (const CtlVars<P>& v : vars) {
        P beta = P(v.ch.beta.v), gamma = P(v.ch.gamma.v);
        std::vector<std::vector<P>> evals;
        for (auto* cols : v.columns) {
            std::vector<P> e;
            for (const Column& c : *cols) e.push_back(col_eval_with_next(c, lv, nv));
            evals.push_back(std::move(e));
        }
        eval_helper_columns(v.filter, evals, lv, nv, v.helpers.data(), v.helpers.size(), beta, gamma, yc);
        if (!v.helpers.empty()) {
            P hs = P(0);
            for (const P& h : v.helpers) hs = hs + h;
            yc.constraint_last_row(v.local_z - hs);
            yc.constraint_transition(v.local_z - v.next_z - hs);
        } else if (v.columns.size() > 1) {
            P c0 = combine(evals[0], beta, gamma), c1 = combine(evals[1], beta, gamma);
            P f0 = filter_eval(*v.filter[0], lv, nv), f1 = filter_eval(*v.filter[1], lv, nv);
            yc.constraint_last_row(c0 * c1 * v.local_z - f0 * c1 - f1 * c0);
            yc.constraint_transition(c0 * c1 * (v.local_z - v.next_z) - f0 * c1 - f1 * c0);
        } else {
            P c0 = combine(evals[0], beta, gamma);
            P f0 = filter_eval(*v.filter[0], lv, nv);
            yc.constraint_last_row(c0 * v.local_z - f0);
            yc.constraint_transition(c0 * (v.local_z - v.next_z) - f0);
        }
    }
}

// eval_vanishing_poly (vanishing_poly.rs:17-46)
template <class P>
static inline void eval_vanishing_poly(const TableLayout& L, const P* lv, const P* nv, const P* aux_l, const P* aux_n,
                                       const std::vector<Fp>& lookup_challenges, const std::vector<CtlVars<P>>& ctl_vars, Consumer<P>& yc) {
    RowView<P> l{lv}, n{nv};
    if (!zkm::tables::eval_table<P, RowView<P>, Consumer<P>>(L.kind, l, n, yc))
        throw std::runtime_error(std::string("constraints of table ") + zkm::tables::table_name(L.kind) + " are not available");
    if (!L.lookups.empty()) eval_lookups(L.lookups, lv, nv, aux_l, aux_n, lookup_challenges, yc);
    eval_ctl_checks(lv, nv, ctl_vars, yc);
}

// Prover-side CtlCheckVars from the layout (prover.rs:719-748)
template <class P>
static inline std::vector<CtlVars<P>> ctl_vars_from_layout(const TableLayout& L, const P* aux_l, const P* aux_n,
                                                           const std::vector<GrandProductChallenge>& chs) {
    std::vector<CtlVars<P>> out;
    int start = 0;
    for (size_t i = 0; i < L.zs.size(); i++) {
        CtlVars<P> v;
        const CtlZInfo& z = L.zs[i];
        for (int j = 0; j < z.num_helpers; j++) v.helpers.push_back(aux_l[L.num_lookup_cols + start + j]);
        v.local_z = aux_l[L.num_lookup_cols + L.num_ctl_helpers + i];
        v.next_z = aux_n[L.num_lookup_cols + L.num_ctl_helpers + i];
        v.ch = chs[z.challenge];
        for (const TableWithColumns& p : z.parts) { v.columns.push_back(&p.columns); v.filter.push_back(&p.filter); }
        start += z.num_helpers;
        out.push_back(std::move(v));
    }
    return out;
}

// --------------------------------------------------------------------------------- proof shapes
struct StarkOpeningSet {
    std::vector<Ext2> local_values, next_values, auxiliary_polys, auxiliary_polys_next;
    std::vector<Fp> ctl_zs_first;
    std::vector<Ext2> quotient_polys;
    FriOpenings to_fri_openings() const {              // proof.rs:336-367
        FriOpenings o;
        std::vector<Ext2> a = local_values;
        a.insert(a.end(), auxiliary_polys.begin(), auxiliary_polys.end());
        a.insert(a.end(), quotient_polys.begin(), quotient_polys.end());
        std::vector<Ext2> b = next_values;
        b.insert(b.end(), auxiliary_polys_next.begin(), auxiliary_polys_next.end());
        std::vector<Ext2> c;
        for (Fp z : ctl_zs_first) c.push_back(Ext2::from_base(z));
        o.batches = {a, b, c};
        return o;
    }
};
struct StarkProof {
    MerkleCap trace_cap, auxiliary_polys_cap, quotient_polys_cap;
    StarkOpeningSet openings;
    FriProof opening_proof;
    unsigned recover_degree_bits(const StarkConfig& cfg) const {       // proof.rs:203-211
        if (opening_proof.query_round_proofs.empty() || opening_proof.query_round_proofs[0].initial_trees_proof.empty())
            throw VerifyError("proof has no query rounds");
        return cfg.fri.cap_height + (unsigned)opening_proof.query_round_proofs[0].initial_trees_proof[0].second.siblings.size() -
               cfg.fri.rate_bits;
    }
};
struct StarkProofWithMetadata { PState init_challenger_state; StarkProof proof; };
struct AllProof {
    std::vector<StarkProofWithMetadata> stark_proofs;
    std::vector<GrandProductChallenge> ctl_challenges;
    PublicValues public_values;
};

// Stark::fri_instance (stark.rs:91-153)
static inline FriInstanceInfo fri_instance(const TableLayout& L, Ext2 zeta, Fp g, const StarkConfig& cfg) {
    FriInstanceInfo fi;
    unsigned ntrace = L.ncols, naux = L.num_aux(), nq = zkm::tables::QUOTIENT_DEGREE_FACTOR * cfg.num_challenges;
    fi.oracle_num_polys = {ntrace, naux, nq};
    FriBatchInfo zb, znb, cb;
    zb.point = zeta;
    for (unsigned i = 0; i < ntrace; i++) zb.polynomials.push_back({0, i});
    for (unsigned i = 0; i < naux; i++) zb.polynomials.push_back({1, i});
    for (unsigned i = 0; i < nq; i++) zb.polynomials.push_back({2, i});
    znb.point = zeta * g;
    for (unsigned i = 0; i < ntrace; i++) znb.polynomials.push_back({0, i});
    for (unsigned i = 0; i < naux; i++) znb.polynomials.push_back({1, i});
    cb.point = Ext2::one();
    for (unsigned i = L.num_lookup_cols + L.num_ctl_helpers; i < naux; i++) cb.polynomials.push_back({1, i});
    fi.batches = {zb, znb, cb};
    return fi;
}

static inline void observe_public_values(Challenger& ch, const PublicValues& pv) {        // get_challenges.rs:14-21,91-105
    for (int i = 0; i < 8; i++) ch.observe_element(Fp(pv.roots_before[i]));
    for (int i = 0; i < 8; i++) ch.observe_element(Fp(pv.roots_after[i]));
    for (uint8_t b : pv.userdata) ch.observe_element(Fp(b));
}
static inline void observe_openings(Challenger& ch, const FriOpenings& o) {
    for (auto& b : o.batches) ch.observe_exts(b);
}

// --------------------------------------------------------------------------------------- prover
// compute_quotient_polys (prover.rs:645-789): returns num_challenges polynomials of 2n coefficients.
static inline std::vector<std::vector<Fp>> compute_quotient_polys(const TableLayout& L, const PolynomialBatch& trace_c,
                                                                  const PolynomialBatch& aux_c, const std::vector<GrandProductChallenge>& chs,
                                                                  const std::vector<Fp>& alphas, unsigned degree_bits, const StarkConfig& cfg) {
    const unsigned qdb = 1;                                   // log2_ceil(quotient_degree_factor = 2)
    size_t degree = (size_t)1 << degree_bits;
    unsigned rate_bits = cfg.fri.rate_bits;
    if (qdb > rate_bits) throw std::runtime_error("Having constraints of degree higher than the rate is not supported yet.");
    size_t step = (size_t)1 << (rate_bits - qdb), next_step = (size_t)1 << qdb, size = degree << qdb;
    // Lagrange selectors on the coset, exactly as the reference builds them: selector -> lde_onto_coset
    auto lagrange = [&](size_t k) {
        std::vector<Fp> v(degree);
        v[k] = Fp::one();
        ifft_inplace(v.data(), degree);
        return lde_coset_values(v, qdb);
    };
    std::vector<Fp> l_first = lagrange(0), l_last = lagrange(degree - 1);
    // ZeroPolyOnCoset(degree_bits, qdb): Z_H(7 w^i) = 7^n * (w_{2n}^n)^i - 1, 2 distinct values
    std::vector<Fp> zh_inv(1 << qdb);
    {
        Fp gn = Fp(GL_GENERATOR).exp_power_of_2(degree_bits);
        Fp wn = primitive_root_of_unity(qdb);
        Fp cur = Fp::one();
        for (size_t i = 0; i < zh_inv.size(); i++) { zh_inv[i] = (gn * cur - Fp::one()).inverse(); cur *= wn; }
    }
    Fp last = primitive_root_of_unity(degree_bits).inverse();
    Fp w = primitive_root_of_unity(degree_bits + qdb);
    std::vector<Fp> coset(size);
    { Fp cur(GL_GENERATOR); for (size_t i = 0; i < size; i++) { coset[i] = cur; cur *= w; } }
    std::vector<Fp> lookup_ch;
    for (auto& c : chs) lookup_ch.push_back(c.beta);
    std::vector<std::vector<Fp>> q(alphas.size(), std::vector<Fp>(size));
    parallel_for(size, [&](size_t i) {
        size_t inext = (i + next_step) % size;
        const std::vector<Fp>&lv = trace_c.get_lde_values(i, step), &nv = trace_c.get_lde_values(inext, step);
        const std::vector<Fp>&al = aux_c.get_lde_values(i, step), &an = aux_c.get_lde_values(inext, step);
        Consumer<Fp> yc(alphas, coset[i] - last, l_first[i], l_last[i]);
        auto cv = ctl_vars_from_layout<Fp>(L, al.data(), an.data(), chs);
        eval_vanishing_poly<Fp>(L, lv.data(), nv.data(), al.data(), an.data(), lookup_ch, cv, yc);
        for (size_t j = 0; j < alphas.size(); j++) q[j][i] = yc.accs[j] * zh_inv[i % zh_inv.size()];
    });
    for (auto& p : q) coset_ifft_inplace(p.data(), size, Fp(GL_GENERATOR));
    return q;
}

static inline StarkProofWithMetadata prove_single_table(const TableLayout& L, const StarkConfig& cfg, const Trace& trace,
                                                        const PolynomialBatch& trace_c, const std::vector<GrandProductChallenge>& chs,
                                                        Challenger& ch) {
    size_t degree = trace[0].size();
    unsigned degree_bits = log2_strict(degree);
    FriParams fp = fri_params(cfg.fri, degree_bits);
    if (fp.total_arities() > degree_bits + cfg.fri.rate_bits - cfg.fri.cap_height)
        throw std::runtime_error("FRI total reduction arity is too large.");
    StarkProofWithMetadata out;
    out.init_challenger_state = ch.compact();
    Trace aux = auxiliary_columns(L, trace, chs);
    if (aux.empty()) throw std::runtime_error("No CTL?");
    PolynomialBatch aux_c = PolynomialBatch::from_values(std::move(aux), cfg.fri.rate_bits, cfg.fri.cap_height);
    ch.observe_cap(aux_c.merkle_tree.cap);
    std::vector<Fp> alphas = ch.get_n_challenges(cfg.num_challenges);
    std::vector<std::vector<Fp>> qp = compute_quotient_polys(L, trace_c, aux_c, chs, alphas, degree_bits, cfg);
    std::vector<std::vector<Fp>> chunks;
    for (auto& p : qp)                                         // trim_to_len(2n) is a no-op; chunks(degree)
        for (size_t k = 0; k < p.size(); k += degree) chunks.emplace_back(p.begin() + k, p.begin() + k + degree);
    PolynomialBatch q_c = PolynomialBatch::from_coeffs(std::move(chunks), cfg.fri.rate_bits, cfg.fri.cap_height);
    ch.observe_cap(q_c.merkle_tree.cap);
    Ext2 zeta = ch.get_extension_challenge();
    Fp g = primitive_root_of_unity(degree_bits);
    if (zeta.exp_power_of_2(degree_bits) == Ext2::one()) throw std::runtime_error("Opening point is in the subgroup.");
    // StarkOpeningSet::new (proof.rs:299-334)
    StarkOpeningSet os;
    Ext2 zeta_next = zeta * g;
    auto eval_all = [&](const PolynomialBatch& b, Ext2 z) {
        std::vector<Ext2> r(b.polynomials.size());
        parallel_for(r.size(), [&](size_t i) { r[i] = poly_eval_ext(b.polynomials[i], z); }, 1);
        return r;
    };
    os.local_values = eval_all(trace_c, zeta);
    os.next_values = eval_all(trace_c, zeta_next);
    os.auxiliary_polys = eval_all(aux_c, zeta);
    os.auxiliary_polys_next = eval_all(aux_c, zeta_next);
    for (size_t i = L.num_lookup_cols + L.num_ctl_helpers; i < aux_c.polynomials.size(); i++)
        os.ctl_zs_first.push_back(poly_eval(aux_c.polynomials[i], Fp::one()));
    os.quotient_polys = eval_all(q_c, zeta);
    observe_openings(ch, os.to_fri_openings());
    FriInstanceInfo fi = fri_instance(L, zeta, g, cfg);
    out.proof.opening_proof = prove_openings(fi, {&trace_c, &aux_c, &q_c}, ch, fp);
    out.proof.trace_cap = trace_c.merkle_tree.cap;
    out.proof.auxiliary_polys_cap = aux_c.merkle_tree.cap;
    out.proof.quotient_polys_cap = q_c.merkle_tree.cap;
    out.proof.openings = std::move(os);
    return out;
}

// prove_with_traces (prover.rs:130-232)
static inline AllProof prove_with_traces(const System& sys, const StarkConfig& cfg, const std::vector<Trace>& traces, const PublicValues& pv) {
    if (traces.size() != sys.kinds.size()) throw std::runtime_error("trace count does not match the system");
    std::vector<TableLayout> layout = zkm::tables::derive_layout(sys, cfg.num_challenges);
    std::vector<PolynomialBatch> commits;
    for (size_t t = 0; t < traces.size(); t++) {
        if ((int)traces[t].size() != layout[t].ncols) throw std::runtime_error("wrong number of trace columns");
        commits.push_back(PolynomialBatch::from_values(traces[t], cfg.fri.rate_bits, cfg.fri.cap_height));
    }
    Challenger ch;
    for (auto& c : commits) ch.observe_cap(c.merkle_tree.cap);
    observe_public_values(ch, pv);
    AllProof proof;
    for (unsigned i = 0; i < cfg.num_challenges; i++) {       // get_grand_product_challenge_set (:560-576)
        GrandProductChallenge c;
        c.beta = ch.get_challenge();
        c.gamma = ch.get_challenge();
        proof.ctl_challenges.push_back(c);
    }
    for (size_t t = 0; t < traces.size(); t++)
        proof.stark_proofs.push_back(prove_single_table(layout[t], cfg, traces[t], commits[t], proof.ctl_challenges, ch));
    proof.public_values = pv;
    return proof;
}

// ------------------------------------------------------------------------------------- verifier
static inline void eval_l_0_and_l_last(unsigned log_n, Ext2 x, Ext2& l0, Ext2& ll) {      // verifier.rs:347-354
    Ext2 n = Ext2::from_base(Fp((u64)1 << log_n));
    Ext2 g = Ext2::from_base(primitive_root_of_unity(log_n));
    Ext2 zx = x.exp_power_of_2(log_n) - Ext2::one();
    l0 = zx * (n * (x - Ext2::one())).inverse();
    ll = zx * (n * (g * x - Ext2::one())).inverse();
}

struct StarkChallenges { std::vector<Fp> alphas; Ext2 zeta; FriChallenges fri; };

// num_ctl_helper_columns_by_table (cross_table_lookup.rs:601-631), independent of derive_layout.
static inline std::vector<std::vector<int>> num_ctl_helper_columns_by_table(const System& sys) {
    std::vector<std::vector<int>> res;
    for (const CrossTableLookup& ctl : sys.ctls) {
        std::vector<int> by(sys.kinds.size(), 0);
        size_t i = 0;
        while (i < ctl.looking_tables.size()) {
            int t = ctl.looking_tables[i].table, cnt = 0;
            while (i < ctl.looking_tables.size() && ctl.looking_tables[i].table == t) { cnt++; i++; }
            if (cnt > 1) by[t] = (cnt + 1) / 2;
        }
        res.push_back(by);
    }
    return res;
}

static inline void verify_stark_proof_with_challenges(int kind, const StarkProof& proof, const StarkChallenges& chal,
                                                      const std::vector<CtlVars<Ext2>>& ctl_vars, const std::vector<GrandProductChallenge>& chs,
                                                      const StarkConfig& cfg) {
    // a layout carrying only what the table-local checks need (kind, columns, lookups)
    TableLayout L;
    L.kind = kind; L.ncols = zkm::tables::table_num_columns(kind); L.lookups = zkm::tables::table_lookups(kind);
    for (const Lookup& l : L.lookups) L.num_lookup_cols += l.num_helper_columns(zkm::tables::CONSTRAINT_DEGREE) * cfg.num_challenges;
    size_t num_ctl_polys = 0;
    for (auto& v : ctl_vars) num_ctl_polys += v.helpers.size();
    size_t num_ctl_z = ctl_vars.size();
    const StarkOpeningSet& os = proof.openings;
    unsigned degree_bits = proof.recover_degree_bits(cfg);
    FriParams fp = fri_params(cfg.fri, degree_bits);
    // validate_proof_shape (verifier.rs:294-342)
    size_t num_aux = num_ctl_polys + L.num_lookup_cols + num_ctl_z;
    size_t nq = zkm::tables::QUOTIENT_DEGREE_FACTOR * cfg.num_challenges;
    size_t capn = (size_t)1 << cfg.fri.cap_height;
    ORC_ENSURE(proof.trace_cap.size() == capn && proof.auxiliary_polys_cap.size() == capn && proof.quotient_polys_cap.size() == capn, "cap height");
    ORC_ENSURE(os.local_values.size() == (size_t)L.ncols && os.next_values.size() == (size_t)L.ncols, "shape: trace openings");
    ORC_ENSURE(os.auxiliary_polys.size() == num_aux && os.auxiliary_polys_next.size() == num_aux, "shape: aux openings");
    ORC_ENSURE(os.ctl_zs_first.size() == num_ctl_z, "shape: ctl_zs_first");
    ORC_ENSURE(os.quotient_polys.size() == nq, "shape: quotient openings");

    Ext2 l0, ll;
    eval_l_0_and_l_last(degree_bits, chal.zeta, l0, ll);
    Fp last = primitive_root_of_unity(degree_bits).inverse();
    std::vector<Ext2> alphas;
    for (Fp a : chal.alphas) alphas.push_back(Ext2::from_base(a));
    Consumer<Ext2> yc(alphas, chal.zeta - Ext2::from_base(last), l0, ll);
    std::vector<Fp> lookup_ch;
    for (auto& c : chs) lookup_ch.push_back(c.beta);
    // ctl_vars already carry their openings; the lookup part reads the first num_lookup_cols aux openings
    L.zs.clear();
    RowView<Ext2> lrow{os.local_values.data()}, nrow{os.next_values.data()};
    if (!zkm::tables::eval_table<Ext2, RowView<Ext2>, Consumer<Ext2>>(kind, lrow, nrow, yc))
        throw VerifyError("constraints of this table are not available");
    if (!L.lookups.empty())
        eval_lookups(L.lookups, os.local_values.data(), os.next_values.data(), os.auxiliary_polys.data(), os.auxiliary_polys_next.data(),
                     lookup_ch, yc);
    eval_ctl_checks(os.local_values.data(), os.next_values.data(), ctl_vars, yc);
    Ext2 zeta_pow_deg = chal.zeta.exp_power_of_2(degree_bits);
    Ext2 z_h_zeta = zeta_pow_deg - Ext2::one();
    for (size_t i = 0; i < cfg.num_challenges; i++) {
        Ext2 acc;
        for (size_t k = zkm::tables::QUOTIENT_DEGREE_FACTOR; k-- > 0;)
            acc = acc * zeta_pow_deg + os.quotient_polys[i * zkm::tables::QUOTIENT_DEGREE_FACTOR + k];
        ORC_ENSURE(yc.accs[i] == z_h_zeta * acc, "Mismatch between evaluation and opening of quotient polynomial");
    }
    // FRI
    TableLayout LF = L;
    LF.num_ctl_helpers = (int)num_ctl_polys;
    LF.zs.resize(num_ctl_z);
    FriInstanceInfo fi = fri_instance(LF, chal.zeta, primitive_root_of_unity(degree_bits), cfg);
    verify_fri_proof(fi, os.to_fri_openings(), chal.fri, {proof.trace_cap, proof.auxiliary_polys_cap, proof.quotient_polys_cap},
                     proof.opening_proof, fp);
}

// verify_proof (verifier.rs:27-176) incl. AllProof::get_challenges (get_challenges.rs:124-148),
// CtlCheckVars::from_proofs (cross_table_lookup.rs:892-1002) and verify_cross_table_lookups (:1415-1452).
static inline void verify_proof(const System& sys, const AllProof& ap, const StarkConfig& cfg) {
    size_t T = sys.kinds.size();
    ORC_ENSURE(ap.stark_proofs.size() == T, "wrong number of table proofs");
    Challenger ch;
    for (auto& p : ap.stark_proofs) ch.observe_cap(p.proof.trace_cap);
    observe_public_values(ch, ap.public_values);
    std::vector<GrandProductChallenge> chs;
    for (unsigned i = 0; i < cfg.num_challenges; i++) { GrandProductChallenge c; c.beta = ch.get_challenge(); c.gamma = ch.get_challenge(); chs.push_back(c); }
    std::vector<StarkChallenges> sc(T);
    for (size_t t = 0; t < T; t++) {
        ch.compact();
        const StarkProof& p = ap.stark_proofs[t].proof;
        unsigned degree_bits = p.recover_degree_bits(cfg);
        ch.observe_cap(p.auxiliary_polys_cap);
        sc[t].alphas = ch.get_n_challenges(cfg.num_challenges);
        ch.observe_cap(p.quotient_polys_cap);
        sc[t].zeta = ch.get_extension_challenge();
        observe_openings(ch, p.openings.to_fri_openings());
        sc[t].fri = fri_challenges(ch, p.opening_proof.commit_phase_merkle_caps, p.opening_proof.final_poly, p.opening_proof.pow_witness,
                                   degree_bits, cfg.fri);
    }
    // from_proofs
    std::vector<size_t> num_lookup(T, 0);
    for (size_t t = 0; t < T; t++)
        for (const Lookup& l : zkm::tables::table_lookups(sys.kinds[t])) num_lookup[t] += l.num_helper_columns(zkm::tables::CONSTRAINT_DEGREE) * cfg.num_challenges;
    std::vector<std::vector<int>> nh = num_ctl_helper_columns_by_table(sys);
    std::vector<size_t> total_helpers(T, 0), start(T, 0), zidx(T, 0);
    for (auto& by : nh) for (size_t t = 0; t < T; t++) total_helpers[t] += (size_t)by[t] * cfg.num_challenges;
    auto aux_at = [&](size_t t, size_t i, bool next) -> Ext2 {
        const StarkOpeningSet& os = ap.stark_proofs[t].proof.openings;
        const std::vector<Ext2>& v = next ? os.auxiliary_polys_next : os.auxiliary_polys;
        ORC_ENSURE(num_lookup[t] + i < v.size(), "auxiliary openings too short");
        return v[num_lookup[t] + i];
    };
    std::vector<std::vector<CtlVars<Ext2>>> vars(T);
    for (size_t k = 0; k < sys.ctls.size(); k++) {
        const CrossTableLookup& ctl = sys.ctls[k];
        for (unsigned c = 0; c < cfg.num_challenges; c++) {
            std::vector<int> distinct;
            for (auto& lt : ctl.looking_tables) if (std::find(distinct.begin(), distinct.end(), lt.table) == distinct.end()) distinct.push_back(lt.table);
            for (int t : distinct) {
                CtlVars<Ext2> v;
                v.local_z = aux_at(t, total_helpers[t] + zidx[t], false);
                v.next_z = aux_at(t, total_helpers[t] + zidx[t], true);
                for (auto& lt : ctl.looking_tables) if (lt.table == t) { v.columns.push_back(&lt.columns); v.filter.push_back(&lt.filter); }
                for (int j = 0; j < nh[k][t]; j++) v.helpers.push_back(aux_at(t, start[t] + j, false));
                start[t] += nh[k][t];
                zidx[t]++;
                v.ch = chs[c];
                vars[t].push_back(std::move(v));
            }
            int lt = ctl.looked_table.table;
            CtlVars<Ext2> v;
            v.local_z = aux_at(lt, total_helpers[lt] + zidx[lt], false);
            v.next_z = aux_at(lt, total_helpers[lt] + zidx[lt], true);
            zidx[lt]++;
            v.columns.push_back(&ctl.looked_table.columns);
            v.filter.push_back(&ctl.looked_table.filter);
            v.ch = chs[c];
            vars[lt].push_back(std::move(v));
        }
    }
    for (size_t t = 0; t < T; t++)
        verify_stark_proof_with_challenges(sys.kinds[t], ap.stark_proofs[t].proof, sc[t], vars[t], chs, cfg);
    // verify_cross_table_lookups
    std::vector<size_t> it(T, 0);
    for (size_t k = 0; k < sys.ctls.size(); k++) {
        const CrossTableLookup& ctl = sys.ctls[k];
        std::vector<int> distinct;
        for (auto& lt : ctl.looking_tables) if (std::find(distinct.begin(), distinct.end(), lt.table) == distinct.end()) distinct.push_back(lt.table);
        for (unsigned c = 0; c < cfg.num_challenges; c++) {
            Fp sum;
            for (int t : distinct) {
                const auto& z = ap.stark_proofs[t].proof.openings.ctl_zs_first;
                ORC_ENSURE(it[t] < z.size(), "ctl_zs_first too short");
                sum += z[it[t]++];
            }
            int lt = ctl.looked_table.table;
            const auto& z = ap.stark_proofs[lt].proof.openings.ctl_zs_first;
            ORC_ENSURE(it[lt] < z.size(), "ctl_zs_first too short");
            ORC_ENSURE(sum == z[it[lt]++], "Cross-table lookup " + std::to_string(k) + " verification failed.");
        }
    }
}

}  // namespace orc
