// ORACLE (test infrastructure): C entry points over the CPU restatement, loaded with ctypes by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs ONLY.
// The product library (zkm_b200/libzkm_b200.so) never links or loads this.
#include "field.h"
#include "poseidon.h"
#include "fft.h"
#include "plonky2_restated.h"
#include "stark.h"
#include "proof_io.h"
#include "../zkm_b200/csrc/tables/systems.h"
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
    if (fast == 2) poseidon_opt(s); else if (fast) poseidon_fast(s); else poseidon_naive(s);
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


// ---- STARK layer: prove / verify a System (oracle/stark.h) ----
static StarkConfig make_cfg(const uint32_t* c) {
    StarkConfig cfg;
    if (c) {
        cfg.fri.rate_bits = c[0]; cfg.fri.cap_height = c[1]; cfg.fri.proof_of_work_bits = c[2]; cfg.fri.num_query_rounds = c[3];
        cfg.num_challenges = c[4]; cfg.fri.arity_bits = c[5]; cfg.fri.final_poly_bits = c[6];
    }
    return cfg;
}
// tables[t] = column pointers of table t (ncols[t] columns of 2^log_n[t] values).  Returns a malloc'ed
// proof buffer (u64 words) or NULL (see orc_last_error).
uint64_t* orc_prove_system(int system_id, const uint64_t* const* const* tables, const uint32_t* ncols, const uint32_t* log_n,
                           const uint32_t* roots_before, const uint32_t* roots_after, const uint8_t* userdata, uint32_t userdata_len,
                           const uint32_t* cfg_words, size_t* out_words) {
    try {
        System sys = zkm::tables::make_system(system_id);
        StarkConfig cfg = make_cfg(cfg_words);
        std::vector<Trace> traces(sys.kinds.size());
        for (size_t t = 0; t < sys.kinds.size(); t++) {
            size_t n = (size_t)1 << log_n[t];
            traces[t].assign(ncols[t], std::vector<Fp>(n));
            for (uint32_t c = 0; c < ncols[t]; c++)
                for (size_t i = 0; i < n; i++) traces[t][c][i] = Fp(tables[t][c][i]);
        }
        PublicValues pv;
        for (int i = 0; i < 8; i++) { pv.roots_before[i] = roots_before[i]; pv.roots_after[i] = roots_after[i]; }
        pv.userdata.assign(userdata, userdata + userdata_len);
        AllProof ap = prove_with_traces(sys, cfg, traces, pv);
        std::vector<u64> w = serialize(ap);
        uint64_t* out = (uint64_t*)malloc(w.size() * sizeof(u64));
        memcpy(out, w.data(), w.size() * sizeof(u64));
        *out_words = w.size();
        return out;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
// Stage outputs of prove_single_table for one table under given challenges (the product's zkm_b200_stage_table): auxiliary
// columns (auxiliary_columns = lookup_helper_columns + cross_table_lookup_data), quotient coefficients (compute_quotient_polys)
// and StarkOpeningSet::new.  aux_out: num_aux x n, quot_out: num_challenges x 2n, open_out: see include/zkm_b200.h.
// Returns the number of opening words written, or -1.
long orc_stage_table(int system_id, uint32_t table_index, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n, const uint32_t* cfg_words,
                     const uint64_t* ctl_challenges, const uint64_t* alphas_in, const uint64_t* zeta_in, uint64_t* aux_out, uint32_t* num_aux_out,
                     uint64_t* quot_out, uint64_t* open_out) {
    try {
        System sys = zkm::tables::make_system(system_id);
        StarkConfig cfg = make_cfg(cfg_words);
        const TableLayout L = zkm::tables::derive_layout(sys, cfg.num_challenges).at(table_index);
        if ((int)ncols != L.ncols) throw std::runtime_error("wrong number of trace columns");
        const size_t n = (size_t)1 << log_n;
        Trace trace(ncols, std::vector<Fp>(n));
        for (uint32_t c = 0; c < ncols; c++) for (size_t i = 0; i < n; i++) trace[c][i] = Fp(cols[c][i]);
        std::vector<GrandProductChallenge> chs(cfg.num_challenges);
        std::vector<Fp> alphas(cfg.num_challenges);
        for (unsigned k = 0; k < cfg.num_challenges; k++) { chs[k].beta = Fp(ctl_challenges[2 * k]); chs[k].gamma = Fp(ctl_challenges[2 * k + 1]); alphas[k] = Fp(alphas_in[k]); }
        Trace aux = auxiliary_columns(L, trace, chs);
        if (aux.empty()) throw std::runtime_error("No CTL?");
        *num_aux_out = (uint32_t)aux.size();
        for (size_t c = 0; c < aux.size(); c++) for (size_t i = 0; i < n; i++) aux_out[c * n + i] = aux[c][i].v;
        PolynomialBatch trace_c = PolynomialBatch::from_values(trace, cfg.fri.rate_bits, cfg.fri.cap_height);
        PolynomialBatch aux_c = PolynomialBatch::from_values(std::move(aux), cfg.fri.rate_bits, cfg.fri.cap_height);
        std::vector<std::vector<Fp>> qp = compute_quotient_polys(L, trace_c, aux_c, chs, alphas, log_n, cfg);
        std::vector<std::vector<Fp>> chunks;
        for (size_t a = 0; a < qp.size(); a++) {
            for (size_t i = 0; i < 2 * n; i++) quot_out[a * 2 * n + i] = qp[a][i].v;
            for (size_t k = 0; k < qp[a].size(); k += n) chunks.emplace_back(qp[a].begin() + k, qp[a].begin() + k + n);
        }
        const Ext2 zeta = Ext2(Fp(zeta_in[0]), Fp(zeta_in[1]));
        const Ext2 zeta_next = zeta * primitive_root_of_unity(log_n);
        size_t w = 0;
        auto put_all = [&](const std::vector<std::vector<Fp>>& polys, Ext2 z) {
            for (auto& p : polys) { Ext2 r = poly_eval_ext(p, z); open_out[w++] = r.a.v; open_out[w++] = r.b.v; }
        };
        put_all(trace_c.polynomials, zeta); put_all(trace_c.polynomials, zeta_next);
        put_all(aux_c.polynomials, zeta); put_all(aux_c.polynomials, zeta_next);
        for (size_t i = L.num_lookup_cols + L.num_ctl_helpers; i < aux_c.polynomials.size(); i++) open_out[w++] = poly_eval(aux_c.polynomials[i], Fp::one()).v;
        put_all(chunks, zeta);
        return (long)w;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
void orc_free(void* p) { free(p); }
// 0 = accepted, -1 = rejected / malformed (orc_last_error says why).
int orc_verify_system(int system_id, const uint64_t* proof, size_t words, const uint32_t* cfg_words) {
    try {
        System sys = zkm::tables::make_system(system_id);
        StarkConfig cfg = make_cfg(cfg_words);
        AllProof ap = deserialize(proof, words);
        verify_proof(sys, ap, cfg);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// Constraint check on H (the reference's check_constraints, prover.rs:793-910, and the
// generate => constraints-vanish tests): evaluates the table constraints of `kind` on every row of
// the trace (next row cyclic) with alpha = 1-free accumulation per constraint; returns the number of
// rows with a non-zero constraint, or -1 on error.  Transition constraints are skipped on the last row
// and first/last-row constraints only apply there, as in the reference.
long orc_check_table_constraints(int kind, const uint64_t* const* cols, uint32_t ncols, uint32_t log_n) {
    try {
        size_t n = (size_t)1 << log_n;
        if ((int)ncols != zkm::tables::table_num_columns(kind)) throw std::runtime_error("wrong column count");
        std::atomic<long> bad(0);
        Fp last = primitive_root_of_unity(log_n).inverse();
        parallel_for(n, [&](size_t i) {
            std::vector<Fp> lv(ncols), nv(ncols);
            for (uint32_t c = 0; c < ncols; c++) { lv[c] = Fp(cols[c][i]); nv[c] = Fp(cols[c][(i + 1) % n]); }
            Fp x = primitive_root_of_unity(log_n).pow(i);
            // two random-ish alphas so that a cancelling combination is implausible
            Consumer<Fp> yc({Fp(0x9E3779B97F4A7C15ULL % GL_P), Fp(0xC2B2AE3D27D4EB4FULL % GL_P)}, x - last,
                            i == 0 ? Fp::one() : Fp::zero(), i == n - 1 ? Fp::one() : Fp::zero());
            RowView<Fp> l{lv.data()}, nx{nv.data()};
            if (!zkm::tables::eval_table<Fp, RowView<Fp>, Consumer<Fp>>(kind, l, nx, yc)) { bad = -(long)n - 1; return; }
            if (!yc.accs[0].is_zero() || !yc.accs[1].is_zero()) bad++;
        });
        if (bad < 0) throw std::runtime_error("constraints of this table are not available");
        return bad;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// Poseidon table row generator (reference poseidon_stark.rs:51-95,105-145: poseidon_with_witness +
// generate_trace_rows_for_perm): inputs[k*12..], timestamps[k] -> row-major rows[k*262..].  Test input
// generator only.
void orc_gen_poseidon_rows(const uint64_t* inputs, const uint64_t* timestamps, size_t count, uint64_t* rows) {
    namespace pz = zkm::tables::poseidon;
    parallel_for(count, [&](size_t k) {
        uint64_t* row = rows + k * pz::NUM_COLUMNS;
        for (int i = 0; i < pz::NUM_COLUMNS; i++) row[i] = 0;
        row[pz::FILTER] = 1;
        row[pz::TIMESTAMP] = timestamps[k];
        PState s;
        for (int i = 0; i < 12; i++) { s[i] = Fp(inputs[k * 12 + i]); row[pz::reg_in(i)] = s[i].v; }
        int round = 0;
        auto full = [&](int r, bool second) {
            for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_ALL_ROUND_CONSTANTS[12 * round + i]);
            for (int i = 0; i < 12; i++) {
                Fp x3 = s[i] * s[i] * s[i], x7 = x3 * x3 * s[i];
                int base = second ? pz::reg_full1_s0(r, i) : pz::reg_full0_s0(r, i);
                row[base] = x3.v; row[base + 1] = x7.v;
                s[i] = x7;
            }
            mds_layer(s);
            round++;
        };
        for (int r = 0; r < 4; r++) full(r, false);
        for (int i = 0; i < 12; i++) s[i] += Fp(POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]);
        {
            PState o; o[0] = s[0];
            for (int c = 0; c < 11; c++) { Fp acc; for (int r = 0; r < 11; r++) acc += s[r + 1] * Fp(POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX[r * 11 + c]); o[c + 1] = acc; }
            s = o;
        }
        for (int r = 0; r < 22; r++) {
            Fp x3 = s[0] * s[0] * s[0], x7 = x3 * x3 * s[0];
            row[pz::reg_partial_s0(r)] = x3.v; row[pz::reg_partial_s0(r) + 1] = x7.v;
            s[0] = x7;
            if (r < 21) s[0] += Fp(POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS[r]);
            Fp d = s[0] * Fp(POSEIDON_MDS_CIRC[0] + POSEIDON_MDS_DIAG[0]);
            for (int j = 1; j < 12; j++) d += s[j] * Fp(POSEIDON_FAST_PARTIAL_ROUND_W_HATS[r * 11 + j - 1]);
            PState o; o[0] = d;
            for (int j = 1; j < 12; j++) o[j] = s[j] + s[0] * Fp(POSEIDON_FAST_PARTIAL_ROUND_VS[r * 11 + j - 1]);
            s = o;
        }
        round += 22;
        for (int r = 0; r < 4; r++) full(r, true);
        for (int i = 0; i < 12; i++) row[pz::reg_out(i)] = s[i].v;
    });
}


// ---- transcription fingerprints (tests/golden/constraint_fingerprints_v1.json; INTEGRATION.md section 3 holds the Rust test
// that prints the same numbers from eval_packed_generic / all_cross_table_lookups on the reference side).
// Rows: cell c of the local row = fp_cell(seed + 2c), of the next row = fp_cell(seed + 2c + 1), SplitMix64 reduced mod p.
static inline uint64_t fp_cell(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return z % GL_P;
}
static const uint64_t FP_ALPHA0 = 0x9E3779B97F4A7C15ULL % GL_P, FP_ALPHA1 = 0xC2B2AE3D27D4EB4FULL % GL_P;
struct CountingConsumer : Consumer<Fp> {
    uint64_t count = 0;
    CountingConsumer() : Consumer<Fp>({Fp(FP_ALPHA0), Fp(FP_ALPHA1)}, Fp(3), Fp(5), Fp(7)) {}
    void constraint(Fp c) { count++; Consumer<Fp>::constraint(c); }
    void constraint_transition(Fp c) { constraint(c * z_last); }
    void constraint_first_row(Fp c) { constraint(c * l_first); }
    void constraint_last_row(Fp c) { constraint(c * l_last); }
};
static void fp_rows(uint64_t seed, int ncols, std::vector<Fp>& lv, std::vector<Fp>& nv) {
    lv.resize(ncols); nv.resize(ncols);
    for (int c = 0; c < ncols; c++) { lv[c] = Fp(fp_cell(seed + 2 * (uint64_t)c)); nv[c] = Fp(fp_cell(seed + 2 * (uint64_t)c + 1)); }
}
// All constraints of table `kind` (the reference's eval_packed_generic, emission order) on one pseudo-random frame, folded
// acc <- acc * alpha + c with alphas (FP_ALPHA0, FP_ALPHA1), z_last = 3, lagrange_first = 5, lagrange_last = 7
// (constraint_consumer.rs:52-75).  out = {number of constraints, acc0, acc1}.  An omitted, added or reordered constraint, a
// wrong column or a wrong coefficient changes acc.
int orc_table_fingerprint(int kind, uint64_t seed, uint64_t out[3]) {
    try {
        std::vector<Fp> lv, nv;
        fp_rows(seed, zkm::tables::table_num_columns(kind), lv, nv);
        CountingConsumer yc;
        RowView<Fp> l{lv.data()}, nx{nv.data()};
        if (!zkm::tables::eval_table<Fp, RowView<Fp>, CountingConsumer>(kind, l, nx, yc)) throw std::runtime_error("no constraints for this table");
        out[0] = yc.count; out[1] = yc.accs[0].v; out[2] = yc.accs[1].v;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// The individual constraint values behind orc_table_fingerprint, in emission order (after the z_last / lagrange factors):
// writes up to max_out values, returns their total number (or -1).  tests/test_independent_transcription.py compares them one
// by one with a second, Python transcription of the reference.
struct ListingConsumer : CountingConsumer {
    std::vector<uint64_t> values;
    void constraint(Fp c) { values.push_back(c.v); CountingConsumer::constraint(c); }
    void constraint_transition(Fp c) { constraint(c * z_last); }
    void constraint_first_row(Fp c) { constraint(c * l_first); }
    void constraint_last_row(Fp c) { constraint(c * l_last); }
};
long orc_table_constraint_values(int kind, uint64_t seed, uint64_t* out, size_t max_out) {
    try {
        std::vector<Fp> lv, nv;
        fp_rows(seed, zkm::tables::table_num_columns(kind), lv, nv);
        ListingConsumer yc;
        RowView<Fp> l{lv.data()}, nx{nv.data()};
        if (!zkm::tables::eval_table<Fp, RowView<Fp>, ListingConsumer>(kind, l, nx, yc)) throw std::runtime_error("no constraints for this table");
        for (size_t i = 0; i < yc.values.size() && i < max_out; i++) out[i] = yc.values[i];
        return (long)yc.values.size();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// fold of (filter value, column values...) of one TableWithColumns on the frame above: acc <- acc * FP_ALPHA0 + v
static Fp fp_table_with_columns(const zkm::tables::TableWithColumns& t, const Fp* lv, const Fp* nv) {
    Fp acc = filter_eval<Fp>(t.filter, lv, nv);
    for (const auto& c : t.columns) acc = acc * Fp(FP_ALPHA0) + col_eval_with_next<Fp>(c, lv, nv);
    return acc;
}
// CTL number `ctl` of System `system_id` (all_stark.rs:136-542 for system 0).  entry < num_looking: that looking table, entry ==
// num_looking: the looked table.  out = {table index in the System, number of columns, fingerprint}; returns the number of
// looking tables, or -1.  The frame is seeded with seed + 0x10000 * table.
int orc_ctl_fingerprint(int system_id, int ctl, int entry, uint64_t seed, uint64_t out[3]) {
    try {
        zkm::tables::System sys = zkm::tables::make_system(system_id);
        if (ctl < 0 || ctl >= (int)sys.ctls.size()) throw std::runtime_error("no such CTL");
        const auto& c = sys.ctls[ctl];
        const int nl = (int)c.looking_tables.size();
        if (entry < 0 || entry > nl) throw std::runtime_error("no such CTL entry");
        const auto& t = entry < nl ? c.looking_tables[entry] : c.looked_table;
        std::vector<Fp> lv, nv;
        fp_rows(seed + 0x10000ULL * (uint64_t)t.table, zkm::tables::table_num_columns(sys.kinds[t.table]), lv, nv);
        out[0] = (uint64_t)t.table; out[1] = t.columns.size(); out[2] = fp_table_with_columns(t, lv.data(), nv.data()).v;
        return nl;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
int orc_num_ctls(int system_id) {
    try { return (int)zkm::tables::make_system(system_id).ctls.size(); } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// In-table logUp lookup number `idx` of table `kind` (lookup.rs:20-31; arithmetic_stark.rs:269-276, memory_stark.rs:476-483):
// out = {number of looked-up columns, fingerprint over (columns..., table_column, frequencies_column, filters...)}; returns the
// number of lookups of the table, or -1.
int orc_lookup_fingerprint(int kind, int idx, uint64_t seed, uint64_t out[2]) {
    try {
        auto ls = zkm::tables::table_lookups(kind);
        if (idx < 0 || idx >= (int)ls.size()) { out[0] = out[1] = 0; return (int)ls.size(); }
        std::vector<Fp> lv, nv;
        fp_rows(seed, zkm::tables::table_num_columns(kind), lv, nv);
        const auto& l = ls[idx];
        Fp acc = Fp(0);
        auto push = [&](Fp v) { acc = acc * Fp(FP_ALPHA0) + v; };
        for (const auto& c : l.columns) push(col_eval_with_next<Fp>(c, lv.data(), nv.data()));
        push(col_eval_with_next<Fp>(l.table_column, lv.data(), nv.data()));
        push(col_eval_with_next<Fp>(l.frequencies_column, lv.data(), nv.data()));
        for (const auto& f : l.filter_columns) push(filter_eval<Fp>(f, lv.data(), nv.data()));
        out[0] = l.columns.size(); out[1] = acc.v;
        return (int)ls.size();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

}  // extern "C"
