// Host driver of the STARK prover over a System of tables: the B200 counterpart of the reference's
// prove_with_traces (prover/src/prover.rs:130-232), prove_with_commitments (:234-438) and
// prove_single_table (:441-641).  All polynomial data stays on the device; the host runs only the
// Fiat-Shamir transcript (a few hundred Poseidon permutations per table) and assembles
// the proof buffer (layout: include/zkm_b200.h "Proof buffer layout").
#include "prover.cuh"
#include "fri.cuh"
#include "shard.cuh"
#include "poseidon.cuh"
#include "poseidon_host.h"
#include "tables/systems.h"
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

namespace zkm {

// ZKM_TRACE=1: host wall-clock per prover phase on stderr (device idle gaps show up here, not in the
// per-kernel event timings).
struct PhaseTimer {
    bool on; cudaStream_t s; std::chrono::steady_clock::time_point t0; const char* table;
    PhaseTimer(cudaStream_t s_, const char* table_) : s(s_), table(table_) {
        on = std::getenv("ZKM_TRACE") != nullptr;
        if (on) { cudaStreamSynchronize(s); t0 = std::chrono::steady_clock::now(); }
    }
    void mark(const char* phase) {
        if (!on) return;
        cudaStreamSynchronize(s);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[zkm_b200] %-18s %-22s %9.3f ms\n", table, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// ---------------------------------------------------------------- Challenger
struct HostChallenger {
    u64 state[12];
    std::vector<u64> in, out;
    HostChallenger() { memset(state, 0, sizeof(state)); }
    void duplexing() {
        for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
        in.clear();
        poseidon_permute_host(state);      // fast CPU schedule (poseidon_host.h): the transcript is serial host work
        out.assign(state, state + 8);
    }
    void observe(u64 x) {
        out.clear();
        in.push_back(x);
        if (in.size() == 8) duplexing();
    }
    void observe(gl x) { observe(x.v); }
    void observe(gl2 x) { observe(x.a.v); observe(x.b.v); }
    void observe_cap(const std::vector<u64>& cap) { for (u64 w : cap) observe(w); }
    gl get_challenge() {
        if (!in.empty() || out.empty()) duplexing();
        u64 r = out.back();
        out.pop_back();
        return gl(r);
    }
    gl2 get_ext_challenge() { gl a = get_challenge(); gl b = get_challenge(); return gl2(a, b); }
    void compact(u64* st) {
        if (!in.empty()) duplexing();
        out.clear();
        memcpy(st, state, sizeof(state));
    }
};

// ------------------------------------------------------------------------------ proof writer
struct ProofWriter {
    std::vector<u64> w;
    void u(u64 x) { w.push_back(x); }
    void e(gl2 x) { w.push_back(x.a.v); w.push_back(x.b.v); }
    void words(const u64* p, size_t n) { w.insert(w.end(), p, p + n); }
    void cap(const std::vector<u64>& c) { u(c.size() / 4); words(c.data(), c.size()); }
    void exts(const std::vector<gl2>& v) { u(v.size()); for (gl2 x : v) e(x); }
};
static const u64 PROOF_MAGIC = 0x464F4F52504D4B5AULL;      // "ZKMPROOF"

// plonky2 FriConfig::fri_params with ConstantArityBits(arity_bits, final_poly_bits) (Appendix A.7)
static std::vector<int> fri_reduction_arity_bits(const StarkCfg& c, int degree_bits) {
    std::vector<int> r;
    int d = degree_bits;
    while (d > (int)c.final_poly_bits && d + (int)c.rate_bits - (int)c.arity_bits >= (int)c.cap_height) {
        r.push_back(c.arity_bits);
        d -= c.arity_bits;
    }
    return r;
}

__global__ void gather_rowmajor_kernel(const u64* __restrict__ rows, int width, const u32* __restrict__ idx, int shift, int nq,
                                       u64* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * width) return;
    int q = t / width, c = t - q * width;
    out[t] = rows[(size_t)(idx[q] >> shift) * width + c];
}
__global__ void shift_idx_kernel(const u32* in, int shift, int nq, u32* out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq) out[t] = in[t] >> shift;
}

struct FriRound {
    DevBuf rows;                 // leaves: (N_r / 16) x 32 u64
    MerkleTreeDev tree;
    int arity_bits = 0;
};

struct TableJob {
    int kind = 0, log_n = 0;
    tables::TableLayout layout;
    DevBuf values;               // trace values on H (kept until the auxiliary columns are built)
    Batch trace;
};

// the table names the reference's scope strings use (prover.rs:250-411 "prove {} STARK")
static const char* reference_scope_name(int kind) {
    static const char* N[tables::NUM_TABLE_KINDS] = {"Arithmetic", "CPU", "Poseidon", "Poseidon sponge", "Keccak", "Keccak sponge", "SHA Extend",
                                                     "SHA Extend sponge", "SHA Compress", "SHA Compress sponge", "Logic", "Memory"};
    return kind >= 0 && kind < tables::NUM_TABLE_KINDS ? N[kind] : "?";
}
// `Table` Debug names (all_stark.rs:97-110), as in "compute trace commitment for {:?}" (prover.rs:152)
static const char* reference_table_debug_name(int kind) { return tables::table_name(kind); }

// StarkOpeningSet::new evaluations, column-sharded when the table is proved by a shard group (shard.cuh): rank r evaluates
// columns [r * per, (r + 1) * per) and the (tiny) results are all-gathered, so every rank still observes all openings.
static void eval_polys_sharded(bool sharded, const u64* d_coeffs, int ncols, int log_n, const gl2* pts, int npts, u64* h_out, cudaStream_t s) {
    const Shard& sh = shard();
    if (!sharded || !sh.active() || ncols < 2 * sh.world) { eval_polys_at_points(d_coeffs, ncols, log_n, pts, npts, h_out, s); return; }
    const size_t n = (size_t)1 << log_n;
    const int per = (ncols + sh.world - 1) / sh.world;
    const int c0 = std::min(ncols, sh.rank * per), c1 = std::min(ncols, c0 + per);
    const size_t words = (size_t)per * npts * 2;
    std::vector<u64> mine(words, 0), all(words * sh.world);
    if (c1 > c0) eval_polys_at_points(d_coeffs + (size_t)c0 * n, c1 - c0, log_n, pts, npts, mine.data(), s);
    DevBuf send(words, s), recv(words * sh.world, s);
    send.upload(mine.data(), words);
    shard_all_gather(send.p, recv.p, words, s);
    recv.download(all.data(), all.size());
    memcpy(h_out, all.data(), (size_t)ncols * npts * 2 * sizeof(u64));      // rank slices are contiguous in column order
}

static void prove_single_table(TableJob& job, const StarkCfg& cfg, const AuxChallenges& ctl_ch, HostChallenger& ch, ProofWriter& W) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    const tables::TableLayout& L = job.layout;
    const int log_n = job.log_n;
    const size_t n = (size_t)1 << log_n;
    const int na = cfg.num_challenges;
    std::vector<int> arities = fri_reduction_arity_bits(cfg, log_n);
    int total_arity = 0;
    for (int a : arities) total_arity += a;
    ZKM_CHECK(total_arity <= log_n + (int)cfg.rate_bits - (int)cfg.cap_height, "FRI total reduction arity is too large.");

    u64 init_state[12];
    ch.compact(init_state);
    PhaseTimer pt(s, tables::table_name(job.kind));
    TimedScope ts_table(std::string("prove ") + reference_scope_name(job.kind) + " STARK", s);      // prover.rs:250..411

    // ---- auxiliary polynomials (prover.rs:469-522)
    DProgram prog;
    prog.build(L, na);
    prog.upload(s);
    const int naux = L.num_aux();
    ZKM_CHECK(naux > 0, "No CTL?");
    Batch aux;
    {
        DevBuf auxv((size_t)naux * n, s);
        {
            // the reference builds the CTL columns of all tables up front ("compute CTL data", prover.rs:193) and the logUp
            // columns per table (:479); here both are one pass over this table's trace
            TimedScope ts("compute CTL data + lookup helper columns", s);
            compute_aux_columns(prog, L, job.values.p, log_n, ctl_ch, auxv.p, s);
        }
        job.values.release();
        TimedScope ts("compute auxiliary polynomials commitment", s);                                 // prover.rs:513
        batch_from_values_dev(aux, std::move(auxv), naux, log_n, cfg.rate_bits, cfg.cap_height);
    }
    pt.mark("aux columns+commit");
    ch.observe_cap(aux.tree.cap);
    std::vector<u64> alphas;
    for (int a = 0; a < na; a++) alphas.push_back(ch.get_challenge().v);

    // ---- quotient polynomials (prover.rs:543-589)
    Batch quot;
    {
        DevBuf q((size_t)na * 2 * n, s);
        {
            TimedScope ts("compute quotient polys", s);                                               // prover.rs:545
            compute_quotient_values(job.kind, prog, L, job.trace, aux, ctl_ch, alphas.data(), na, q.p, s);
            coset_intt(c.ntt, q.p, 2 * n, q.p, 2 * n, na, log_n + 1, s);
        }
        // "split quotient polys" (prover.rs:562): column a, chunk k (n coefficients) is polynomial 2a + k: already contiguous
        TimedScope ts("compute quotient commitment", s);                                              // prover.rs:578
        batch_from_coeffs_dev(quot, std::move(q), 2 * na, log_n, cfg.rate_bits, cfg.cap_height);
    }
    pt.mark("quotient+commit");
    ch.observe_cap(quot.tree.cap);
    gl2 zeta = ch.get_ext_challenge();
    gl g = gl_root_of_unity(log_n);
    ZKM_CHECK(gl2_exp2(zeta, log_n) != gl2::one(), "Opening point is in the subgroup.");
    gl2 zeta_next = zeta * g;

    // ---- openings (proof.rs:299-334)
    const int C = L.ncols, Q = 2 * na;
    const int zstart = L.num_lookup_cols + L.num_ctl_helpers, nz = naux - zstart;
    std::vector<gl2> local_values(C), next_values(C), aux_local(naux), aux_next(naux), quot_open(Q);
    std::vector<gl> ctl_zs_first(nz);
    {
        TimedScope ts("compute openings", s);                          // StarkOpeningSet::new (prover.rs:601; untimed upstream)
        gl2 pts[3] = {zeta, zeta_next, gl2::one()};
        std::vector<u64> h((size_t)std::max(std::max(C, naux), Q) * 3 * 2);
        eval_polys_sharded(job.trace.sharded, job.trace.coeffs.p, C, log_n, pts, 2, h.data(), s);
        for (int i = 0; i < C; i++) {
            local_values[i] = gl2(gl(h[(i * 2 + 0) * 2]), gl(h[(i * 2 + 0) * 2 + 1]));
            next_values[i] = gl2(gl(h[(i * 2 + 1) * 2]), gl(h[(i * 2 + 1) * 2 + 1]));
        }
        eval_polys_sharded(job.trace.sharded, aux.coeffs.p, naux, log_n, pts, 3, h.data(), s);
        for (int i = 0; i < naux; i++) {
            aux_local[i] = gl2(gl(h[(i * 3 + 0) * 2]), gl(h[(i * 3 + 0) * 2 + 1]));
            aux_next[i] = gl2(gl(h[(i * 3 + 1) * 2]), gl(h[(i * 3 + 1) * 2 + 1]));
            if (i >= zstart) ctl_zs_first[i - zstart] = gl(h[(i * 3 + 2) * 2]);
        }
        eval_polys_at_points(quot.coeffs.p, Q, log_n, pts, 1, h.data(), s);
        for (int i = 0; i < Q; i++) quot_open[i] = gl2(gl(h[i * 2]), gl(h[i * 2 + 1]));
    }
    pt.mark("openings");
    // observe_openings: zeta batch, zeta_next batch, ctl_zs_first as extension elements (proof.rs:336-367)
    for (gl2 x : local_values) ch.observe(x);
    for (gl2 x : aux_local) ch.observe(x);
    for (gl2 x : quot_open) ch.observe(x);
    for (gl2 x : next_values) ch.observe(x);
    for (gl2 x : aux_next) ch.observe(x);
    for (gl x : ctl_zs_first) ch.observe(gl2(x));

    // ---- FRI (prover.rs:618-628)
    TimedScope ts_fri("compute openings proof", s);                                                   // prover.rs:620
    gl2 alpha = ch.get_ext_challenge();
    const int K0 = C + naux + Q, K1 = C + naux;
    std::vector<gl2> apow(K0);
    { gl2 cur = gl2::one(); for (int k = 0; k < K0; k++) { apow[k] = cur; cur = cur * alpha; } }
    gl2 v0 = gl2::zero(), v1 = gl2::zero(), v2 = gl2::zero();
    for (int k = 0; k < C; k++) { v0 += apow[k] * local_values[k]; v1 += apow[k] * next_values[k]; }
    for (int k = 0; k < naux; k++) { v0 += apow[C + k] * aux_local[k]; v1 += apow[C + k] * aux_next[k]; }
    for (int k = 0; k < Q; k++) v0 += apow[K1 + k] * quot_open[k];
    for (int k = 0; k < nz; k++) v2 += apow[k] * ctl_zs_first[k];
    gl2 a1 = gl2_pow(alpha, (u64)nz), a0 = gl2_pow(alpha, (u64)(K1 + nz));

    DevBuf cur(2 * n, s);                          // F coefficients, 2 columns
    {
        DevBuf R(6 * n, s), Rv(6 * n, s);
        fri_reduce_batches(job.trace, aux, quot, zstart, apow, R.p, s);
        lde_coset(c.ntt, R.p, n, Rv.p, n, 6, log_n, 0, s);
        fri_combine(Rv.p, log_n, zeta, zeta_next, v0, v1, v2, a0, a1, cur.p, s);
        coset_intt(c.ntt, cur.p, n, cur.p, n, 2, log_n, s);
    }
    pt.mark("fri reduce/combine");
    std::vector<FriRound> rounds(arities.size());
    int log_nr = log_n, shift_bits = 0;
    for (size_t r = 0; r < arities.size(); r++) {
        FriRound& fr = rounds[r];
        fr.arity_bits = arities[r];
        size_t nr = (size_t)1 << log_nr, Nr = nr << cfg.rate_bits;
        {
            DevBuf lde(2 * Nr, s);
            lde_coset(c.ntt, cur.p, nr, lde.p, Nr, 2, log_nr, cfg.rate_bits, s, shift_bits);
            fr.rows.alloc(2 * Nr, s);
            fri_leaf_rows(lde.p, Nr, log_nr, cfg.rate_bits, fr.arity_bits, fr.rows.p, s);
        }
        int log_leaves = log_nr + cfg.rate_bits - fr.arity_bits;
        merkle_alloc(fr.tree, log_leaves, cfg.cap_height, s);
        rows_leaf_hash(fr.rows.p, 2 << fr.arity_bits, (size_t)1 << log_leaves, fr.tree.digests.p, s);
        merkle_build_from_leaf_digests(fr.tree, s);
        ch.observe_cap(fr.tree.cap);
        gl2 beta = ch.get_ext_challenge();
        DevBuf next(2 * (nr >> fr.arity_bits), s);
        fri_fold(cur.p, nr, fr.arity_bits, beta, next.p, s);
        cur = std::move(next);
        log_nr -= fr.arity_bits;
        shift_bits += fr.arity_bits;
    }
    pt.mark("fri commit phase");
    size_t nfinal = (size_t)1 << log_nr;
    std::vector<u64> fin(2 * nfinal);
    cur.download(fin.data(), 2 * nfinal);
    std::vector<gl2> final_poly(nfinal);
    for (size_t i = 0; i < nfinal; i++) { final_poly[i] = gl2(gl(fin[i]), gl(fin[nfinal + i])); ch.observe(final_poly[i]); }
    // proof of work (Appendix A.9): minimum witness
    u64 pow_witness;
    {
        u64 st[12];
        memcpy(st, ch.state, sizeof(st));
        for (size_t i = 0; i < ch.in.size(); i++) st[i] = ch.in[i];
        pow_witness = fri_pow_grind(st, (int)ch.in.size(), cfg.pow_bits, s);
        ch.observe(pow_witness);
        gl resp = ch.get_challenge();
        int lz = resp.v ? __builtin_clzll(resp.v) : 64;
        ZKM_CHECK(lz >= (int)cfg.pow_bits, "proof-of-work self check failed");
    }
    pt.mark("fri pow");
    // queries (Appendix A.10)
    const int nq = cfg.num_queries;
    const size_t N = n << cfg.rate_bits;
    std::vector<u32> qidx(nq);
    for (int q = 0; q < nq; q++) qidx[q] = (u32)(ch.get_challenge().v % (u64)N);
    DevBuf didx((nq + 1) / 2 + 1, s);
    ZKM_CUDA(cudaMemcpyAsync(didx.p, qidx.data(), nq * sizeof(u32), cudaMemcpyHostToDevice, s));
    const Batch* oracles[3] = {&job.trace, &aux, &quot};
    std::vector<std::vector<u64>> o_rows(3), o_paths(3);
    const int plen = log_n + cfg.rate_bits - cfg.cap_height;
    for (int o = 0; o < 3; o++) {
        const Batch& b = *oracles[o];
        const size_t nrow = (size_t)nq * b.ncols, npath = (size_t)nq * plen * 4;
        o_rows[o].resize(nrow);
        o_paths[o].resize(npath);
        if (b.sharded) {
            // in-segment sharding: a queried leaf and its authentication path live on the rank that owns its coset; every
            // rank answers all queries from what it holds, the answers are all-gathered and the owner's copy is kept
            const Shard& sh = shard();
            const size_t per = nrow + npath;
            DevBuf pack(per, s), all(per * sh.world, s);
            lde_gather_rows(b.lde.p, b.lde_n(), b.ncols, b.log_n, b.rate_bits, (const u32*)didx.p, nq, pack.p, s);
            merkle_gather_paths(b.tree, (const u32*)didx.p, nq, pack.p + nrow, s);
            shard_all_gather(pack.p, all.p, per, s);
            std::vector<u64> h(per * sh.world);
            all.download(h.data(), h.size());
            for (int q = 0; q < nq; q++) {
                const int quarter = (int)(qidx[q] >> (b.lde_bits() - 2));
                const int part = (int)(qidx[q] >> (b.lde_bits() - 2 - sh.log_parts())) & (sh.parts() - 1);
                const u64* src = h.data() + (size_t)Shard::rank_of(bitrev2(quarter), part, sh.world) * per;
                memcpy(o_rows[o].data() + (size_t)q * b.ncols, src + (size_t)q * b.ncols, (size_t)b.ncols * sizeof(u64));
                memcpy(o_paths[o].data() + (size_t)q * plen * 4, src + nrow + (size_t)q * plen * 4, (size_t)plen * 4 * sizeof(u64));
            }
            continue;
        }
        DevBuf rows(nrow, s), paths(npath + 1, s);
        lde_gather_rows(b.lde.p, b.lde_n(), b.ncols, b.log_n, b.rate_bits, (const u32*)didx.p, nq, rows.p, s);
        merkle_gather_paths(b.tree, (const u32*)didx.p, nq, paths.p, s);
        rows.download(o_rows[o].data(), nrow);
        if (plen) paths.download(o_paths[o].data(), npath);
    }
    std::vector<std::vector<u64>> s_rows(rounds.size()), s_paths(rounds.size());
    std::vector<int> s_plen(rounds.size());
    {
        int shift = 0;
        DevBuf sidx((nq + 1) / 2 + 1, s);
        for (size_t r = 0; r < rounds.size(); r++) {
            FriRound& fr = rounds[r];
            shift += fr.arity_bits;
            int width = 2 << fr.arity_bits;
            int pl = fr.tree.log_leaves - fr.tree.cap_height;
            s_plen[r] = pl;
            DevBuf rows((size_t)nq * width, s), paths((size_t)nq * pl * 4 + 1, s);
            gather_rowmajor_kernel<<<(nq * width + 255) / 256, 256, 0, s>>>(fr.rows.p, width, (const u32*)didx.p, shift, nq, rows.p);
            ZKM_LAUNCHED();
            shift_idx_kernel<<<1, 64, 0, s>>>((const u32*)didx.p, shift, nq, (u32*)sidx.p);
            ZKM_LAUNCHED();
            merkle_gather_paths(fr.tree, (const u32*)sidx.p, nq, paths.p, s);
            s_rows[r].resize((size_t)nq * width);
            rows.download(s_rows[r].data(), s_rows[r].size());
            s_paths[r].resize((size_t)nq * pl * 4);
            if (pl) paths.download(s_paths[r].data(), s_paths[r].size());
        }
    }

    pt.mark("fri queries");
    // ---- serialise StarkProofWithMetadata
    W.words(init_state, 12);
    W.cap(job.trace.tree.cap); W.cap(aux.tree.cap); W.cap(quot.tree.cap);
    W.exts(local_values); W.exts(next_values); W.exts(aux_local); W.exts(aux_next);
    W.u(ctl_zs_first.size());
    for (gl x : ctl_zs_first) W.u(x.v);
    W.exts(quot_open);
    W.u(rounds.size());
    for (FriRound& fr : rounds) W.cap(fr.tree.cap);
    W.u(nq);
    for (int q = 0; q < nq; q++) {
        W.u(3);
        for (int o = 0; o < 3; o++) {
            int nc = oracles[o]->ncols;
            W.u(nc); W.words(o_rows[o].data() + (size_t)q * nc, nc);
            W.u(plen); W.words(o_paths[o].data() + (size_t)q * plen * 4, (size_t)plen * 4);
        }
        W.u(rounds.size());
        for (size_t r = 0; r < rounds.size(); r++) {
            int width = 2 << rounds[r].arity_bits;
            W.u(width / 2); W.words(s_rows[r].data() + (size_t)q * width, width);
            W.u(s_plen[r]); W.words(s_paths[r].data() + (size_t)q * s_plen[r] * 4, (size_t)s_plen[r] * 4);
        }
    }
    W.exts(final_poly);
    W.u(pow_witness);
    pt.mark("serialise");
}

// Stage-by-stage outputs of prove_single_table for ONE table of a System under caller-given challenges (no transcript): what
// SURVEY section 8(b) calls the finer-grained seam (cross_table_lookup_data + lookup_helper_columns, compute_quotient_polys +
// post-processing, StarkOpeningSet::new), for stage-level parity tests against the oracle.
void stage_single_table(int system_id, int table_index, const StarkCfg& cfg, DevBuf&& values, int ncols, int log_n, const AuxChallenges& ctl_ch,
                        const u64* alphas, gl2 zeta, std::vector<u64>& aux_out, std::vector<u64>& quot_out, std::vector<u64>& open_out) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    tables::System sys = tables::make_system(system_id);
    ZKM_CHECK(table_index >= 0 && (size_t)table_index < sys.kinds.size(), "no such table in this system");
    ZKM_CHECK(cfg.num_challenges >= 1 && cfg.num_challenges <= MAX_CHALLENGES && ctl_ch.count == (int)cfg.num_challenges, "unsupported num_challenges");
    ZKM_CHECK(cfg.rate_bits == 2, "only rate_bits = 2 is supported (quotient kernel layout)");
    const tables::TableLayout L = tables::derive_layout(sys, cfg.num_challenges)[table_index];
    const int kind = sys.kinds[table_index], na = cfg.num_challenges;
    ZKM_CHECK(ncols == L.ncols, "wrong number of trace columns");
    ZKM_CHECK(log_n + (int)cfg.rate_bits >= (int)cfg.cap_height && log_n >= 1, "trace too short");
    const size_t n = (size_t)1 << log_n;
    Batch trace;
    {
        DevBuf coeffs((size_t)ncols * n, s);
        ntt_inverse(c.ntt, values.p, n, coeffs.p, n, ncols, log_n, s);
        batch_from_coeffs_dev(trace, std::move(coeffs), ncols, log_n, cfg.rate_bits, cfg.cap_height);
    }
    DProgram prog;
    prog.build(L, na);
    prog.upload(s);
    const int naux = L.num_aux();
    ZKM_CHECK(naux > 0, "No CTL?");
    Batch aux;
    {
        DevBuf auxv((size_t)naux * n, s);
        compute_aux_columns(prog, L, values.p, log_n, ctl_ch, auxv.p, s);
        aux_out.resize((size_t)naux * n);
        auxv.download(aux_out.data(), aux_out.size());
        batch_from_values_dev(aux, std::move(auxv), naux, log_n, cfg.rate_bits, cfg.cap_height);
    }
    Batch quot;
    {
        DevBuf q((size_t)na * 2 * n, s);
        compute_quotient_values(kind, prog, L, trace, aux, ctl_ch, alphas, na, q.p, s);
        coset_intt(c.ntt, q.p, 2 * n, q.p, 2 * n, na, log_n + 1, s);
        quot_out.resize((size_t)na * 2 * n);
        q.download(quot_out.data(), quot_out.size());
        batch_from_coeffs_dev(quot, std::move(q), 2 * na, log_n, cfg.rate_bits, cfg.cap_height);
    }
    ZKM_CHECK(gl2_exp2(zeta, log_n) != gl2::one(), "Opening point is in the subgroup.");
    const gl2 zeta_next = zeta * gl_root_of_unity(log_n);
    const int C = L.ncols, Q = 2 * na, zstart = L.num_lookup_cols + L.num_ctl_helpers;
    gl2 pts[3] = {zeta, zeta_next, gl2::one()};
    std::vector<u64> h((size_t)std::max(std::max(C, naux), Q) * 3 * 2);
    // layout of open_out: local[C], next[C], aux[naux], aux_next[naux] (2 words each), ctl_zs_first (1 word each), quot[Q] (2 words)
    open_out.clear();
    eval_polys_at_points(trace.coeffs.p, C, log_n, pts, 2, h.data(), s);
    for (int p = 0; p < 2; p++) for (int i = 0; i < C; i++) { open_out.push_back(h[(i * 2 + p) * 2]); open_out.push_back(h[(i * 2 + p) * 2 + 1]); }
    eval_polys_at_points(aux.coeffs.p, naux, log_n, pts, 3, h.data(), s);
    for (int p = 0; p < 2; p++) for (int i = 0; i < naux; i++) { open_out.push_back(h[(i * 3 + p) * 2]); open_out.push_back(h[(i * 3 + p) * 2 + 1]); }
    for (int i = zstart; i < naux; i++) open_out.push_back(h[(i * 3 + 2) * 2]);
    eval_polys_at_points(quot.coeffs.p, Q, log_n, pts, 1, h.data(), s);
    for (int i = 0; i < Q; i++) { open_out.push_back(h[i * 2]); open_out.push_back(h[i * 2 + 1]); }
    ZKM_CUDA(stream_sync(s));
}

std::vector<u64> prove_system(int system_id, const StarkCfg& cfg, std::vector<TableInput>& inputs, const PublicInputs& pv) {
    Ctx& c = ctx();
    cudaStream_t s = c.stream;
    tables::System sys = tables::make_system(system_id);
    ZKM_CHECK(inputs.size() == sys.kinds.size(), "wrong number of tables for this system");
    ZKM_CHECK(cfg.num_challenges >= 1 && cfg.num_challenges <= MAX_CHALLENGES, "unsupported num_challenges");
    ZKM_CHECK(cfg.rate_bits == 2, "only rate_bits = 2 is supported (quotient kernel layout)");
    ZKM_CHECK(cfg.arity_bits >= 1 && cfg.arity_bits <= 4, "unsupported FRI arity");
    std::vector<tables::TableLayout> layout = tables::derive_layout(sys, cfg.num_challenges);
    std::vector<TableJob> jobs(inputs.size());
    PhaseTimer pt0(s, "all");
    scopes_begin();
    struct ScopesEnd { ~ScopesEnd() { scopes_finish(); } } scopes_end;       // also on the error paths
    std::unique_ptr<TimedScope> ts_commit(new TimedScope("compute all trace commitments", s));        // prover.rs:146
    // trace commitments (prover.rs:144-167)
    // The 12 commitments are independent (their caps enter the transcript afterwards, in table order), so they run smallest
    // table first: with host inputs the uploader streams the tables in this same order (capi.cu) and the large tables arrive
    // while the small ones are being committed.
    std::vector<size_t> in_bytes(inputs.size());
    for (size_t t = 0; t < inputs.size(); t++) in_bytes[t] = (size_t)inputs[t].ncols << inputs[t].log_n;
    for (size_t t : commit_order(in_bytes)) {
        TableJob& j = jobs[t];
        j.kind = sys.kinds[t];
        j.layout = layout[t];
        j.log_n = inputs[t].log_n;
        ZKM_CHECK(tables::table_implemented(j.kind), std::string("constraints of table ") + tables::table_name(j.kind) + " are not available");
        ZKM_CHECK(inputs[t].ncols == j.layout.ncols, std::string("wrong number of trace columns for table ") + tables::table_name(j.kind));
        ZKM_CHECK(j.log_n + (int)cfg.rate_bits >= (int)cfg.cap_height && j.log_n >= 1, "trace too short");
        size_t n = (size_t)1 << j.log_n;
        j.values = std::move(inputs[t].values);
        TimedScope ts(std::string("compute trace commitment for ") + reference_table_debug_name(j.kind), s);   // prover.rs:152
        DevBuf coeffs((size_t)j.layout.ncols * n, s);
        if (!inputs[t].group_ends.empty()) {
            TableInput& in = inputs[t];
            batch_from_values_grouped_dev(j.trace, j.values.p, std::move(coeffs), j.layout.ncols, j.log_n, cfg.rate_bits, cfg.cap_height,
                                          in.group_ends, [&in, s](size_t k) {
                                              in.wait_group(k);
                                              ZKM_CUDA(cudaStreamWaitEvent(s, in.group_ready[k], 0));
                                          });
            continue;
        }
        if (inputs[t].wait_recorded) inputs[t].wait_recorded();
        if (inputs[t].ready) ZKM_CUDA(cudaStreamWaitEvent(s, inputs[t].ready, 0));
        ntt_inverse(c.ntt, j.values.p, n, coeffs.p, n, j.layout.ncols, j.log_n, s);
        batch_from_coeffs_dev(j.trace, std::move(coeffs), j.layout.ncols, j.log_n, cfg.rate_bits, cfg.cap_height);
    }
    pt0.mark("trace commitments");
    ts_commit.reset();
    HostChallenger ch;
    for (TableJob& j : jobs) ch.observe_cap(j.trace.tree.cap);
    for (int i = 0; i < 8; i++) ch.observe((u64)pv.roots_before[i]);
    for (int i = 0; i < 8; i++) ch.observe((u64)pv.roots_after[i]);
    for (uint8_t b : pv.userdata) ch.observe((u64)b);
    AuxChallenges ctl = {};
    ctl.count = cfg.num_challenges;
    for (unsigned k = 0; k < cfg.num_challenges; k++) { ctl.beta[k] = ch.get_challenge().v; ctl.gamma[k] = ch.get_challenge().v; }

    ProofWriter W;
    W.u(PROOF_MAGIC); W.u(1); W.u(jobs.size());
    W.u(cfg.num_challenges);
    for (unsigned k = 0; k < cfg.num_challenges; k++) { W.u(ctl.beta[k]); W.u(ctl.gamma[k]); }
    for (int i = 0; i < 8; i++) W.u(pv.roots_before[i]);
    for (int i = 0; i < 8; i++) W.u(pv.roots_after[i]);
    W.u(pv.userdata.size());
    for (uint8_t b : pv.userdata) W.u(b);
    {
        TimedScope ts("compute all proofs given commitments", s);                                    // prover.rs:204
        for (TableJob& j : jobs) {
            prove_single_table(j, cfg, ctl, ch, W);
            j.trace = Batch();                                  // release this table's device memory
        }
    }
    ZKM_CUDA(stream_sync(s));
    return std::move(W.w);
}

}  // namespace zkm
