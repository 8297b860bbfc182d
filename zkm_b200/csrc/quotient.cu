// Quotient (constraint) evaluation on the degree-2n coset: one thread per point.
// Replaces reference compute_quotient_polys (prover/src/prover.rs:645-789) + eval_vanishing_poly
// (vanishing_poly.rs:17-46): table constraints (tables/*.h templates, the analogue of each table's
// eval_packed_generic), logUp checks (lookup.rs:138-198) and CTL checks (cross_table_lookup.rs:
// 1006-1150), folded with the alphas in emission order (constraint_consumer.rs:52-75), then
// multiplied by 1/Z_H(x) (ZeroPolyOnCoset, 2 distinct values).
//
// Data layout: the trace / auxiliary LDEs are column-major and coset-major (ntt.cuh): LDE natural index
// m = 4*idx + j sits at j*n + idx.  The quotient domain is the even natural indices, i.e. cosets
// j = 0 and j = 2, and the "next" row (m + 4) is idx + 1 of the same coset: every column read is a
// fully coalesced 8-byte-per-lane access and the next-row read hits the neighbouring lane's line.
// Only half of the LDE (2 of 4 cosets) is read: 16*n*(C+A) algorithmic bytes.
#include "aux.cuh"
#include "shard.cuh"
#include "tables/registry.h"


namespace zkm {

struct LdeRow {
    const u64* base; size_t stride;
    __device__ __forceinline__ gl operator[](int c) const { return gl(__ldg(base + (size_t)c * stride)); }
};

// Number of alpha-folds compiled into the kernels: StarkConfig::standard_fast_config has num_challenges = 2
// (config.rs:17-29); every extra fold adds ~30 instructions to each of the hundreds of constraints.
constexpr int QUOTIENT_ALPHAS = 2;

// The alpha-fold acc <- acc * alpha + c runs once per constraint (hundreds of times per point).  It is a
// NON-inlined function so that its ~55 instructions exist once in the instruction cache instead of once per
// constraint (the straight-line constraint code is instruction-fetch bound).  The alphas arrive as kernel parameters
// (QParams::alphas, i.e. the kernel's own constant bank) and are handed to the fold by the consumer: no device-wide
// state, so proofs on different worker contexts never serialise on it.
struct FoldPair { u64 a0, a1; };
static __device__ __noinline__ FoldPair quotient_fold(u64 a0, u64 a1, u64 c, u64 alpha0, u64 alpha1) {
    FoldPair r;
    r.a0 = (gl(a0) * gl(alpha0) + gl(c)).v;
    r.a1 = (gl(a1) * gl(alpha1) + gl(c)).v;
    return r;
}

// Small tables (2n points fill a few CTAs at most) are latency bound: one thread walks the whole interpreted CTL/lookup
// program, two dependent loads per term, with one warp per scheduler and nothing to hide the latency (KeccakSponge at
// 2^6 rows: 8.7 ms for 128 points).  The cooperative variant gives every point W worker threads (threadIdx.y): the
// helper-column constraint VALUES, which are independent of each other, are computed by the workers in parallel into
// shared memory, and every worker then folds them in emission order, so the result is bit-identical.
constexpr int COOP_CHUNK = 64;          // helper constraints per shared-memory round
constexpr int COOP_PX = 32;             // points per CTA in cooperative mode (one warp per worker)
struct DevConsumer {
    gl alpha[QUOTIENT_ALPHAS], acc[QUOTIENT_ALPHAS];
    int na;
    gl z_last, l_first, l_last;
    int worker = 0, workers = 1, px = 0, npx = 0;      // cooperative mode: this thread's worker id, point slot
    u64* stage = nullptr;                              // shared memory [COOP_CHUNK][npx]
    __device__ __forceinline__ void constraint(gl c) {
        FoldPair r = quotient_fold(acc[0].v, acc[1].v, c.v, alpha[0].v, alpha[1].v);
        acc[0] = gl(r.a0); acc[1] = gl(r.a1);
    }
    // Re-converges the CTA: the constraint code is straight-line and instruction-fetch bound (ncu: "no instruction"
    // is the top stall), so keeping all warps of a CTA inside the same code window lets them share fetched lines.
    __device__ __forceinline__ void checkpoint() { __syncthreads(); }
    __device__ __forceinline__ void constraint_transition(gl c) { constraint(c * z_last); }
    __device__ __forceinline__ void constraint_first_row(gl c) { constraint(c * l_first); }
    __device__ __forceinline__ void constraint_last_row(gl c) { constraint(c * l_last); }
};

struct QParams {
    const u64* tr; size_t tr_cs;
    const u64* ax; size_t ax_cs;
    int log_n, na;
    u64 alphas[MAX_CHALLENGES];
    AuxChallenges ch;
    DProgramView prog;
    PowTable w2n;                 // powers of w_{2n}
    u64 zh[2], zh_inv[2];         // Z_H(x_i) for i even / odd, and inverses
    u64 last, g, n_inv;           // g^(n-1) = g^-1, g = w_n, 1/n
    u64* q;
    // in-segment sharding: the launch covers one half of the quotient domain (coset 0 or coset 2 of the LDE) and writes it
    // half-major, q[(half * na + a) * n + idx], for the exchange; otherwise both halves, natural order q[a * 2n + 2 idx + half]
    int half_base, half_major;
    size_t idx_base;              // sharded over 8 ranks: the two owners of a coset evaluate one half of its points each
};

// Column::eval (no next-row terms), used by the logUp Z check (lookup.rs:182-187).
__device__ __forceinline__ gl dcol_eval_local(const DProgramView& P, int ci, const LdeRow& lv) {
    const DColumn c = P.cols[ci];
    gl r(c.constant);
    for (int k = 0; k < c.lin_cnt; k++) { DTerm t = P.terms[c.lin_off + k]; r = r + lv[t.col] * gl(t.coef); }
    return r;
}

// eval_helper_columns (cross_table_lookup.rs:1006-1057) for parts [part_off, part_off + part_cnt).
__device__ __forceinline__ gl dev_helper_value(const DProgramView& P, int part_off, int part_cnt, int j, const LdeRow& lv, const LdeRow& nv,
                                               const LdeRow& al, int helper_aux, gl beta, gl gamma) {
    gl h = al[helper_aux + j];
    const DPart p0 = P.parts[part_off + 2 * j];
    gl c0 = dpart_combine(P, p0, lv, nv, beta, gamma);
    gl f0 = dfilter_eval(P, p0.filter, lv, nv);
    if (2 * j + 1 < part_cnt) {
        const DPart p1 = P.parts[part_off + 2 * j + 1];
        gl c1 = dpart_combine(P, p1, lv, nv, beta, gamma);
        gl f1 = dfilter_eval(P, p1.filter, lv, nv);
        return c1 * c0 * h - f0 * c1 - f1 * c0;
    }
    return c0 * h - f0;
}
template <bool COOP>
__device__ __forceinline__ void dev_eval_helper_columns(const DProgramView& P, int part_off, int part_cnt, int num_helpers, const LdeRow& lv,
                                                        const LdeRow& nv, const LdeRow& al, int helper_aux, gl beta, gl gamma, DevConsumer& yc) {
    if (num_helpers == 0) return;
    const int nh = (part_cnt + 1) / 2;
    if (!COOP) {
        for (int j = 0; j < nh; j++) yc.constraint(dev_helper_value(P, part_off, part_cnt, j, lv, nv, al, helper_aux, beta, gamma));
        return;
    }
    for (int j0 = 0; j0 < nh; j0 += COOP_CHUNK) {
        const int cnt = nh - j0 < COOP_CHUNK ? nh - j0 : COOP_CHUNK;
        for (int jj = yc.worker; jj < cnt; jj += yc.workers)
            yc.stage[jj * yc.npx + yc.px] = dev_helper_value(P, part_off, part_cnt, j0 + jj, lv, nv, al, helper_aux, beta, gamma).v;
        __syncthreads();
        for (int jj = 0; jj < cnt; jj++) yc.constraint(gl(yc.stage[jj * yc.npx + yc.px]));
        __syncthreads();
    }
}

template <bool COOP>
__device__ __forceinline__ void dev_eval_lookups(const QParams& q, const LdeRow& lv, const LdeRow& nv, const LdeRow& al, const LdeRow& an,
                                                 DevConsumer& yc) {
    const DProgramView& P = q.prog;
    for (int li = 0; li < P.num_lookups; li++) {
        const DLookup l = P.lookups[li];
        gl challenge(q.ch.beta[l.challenge]);
        dev_eval_helper_columns<COOP>(P, l.part_off, l.part_cnt, l.num_helpers, lv, nv, al, l.aux_start, gl::one(), challenge, yc);
        gl z = al[l.aux_start + l.num_helpers], next_z = an[l.aux_start + l.num_helpers];
        gl twc = dcol_eval_local(P, l.table_col, lv) + challenge;
        gl hs = gl::zero();
        for (int j = 0; j < l.num_helpers; j++) hs = hs + al[l.aux_start + j];
        gl y = hs * twc - dcol_eval_local(P, l.freq_col, lv);
        yc.constraint_first_row(z);
        yc.constraint((next_z - z) * twc - y);
    }
}

template <bool COOP>
__device__ __forceinline__ void dev_eval_ctl_checks(const QParams& q, const LdeRow& lv, const LdeRow& nv, const LdeRow& al, const LdeRow& an,
                                                    DevConsumer& yc) {
    const DProgramView& P = q.prog;
    for (int zi = 0; zi < P.num_zs; zi++) {
        const DZ z = P.zs[zi];
        gl beta(q.ch.beta[z.challenge]), gamma(q.ch.gamma[z.challenge]);
        dev_eval_helper_columns<COOP>(P, z.part_off, z.part_cnt, z.num_helpers, lv, nv, al, z.helper_aux, beta, gamma, yc);
        gl local_z = al[z.z_aux], next_z = an[z.z_aux];
        if (z.num_helpers) {
            gl hs = gl::zero();
            for (int j = 0; j < z.num_helpers; j++) hs = hs + al[z.helper_aux + j];
            yc.constraint_last_row(local_z - hs);
            yc.constraint_transition(local_z - next_z - hs);
        } else if (z.part_cnt > 1) {
            const DPart p0 = P.parts[z.part_off], p1 = P.parts[z.part_off + 1];
            gl c0 = dpart_combine(P, p0, lv, nv, beta, gamma), c1 = dpart_combine(P, p1, lv, nv, beta, gamma);
            gl f0 = dfilter_eval(P, p0.filter, lv, nv), f1 = dfilter_eval(P, p1.filter, lv, nv);
            yc.constraint_last_row(c0 * c1 * local_z - f0 * c1 - f1 * c0);
            yc.constraint_transition(c0 * c1 * (local_z - next_z) - f0 * c1 - f1 * c0);
        } else {
            const DPart p0 = P.parts[z.part_off];
            gl c0 = dpart_combine(P, p0, lv, nv, beta, gamma);
            gl f0 = dfilter_eval(P, p0.filter, lv, nv);
            yc.constraint_last_row(c0 * local_z - f0);
            yc.constraint_transition(c0 * (local_z - next_z) - f0);
        }
    }
}

__device__ __forceinline__ gl gl_inv_q(gl x) {
    gl x2 = x * x * x;
    gl x4 = gl_exp2(x2, 2) * x2;
    gl x8 = gl_exp2(x4, 4) * x4;
    gl x16 = gl_exp2(x8, 8) * x8;
    gl x32 = gl_exp2(x16, 16) * x16;
    gl x31 = gl_exp2(x16, 8) * x8;
    x31 = gl_exp2(x31, 4) * x4;
    x31 = gl_exp2(x31, 2) * x2;
    x31 = gl_exp2(x31, 1) * x;
    return gl_exp2(x31, 33) * x32;
}

// Register budget, measured on U20 (profiles/r2a_quotient_regs.txt): 64 registers (2 CTAs of 512 threads per SM, ~500 B of
// spills per thread in the CPU table's kernel) 27.4 ms per proof; 128 registers (no spills, half the resident warps) 33.9 ms:
// the kernel is latency/issue bound and wants the warps more than the registers.  ZKM_Q_THREADS / ZKM_Q_MINBLOCKS for A/B builds.
#ifndef ZKM_Q_THREADS
#define ZKM_Q_THREADS 512
#endif
#ifndef ZKM_Q_MINBLOCKS
#define ZKM_Q_MINBLOCKS 2
#endif
template <int KIND, bool COOP>
__global__ void __launch_bounds__(COOP ? 512 : ZKM_Q_THREADS, COOP ? 1 : ZKM_Q_MINBLOCKS) quotient_kernel(QParams q) {
    __shared__ u64 coop_stage[COOP ? COOP_CHUNK * COOP_PX : 1];
    const size_t n = (size_t)1 << q.log_n;
    // the grid covers the 2n points (or, sharded, the n points of one half) exactly: no early exit (checkpoints)
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x + ((size_t)q.half_base << q.log_n) + q.idx_base;
    const size_t half = t >> q.log_n, idx = t & (n - 1);
    const size_t i = 2 * idx + half;                        // index in the quotient domain 7*H_{2n}
    const size_t pos = (2 * half) * n + idx, pos_next = (2 * half) * n + ((idx + 1) & (n - 1));
    LdeRow lv{q.tr + pos, q.tr_cs}, nv{q.tr + pos_next, q.tr_cs};
    LdeRow al{q.ax + pos, q.ax_cs}, an{q.ax + pos_next, q.ax_cs};

    gl x = gl(GL_GENERATOR) * pow_lookup(q.w2n, i);
    DevConsumer yc;
    yc.na = q.na;
    if (COOP) { yc.worker = threadIdx.y; yc.workers = blockDim.y; yc.px = threadIdx.x; yc.npx = blockDim.x; yc.stage = coop_stage; }
#pragma unroll
    for (int a = 0; a < QUOTIENT_ALPHAS; a++) { yc.alpha[a] = gl(q.alphas[a]); yc.acc[a] = gl::zero(); }
    yc.z_last = x - gl(q.last);
    // L_first(x) = (x^n - 1) / (n (x - 1)),  L_last(x) = (x^n - 1) / (n (g x - 1))   (verifier.rs:347-354;
    // the reference prover tabulates the same two polynomials with an LDE of the selectors, prover.rs:677-681)
    gl d0 = x - gl::one(), d1 = gl(q.g) * x - gl::one();
    gl inv = gl_inv_q(d0 * d1);
    gl zhx = gl(q.zh[i & 1]) * gl(q.n_inv);
    yc.l_first = zhx * (inv * d1);
    yc.l_last = zhx * (inv * d0);

    tables::eval_table<gl, LdeRow, DevConsumer>(KIND, lv, nv, yc);
    if (q.prog.num_lookups) dev_eval_lookups<COOP>(q, lv, nv, al, an, yc);
    dev_eval_ctl_checks<COOP>(q, lv, nv, al, an, yc);

    if (COOP && threadIdx.y != 0) return;               // every worker holds the same accumulators
    gl zi(q.zh_inv[i & 1]);
#pragma unroll
    for (int a = 0; a < QUOTIENT_ALPHAS; a++)
        if (a < q.na) q.q[q.half_major ? (half * q.na + a) * n + idx : (size_t)a * 2 * n + i] = (yc.acc[a] * zi).v;
}

// The 12 x 2 kernel instantiations are spread over four translation units (quotient.cu = part 0 plus the launcher,
// quotient_p1..3.cu include this file with ZKM_QPART set) so that they compile in parallel.
typedef void (*quotient_kernel_t)(QParams);
#ifndef ZKM_QPART
#define ZKM_QPART 0
#endif
#define ZKM_QK(k) case tables::k: kern = coop ? quotient_kernel<tables::k, true> : quotient_kernel<tables::k, false>; break;
// Returns the kernel for `kind` if this translation unit holds it, else nullptr.
#define ZKM_QPART_FN(name, cases)                                                                          \
    quotient_kernel_t name(int kind, bool coop) {                                                          \
        quotient_kernel_t kern = nullptr;                                                                  \
        switch (kind) { cases default: break; }                                                            \
        return kern;                                                                                       \
    }
#if ZKM_QPART == 0
ZKM_QPART_FN(quotient_kernels_part0, ZKM_QK(T_CPU))
#elif ZKM_QPART == 1
ZKM_QPART_FN(quotient_kernels_part1, ZKM_QK(T_ARITHMETIC) ZKM_QK(T_KECCAK) ZKM_QK(T_LOGIC))
#elif ZKM_QPART == 2
ZKM_QPART_FN(quotient_kernels_part2, ZKM_QK(T_POSEIDON) ZKM_QK(T_POSEIDON_SPONGE) ZKM_QK(T_KECCAK_SPONGE) ZKM_QK(T_MEMORY))
#else
ZKM_QPART_FN(quotient_kernels_part3, ZKM_QK(T_SHA_EXTEND) ZKM_QK(T_SHA_EXTEND_SPONGE) ZKM_QK(T_SHA_COMPRESS) ZKM_QK(T_SHA_COMPRESS_SPONGE))
#endif
#undef ZKM_QK

#if ZKM_QPART == 0
quotient_kernel_t quotient_kernels_part1(int kind, bool coop);
quotient_kernel_t quotient_kernels_part2(int kind, bool coop);
quotient_kernel_t quotient_kernels_part3(int kind, bool coop);
static quotient_kernel_t quotient_kernel_for(int kind, bool coop) {
    quotient_kernel_t k = quotient_kernels_part0(kind, coop);
    if (!k) k = quotient_kernels_part1(kind, coop);
    if (!k) k = quotient_kernels_part2(kind, coop);
    if (!k) k = quotient_kernels_part3(kind, coop);
    if (!k) throw std::runtime_error(std::string("constraints of table ") + tables::table_name(kind) + " are not available on the device");
    return k;
}

// half-major quotient values (sharded evaluation) -> natural order: q[a * 2n + 2 idx + half] = halves[(half * na + a) * n + idx]
__global__ void quotient_interleave_kernel(const u64* __restrict__ halves, u64* __restrict__ q, int log_n, int na) {
    const size_t n = (size_t)1 << log_n;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n * na) return;
    const size_t a = t / (2 * n), i = t - a * 2 * n;
    q[t] = halves[((i & 1) * na + a) * n + (i >> 1)];
}

void compute_quotient_values(int kind, const DProgram& prog, const tables::TableLayout& L, const Batch& trace, const Batch& aux,
                             const AuxChallenges& ch, const u64* alphas, int num_alphas, u64* d_q, cudaStream_t s) {
    ZKM_CHECK(num_alphas >= 1 && num_alphas <= QUOTIENT_ALPHAS, "num_challenges above 2 is not supported by the quotient kernels");
    ZKM_CHECK(trace.rate_bits == 2 && aux.rate_bits == 2, "quotient kernel expects rate_bits = 2");
    ZKM_CHECK(trace.log_n == aux.log_n && trace.ncols == L.ncols && aux.ncols == L.num_aux(), "quotient: batch shape mismatch");
    int log_n = trace.log_n;
    size_t n = (size_t)1 << log_n;
    QParams q = {};
    q.tr = trace.lde.p; q.tr_cs = trace.lde_n();
    q.ax = aux.lde.p; q.ax_cs = aux.lde_n();
    q.log_n = log_n; q.na = num_alphas;
    for (int a = 0; a < num_alphas; a++) q.alphas[a] = alphas[a];
    q.ch = ch;
    q.prog = prog.view;
    q.w2n = ntt_root_table(ctx().ntt, log_n + 1, 0, s);      // cached per context
    gl gn = gl_exp2(gl(GL_GENERATOR), log_n);                 // 7^n
    gl z0 = gn - gl::one(), z1 = -gn - gl::one();              // x^n = 7^n * (-1)^i
    q.zh[0] = z0.v; q.zh[1] = z1.v;
    q.zh_inv[0] = gl_inv(z0).v; q.zh_inv[1] = gl_inv(z1).v;
    gl g = gl_root_of_unity(log_n);
    q.g = g.v; q.last = gl_inv(g).v; q.n_inv = gl_inv(gl((u64)n)).v;
    q.q = d_q;
    // cooperative variant while the 2n points cannot fill the machine: 32 points x 16 workers per CTA
    const bool coop = 2 * n <= 8192;
    quotient_kernel_t k = quotient_kernel_for(kind, coop);
    const Shard& sh = shard();
    if (trace.sharded) {
        // each half of the quotient domain is evaluated by the rank that owns its coset (half h = LDE coset 2h), then both
        // halves are broadcast so that every rank continues with the complete quotient values (shard.cuh)
        ZKM_CHECK(aux.sharded && !coop && sh.active(), "quotient: inconsistent sharding");
        DevBuf halves((size_t)2 * num_alphas * n, s);
        q.q = halves.p; q.half_major = 1;
        const size_t per = n / sh.parts();                 // points of one half evaluated by one rank
        for (int h = 0; h < 2; h++) {
            if (!sh.owns_coset(2 * h)) continue;
            q.half_base = h;
            q.idx_base = (size_t)sh.part() * per;
            ProfScope ps("quotient", s, 8.0 * (double)per * (L.ncols + L.num_aux()) + 8.0 * (double)per * num_alphas);
            k<<<(unsigned)(per / ZKM_Q_THREADS), ZKM_Q_THREADS, 0, s>>>(q);
            ZKM_LAUNCHED();
        }
        for (int h = 0; h < 2; h++)
            for (int p = 0; p < sh.parts(); p++)
                for (int a = 0; a < num_alphas; a++)
                    shard_broadcast(halves.p + ((size_t)h * num_alphas + a) * n + (size_t)p * per, per, Shard::rank_of(2 * h, p, sh.world), s);
        quotient_interleave_kernel<<<(unsigned)((2 * n * num_alphas + 255) / 256), 256, 0, s>>>(halves.p, d_q, log_n, num_alphas);
        ZKM_LAUNCHED();
        return;
    }
    ProfScope ps("quotient", s, 16.0 * (double)n * (L.ncols + L.num_aux()) + 16.0 * (double)n * num_alphas);
    if (coop) {
        const unsigned px = 2 * n >= (size_t)COOP_PX ? COOP_PX : (unsigned)(2 * n);
        k<<<(unsigned)(2 * n / px), dim3(px, 16), 0, s>>>(q);
    } else {
        const unsigned threads = ZKM_Q_THREADS;
        k<<<(unsigned)(2 * n / threads), threads, 0, s>>>(q);
    }
    ZKM_LAUNCHED();
}

#endif  // ZKM_QPART == 0

}  // namespace zkm
