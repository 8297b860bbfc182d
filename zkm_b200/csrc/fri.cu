// FRI opening proof kernels.  Replaces plonky2's PolynomialBatch::prove_openings + fri_proof as called
// from the reference (prover/src/prover.rs:618-628; semantics restated in SURVEY Appendix A.8-A.10):
//   1. fri_reduce      R_b(X) = sum_k alpha^k p_k(X) for the three opening batches {zeta, g*zeta, 1},
//                      in coefficient form, one pass over all committed coefficients (8n(C+A+4) bytes);
//                      R_1 is a prefix of R_0's sum (both start with trace then aux polynomials).
//   2. coset NTT of the 3 extension polynomials (6 base columns) onto 7*H_n, then fri_combine:
//                      F(x) = a0 (R_0(x)-v_0)/(x-zeta) + a1 (R_1(x)-v_1)/(x-g zeta) + (R_2(x)-v_2)/(x-1)
//                      (equals the reference's coefficient-space divide_by_linear + shift_poly chain because
//                      deg F < n and the n coset points determine it), one batched inversion per point;
//                      coset iNTT gives F's coefficients.
//   3. per arity-16 round: LDE x4 on the shifted coset, 16-value leaves (bit-reversed order), Poseidon
//                      Merkle tree, fold of the coefficients by beta.
//   4. proof-of-work grind (minimum witness), query row/path gathers.
#include "fri.cuh"
#include "poseidon_v2.cuh"

namespace zkm {

__device__ __forceinline__ gl gl_inv_f(gl x) {
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

// ---- 1. batch reduction in coefficient space -------------------------------------------------
struct ReduceParams {
    const u64* tr; int ntr;
    const u64* ax; int nax; int zstart;       // CTL Z polynomials are aux columns [zstart, nax)
    const u64* qt; int nqt;
    const u64* apow;                          // alpha^k as (a, b) pairs, k < ntr + nax + nqt
    size_t n;
    u64* out;                                 // 6 columns of n: R0.a R0.b R1.a R1.b R2.a R2.b
};

__global__ void __launch_bounds__(256) fri_reduce_kernel(ReduceParams p) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(p.apow);
    gl2 acc = gl2::zero(), accz = gl2::zero();
    int k = 0;
    for (int c = 0; c < p.ntr; c++, k++) {
        ulonglong2 a = __ldg(ap + k);
        acc = acc + gl2(gl(a.x), gl(a.y)) * gl(__ldg(p.tr + (size_t)c * p.n + i));
    }
    for (int c = 0; c < p.nax; c++, k++) {
        ulonglong2 a = __ldg(ap + k);
        gl v(__ldg(p.ax + (size_t)c * p.n + i));
        acc = acc + gl2(gl(a.x), gl(a.y)) * v;
        if (c >= p.zstart) {
            ulonglong2 b = __ldg(ap + (c - p.zstart));
            accz = accz + gl2(gl(b.x), gl(b.y)) * v;
        }
    }
    gl2 r1 = acc;
    for (int c = 0; c < p.nqt; c++, k++) {
        ulonglong2 a = __ldg(ap + k);
        acc = acc + gl2(gl(a.x), gl(a.y)) * gl(__ldg(p.qt + (size_t)c * p.n + i));
    }
    p.out[i] = acc.a.v; p.out[p.n + i] = acc.b.v;
    p.out[2 * p.n + i] = r1.a.v; p.out[3 * p.n + i] = r1.b.v;
    p.out[4 * p.n + i] = accz.a.v; p.out[5 * p.n + i] = accz.b.v;
}

// Short polynomials (n < 2^12, the 2^6-row tables with up to 2431 columns): one CTA per coefficient index,
// threads stride over the columns, tree reduction in shared memory.
__global__ void __launch_bounds__(256) fri_reduce_small_kernel(ReduceParams p) {
    __shared__ u64 sh[3][256][2];
    const size_t i = blockIdx.x;
    const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(p.apow);
    gl2 acc0 = gl2::zero(), acc1 = gl2::zero(), accz = gl2::zero();
    const int total = p.ntr + p.nax + p.nqt;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        ulonglong2 a = __ldg(ap + k);
        gl2 al = mk2(a.x, a.y);
        if (k < p.ntr) {
            gl2 t = al * gl(__ldg(p.tr + (size_t)k * p.n + i));
            acc0 = acc0 + t; acc1 = acc1 + t;
        } else if (k < p.ntr + p.nax) {
            int c = k - p.ntr;
            gl v(__ldg(p.ax + (size_t)c * p.n + i));
            gl2 t = al * v;
            acc0 = acc0 + t; acc1 = acc1 + t;
            if (c >= p.zstart) { ulonglong2 b = __ldg(ap + (c - p.zstart)); accz = accz + mk2(b.x, b.y) * v; }
        } else {
            acc0 = acc0 + al * gl(__ldg(p.qt + (size_t)(k - p.ntr - p.nax) * p.n + i));
        }
    }
    const int t = threadIdx.x;
    sh[0][t][0] = acc0.a.v; sh[0][t][1] = acc0.b.v; sh[1][t][0] = acc1.a.v; sh[1][t][1] = acc1.b.v; sh[2][t][0] = accz.a.v; sh[2][t][1] = accz.b.v;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (t < off)
            for (int q = 0; q < 3; q++) {
                gl2 r = mk2(sh[q][t][0], sh[q][t][1]) + mk2(sh[q][t + off][0], sh[q][t + off][1]);
                sh[q][t][0] = r.a.v; sh[q][t][1] = r.b.v;
            }
        __syncthreads();
    }
    if (t < 6) p.out[(size_t)t * p.n + i] = sh[t >> 1][0][t & 1];
}

// ---- 2. pointwise combination on the coset 7*H_n ---------------------------------------------
struct CombineParams {
    const u64* r;                 // 6 columns of n: values of R0, R1, R2 on 7*w_n^i
    size_t n;
    PowTable wn;
    u64 zeta[2], zeta_next[2];
    u64 v0[2], v1[2], v2[2];      // R_b(point_b)
    u64 a0[2], a1[2];             // alpha^(|batch1|+|batch2|), alpha^|batch2|
    u64* out;                     // 2 columns of n
};

__global__ void __launch_bounds__(256) fri_combine_kernel(CombineParams p) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    gl x = gl(GL_GENERATOR) * pow_lookup(p.wn, i);
    // denominators: d0 = x - zeta, d1 = x - g zeta (extension), d2 = x - 1 (base)
    gl d0a = x - gl(p.zeta[0]), d0b = -gl(p.zeta[1]);
    gl d1a = x - gl(p.zeta_next[0]), d1b = -gl(p.zeta_next[1]);
    gl d2 = x - gl::one();
    gl n0 = d0a * d0a - gl(7) * (d0b * d0b), n1 = d1a * d1a - gl(7) * (d1b * d1b);
    gl inv = gl_inv_f(n0 * n1 * d2);
    gl i0 = inv * n1 * d2, i1 = inv * n0 * d2, i2 = inv * n0 * n1;
    gl2 q0 = (gl2(gl(p.r[i]), gl(p.r[p.n + i])) - gl2(gl(p.v0[0]), gl(p.v0[1]))) * gl2(d0a * i0, -(d0b * i0));
    gl2 q1 = (gl2(gl(p.r[2 * p.n + i]), gl(p.r[3 * p.n + i])) - gl2(gl(p.v1[0]), gl(p.v1[1]))) * gl2(d1a * i1, -(d1b * i1));
    gl2 q2 = (gl2(gl(p.r[4 * p.n + i]), gl(p.r[5 * p.n + i])) - gl2(gl(p.v2[0]), gl(p.v2[1]))) * i2;
    gl2 f = q0 * gl2(gl(p.a0[0]), gl(p.a0[1])) + q1 * gl2(gl(p.a1[0]), gl(p.a1[1])) + q2;
    p.out[i] = f.a.v; p.out[p.n + i] = f.b.v;
}

// ---- 3. commit phase ---------------------------------------------------------------------------
// rows[leaf*32 + 2k + {0,1}] = value at natural index bitrev(16*leaf + k) of the coset-major LDE.
__global__ void __launch_bounds__(256) fri_leaf_rows_kernel(const u64* __restrict__ lde, size_t cs, int log_nr, int rate_bits, int arity_bits,
                                                            u64* __restrict__ rows) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;         // = leaf * arity + k
    int L = log_nr + rate_bits;
    if (t >= ((size_t)1 << L)) return;
    u32 m = bitrev32((u32)t, L);
    u32 j = m & ((1u << rate_bits) - 1), idx = m >> rate_bits;
    size_t pos = ((size_t)j << log_nr) + idx;
    ulonglong2 v = make_ulonglong2(lde[pos], lde[cs + pos]);
    reinterpret_cast<ulonglong2*>(rows)[t] = v;
}

// out[i] = sum_j c[16 i + j] beta^j  (reduce_with_powers over coefficient chunks)
__global__ void __launch_bounds__(256) fri_fold_kernel(const u64* __restrict__ ca, const u64* __restrict__ cb, size_t n_out, int arity,
                                                       u64 beta_a, u64 beta_b, u64* __restrict__ oa, u64* __restrict__ ob) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    gl2 beta = mk2(beta_a, beta_b), acc = gl2::zero();
    for (int j = arity - 1; j >= 0; j--) acc = acc * beta + gl2(gl(ca[i * arity + j]), gl(cb[i * arity + j]));
    oa[i] = acc.a.v; ob[i] = acc.b.v;
}

// ---- 4. proof of work -------------------------------------------------------------------------
struct PowParams { u64 state[12]; int pos; int min_lz; u64 start; };
__global__ void __launch_bounds__(128) fri_pow_kernel(PowParams p, unsigned long long* best) {
    u64 w = p.start + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= GL_P) return;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = p.state[i];
#pragma unroll
    for (int i = 0; i < 12; i++) if (i == p.pos) s[i] = w;
    poseidon_permute_dev(s);
    int lz = s[7] ? __clzll((long long)s[7]) : 64;
    if (lz >= p.min_lz) atomicMin(best, (unsigned long long)w);
}

// ------------------------------------------------------------------------------------ launchers
void fri_reduce_batches(const Batch& trace, const Batch& aux, const Batch& quot, int zstart, const std::vector<gl2>& apow, u64* d_out,
                        cudaStream_t s) {
    size_t n = trace.n();
    ZKM_CHECK(aux.n() == n && quot.n() == n, "fri: oracle sizes differ");
    ZKM_CHECK((int)apow.size() >= trace.ncols + aux.ncols + quot.ncols, "fri: alpha power table too short");
    DevBuf ap(apow.size() * 2, s);
    std::vector<u64> h(apow.size() * 2);
    for (size_t k = 0; k < apow.size(); k++) { h[2 * k] = apow[k].a.v; h[2 * k + 1] = apow[k].b.v; }
    ap.upload(h.data(), h.size());
    ReduceParams p;
    p.tr = trace.coeffs.p; p.ntr = trace.ncols;
    p.ax = aux.coeffs.p; p.nax = aux.ncols; p.zstart = zstart;
    p.qt = quot.coeffs.p; p.nqt = quot.ncols;
    p.apow = ap.p; p.n = n; p.out = d_out;
    ProfScope ps("fri_reduce", s, 8.0 * (double)n * (trace.ncols + aux.ncols + quot.ncols) + 48.0 * (double)n);
    if (n < 4096) fri_reduce_small_kernel<<<(unsigned)n, 256, 0, s>>>(p);
    else fri_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
    ZKM_LAUNCHED();
    ZKM_CUDA(stream_sync(s));           // h / ap lifetime
}

void fri_combine(const u64* d_r, int log_n, gl2 zeta, gl2 zeta_next, gl2 v0, gl2 v1, gl2 v2, gl2 a0, gl2 a1, u64* d_out, cudaStream_t s) {
    size_t n = (size_t)1 << log_n;
    CombineParams p;
    p.r = d_r; p.n = n; p.wn = ntt_root_table(ctx().ntt, log_n, 0, s);      // cached per context
    p.zeta[0] = zeta.a.v; p.zeta[1] = zeta.b.v; p.zeta_next[0] = zeta_next.a.v; p.zeta_next[1] = zeta_next.b.v;
    p.v0[0] = v0.a.v; p.v0[1] = v0.b.v; p.v1[0] = v1.a.v; p.v1[1] = v1.b.v; p.v2[0] = v2.a.v; p.v2[1] = v2.b.v;
    p.a0[0] = a0.a.v; p.a0[1] = a0.b.v; p.a1[0] = a1.a.v; p.a1[1] = a1.b.v;
    p.out = d_out;
    ProfScope ps("fri_combine", s, 64.0 * (double)n);
    fri_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
    ZKM_LAUNCHED();
}

void fri_leaf_rows(const u64* d_lde, size_t col_stride, int log_nr, int rate_bits, int arity_bits, u64* d_rows, cudaStream_t s) {
    size_t N = (size_t)1 << (log_nr + rate_bits);
    ProfScope ps("fri_leaf_rows", s, 32.0 * (double)N);
    fri_leaf_rows_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(d_lde, col_stride, log_nr, rate_bits, arity_bits, d_rows);
    ZKM_LAUNCHED();
}

void fri_fold(const u64* d_in, size_t n_in, int arity_bits, gl2 beta, u64* d_out, cudaStream_t s) {
    size_t n_out = n_in >> arity_bits;
    ProfScope ps("fri_fold", s, 16.0 * (double)n_in + 16.0 * (double)n_out);
    fri_fold_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, s>>>(d_in, d_in + n_in, n_out, 1 << arity_bits, beta.a.v, beta.b.v, d_out,
                                                                   d_out + n_out);
    ZKM_LAUNCHED();
}

u64 fri_pow_grind(const u64 state[12], int pos, int min_lz, cudaStream_t s) {
    DevBuf best(1, s);
    const u64 NONE = ~(u64)0;
    best.upload(&NONE, 1);
    PowParams p;
    for (int i = 0; i < 12; i++) p.state[i] = state[i];
    p.pos = pos; p.min_lz = min_lz;
    const u64 batch = (u64)1 << (min_lz + 2 > 24 ? 24 : (min_lz + 2 < 12 ? 12 : min_lz + 2));
    ProfScope ps("fri_pow", s);
    for (u64 start = 0;; start += batch) {
        p.start = start;
        fri_pow_kernel<<<(unsigned)(batch / 128), 128, 0, s>>>(p, (unsigned long long*)best.p);
        ZKM_LAUNCHED();
        u64 h;
        best.download(&h, 1);
        if (h != NONE) return h;                  // the minimum over [0, start + batch): smaller batches were exhausted
        ZKM_CHECK(start + batch > start && start + batch < GL_P, "Proof of work failed. This is highly unlikely!");
    }
}

}  // namespace zkm
