// Batched Goldilocks NTT / inverse NTT / coset low-degree extension on column-major buffers.
// Replaces plonky2's per-column fft/ifft/coset_fft as reached from the reference's
// PolynomialBatch::from_values / from_coeffs call sites (prover/src/prover.rs:154-163,514-521,
// 576-587) and coset_ifft (:787).  Results are mathematically the same transforms, so outputs are
// bit-identical canonical field elements.
//
// Layout: a column is n contiguous u64.  A size-n transform (n = n1*n2) runs as two shared-memory
// passes, natural order in -> natural order out:
//   pass A  tile = n1 strided points x T consecutive i2   (loads/stores T*8-byte contiguous runs)
//           optional coset pre-scale, radix-2 DIF over i1, inter-pass twiddle w_n^(i2*k1)
//   pass B  tile = T consecutive k1 rows x n2 contiguous points, DIF over i2, transposed store
//           out[k1 + n1*k2]  (T*8-byte contiguous runs)
// n <= 2^12 runs as a single pass with T columns per CTA.
// The LDE onto 7*H_{4n} is four independent size-n coset transforms (shift 7*w_{4n}^j); the output is
// kept "coset-major": LDE natural index m = 4i + j lives at lde[j*n + i].
#include "ntt.cuh"
#include "dev.cuh"
#include "poseidon_v2.cuh"
#include <mutex>
#include <cstdlib>

namespace zkm {

unsigned long long g_launch_count = 0;

std::shared_ptr<PowTableOwner> make_pow_table(gl g, int max_bits, cudaStream_t s) {
    auto o = std::make_shared<PowTableOwner>();
    int lo_bits = max_bits < 10 ? max_bits : 10;
    size_t nlo = (size_t)1 << lo_bits;
    size_t nhi = max_bits > lo_bits ? ((size_t)1 << (max_bits - lo_bits)) : 1;
    std::vector<u64> lo(nlo), hi(nhi);
    gl cur = gl::one();
    for (size_t i = 0; i < nlo; i++) { lo[i] = cur.v; cur = cur * g; }
    gl step = cur;  // g^(2^lo_bits)
    cur = gl::one();
    for (size_t i = 0; i < nhi; i++) { hi[i] = cur.v; cur = cur * step; }
    o->lo.alloc(nlo, s); o->lo.upload(lo.data(), nlo);
    o->hi.alloc(nhi, s); o->hi.upload(hi.data(), nhi);
    ZKM_CUDA(cudaStreamSynchronize(s));
    o->view.lo = o->lo.p; o->view.hi = o->hi.p; o->view.lo_bits = lo_bits;
    return o;
}

struct NttPassParams {
    const u64* in; u64* out;
    size_t in_c, in_b, in_r, in_t, in_z;
    size_t out_c, out_b, out_r, out_t, out_z;
    int T, log_T;                 // tile width (independent transforms per CTA), a power of two
    u64 w16[8];                   // w_16^e (forward or inverse), e < 8
    int ncols_total;              // valid range for the "t = column" single-pass mode
    int t_is_column;              // single-pass mode: t indexes columns; guard against ncols
    int load_t_fast, store_t_fast;
    PowTable pre[8]; int has_pre; size_t pre_b, pre_r, pre_t;     // has_pre = 1: input scale  pre[z]^(b*pre_b+r*pre_r+t*pre_t)
    PowTable post; int has_post; size_t post_b, post_t;           // has_post = 1: output scale post^((b*post_b+t*post_t)*k)
    // Full-size factor tables of the two-pass transforms (has_pre / has_post = 2, FullTables below): the input is scaled by
    // pre_row[z][r] (the row part of the coset shift), the output by post_full[z][position inside the column] -- the inter-pass
    // twiddle times the column part of the coset shift times the constant output factor, ONE load + ONE multiply per element
    // where the two-level power tables cost two loads + two multiplies at each end.
    const u64* pre_row[8]; const u64* post_full[8];
    u64 scale;                    // constant output multiplier (1 = none)
    const u64* tw;                // w_R^k, k < R/2
};

constexpr int NTT_MAX_THREADS = 1024;

// a * b mod p (canonical) with the hand-scheduled product/reduction of poseidon_v2.cuh
__device__ __forceinline__ gl fmul(gl a, gl b) { return a * b; }

// Size-2^B DIF network on registers; v[j] ends up holding frequency bitrev_B(j).  w16[e] = w_16^e.
template <int B>
__device__ __forceinline__ void dft_regs(gl* v, const u64* w16) {
    constexpr int Q = 1 << B;
#pragma unroll
    for (int m = 0; m < B; m++) {
        const int half = Q >> (m + 1);
#pragma unroll
        for (int g = 0; g < (1 << m); g++) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const int i0 = g * 2 * half + j, i1 = i0 + half;
                gl a = v[i0], c = v[i1];
                v[i0] = a + c;
                gl d = a - c;
                // w_{2 half}^j = w_16^(j * 16 / (2 half))
                const int e = j * (16 / (2 * half));
                v[i1] = (e == 0) ? d : fmul(d, gl(w16[e]));
            }
        }
    }
}

// One radix-2^B step of the in-place DIF at stride 2^log_S (blocks of 2^(B + log_S) points); `done` = number
// of index bits already processed.  log_S / done are run-time values so that all radix-16 steps of a pass
// share ONE copy of the butterfly code (the fully specialised version was instruction-fetch bound).
template <int LOG_R, int B>
__device__ __forceinline__ void radix_step(u64* x, const u64* tw, const NttPassParams& p, int log_T, int tid, int nth, int log_S, int done) {
    constexpr int R = 1 << LOG_R, Q = 1 << B;
    const int S = 1 << log_S;
    const int T = 1 << log_T, TS = T + 1;
    const int items = (R / Q) << log_T;
    for (int idx = tid; idx < items; idx += nth) {
        const int t = idx & (T - 1), rest = idx >> log_T;
        const int i2 = rest & (S - 1), blk = rest >> log_S;
        const int base = (blk << (B + log_S)) + i2;
        gl v[Q];
#pragma unroll
        for (int j = 0; j < Q; j++) v[j] = gl(x[(base + (j << log_S)) * TS + t]);
        dft_regs<B>(v, p.w16);
#pragma unroll
        for (int j = 0; j < Q; j++) {
            const int k1 = (int)(__brev((unsigned)j) >> (32 - B));
            gl o = v[j];
            if (k1 != 0 && log_S > 0) o = fmul(o, gl(tw[(i2 * k1) << done]));      // w_{Q S}^(i2 k1) = w_R^((R / (Q S)) i2 k1)
            x[(base + (k1 << log_S)) * TS + t] = o.v;
        }
    }
    __syncthreads();
}

template <int LOG_R>
__device__ __forceinline__ void radix_steps(u64* x, const u64* tw, const NttPassParams& p, int log_T, int tid, int nth) {
    int done = 0;
    if constexpr (LOG_R >= 4) {
#pragma unroll 1
        for (; done + 4 <= LOG_R; done += 4) radix_step<LOG_R, 4>(x, tw, p, log_T, tid, nth, LOG_R - done - 4, done);
    }
    constexpr int TAIL = LOG_R % 4;
    if constexpr (TAIL > 0) radix_step<LOG_R, TAIL>(x, tw, p, log_T, tid, nth, 0, LOG_R - TAIL);
}
// position of frequency k after the in-place mixed-radix DIF: digits (radix 16, low digit first) reversed
template <int LOG_R>
__device__ __forceinline__ int freq_position(int k) {
    int pos = 0, done = 0;
#pragma unroll
    for (int rem = LOG_R; rem > 0;) {
        const int b = rem >= 4 ? 4 : rem;
        const int digit = (k >> done) & ((1 << b) - 1);
        rem -= b;
        pos |= digit << rem;
        done += b;
    }
    return pos;
}

// ---- bulk-copy (TMA engine) staging of the per-pass twiddle table: one elected thread issues a single cp.async.bulk of the R
// table entries into shared memory, completion is tracked by an mbarrier (expect_tx / complete_tx), and every thread waits on
// the barrier's phase only after it has issued its own tile loads -- the table arrives while the tile is in flight instead of
// costing each thread R / blockDim LDG + STS pairs up front.  (The tile itself cannot take this path: its coset scale is applied
// in registers on the way in and its rows are padded)  SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64.
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_g2s(void* dst_smem, const void* src_gmem, u32 bytes, u64* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 phase) {
    u32 done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(phase)
                     : "memory");
    } while (!done);
}

// FT: the launch uses the full-size factor tables (has_pre in {0, 2}, has_post = 2) -- a compile-time variant, so that neither
// path pays for the other's branches.
template <int LOG_R, bool FT>
__global__ void __launch_bounds__(NTT_MAX_THREADS) ntt_pass_kernel(NttPassParams p) {
    constexpr int R = 1 << LOG_R;
    extern __shared__ __align__(16) u64 smem[];
    __shared__ __align__(8) u64 tw_bar;
    const int T = p.T, TS = T + 1, log_T = p.log_T;
    u64* tw = smem;                    // R entries: w_R^e
    u64* x = smem + R;
    const int tid = threadIdx.x, nth = blockDim.x;
    const size_t b = blockIdx.x, c = blockIdx.y, z = blockIdx.z;
    constexpr bool BULK_TW = LOG_R >= 1;          // cp.async.bulk moves multiples of 16 bytes
    if (BULK_TW) {
        if (tid == 0) mbar_init(&tw_bar, 1);
        __syncthreads();
        if (tid == 0) bulk_load_g2s(tw, p.tw, (u32)(R * sizeof(u64)), &tw_bar);
    } else {
        if (tid == 0) tw[0] = p.tw[0];
    }
    const u64* in = p.in + c * p.in_c + b * p.in_b + z * p.in_z;
    u64* out = p.out + c * p.out_c + b * p.out_b + z * p.out_z;
    const int total = R << log_T;
    int tmax = T;
    if (p.t_is_column) { int rem = p.ncols_total - (int)(b * T); tmax = rem < T ? rem : T; }
    // ---- load (+ optional coset scale) ----  unrolled x8: +1.9 % (x16, or __ldg loads, measured slower)
#pragma unroll 8
    for (int idx = tid; idx < total; idx += nth) {
        int r, t;
        if (p.load_t_fast) { r = idx >> log_T; t = idx & (T - 1); } else { t = idx >> LOG_R; r = idx & (R - 1); }
        u64 v = 0;
        if (t < tmax) {
            v = in[(size_t)r * p.in_r + (size_t)t * p.in_t];
            if (FT) {
                if (p.has_pre) v = fmul(gl(v), gl(__ldg(p.pre_row[z] + r))).v;
            } else if (p.has_pre) {
                u64 e = b * p.pre_b + (size_t)r * p.pre_r + (size_t)t * p.pre_t;
                v = fmul(gl(v), pow_lookup(p.pre[z], e)).v;
            }
        }
        x[r * TS + t] = v;
    }
    if (BULK_TW) mbar_wait(&tw_bar, 0);           // the twiddle table has landed (phase 0 of its barrier completed)
    __syncthreads();
    // ---- mixed-radix (16) DIF over r, butterflies in registers ----
    radix_steps<LOG_R>(x, tw, p, log_T, tid, nth);
    // ---- store ----
    const gl scale(p.scale);
    if (FT) {
        // pass A of a two-pass transform: k-major tile (store_t_fast), every t valid; the factor loads of four elements are issued
        // together ahead of their multiplies
        const u64* __restrict__ ftab = p.post_full[z] + b * p.out_b;
#pragma unroll 4
        for (int idx = tid; idx < total; idx += nth) {
            const int k = idx >> log_T, t = idx & (T - 1);
            const size_t o = (size_t)k * p.out_r + (size_t)t * p.out_t;
            const gl f(__ldg(ftab + o));
            out[o] = fmul(gl(x[freq_position<LOG_R>(k) * TS + t]), f).v;
        }
        return;
    }
    for (int idx = tid; idx < total; idx += nth) {
        int k, t;
        if (p.store_t_fast) { k = idx >> log_T; t = idx & (T - 1); } else { t = idx >> LOG_R; k = idx & (R - 1); }
        if (t >= tmax) continue;
        gl v(x[freq_position<LOG_R>(k) * TS + t]);
        if (p.has_post) {
            u64 e = (b * p.post_b + (size_t)t * p.post_t) * (u64)k;
            v = fmul(v, pow_lookup(p.post, e));
        }
        if (p.scale != 1) v = fmul(v, scale);
        out[(size_t)k * p.out_r + (size_t)t * p.out_t] = v.v;
    }
}

typedef void (*ntt_kernel_t)(NttPassParams);
static ntt_kernel_t kernel_for(int log_r, bool ft) {
    switch (log_r) {
#define K(i) case i: return ft ? ntt_pass_kernel<i, true> : ntt_pass_kernel<i, false>;
        K(0) K(1) K(2) K(3) K(4) K(5) K(6) K(7) K(8) K(9) K(10) K(11) K(12)
#undef K
    }
    throw std::runtime_error("unsupported NTT radix");
}

static void launch_pass(int log_r, NttPassParams p, dim3 grid, cudaStream_t s, double alg_bytes) {
    size_t R = (size_t)1 << log_r;
    p.log_T = 0;
    while ((1 << p.log_T) < p.T) p.log_T++;
    ZKM_CHECK((1 << p.log_T) == p.T, "NTT tile width must be a power of two");
    size_t smem = (R + R * (p.T + 1)) * sizeof(u64);
    ntt_kernel_t k = kernel_for(log_r, p.has_post == 2);
    // alg_bytes: this launch's share of the transform's algorithmic bytes (inputs read once, outputs
    // written once -- charged to the first pass); the second counter is the traffic of the pass itself, 16 B per element
    ProfScope ps("ntt_pass", s, alg_bytes, 16.0 * (double)R * p.T * grid.x * grid.y * grid.z);
    // the kernel also holds a few bytes of static shared memory (the twiddle table's mbarrier): opt in to large dynamic shared
    // memory as soon as static + dynamic could pass the 48 KB default (R = 1024, T = 4 asks for exactly 48 KB of dynamic)
    if (smem + 256 > 48 * 1024) ZKM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one radix-16 work item per thread per step when the tile is large enough (8 warps per SM sub-partition)
    size_t items = (R >> (log_r >= 4 ? 4 : log_r)) * (size_t)p.T;
    int threads = items >= 1024 ? 1024 : (items >= 512 ? 512 : 256);
    static const int tune_threads = std::getenv("ZKM_NTT_THREADS") ? atoi(std::getenv("ZKM_NTT_THREADS")) : 0;   // tuning sweeps only
    if (tune_threads && (size_t)tune_threads <= items) threads = tune_threads;
    k<<<grid, threads, smem, s>>>(p);
    ZKM_LAUNCHED();
}

// ---------------------------------------------------------------------------------- tables
struct NttTables::Impl {
    std::mutex mu;
    // [log_n][inverse] -> powers of w_n (or w_n^-1) for exponents < n
    std::map<std::pair<int, int>, std::shared_ptr<PowTableOwner>> roots;
    // small per-radix twiddle arrays  w_R^k k<R/2
    std::map<std::pair<int, int>, std::shared_ptr<DevBuf>> tw;
    // coset shift tables: key (log_n, rate_bits, j, inverse, shift_exp_bits)
    std::map<std::tuple<int, int, int, int, int>, std::shared_ptr<PowTableOwner>> shifts;
    // full-size factor tables of the two-pass transforms: key (log_n, inverse, rate_bits, j, shift_exp_bits; rate_bits = -1: no shift)
    struct FullTab { DevBuf row, full; };
    std::map<std::tuple<int, int, int, int, int>, std::shared_ptr<FullTab>> full;
};
NttTables::NttTables() : impl(new Impl) {}
NttTables::~NttTables() { delete impl; }

static std::shared_ptr<PowTableOwner> get_roots(NttTables& T, int log_n, int inverse, cudaStream_t s) {
    std::lock_guard<std::mutex> g(T.impl->mu);
    auto key = std::make_pair(log_n, inverse);
    auto it = T.impl->roots.find(key);
    if (it != T.impl->roots.end()) return it->second;
    gl w = gl_root_of_unity(log_n);
    if (inverse) w = gl_inv(w);
    auto t = make_pow_table(w, log_n, s);
    T.impl->roots[key] = t;
    return t;
}
PowTable ntt_root_table(NttTables& t, int log_n, int inverse, cudaStream_t s) { return get_roots(t, log_n, inverse, s)->view; }
static const u64* get_tw(NttTables& T, int log_r, int inverse, cudaStream_t s) {
    std::lock_guard<std::mutex> g(T.impl->mu);
    auto key = std::make_pair(log_r, inverse);
    auto it = T.impl->tw.find(key);
    if (it != T.impl->tw.end()) return it->second->p;
    size_t half = (size_t)1 << log_r;               // all R powers w_R^e
    std::vector<u64> h(half);
    gl w = gl_root_of_unity(log_r);
    if (inverse) w = gl_inv(w);
    gl cur = gl::one();
    for (size_t i = 0; i < half; i++) { h[i] = cur.v; cur = cur * w; }
    auto b = std::make_shared<DevBuf>(half, s);
    b->upload(h.data(), half);
    ZKM_CUDA(cudaStreamSynchronize(s));
    T.impl->tw[key] = b;
    return b->p;
}
// powers of shift^(+-1) where shift = base * w_{n*2^rate_bits}^j
static std::shared_ptr<PowTableOwner> get_shift(NttTables& T, int log_n, int rate_bits, int j, int inverse, cudaStream_t s,
                                                int shift_exp_bits = 0) {
    std::lock_guard<std::mutex> g(T.impl->mu);
    auto key = std::make_tuple(log_n, rate_bits, j, inverse, shift_exp_bits);
    auto it = T.impl->shifts.find(key);
    if (it != T.impl->shifts.end()) return it->second;
    gl sh = gl_exp2(gl(GL_GENERATOR), shift_exp_bits) * gl_pow(gl_root_of_unity(log_n + rate_bits), (u64)j);
    if (inverse) sh = gl_inv(sh);
    auto t = make_pow_table(sh, log_n, s);
    T.impl->shifts[key] = t;
    return t;
}

// full[k1 * n2 + i2] = w_n^(i2 k1) * shift^i2 * scale   (inter-pass twiddle x column part of the coset shift x constant);
// row[r] = shift^(r n2)   (row part of the coset shift).  Field arithmetic is exact: the transforms give the same words.
__global__ void build_full_table_kernel(u64* full, u64* row, int l1, int l2, PowTable roots, PowTable shift, int has_shift, u64 scale) {
    const size_t n1 = (size_t)1 << l1, n2 = (size_t)1 << l2;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n1 * n2) return;
    const size_t k1 = idx >> l2, i2 = idx & (n2 - 1);
    gl v = pow_lookup(roots, (u64)i2 * k1);
    if (has_shift) v = v * pow_lookup(shift, i2);
    if (scale != 1) v = v * gl(scale);
    full[idx] = v.v;
    if (has_shift && idx < n1) row[idx] = pow_lookup(shift, (u64)idx * n2).v;
}
static void plan(int log_n, int& l1, int& l2, int& T);
struct ShiftKey { int rate_bits, j, shift_exp_bits; };               // coset shift 7^(2^shift_exp_bits) * w_{n 2^rate_bits}^j
static std::shared_ptr<NttTables::Impl::FullTab> get_full(NttTables& T, int log_n, int inverse, const ShiftKey* sk, u64 scale, cudaStream_t s) {
    auto key = sk ? std::make_tuple(log_n, inverse, sk->rate_bits, sk->j, sk->shift_exp_bits) : std::make_tuple(log_n, inverse, -1, 0, 0);
    {
        std::lock_guard<std::mutex> g(T.impl->mu);
        auto it = T.impl->full.find(key);
        if (it != T.impl->full.end()) return it->second;
    }
    int l1, l2, tw;
    plan(log_n, l1, l2, tw);
    auto roots = get_roots(T, log_n, inverse, s);
    std::shared_ptr<PowTableOwner> sh;
    if (sk) sh = get_shift(T, log_n, sk->rate_bits, sk->j, 0, s, sk->shift_exp_bits);
    auto t = std::make_shared<NttTables::Impl::FullTab>();
    const size_t n = (size_t)1 << log_n;
    t->full.alloc(n, s);
    t->row.alloc((size_t)1 << l1, s);
    build_full_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t->full.p, t->row.p, l1, l2, roots->view, sk ? sh->view : roots->view, sk != nullptr, scale);
    ZKM_LAUNCHED();
    ZKM_CUDA(cudaStreamSynchronize(s));
    std::lock_guard<std::mutex> g(T.impl->mu);
    T.impl->full[key] = t;
    return t;
}
// ZKM_NTT_FULLTAB=0 falls back to the two-level power tables on both ends of pass A (A/B switch)
static bool use_full_tables() {
    static const bool on = !(std::getenv("ZKM_NTT_FULLTAB") && atoi(std::getenv("ZKM_NTT_FULLTAB")) == 0);
    return on;
}

// Split log_n into (log_n1, log_n2) and choose the tile width.
static void plan(int log_n, int& l1, int& l2, int& T) {
    if (log_n <= 12) { l1 = log_n; l2 = 0; T = (1 << 12) >> log_n; if (T > 16) T = 16; if (T < 1) T = 1; return; }
    l2 = log_n / 2; l1 = log_n - l2;
    int lmax = l1;   // l1 >= l2
    T = (1 << 13) >> lmax;      // <= 8192 elements (66 KB + pad) per tile: 3 CTAs per SM overlap their load/compute/store phases
    if (T > 4) T = 4;           // sweep on B200 (tools/ntt_sweep.sh, 54 x 2^20): T = 4: 8.65 ms, 8: 8.90, 2: 9.18, 16: 10.2
    if (T < 2) T = 2;
    static const int tune_T = std::getenv("ZKM_NTT_T") ? atoi(std::getenv("ZKM_NTT_T")) : 0;                     // tuning sweeps only
    if (tune_T) T = tune_T;
}

// One size-n transform per (column, z).  shifts: optional pre-scale tables per z (coset).
// alg_bytes_per_col: algorithmic bytes of the whole call per column.
static void ntt_generic(NttTables& tabs, const u64* in, size_t in_cs, size_t in_zs, u64* out, size_t out_cs, size_t out_zs,
                        int ncols, int nz, int log_n, int inverse, const PowTable* pre, u64 final_scale, cudaStream_t s,
                        double alg_bytes_per_col, const ShiftKey* shift_keys = nullptr) {
    if (ncols == 0) return;
    int l1, l2, T;
    plan(log_n, l1, l2, T);
    size_t n = (size_t)1 << log_n;
    NttPassParams p = {};
    {
        gl w = gl_root_of_unity(4);
        if (inverse) w = gl_inv(w);
        gl cur = gl::one();
        for (int e = 0; e < 8; e++) { p.w16[e] = cur.v; cur = cur * w; }
    }
    p.has_pre = pre != nullptr;
    if (pre) for (int z = 0; z < nz; z++) p.pre[z] = pre[z];
    if (l2 == 0) {
        // single pass; t indexes columns
        p.in = in; p.out = out;
        p.in_c = 0; p.in_b = (size_t)T * in_cs; p.in_r = 1; p.in_t = in_cs; p.in_z = in_zs;
        p.out_c = 0; p.out_b = (size_t)T * out_cs; p.out_r = 1; p.out_t = out_cs; p.out_z = out_zs;
        p.T = T; p.t_is_column = 1; p.ncols_total = ncols;
        p.load_t_fast = 0; p.store_t_fast = 0;
        p.pre_b = 0; p.pre_r = 1; p.pre_t = 0;
        p.has_post = 0; p.scale = final_scale;
        p.tw = get_tw(tabs, l1, inverse, s);
        dim3 grid((ncols + T - 1) / T, 1, nz);
        launch_pass(l1, p, grid, s, alg_bytes_per_col * ncols);
        return;
    }
    size_t n1 = (size_t)1 << l1, n2 = (size_t)1 << l2;
    auto roots = get_roots(tabs, log_n, inverse, s);
    // full-size factor tables (one load + one multiply per element at each end of pass A; the constant output factor rides along)
    const bool fulltab = use_full_tables() && (!pre || shift_keys);
    std::shared_ptr<NttTables::Impl::FullTab> ft[8];
    if (fulltab)
        for (int z = 0; z < nz; z++) ft[z] = get_full(tabs, log_n, inverse, pre ? &shift_keys[z] : nullptr, final_scale, s);
    const u64* twA = get_tw(tabs, l1, inverse, s);
    const u64* twB = get_tw(tabs, l2, inverse, s);
    // Column chunks sized so the inter-pass scratch (pass A output) stays L2-resident (~64 MB).
    size_t per_col = (size_t)nz * n * sizeof(u64);
    int chunk = (int)((64u << 20) / per_col);
    if (chunk < 1) chunk = 1;
    if (chunk > ncols) chunk = ncols;
    DevBuf scratch((size_t)chunk * nz * n, s);
    for (int c0 = 0; c0 < ncols; c0 += chunk) {
        int nc = ncols - c0 < chunk ? ncols - c0 : chunk;
        // pass A: in -> scratch[(c*nz+z)*n + k1*n2 + i2], strided radix-n1 transforms
        p.in = in + (size_t)c0 * in_cs; p.out = scratch.p;
        p.in_c = in_cs; p.in_b = T; p.in_r = n2; p.in_t = 1; p.in_z = in_zs;
        p.out_c = (size_t)nz * n; p.out_b = T; p.out_r = n2; p.out_t = 1; p.out_z = n;
        p.T = T; p.t_is_column = 0; p.ncols_total = nc;
        p.load_t_fast = 1; p.store_t_fast = 1;
        p.pre_b = T; p.pre_r = n2; p.pre_t = 1;
        p.post = roots->view; p.has_post = 1; p.post_b = T; p.post_t = 1;
        p.scale = 1;
        if (fulltab) {
            p.has_post = 2;
            if (pre) p.has_pre = 2;
            for (int z = 0; z < nz; z++) { p.post_full[z] = ft[z]->full.p; p.pre_row[z] = ft[z]->row.p; }
        }
        p.tw = twA;
        launch_pass(l1, p, dim3((unsigned)(n2 / T), nc, nz), s, alg_bytes_per_col * nc);
        // pass B: scratch rows k1 (contiguous i2) -> out[k1 + n1*k2]
        NttPassParams q = {};
        for (int e = 0; e < 8; e++) q.w16[e] = p.w16[e];
        q.in = scratch.p; q.out = out + (size_t)c0 * out_cs;
        q.in_c = (size_t)nz * n; q.in_b = (size_t)T * n2; q.in_r = 1; q.in_t = n2; q.in_z = n;
        q.out_c = out_cs; q.out_b = T; q.out_r = n1; q.out_t = 1; q.out_z = out_zs;
        q.T = T; q.t_is_column = 0; q.ncols_total = nc;
        q.load_t_fast = 0; q.store_t_fast = 1;
        q.has_pre = 0; q.has_post = 0; q.scale = fulltab ? 1 : final_scale;
        q.tw = twB;
        launch_pass(l2, q, dim3((unsigned)(n1 / T), nc, nz), s, 0.0);
    }
}

void ntt_forward(NttTables& t, const u64* in, size_t in_cs, u64* out, size_t out_cs, int ncols, int log_n, cudaStream_t s) {
    ntt_generic(t, in, in_cs, 0, out, out_cs, 0, ncols, 1, log_n, 0, nullptr, 1, s, 16.0 * ((size_t)1 << log_n));
}
void ntt_inverse(NttTables& t, const u64* in, size_t in_cs, u64* out, size_t out_cs, int ncols, int log_n, cudaStream_t s) {
    gl ninv = gl_inv(gl((u64)1 << log_n));
    ntt_generic(t, in, in_cs, 0, out, out_cs, 0, ncols, 1, log_n, 1, nullptr, ninv.v, s, 16.0 * ((size_t)1 << log_n));
}
void lde_coset(NttTables& t, const u64* coeffs, size_t in_cs, u64* lde, size_t out_cs, int ncols, int log_n, int rate_bits,
               cudaStream_t s, int shift_exp_bits, int coset_begin, int coset_count) {
    ZKM_CHECK(rate_bits <= 3, "rate_bits > 3 unsupported");        // 3 = plonky2's standard_recursion_config (blow-up 8)
    const int all = 1 << rate_bits;
    if (coset_count < 0) coset_count = all - coset_begin;
    ZKM_CHECK(coset_begin >= 0 && coset_count >= 1 && coset_begin + coset_count <= all, "bad coset range");
    const int nz = coset_count;
    const size_t n = (size_t)1 << log_n;
    PowTable pre[8];
    std::shared_ptr<PowTableOwner> keep[8];
    ShiftKey keys[8];
    for (int z = 0; z < nz; z++) {
        keep[z] = get_shift(t, log_n, rate_bits, coset_begin + z, 0, s, shift_exp_bits); pre[z] = keep[z]->view;
        keys[z] = {rate_bits, coset_begin + z, shift_exp_bits};
    }
    // section 8(d): the LDE writes 8*n bytes per coset and column; its input is the coefficient vector the preceding iNTT (or
    // fold) just wrote, which the survey's 48*n*C figure for iNTT + LDE does not count a second time
    ntt_generic(t, coeffs, in_cs, 0, lde + (size_t)coset_begin * n, out_cs, n, ncols, nz, log_n, 0, pre, 1, s, 8.0 * nz * n, keys);
}

// values on the coset 7*H_n (natural order) -> coefficients:  ifft then scale coefficient i by 7^-i.
__global__ void scale_pow_kernel(u64* data, size_t cs, size_t n, PowTable tab) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64* col = data + (size_t)blockIdx.y * cs;
    col[i] = (gl(col[i]) * pow_lookup(tab, i)).v;
}
void coset_intt(NttTables& t, const u64* in, size_t in_cs, u64* out, size_t out_cs, int ncols, int log_n, cudaStream_t s) {
    ntt_inverse(t, in, in_cs, out, out_cs, ncols, log_n, s);
    auto sh = get_shift(t, log_n, 0, 0, 1, s);   // powers of 7^-1
    size_t n = (size_t)1 << log_n;
    dim3 grid((unsigned)((n + 255) / 256), ncols);
    ProfScope ps("scale_pow", s, 16.0 * (double)n * ncols);
    scale_pow_kernel<<<grid, 256, 0, s>>>(out, out_cs, n, sh->view);
    ZKM_LAUNCHED();
}

}  // namespace zkm
