// Auxiliary polynomials of one table: logUp lookup helper columns and cross-table-lookup helper /
// Z columns, generated on the device from the resident trace values.
// Replaces the reference's serial, per-row-allocating generators:
//   get_helper_cols      cross_table_lookup.rs:709-795   (1/combine(columns) where filter = 1, else 0)
//   partial_sums         cross_table_lookup.rs:841-872   (helpers summed pairwise, Z = suffix sums)
//   lookup_helper_columns lookup.rs:46-124               (helpers, 1/(table+x), Z = prefix sums)
// Kernels: aux_terms (row-parallel linear combinations), batch_inverse (Montgomery trick, 8 per
// thread, one Fermat inversion per thread), aux_assemble (pairwise sums), column scans.
#include "devprog.cuh"
#include "aux.cuh"
#include <cstring>

namespace zkm {

// ------------------------------------------------------------------------------- program build
void DProgram::build(const tables::TableLayout& L, int num_challenges) {
    using namespace tables;
    cols.clear(); terms.clear(); filters.clear(); pairs.clear(); idx.clear(); parts.clear(); zs.clear(); lookups.clear();
    uses_next = false;
    auto add_column = [&](const Column& c) {
        DColumn d;
        d.lin_off = (int)terms.size(); d.lin_cnt = (int)c.lin.size();
        for (auto& p : c.lin) terms.push_back({p.first, p.second});
        d.next_off = (int)terms.size(); d.next_cnt = (int)c.next.size();
        for (auto& p : c.next) terms.push_back({p.first, p.second});
        if (!c.next.empty()) uses_next = true;
        d.constant = c.constant;
        cols.push_back(d);
        return (int)cols.size() - 1;
    };
    auto add_filter = [&](const Filter& f) {
        DFilter d;
        d.present = f.present ? 1 : 0;
        std::vector<int> pr, cs;
        for (auto& p : f.products) { pr.push_back(add_column(p.first)); pr.push_back(add_column(p.second)); }
        for (auto& c : f.constants) cs.push_back(add_column(c));
        d.prod_off = (int)pairs.size(); d.prod_cnt = (int)f.products.size();
        pairs.insert(pairs.end(), pr.begin(), pr.end());
        d.const_off = (int)idx.size(); d.const_cnt = (int)cs.size();
        idx.insert(idx.end(), cs.begin(), cs.end());
        filters.push_back(d);
        return (int)filters.size() - 1;
    };
    auto add_part = [&](const std::vector<Column>& cs, const Filter& f, int challenge, int is_lookup) {
        std::vector<int> ci;
        for (auto& c : cs) ci.push_back(add_column(c));
        DPart p;
        p.filter = add_filter(f);
        p.col_off = (int)idx.size(); p.col_cnt = (int)ci.size();
        idx.insert(idx.end(), ci.begin(), ci.end());
        p.challenge = challenge; p.is_lookup = is_lookup;
        parts.push_back(p);
        return (int)parts.size() - 1;
    };
    int aux = 0;
    for (const Lookup& lk : L.lookups) {
        ZKM_CHECK(lk.columns.size() == lk.filter_columns.size(), "lookup: columns/filters length mismatch");
        for (int ch = 0; ch < num_challenges; ch++) {
            DLookup d;
            d.part_off = (int)parts.size(); d.part_cnt = (int)lk.columns.size();
            for (size_t i = 0; i < lk.columns.size(); i++) add_part({lk.columns[i]}, lk.filter_columns[i], ch, 1);
            d.table_part = add_part({lk.table_column}, Filter::none(), ch, 1);
            d.table_col = add_column(lk.table_column);
            d.freq_col = add_column(lk.frequencies_column);
            d.num_helpers = (d.part_cnt + 1) / 2;
            d.aux_start = aux;
            d.challenge = ch;
            aux += d.num_helpers + 1;
            lookups.push_back(d);
        }
    }
    ZKM_CHECK(aux == L.num_lookup_cols, "lookup column count mismatch");
    for (size_t z = 0; z < L.zs.size(); z++) {
        const CtlZInfo& zi = L.zs[z];
        DZ d;
        d.part_off = (int)parts.size(); d.part_cnt = (int)zi.parts.size();
        for (auto& p : zi.parts) add_part(p.columns, p.filter, zi.challenge, 0);
        d.num_helpers = zi.num_helpers;
        d.helper_aux = L.helper_col((int)z, 0);
        d.z_aux = L.z_col((int)z);
        d.challenge = zi.challenge;
        zs.push_back(d);
    }
}

template <class T>
static size_t place(std::vector<unsigned char>& blob, const std::vector<T>& v) {
    size_t off = (blob.size() + 15) & ~(size_t)15;
    blob.resize(off + v.size() * sizeof(T));
    if (!v.empty()) memcpy(blob.data() + off, v.data(), v.size() * sizeof(T));
    return off;
}

void DProgram::upload(cudaStream_t s) {
    std::vector<unsigned char> b;
    size_t o_cols = place(b, cols), o_terms = place(b, terms), o_filters = place(b, filters), o_pairs = place(b, pairs),
           o_idx = place(b, idx), o_parts = place(b, parts), o_zs = place(b, zs), o_lookups = place(b, lookups);
    b.resize((b.size() + 15) & ~(size_t)15);
    blob.alloc(b.size() / 8 + 2, s);
    ZKM_CUDA(cudaMemcpyAsync(blob.p, b.data(), b.size(), cudaMemcpyHostToDevice, s));
    ZKM_CUDA(stream_sync(s));
    unsigned char* base = (unsigned char*)blob.p;
    view.cols = (const DColumn*)(base + o_cols); view.terms = (const DTerm*)(base + o_terms);
    view.filters = (const DFilter*)(base + o_filters); view.pairs = (const int*)(base + o_pairs);
    view.idx = (const int*)(base + o_idx); view.parts = (const DPart*)(base + o_parts);
    view.zs = (const DZ*)(base + o_zs); view.lookups = (const DLookup*)(base + o_lookups);
    view.num_parts = (int)parts.size(); view.num_zs = (int)zs.size(); view.num_lookups = (int)lookups.size();
}

// ------------------------------------------------------------------------------------ kernels
// Trace values on H, column-major, natural order.  "next" terms read 0 on the last row
// (Column::eval_table, cross_table_lookup.rs:266-285).
struct ValRow {
    const u64* base; size_t stride; bool valid;
    __device__ __forceinline__ gl operator[](int c) const { return valid ? gl(__ldg(base + (size_t)c * stride)) : gl::zero(); }
};

__global__ void __launch_bounds__(256) aux_terms_kernel(DProgramView P, const u64* __restrict__ values, size_t n, AuxChallenges ch,
                                                        u64* __restrict__ den, unsigned char* __restrict__ mask, int* __restrict__ err) {
    size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const DPart p = P.parts[blockIdx.y];
    ValRow lv{values + d, n, true}, nv{values + d + 1, n, d + 1 < n};
    gl f = dfilter_eval(P, p.filter, lv, nv);
    gl beta = p.is_lookup ? gl::one() : gl(ch.beta[p.challenge]);
    gl gamma = p.is_lookup ? gl(ch.beta[p.challenge]) : gl(ch.gamma[p.challenge]);
    gl v = gl::one();
    unsigned char m = 0;
    if (f.v == 1) {
        v = dpart_combine(P, p, lv, nv, beta, gamma);
        m = 1;
        if (v.v == 0) atomicMax(err, 2);
    } else if (f.v != 0) {
        atomicMax(err, 1);                           // "Non-binary filter?" (cross_table_lookup.rs:741)
    }
    size_t o = (size_t)blockIdx.y * n + d;
    den[o] = v.v;
    mask[o] = m;
}

// x^(p-2) by square-and-multiply over the fixed exponent 0xFFFFFFFEFFFFFFFF.
__device__ __forceinline__ gl gl_inv_dev(gl x) {
    // p - 2 = 2^64 - 2^32 - 1: x^(2^32-1) then shift, standard addition chain
    gl x2 = x * x * x;                               // 2 bits
    gl x4 = gl_exp2(x2, 2) * x2;                     // 4 bits
    gl x8 = gl_exp2(x4, 4) * x4;
    gl x16 = gl_exp2(x8, 8) * x8;
    gl x32 = gl_exp2(x16, 16) * x16;                 // x^(2^32-1)
    gl x31 = gl_exp2(x16, 8) * x8;                   // 24 bits
    x31 = gl_exp2(x31, 4) * x4;                      // 28 bits
    x31 = gl_exp2(x31, 2) * x2;                      // 30 bits
    x31 = gl_exp2(x31, 1) * x;                       // x^(2^31-1)
    // exponent bits: 31 ones, a zero, 32 ones  ->  ((x^(2^31-1))^2)^(2^32) * x^(2^32-1)
    gl r = gl_exp2(x31, 33) * x32;
    return r;
}

__global__ void __launch_bounds__(256) batch_inverse_kernel(u64* __restrict__ data, const unsigned char* __restrict__ mask, size_t total) {
    constexpr int K = 8;
    size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * K;
    if (base >= total) return;
    gl x[K], pre[K];
    gl acc = gl::one();
#pragma unroll
    for (int k = 0; k < K; k++) {
        x[k] = base + k < total ? gl(data[base + k]) : gl::one();
        pre[k] = acc;
        acc = acc * x[k];
    }
    gl inv = gl_inv_dev(acc);
#pragma unroll
    for (int k = K - 1; k >= 0; k--) {
        gl r = inv * pre[k];
        inv = inv * x[k];
        if (base + k < total) data[base + k] = (mask == nullptr || mask[base + k]) ? r.v : 0;
    }
}

// Writes the helper columns and the (unscanned) per-row increments of every Z column.
__global__ void __launch_bounds__(256) aux_assemble_kernel(DProgramView P, const u64* __restrict__ values, const u64* __restrict__ inv, size_t n,
                                                           u64* __restrict__ aux) {
    size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    int item = blockIdx.y;
    if (item < P.num_zs) {
        const DZ z = P.zs[item];
        gl sum = gl::zero();
        for (int j = 0; 2 * j < z.part_cnt; j++) {
            gl h(inv[(size_t)(z.part_off + 2 * j) * n + d]);
            if (2 * j + 1 < z.part_cnt) h = h + gl(inv[(size_t)(z.part_off + 2 * j + 1) * n + d]);
            if (z.num_helpers) aux[(size_t)(z.helper_aux + j) * n + d] = h.v;
            sum = sum + h;
        }
        aux[(size_t)z.z_aux * n + d] = sum.v;
    } else {
        const DLookup l = P.lookups[item - P.num_zs];
        gl sum = gl::zero();
        for (int j = 0; j < l.num_helpers; j++) {
            gl h(inv[(size_t)(l.part_off + 2 * j) * n + d]);
            if (2 * j + 1 < l.part_cnt) h = h + gl(inv[(size_t)(l.part_off + 2 * j + 1) * n + d]);
            aux[(size_t)(l.aux_start + j) * n + d] = h.v;
            sum = sum + h;
        }
        ValRow lv{values + d, n, true}, nv{values + d + 1, n, d + 1 < n};
        gl freq = dcol_eval(P, l.freq_col, lv, nv);
        sum = sum - freq * gl(inv[(size_t)l.table_part * n + d]);
        aux[(size_t)(l.aux_start + l.num_helpers) * n + d] = sum.v;
    }
}

// ---- column scans over field addition.  mode 0: z[d] = sum_{j>=d} s[j] (inclusive suffix, CTL Z);
//      mode 1: z[d] = sum_{j<d} s[j] (exclusive prefix, logUp Z).  In place.
struct ScanCols { int col[64]; int mode[64]; };
constexpr int SCAN_T = 256, SCAN_K = 8, SCAN_CHUNK = SCAN_T * SCAN_K;

__device__ __forceinline__ size_t scan_pos(size_t k, size_t n, int mode) { return mode == 0 ? n - 1 - k : k; }

__global__ void __launch_bounds__(SCAN_T) scan_reduce_kernel(const u64* __restrict__ aux, size_t n, ScanCols sc, u64* __restrict__ partial,
                                                             size_t nblk) {
    __shared__ u64 sh[SCAN_T];
    const u64* col = aux + (size_t)sc.col[blockIdx.y] * n;
    int mode = sc.mode[blockIdx.y];
    size_t k0 = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_K;
    gl s = gl::zero();
#pragma unroll
    for (int k = 0; k < SCAN_K; k++) if (k0 + k < n) s = s + gl(col[scan_pos(k0 + k, n, mode)]);
    sh[threadIdx.x] = s.v;
    __syncthreads();
    for (int off = SCAN_T / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] = (gl(sh[threadIdx.x]) + gl(sh[threadIdx.x + off])).v;
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * nblk + blockIdx.x] = sh[0];
}
__global__ void scan_partials_kernel(u64* partial, size_t nblk, int ncols) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    gl run = gl::zero();
    for (size_t b = 0; b < nblk; b++) { gl v(partial[(size_t)c * nblk + b]); partial[(size_t)c * nblk + b] = run.v; run = run + v; }
}
__global__ void __launch_bounds__(SCAN_T) scan_apply_kernel(u64* __restrict__ aux, size_t n, ScanCols sc, const u64* __restrict__ partial,
                                                            size_t nblk) {
    __shared__ u64 sh[SCAN_T];
    u64* col = aux + (size_t)sc.col[blockIdx.y] * n;
    int mode = sc.mode[blockIdx.y];
    size_t k0 = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_K;
    gl v[SCAN_K];
    gl s = gl::zero();
#pragma unroll
    for (int k = 0; k < SCAN_K; k++) { v[k] = k0 + k < n ? gl(col[scan_pos(k0 + k, n, mode)]) : gl::zero(); s = s + v[k]; }
    sh[threadIdx.x] = s.v;
    __syncthreads();
    // inclusive Hillis-Steele scan of the thread sums
    for (int off = 1; off < SCAN_T; off <<= 1) {
        gl add = threadIdx.x >= off ? gl(sh[threadIdx.x - off]) : gl::zero();
        __syncthreads();
        sh[threadIdx.x] = (gl(sh[threadIdx.x]) + add).v;
        __syncthreads();
    }
    gl run = gl(partial[(size_t)blockIdx.y * nblk + blockIdx.x]) + (threadIdx.x ? gl(sh[threadIdx.x - 1]) : gl::zero());
#pragma unroll
    for (int k = 0; k < SCAN_K; k++) {
        if (k0 + k >= n) break;
        gl out;
        if (mode == 0) { run = run + v[k]; out = run; } else { out = run; run = run + v[k]; }
        col[scan_pos(k0 + k, n, mode)] = out.v;
    }
}

void compute_aux_columns(const DProgram& prog, const tables::TableLayout& L, const u64* d_values, int log_n, const AuxChallenges& ch,
                         u64* d_aux, cudaStream_t s) {
    size_t n = (size_t)1 << log_n;
    int np = prog.view.num_parts;
    ZKM_CHECK(np > 0, "No CTL?");
    DevBuf den((size_t)np * n, s), maskbuf(((size_t)np * n + 7) / 8 + 1, s), errbuf(1, s);
    errbuf.zero();
    unsigned gx = (unsigned)((n + 255) / 256);
    {
        ProfScope ps("aux_terms", s);
        aux_terms_kernel<<<dim3(gx, np), 256, 0, s>>>(prog.view, d_values, n, ch, den.p, (unsigned char*)maskbuf.p, (int*)errbuf.p);
        ZKM_LAUNCHED();
    }
    {
        ProfScope ps("batch_inverse", s);
        size_t total = (size_t)np * n;
        batch_inverse_kernel<<<(unsigned)((total + 8 * 256 - 1) / (8 * 256)), 256, 0, s>>>(den.p, (unsigned char*)maskbuf.p, total);
        ZKM_LAUNCHED();
    }
    int items = prog.view.num_zs + prog.view.num_lookups;
    {
        ProfScope ps("aux_assemble", s);
        aux_assemble_kernel<<<dim3(gx, items), 256, 0, s>>>(prog.view, d_values, den.p, n, d_aux);
        ZKM_LAUNCHED();
    }
    u64 herr = 0;
    errbuf.download(&herr, 1);
    int e = (int)(herr & 0xffffffffu);
    ZKM_CHECK(e != 1, "Non-binary filter?");
    ZKM_CHECK(e != 2, "lookup denominator is zero (batch inverse of zero)");
    // scans, 64 columns per launch
    std::vector<std::pair<int, int>> todo;
    for (const DZ& z : prog.zs) todo.push_back({z.z_aux, 0});
    for (const DLookup& l : prog.lookups) todo.push_back({l.aux_start + l.num_helpers, 1});
    size_t nblk = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    DevBuf partial((size_t)64 * nblk, s);
    ProfScope ps("aux_scan", s);
    for (size_t i = 0; i < todo.size(); i += 64) {
        ScanCols sc;
        int cnt = (int)std::min<size_t>(64, todo.size() - i);
        for (int k = 0; k < cnt; k++) { sc.col[k] = todo[i + k].first; sc.mode[k] = todo[i + k].second; }
        scan_reduce_kernel<<<dim3((unsigned)nblk, cnt), SCAN_T, 0, s>>>(d_aux, n, sc, partial.p, nblk);
        ZKM_LAUNCHED();
        scan_partials_kernel<<<1, 64, 0, s>>>(partial.p, nblk, cnt);
        ZKM_LAUNCHED();
        scan_apply_kernel<<<dim3((unsigned)nblk, cnt), SCAN_T, 0, s>>>(d_aux, n, sc, partial.p, nblk);
        ZKM_LAUNCHED();
    }
}

}  // namespace zkm
