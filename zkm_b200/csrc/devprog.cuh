// Flattened, device-readable form of one table's lookup / cross-table-lookup description
// (tables/dsl.h Column, Filter, Lookup and tables/system.h CtlZInfo), consumed by the auxiliary-
// column kernels (ctl.cu) and by the generic part of the quotient kernels (quotient.cu).
// Mirrors reference cross_table_lookup.rs:33-70 (Filter), :120-124 (Column), :427-438 (CtlZData)
// and lookup.rs:20-31 (Lookup).
#pragma once
#include "dev.cuh"
#include "tables/system.h"

namespace zkm {

struct DColumn {                 // sum lin + sum next + constant
    int lin_off, lin_cnt;        // into terms[]
    int next_off, next_cnt;
    u64 constant;
};
struct DTerm { int col; u64 coef; };
struct DFilter {                 // sum prod(a*b) + sum consts ; present == 0 -> constant 1
    int present;
    int prod_off, prod_cnt;      // into pairs[] (2 DColumn indices per product)
    int const_off, const_cnt;    // into idx[]   (DColumn indices)
};
// One (columns, filter) set: value = combine(columns) = sum_i col_i * beta^i + gamma.
struct DPart {
    int col_off, col_cnt;        // into idx[] (DColumn indices)
    int filter;                  // DFilter index
    int challenge;               // challenge set index: beta/gamma = ctl challenge, or (1, beta) for lookups
    int is_lookup;               // 1: logUp part (beta = 1, gamma = ctl beta[challenge])
};
struct DZ {                      // one CTL Z polynomial
    int part_off, part_cnt;      // into parts[]
    int num_helpers;
    int helper_aux;              // aux column index of the first helper
    int z_aux;                   // aux column index of Z
    int challenge;
};
struct DLookup {                 // one (Lookup, challenge) instance
    int part_off, part_cnt;      // looked-up columns as parts (single column each)
    int table_part;              // part index of (table_column, no filter): 1/(table + challenge)
    int table_col, freq_col;     // DColumn indices
    int aux_start;               // first helper aux column; Z at aux_start + num_helpers
    int num_helpers;             // ceil(part_cnt / 2)
    int challenge;
};

struct DProgramView {
    const DColumn* cols; const DTerm* terms; const DFilter* filters; const int* pairs; const int* idx;
    const DPart* parts; const DZ* zs; const DLookup* lookups;
    int num_parts, num_zs, num_lookups;
};

struct DProgram {
    // host copies
    std::vector<DColumn> cols; std::vector<DTerm> terms; std::vector<DFilter> filters; std::vector<int> pairs, idx;
    std::vector<DPart> parts; std::vector<DZ> zs; std::vector<DLookup> lookups;
    bool uses_next = false;       // some Column reads the next row
    DevBuf blob;
    DProgramView view{};
    void build(const tables::TableLayout& L, int num_challenges);
    void upload(cudaStream_t s);
};

#ifdef __CUDACC__
// Column on a row accessor pair (Column::eval_with_next, cross_table_lookup.rs:247-263).
template <class V>
__device__ __forceinline__ gl dcol_eval(const DProgramView& P, int ci, const V& lv, const V& nv) {
    const DColumn c = P.cols[ci];
    gl r(c.constant);
    for (int k = 0; k < c.lin_cnt; k++) { DTerm t = P.terms[c.lin_off + k]; r = r + lv[t.col] * gl(t.coef); }
    for (int k = 0; k < c.next_cnt; k++) { DTerm t = P.terms[c.next_off + k]; r = r + nv[t.col] * gl(t.coef); }
    return r;
}
template <class V>
__device__ __forceinline__ gl dfilter_eval(const DProgramView& P, int fi, const V& lv, const V& nv) {
    const DFilter f = P.filters[fi];
    if (!f.present) return gl::one();
    gl s = gl::zero();
    for (int k = 0; k < f.prod_cnt; k++)
        s = s + dcol_eval(P, P.pairs[f.prod_off + 2 * k], lv, nv) * dcol_eval(P, P.pairs[f.prod_off + 2 * k + 1], lv, nv);
    for (int k = 0; k < f.const_cnt; k++) s = s + dcol_eval(P, P.idx[f.const_off + k], lv, nv);
    return s;
}
// combine(columns) = reduce_with_powers(evals, beta) + gamma (cross_table_lookup.rs:494-504)
template <class V>
__device__ __forceinline__ gl dpart_combine(const DProgramView& P, const DPart& p, const V& lv, const V& nv, gl beta, gl gamma) {
    gl acc = gl::zero();
    for (int k = p.col_cnt - 1; k >= 0; k--) acc = acc * beta + dcol_eval(P, P.idx[p.col_off + k], lv, nv);
    return acc + gamma;
}
#endif

}  // namespace zkm
