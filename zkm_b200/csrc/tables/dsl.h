// Static description language for cross-table lookups (CTL) and in-table logUp lookups: the
// B200 build's counterpart of the reference's `Column` / `Filter` / `TableWithColumns` /
// `CrossTableLookup` (prover/src/cross_table_lookup.rs:33-116,120-345,347-382) and `Lookup`
// (prover/src/lookup.rs:20-39).  Host-only plain data; the device side consumes a flattened copy
// (ctl.cu), the CPU oracle walks these structs directly.
#pragma once
#include <stdint.h>
#include <vector>
#include <utility>
#include <initializer_list>

namespace zkm {
namespace tables {

typedef uint64_t u64;
static const u64 FP = 0xFFFFFFFF00000001ULL;

inline u64 fmul(u64 a, u64 b) { return (u64)((unsigned __int128)a * b % FP); }
inline u64 fadd(u64 a, u64 b) { return (u64)(((unsigned __int128)a + b) % FP); }
inline u64 fneg(u64 a) { return a ? FP - a : 0; }

// Linear combination of columns of the current row, of the next row, plus a constant
// (cross_table_lookup.rs:120-124).
struct Column {
    std::vector<std::pair<int, u64>> lin;
    std::vector<std::pair<int, u64>> next;
    u64 constant = 0;

    static Column single(int c) { Column r; r.lin.push_back({c, 1}); return r; }
    static Column single_next_row(int c) { Column r; r.next.push_back({c, 1}); return r; }
    static Column constant_(u64 v) { Column r; r.constant = v; return r; }
    static Column zero() { return constant_(0); }
    static Column one() { return constant_(1); }
    static Column linear_combination(std::vector<std::pair<int, u64>> v) { Column r; r.lin = std::move(v); return r; }
    static Column linear_combination_with_constant(std::vector<std::pair<int, u64>> v, u64 c) {
        Column r; r.lin = std::move(v); r.constant = c; return r;
    }
    static Column linear_combination_and_next_row_with_constant(std::vector<std::pair<int, u64>> v,
                                                                std::vector<std::pair<int, u64>> nv, u64 c) {
        Column r; r.lin = std::move(v); r.next = std::move(nv); r.constant = c; return r;
    }
    // sum_i cols[i] * 2^i
    static Column le_bits(const std::vector<int>& cols) {
        Column r; u64 p = 1;
        for (int c : cols) { r.lin.push_back({c, p}); p = fmul(p, 2); }
        return r;
    }
    // sum_i cols[i] * 256^i
    static Column le_bytes(const std::vector<int>& cols) {
        Column r; u64 p = 1;
        for (int c : cols) { r.lin.push_back({c, p}); p = fmul(p, 256); }
        return r;
    }
    static Column sum(const std::vector<int>& cols) {
        Column r;
        for (int c : cols) r.lin.push_back({c, 1});
        return r;
    }
    static std::vector<Column> singles(const std::vector<int>& cols) {
        std::vector<Column> r;
        for (int c : cols) r.push_back(single(c));
        return r;
    }
};

inline std::vector<int> range(int a, int b) { std::vector<int> r; for (int i = a; i < b; i++) r.push_back(i); return r; }

// sum_i products[i].0 * products[i].1 + sum_j constants[j]   (cross_table_lookup.rs:33-70)
struct Filter {
    std::vector<std::pair<Column, Column>> products;
    std::vector<Column> constants;
    bool present = false;          // Option<Filter>: false = None (always selected)
    static Filter none() { return Filter(); }
    static Filter new_simple(Column c) { Filter f; f.constants.push_back(std::move(c)); f.present = true; return f; }
    static Filter new_(std::vector<std::pair<Column, Column>> p, std::vector<Column> c) {
        Filter f; f.products = std::move(p); f.constants = std::move(c); f.present = true; return f;
    }
};

struct TableWithColumns {
    int table = 0;
    std::vector<Column> columns;
    Filter filter;
    TableWithColumns() {}
    TableWithColumns(int t, std::vector<Column> c, Filter f) : table(t), columns(std::move(c)), filter(std::move(f)) {}
};

struct CrossTableLookup {
    std::vector<TableWithColumns> looking_tables;
    TableWithColumns looked_table;
};

// logUp range check inside one table (lookup.rs:20-31).
struct Lookup {
    std::vector<Column> columns;
    Column table_column;
    Column frequencies_column;
    std::vector<Filter> filter_columns;      // one per column (present=false = None)
    int num_helper_columns(int constraint_degree) const {
        return ((int)columns.size() + constraint_degree - 2) / (constraint_degree - 1) + 1;
    }
};

}  // namespace tables
}  // namespace zkm
