// A "system" = an ordered list of STARK tables plus the cross-table lookups between them: the
// B200 build's counterpart of the reference's `AllStark` (prover/src/all_stark.rs:39-94, Table enum
// :97-110, all_cross_table_lookups :136-154).  The full 12-table AllStark is one System
// (all_stark.h); tests also prove smaller Systems through the very same code path.
//
// derive_layout() fixes, per table, which auxiliary polynomials exist and in which order — the
// prover-side bookkeeping of cross_table_lookup_data (cross_table_lookup.rs:634-703), partial_sums
// (:841-872) and prove_single_table's auxiliary_polys assembly (prover.rs:469-509):
//     [ lookup columns: per Lookup, per challenge: ceil(k/2) helpers then Z ]
//     [ CTL helper columns, in zs_columns push order ]
//     [ CTL Z columns, in zs_columns push order ]
#pragma once
#include "dsl.h"
#include "hd.h"
#include <stdexcept>
#include <string>

namespace zkm {
namespace tables {

enum TableKind {
    T_ARITHMETIC = 0, T_CPU, T_POSEIDON, T_POSEIDON_SPONGE, T_KECCAK, T_KECCAK_SPONGE, T_SHA_EXTEND,
    T_SHA_EXTEND_SPONGE, T_SHA_COMPRESS, T_SHA_COMPRESS_SPONGE, T_LOGIC, T_MEMORY, NUM_TABLE_KINDS
};
static const int CONSTRAINT_DEGREE = 3;            // every table's constraint_degree() (e.g. cpu_stark.rs:319-321)
static const int QUOTIENT_DEGREE_FACTOR = 2;       // stark.rs:82-84

struct System {
    std::vector<int> kinds;                        // table i is of kind kinds[i]
    std::vector<CrossTableLookup> ctls;            // TableWithColumns::table indexes `kinds`
};

// One Z polynomial of one table (CtlZData, cross_table_lookup.rs:427-438).
struct CtlZInfo {
    int challenge = 0;                             // index into ctl_challenges
    std::vector<TableWithColumns> parts;           // (columns, filter) sets summed into this Z
    int num_helpers = 0;                           // ceil(parts/2) if parts > 1 else 0
};

struct TableLayout {
    int kind = 0, ncols = 0;
    std::vector<Lookup> lookups;
    int num_lookup_cols = 0;                       // Stark::num_lookup_helper_columns (stark.rs:217-223)
    std::vector<CtlZInfo> zs;
    int num_ctl_helpers = 0;
    int num_aux() const { return num_lookup_cols + num_ctl_helpers + (int)zs.size(); }
    int helper_col(int z, int j) const {           // aux index of helper j of Z number z
        int off = num_lookup_cols;
        for (int i = 0; i < z; i++) off += zs[i].num_helpers;
        return off + j;
    }
    int z_col(int z) const { return num_lookup_cols + num_ctl_helpers + z; }
};

int table_num_columns(int kind);                   // defined in registry.h
std::vector<Lookup> table_lookups(int kind);
const char* table_name(int kind);

inline std::vector<TableLayout> derive_layout(const System& sys, int num_challenges) {
    std::vector<TableLayout> out(sys.kinds.size());
    for (size_t t = 0; t < sys.kinds.size(); t++) {
        out[t].kind = sys.kinds[t];
        out[t].ncols = table_num_columns(sys.kinds[t]);
        out[t].lookups = table_lookups(sys.kinds[t]);
        for (const Lookup& l : out[t].lookups) out[t].num_lookup_cols += l.num_helper_columns(CONSTRAINT_DEGREE) * num_challenges;
    }
    for (const CrossTableLookup& ctl : sys.ctls) {
        for (const TableWithColumns& lt : ctl.looking_tables)
            if (lt.columns.size() != ctl.looked_table.columns.size()) throw std::runtime_error("CTL column count mismatch");
        for (int ch = 0; ch < num_challenges; ch++) {
            // looking tables grouped by consecutive equal table (itertools group_by, :808)
            std::vector<bool> seen(sys.kinds.size(), false);
            size_t i = 0;
            while (i < ctl.looking_tables.size()) {
                int t = ctl.looking_tables[i].table;
                if (t < 0 || t >= (int)sys.kinds.size()) throw std::runtime_error("CTL references a table outside the system");
                if (seen[t]) throw std::runtime_error("CTL lists a looking table in two separate groups (unsupported)");
                seen[t] = true;
                CtlZInfo z;
                z.challenge = ch;
                while (i < ctl.looking_tables.size() && ctl.looking_tables[i].table == t) z.parts.push_back(ctl.looking_tables[i++]);
                z.num_helpers = z.parts.size() > 1 ? (int)(z.parts.size() + CONSTRAINT_DEGREE - 2) / (CONSTRAINT_DEGREE - 1) : 0;
                out[t].num_ctl_helpers += z.num_helpers;
                out[t].zs.push_back(std::move(z));
            }
            CtlZInfo z;
            z.challenge = ch;
            z.parts.push_back(ctl.looked_table);
            int t = ctl.looked_table.table;
            if (t < 0 || t >= (int)sys.kinds.size()) throw std::runtime_error("CTL looked table outside the system");
            out[t].zs.push_back(std::move(z));
        }
    }
    for (size_t t = 0; t < out.size(); t++)
        if (out[t].num_aux() == 0) throw std::runtime_error(std::string("No CTL? table ") + table_name(out[t].kind));   // prover.rs:509
    return out;
}

}  // namespace tables
}  // namespace zkm
