// Table registry: column counts, in-table lookups and the constraint evaluator of every table kind
// (the reference's 12 `Stark` impls held by AllStark, prover/src/all_stark.rs:39-74).
// eval_table<P>() is the single source the CUDA quotient kernels (P = device Goldilocks scalar) and
// the CPU oracle (P = scalar at H / coset points, quadratic extension at zeta) are instantiated from —
// the analogue of `Stark::eval_packed_generic<FE, P, D2>` (stark.rs:41-47).
#pragma once
#include "system.h"
#include "arithmetic.h"
#include "cpu.h"
#include "keccak.h"
#include "keccak_sponge.h"
#include "logic.h"
#include "memory.h"
#include "poseidon.h"
#include "poseidon_sponge.h"
#include "sha_compress.h"
#include "sha_compress_sponge.h"
#include "sha_extend.h"
#include "sha_extend_sponge.h"

namespace zkm {
namespace tables {

inline const char* table_name(int kind) {
    static const char* N[NUM_TABLE_KINDS] = {"Arithmetic", "Cpu", "Poseidon", "PoseidonSponge", "Keccak", "KeccakSponge",
                                             "ShaExtend", "ShaExtendSponge", "ShaCompress", "ShaCompressSponge", "Logic", "Memory"};
    return kind >= 0 && kind < NUM_TABLE_KINDS ? N[kind] : "?";
}

// SURVEY Appendix B / each table's column map.
inline int table_num_columns(int kind) {
    static const int C[NUM_TABLE_KINDS] = {arithmetic::NUM_COLUMNS, cpu::NUM_COLUMNS, poseidon::NUM_COLUMNS, poseidon_sponge::NUM_COLUMNS,
                                           keccak::NUM_COLUMNS, keccak_sponge::NUM_COLUMNS, sha_extend::NUM_COLUMNS,
                                           sha_extend_sponge::NUM_COLUMNS, sha_compress::NUM_COLUMNS, sha_compress_sponge::NUM_COLUMNS,
                                           logic::NUM_COLUMNS, memory::NUM_COLUMNS};
    if (kind < 0 || kind >= NUM_TABLE_KINDS) throw std::runtime_error("bad table kind");
    return C[kind];
}

inline bool table_implemented(int kind) {
    return kind >= 0 && kind < NUM_TABLE_KINDS;
}

inline std::vector<Lookup> table_lookups(int kind) {
    switch (kind) {
        case T_ARITHMETIC: return arithmetic::lookups();
        case T_MEMORY: return memory::lookups();
        default: return {};
    }
}

// Table constraints, emitted in the reference's order.  Returns false for a kind whose constraints
// have not been transcribed yet (callers turn that into an error; nothing is silently skipped).
template <class P, class V, class YC>
ZKM_HD bool eval_table(int kind, const V& lv, const V& nv, YC& yc) {
    switch (kind) {
        case T_ARITHMETIC: arithmetic::eval<P, V, YC>(lv, nv, yc); return true;
        case T_CPU: cpu::eval<P, V, YC>(lv, nv, yc); return true;
        case T_POSEIDON: poseidon::eval<P, V, YC>(lv, nv, yc); return true;
        case T_POSEIDON_SPONGE: poseidon_sponge::eval<P, V, YC>(lv, nv, yc); return true;
        case T_KECCAK: keccak::eval<P, V, YC>(lv, nv, yc); return true;
        case T_KECCAK_SPONGE: keccak_sponge::eval<P, V, YC>(lv, nv, yc); return true;
        case T_SHA_EXTEND: sha_extend::eval<P, V, YC>(lv, nv, yc); return true;
        case T_SHA_EXTEND_SPONGE: sha_extend_sponge::eval<P, V, YC>(lv, nv, yc); return true;
        case T_SHA_COMPRESS: sha_compress::eval<P, V, YC>(lv, nv, yc); return true;
        case T_SHA_COMPRESS_SPONGE: sha_compress_sponge::eval<P, V, YC>(lv, nv, yc); return true;
        case T_LOGIC: logic::eval<P, V, YC>(lv, nv, yc); return true;
        case T_MEMORY: memory::eval<P, V, YC>(lv, nv, yc); return true;
        default: return false;
    }
}

}  // namespace tables
}  // namespace zkm
