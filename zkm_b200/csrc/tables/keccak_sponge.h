// KeccakSponge table (470 columns).  Column map: reference prover/src/keccak_sponge/columns.rs:19-70
// (#[repr(C)]); constraints: keccak_sponge_stark.rs:456-567; CTL selectors :29-200.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace keccak_sponge {

constexpr int KECCAK_RATE_BYTES = 136, KECCAK_RATE_U32S = 34, KECCAK_CAPACITY_U32S = 16, KECCAK_DIGEST_BYTES = 32, KECCAK_DIGEST_U32S = 8,
              KECCAK_WIDTH_MINUS_DIGEST_U32S = 42, U8S_PER_CTL = 4;
constexpr int IS_FULL_INPUT_BLOCK = 0, CONTEXT = 1, SEGMENT = 2, VIRT = 3, TIMESTAMP = VIRT + KECCAK_RATE_U32S, LEN = TIMESTAMP + 1,
              ALREADY_ABSORBED_BYTES = LEN + 1, IS_FINAL_INPUT_LEN = ALREADY_ABSORBED_BYTES + 1,
              ORIGINAL_RATE_U32S = IS_FINAL_INPUT_LEN + KECCAK_RATE_BYTES, ORIGINAL_CAPACITY_U32S = ORIGINAL_RATE_U32S + KECCAK_RATE_U32S,
              BLOCK_BYTES = ORIGINAL_CAPACITY_U32S + KECCAK_CAPACITY_U32S, XORED_RATE_U32S = BLOCK_BYTES + KECCAK_RATE_BYTES,
              PARTIAL_UPDATED_STATE_U32S = XORED_RATE_U32S + KECCAK_RATE_U32S,
              UPDATED_DIGEST_STATE_BYTES = PARTIAL_UPDATED_STATE_U32S + KECCAK_WIDTH_MINUS_DIGEST_U32S,
              NUM_COLUMNS = UPDATED_DIGEST_STATE_BYTES + KECCAK_DIGEST_BYTES;
static_assert(NUM_COLUMNS == 470, "keccak sponge layout");

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    const P is_full_input_block = lv[IS_FULL_INPUT_BLOCK];
    yc.constraint(is_full_input_block * (is_full_input_block - P(1)));
    P is_final_block = P(0);
    for (int i = 0; i < KECCAK_RATE_BYTES; i++) is_final_block = is_final_block + lv[IS_FINAL_INPUT_LEN + i];
    yc.constraint(is_final_block * (is_final_block - P(1)));
    for (int i = 0; i < KECCAK_RATE_BYTES; i++) { P f = lv[IS_FINAL_INPUT_LEN + i]; yc.constraint(f * (f - P(1))); }
    yc.constraint(is_final_block * is_full_input_block);
    const P already_absorbed_bytes = lv[ALREADY_ABSORBED_BYTES];
    yc.constraint_first_row(already_absorbed_bytes);
    for (int i = 0; i < KECCAK_RATE_U32S; i++) yc.constraint_first_row(lv[ORIGINAL_RATE_U32S + i]);
    for (int i = 0; i < KECCAK_CAPACITY_U32S; i++) yc.constraint_first_row(lv[ORIGINAL_CAPACITY_U32S + i]);
    yc.constraint_transition(is_final_block * nv[ALREADY_ABSORBED_BYTES]);
    for (int i = 0; i < KECCAK_RATE_U32S; i++) yc.constraint_transition(is_final_block * nv[ORIGINAL_RATE_U32S + i]);
    for (int i = 0; i < KECCAK_CAPACITY_U32S; i++) yc.constraint_transition(is_final_block * nv[ORIGINAL_CAPACITY_U32S + i]);
    yc.constraint_transition(is_full_input_block * (lv[CONTEXT] - nv[CONTEXT]));
    yc.constraint_transition(is_full_input_block * (lv[SEGMENT] - nv[SEGMENT]));
    yc.constraint_transition(is_full_input_block * (lv[TIMESTAMP] - nv[TIMESTAMP]));
    for (int k = 0; k < KECCAK_DIGEST_U32S; k++) {
        P current_after = lv[UPDATED_DIGEST_STATE_BYTES + 4 * k];
        for (int i = 1; i < 4; i++) current_after = current_after + lv[UPDATED_DIGEST_STATE_BYTES + 4 * k + i] * P((uint64_t)1 << (8 * i));
        yc.constraint_transition(is_full_input_block * (nv[ORIGINAL_RATE_U32S + k] - current_after));
    }
    for (int i = 0; i < KECCAK_RATE_U32S - KECCAK_DIGEST_U32S; i++)
        yc.constraint_transition(is_full_input_block * (nv[ORIGINAL_RATE_U32S + KECCAK_DIGEST_U32S + i] - lv[PARTIAL_UPDATED_STATE_U32S + i]));
    for (int i = 0; i < KECCAK_CAPACITY_U32S; i++)
        yc.constraint_transition(is_full_input_block *
                                 (nv[ORIGINAL_CAPACITY_U32S + i] - lv[PARTIAL_UPDATED_STATE_U32S + KECCAK_RATE_U32S - KECCAK_DIGEST_U32S + i]));
    yc.constraint_transition(is_full_input_block * (already_absorbed_bytes + P(KECCAK_RATE_BYTES) - nv[ALREADY_ABSORBED_BYTES]));
    const P is_dummy = P(1) - is_full_input_block - is_final_block;
    P next_is_final_block = P(0);
    for (int i = 0; i < KECCAK_RATE_BYTES; i++) next_is_final_block = next_is_final_block + nv[IS_FINAL_INPUT_LEN + i];
    yc.constraint_transition(is_dummy * (nv[IS_FULL_INPUT_BLOCK] + next_is_final_block));
    const P offset = lv[LEN] - already_absorbed_bytes;
    for (int i = 0; i < KECCAK_RATE_BYTES; i++) yc.constraint(lv[IS_FINAL_INPUT_LEN + i] * (offset - P((uint64_t)i)));
}

inline std::vector<Column> ctl_looked_data() {
    std::vector<Column> res = Column::singles({CONTEXT, SEGMENT, VIRT, LEN, TIMESTAMP});
    for (int i = 7; i >= 0; i--) {
        std::vector<std::pair<int, u64>> lc;
        for (int j = 0; j < 4; j++) lc.push_back({UPDATED_DIGEST_STATE_BYTES + i * 4 + j, (u64)1 << (24 - 8 * j)});
        res.push_back(Column::linear_combination(lc));
    }
    return res;
}
inline std::vector<Column> ctl_looking_keccak_inputs() {
    std::vector<int> c = range(XORED_RATE_U32S, XORED_RATE_U32S + KECCAK_RATE_U32S);
    for (int i = 0; i < KECCAK_CAPACITY_U32S; i++) c.push_back(ORIGINAL_CAPACITY_U32S + i);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looking_keccak_outputs() {
    std::vector<Column> res;
    for (int k = 0; k < KECCAK_DIGEST_U32S; k++) {
        std::vector<std::pair<int, u64>> lc;
        for (int i = 0; i < 4; i++) lc.push_back({UPDATED_DIGEST_STATE_BYTES + 4 * k + i, (u64)1 << (8 * i)});
        res.push_back(Column::linear_combination(lc));
    }
    for (int i = 0; i < KECCAK_WIDTH_MINUS_DIGEST_U32S; i++) res.push_back(Column::single(PARTIAL_UPDATED_STATE_U32S + i));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline std::vector<Column> ctl_looking_memory(int i) {
    std::vector<Column> res = {Column::constant_(1), Column::single(CONTEXT), Column::single(SEGMENT), Column::single(VIRT + i / 4)};
    int start = (i / 4) * 4;
    res.push_back(Column::le_bytes({BLOCK_BYTES + start + 3, BLOCK_BYTES + start + 2, BLOCK_BYTES + start + 1, BLOCK_BYTES + start}));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline int num_logic_ctls() { return (KECCAK_RATE_BYTES + U8S_PER_CTL - 1) / U8S_PER_CTL; }
inline std::vector<Column> ctl_looking_logic(int i) {
    std::vector<Column> res = {Column::constant_(0b100110 * (1 << 6))};
    res.push_back(Column::single(ORIGINAL_RATE_U32S + i));
    res.push_back(Column::le_bytes(range(BLOCK_BYTES + i * U8S_PER_CTL, BLOCK_BYTES + i * U8S_PER_CTL + 4)));
    res.push_back(Column::single(XORED_RATE_U32S + i));
    return res;
}
inline Filter ctl_looked_filter() { return Filter::new_simple(Column::sum(range(IS_FINAL_INPUT_LEN, IS_FINAL_INPUT_LEN + KECCAK_RATE_BYTES))); }
inline Filter ctl_looking_memory_filter(int i) {
    if (i == KECCAK_RATE_BYTES - 1) return Filter::new_simple(Column::single(IS_FULL_INPUT_BLOCK));
    std::vector<int> c = {IS_FULL_INPUT_BLOCK};
    for (int k = i + 1; k < KECCAK_RATE_BYTES; k++) c.push_back(IS_FINAL_INPUT_LEN + k);
    return Filter::new_simple(Column::sum(c));
}
inline Filter ctl_looking_logic_filter() {
    std::vector<int> c = {IS_FULL_INPUT_BLOCK};
    for (int k = 0; k < KECCAK_RATE_BYTES; k++) c.push_back(IS_FINAL_INPUT_LEN + k);
    return Filter::new_simple(Column::sum(c));
}
inline Filter ctl_looking_keccak_filter() { return ctl_looking_logic_filter(); }

}  // namespace keccak_sponge
}  // namespace tables
}  // namespace zkm
