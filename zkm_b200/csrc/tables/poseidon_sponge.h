// PoseidonSponge table (110 columns).  Column map: reference prover/src/poseidon_sponge/columns.rs:19-68
// (#[repr(C)]); constraints: poseidon_sponge_stark.rs:374-478; CTL selectors :25-128.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace poseidon_sponge {

constexpr int SPONGE_RATE = 8, SPONGE_CAPACITY = 4, POSEIDON_DIGEST = 4, POSEIDON_RATE_BYTES = 32, WIDTH_MINUS_DIGEST = 8;
constexpr int IS_FULL_INPUT_BLOCK = 0, CONTEXT = 1, SEGMENT = 2, VIRT = 3, TIMESTAMP = VIRT + SPONGE_RATE, LEN = TIMESTAMP + 1,
              ALREADY_ABSORBED_BYTES = LEN + 1, IS_FINAL_INPUT_LEN = ALREADY_ABSORBED_BYTES + 1,
              ORIGINAL_RATE = IS_FINAL_INPUT_LEN + POSEIDON_RATE_BYTES, ORIGINAL_CAPACITY = ORIGINAL_RATE + SPONGE_RATE,
              BLOCK_BYTES = ORIGINAL_CAPACITY + SPONGE_CAPACITY, NEW_RATE = BLOCK_BYTES + POSEIDON_RATE_BYTES,
              PARTIAL_UPDATED_STATE = NEW_RATE + SPONGE_RATE, UPDATED_DIGEST_STATE = PARTIAL_UPDATED_STATE + WIDTH_MINUS_DIGEST,
              NUM_COLUMNS = UPDATED_DIGEST_STATE + POSEIDON_DIGEST;
static_assert(NUM_COLUMNS == 110, "poseidon sponge layout");

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    const P is_full_input_block = lv[IS_FULL_INPUT_BLOCK];
    yc.constraint(is_full_input_block * (is_full_input_block - P(1)));
    P is_final_block = P(0);
    for (int i = 0; i < POSEIDON_RATE_BYTES; i++) is_final_block = is_final_block + lv[IS_FINAL_INPUT_LEN + i];
    yc.constraint(is_final_block * (is_final_block - P(1)));
    for (int i = 0; i < POSEIDON_RATE_BYTES; i++) { P f = lv[IS_FINAL_INPUT_LEN + i]; yc.constraint(f * (f - P(1))); }
    yc.constraint(is_final_block * is_full_input_block);
    const P already_absorbed_bytes = lv[ALREADY_ABSORBED_BYTES];
    yc.constraint_first_row(already_absorbed_bytes);
    for (int i = 0; i < SPONGE_RATE; i++) yc.constraint_first_row(lv[ORIGINAL_RATE + i]);
    for (int i = 0; i < SPONGE_CAPACITY; i++) yc.constraint_first_row(lv[ORIGINAL_CAPACITY + i]);
    yc.constraint_transition(is_final_block * nv[ALREADY_ABSORBED_BYTES]);
    for (int i = 0; i < SPONGE_RATE; i++) yc.constraint_transition(is_final_block * nv[ORIGINAL_RATE + i]);
    for (int i = 0; i < SPONGE_CAPACITY; i++) yc.constraint_transition(is_final_block * nv[ORIGINAL_CAPACITY + i]);
    yc.constraint_transition(is_full_input_block * (lv[CONTEXT] - nv[CONTEXT]));
    yc.constraint_transition(is_full_input_block * (lv[SEGMENT] - nv[SEGMENT]));
    yc.constraint_transition(is_full_input_block * (lv[TIMESTAMP] - nv[TIMESTAMP]));
    for (int i = 0; i < POSEIDON_DIGEST; i++)
        yc.constraint_transition(is_full_input_block * (nv[ORIGINAL_RATE + i] - lv[UPDATED_DIGEST_STATE + i]));
    for (int i = 0; i < SPONGE_RATE - POSEIDON_DIGEST; i++)
        yc.constraint_transition(is_full_input_block * (nv[ORIGINAL_RATE + POSEIDON_DIGEST + i] - lv[PARTIAL_UPDATED_STATE + i]));
    for (int i = 0; i < SPONGE_CAPACITY; i++)
        yc.constraint_transition(is_full_input_block * (nv[ORIGINAL_CAPACITY + i] - lv[PARTIAL_UPDATED_STATE + SPONGE_RATE - POSEIDON_DIGEST + i]));
    yc.constraint_transition(is_full_input_block * (already_absorbed_bytes + P(POSEIDON_RATE_BYTES) - nv[ALREADY_ABSORBED_BYTES]));
    const P is_dummy = P(1) - is_full_input_block - is_final_block;
    P next_is_final_block = P(0);
    for (int i = 0; i < POSEIDON_RATE_BYTES; i++) next_is_final_block = next_is_final_block + nv[IS_FINAL_INPUT_LEN + i];
    yc.constraint_transition(is_dummy * (nv[IS_FULL_INPUT_BLOCK] + next_is_final_block));
    const P offset = lv[LEN] - already_absorbed_bytes;
    for (int i = 0; i < POSEIDON_RATE_BYTES; i++) yc.constraint(lv[IS_FINAL_INPUT_LEN + i] * (offset - P((uint64_t)i)));
}

inline std::vector<Column> ctl_looked_data() {
    std::vector<int> c = {CONTEXT, SEGMENT, VIRT, LEN, TIMESTAMP};
    for (int i = 0; i < POSEIDON_DIGEST; i++) c.push_back(UPDATED_DIGEST_STATE + i);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looking_poseidon_inputs() {
    std::vector<int> c = range(NEW_RATE, NEW_RATE + SPONGE_RATE);
    for (int i = 0; i < SPONGE_CAPACITY; i++) c.push_back(ORIGINAL_CAPACITY + i);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looking_poseidon_outputs() {
    std::vector<int> c = range(UPDATED_DIGEST_STATE, UPDATED_DIGEST_STATE + POSEIDON_DIGEST);
    for (int i = 0; i < WIDTH_MINUS_DIGEST; i++) c.push_back(PARTIAL_UPDATED_STATE + i);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looking_memory(int i) {
    std::vector<Column> res = {Column::constant_(1), Column::single(CONTEXT), Column::single(SEGMENT), Column::single(VIRT + i / 4)};
    int start = (i / 4) * 4;
    res.push_back(Column::le_bytes({BLOCK_BYTES + start + 3, BLOCK_BYTES + start + 2, BLOCK_BYTES + start + 1, BLOCK_BYTES + start}));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline Filter ctl_looked_filter() { return Filter::new_simple(Column::sum(range(IS_FINAL_INPUT_LEN, IS_FINAL_INPUT_LEN + POSEIDON_RATE_BYTES))); }
inline Filter ctl_looking_memory_filter(int i) {
    if (i == POSEIDON_RATE_BYTES - 1) return Filter::new_simple(Column::single(IS_FULL_INPUT_BLOCK));
    std::vector<int> c = {IS_FULL_INPUT_BLOCK};
    for (int k = i + 1; k < POSEIDON_RATE_BYTES; k++) c.push_back(IS_FINAL_INPUT_LEN + k);
    return Filter::new_simple(Column::sum(c));
}
inline Filter ctl_looking_poseidon_filter() {
    std::vector<int> c = {IS_FULL_INPUT_BLOCK};
    for (int k = 0; k < POSEIDON_RATE_BYTES; k++) c.push_back(IS_FINAL_INPUT_LEN + k);
    return Filter::new_simple(Column::sum(c));
}

}  // namespace poseidon_sponge
}  // namespace tables
}  // namespace zkm
