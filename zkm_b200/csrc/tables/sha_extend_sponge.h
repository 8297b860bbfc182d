// ShaExtendSponge table (76 columns).  Column map: reference prover/src/sha_extend_sponge/columns.rs:7-33
// (fields in declaration order); constraints: sha_extend_sponge_stark.rs:229-327 (note: plain
// `constraint`, not `constraint_transition`, on the next-row relations, as in the reference);
// CTL selectors :28-92.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace sha_extend_sponge {

constexpr int NUM_ROUNDS = 48, NUM_EXTEND_INPUT = 4, SHA_EXTEND_SPONGE_READ_BYTES = 16, NUM_CHANNELS = 10;
constexpr int ROUND = 0, W_I_MINUS_15 = 48, W_I_MINUS_2 = 52, W_I_MINUS_16 = 56, W_I_MINUS_7 = 60, W_I = 64, INPUT_VIRT = 68, OUTPUT_VIRT = 72,
              CONTEXT = 73, SEGMENT = 74, TIMESTAMP = 75, NUM_COLUMNS = 76;

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    for (int i = 0; i < NUM_ROUNDS; i++) { P r = lv[ROUND + i]; yc.constraint(r * (r - P(1))); }
    const P is_final = lv[ROUND + NUM_ROUNDS - 1];
    yc.constraint(is_final * (is_final - P(1)));
    const P not_final = P(1) - is_final;
    P sum_round_flags = P(0), local_round_index = P(0), next_round_index = P(0);
    for (int i = 0; i < NUM_ROUNDS; i++) {
        sum_round_flags = sum_round_flags + lv[ROUND + i];
        local_round_index = local_round_index + lv[ROUND + i] * P((uint64_t)i);
        next_round_index = next_round_index + nv[ROUND + i] * P((uint64_t)i);
    }
    yc.constraint(sum_round_flags * not_final * (nv[TIMESTAMP] - lv[TIMESTAMP] - P(2 * NUM_CHANNELS)));
    yc.constraint(sum_round_flags * not_final * (next_round_index - local_round_index - P(1)));
    for (int i = 0; i < NUM_EXTEND_INPUT; i++)
        yc.constraint(sum_round_flags * not_final * (nv[INPUT_VIRT + i] - lv[INPUT_VIRT + i] - P(4)));
    yc.constraint(sum_round_flags * not_final * (nv[OUTPUT_VIRT] - lv[OUTPUT_VIRT] - P(4)));
    yc.constraint(sum_round_flags * (lv[INPUT_VIRT + 0] - lv[INPUT_VIRT + 2] - P(4)));
    yc.constraint(sum_round_flags * (lv[INPUT_VIRT + 1] - lv[INPUT_VIRT + 2] - P(56)));
    yc.constraint(sum_round_flags * (lv[INPUT_VIRT + 3] - lv[INPUT_VIRT + 2] - P(36)));
    yc.constraint(sum_round_flags * (lv[OUTPUT_VIRT] - lv[INPUT_VIRT + 2] - P(64)));
}

inline std::vector<Column> ctl_looking_sha_extend_inputs() {
    std::vector<int> c;
    for (int s : {W_I_MINUS_15, W_I_MINUS_2, W_I_MINUS_16, W_I_MINUS_7}) for (int i = 0; i < 4; i++) c.push_back(s + i);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looking_sha_extend_outputs() {
    std::vector<int> c = range(W_I, W_I + 4);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looked_data() {
    std::vector<Column> res = Column::singles({CONTEXT, SEGMENT, OUTPUT_VIRT, TIMESTAMP});
    res.push_back(Column::le_bytes(range(W_I, W_I + 4)));
    return res;
}
inline std::vector<Column> ctl_looking_memory(int i) {
    std::vector<Column> res = {Column::constant_(1), Column::single(CONTEXT), Column::single(SEGMENT), Column::single(INPUT_VIRT + i / 4)};
    const int src[4] = {W_I_MINUS_15, W_I_MINUS_2, W_I_MINUS_16, W_I_MINUS_7};
    int start = i / 4;
    res.push_back(Column::le_bytes(range(src[start > 3 ? 3 : start], src[start > 3 ? 3 : start] + 4)));
    res.push_back(Column::single(TIMESTAMP));
    return res;
}
inline Filter ctl_looking_sha_extend_filter() { return Filter::new_simple(Column::sum(range(ROUND, ROUND + NUM_ROUNDS))); }

}  // namespace sha_extend_sponge
}  // namespace tables
}  // namespace zkm
