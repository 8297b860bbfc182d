// ShaExtend table (78 columns): one SHA-256 message-schedule word per row.
// Column map: reference prover/src/sha_extend/columns.rs:8-37 with the gadget structs
// rotate_right.rs:8-12, shift_right.rs:8-12, wrapping_add_4.rs:8-11 (fields in declaration order);
// constraints: sha_extend_stark.rs:246-321 with rotate_right.rs:40-79, shift_right.rs:39-73,
// wrapping_add_4.rs:48-91; CTL selectors sha_extend_stark.rs:33-113.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace sha_extend {

constexpr int W_I_VALUE = 0, W_I_CARRY = 4, W_I_MINUS_15 = 8, W_I_MINUS_2 = 12, W_I_MINUS_16 = 16, W_I_MINUS_7 = 20, S_0_INTER = 24, S_0 = 28,
              S_1_INTER = 32, S_1 = 36, W_I_MINUS_15_RR_7 = 40, W_I_MINUS_15_RR_18 = 46, W_I_MINUS_2_RR_17 = 52, W_I_MINUS_2_RR_19 = 58,
              W_I_MINUS_2_RS_10 = 64, W_I_MINUS_15_RS_3 = 70, TIMESTAMP = 76, IS_REAL_ROUND = 77, NUM_COLUMNS = 78;
constexpr int OP_VALUE = 0, OP_SHIFT = 4, OP_CARRY = 5;      // RotateRightOp / ShiftRightOp field offsets

template <class P, class V>
ZKM_HD P from_bytes(const V& lv, int start) {
    return lv[start] + P(1u << 8) * lv[start + 1] + P(1u << 16) * lv[start + 2] + P(1u << 24) * lv[start + 3];
}
// rotate_right.rs:40-79
template <class P, class V, class YC>
ZKM_HD void rotate_right(const V& lv, int input, int op, int rotation, YC& yc) {
    P rotated = from_bytes<P>(lv, op + OP_VALUE), in = from_bytes<P>(lv, input);
    P carry_multiplier = P((uint64_t)1 << (32 - rotation)), shift_multiplier = P((uint64_t)1 << rotation);
    yc.constraint(rotated - lv[op + OP_CARRY] * carry_multiplier - lv[op + OP_SHIFT]);
    yc.constraint(in - lv[op + OP_SHIFT] * shift_multiplier - lv[op + OP_CARRY]);
}
// shift_right.rs:39-73
template <class P, class V, class YC>
ZKM_HD void shift_right(const V& lv, int input, int op, int rotation, YC& yc) {
    P shifted = from_bytes<P>(lv, op + OP_VALUE), in = from_bytes<P>(lv, input);
    P shift_multiplier = P((uint64_t)1 << rotation);
    yc.constraint(shifted - lv[op + OP_SHIFT]);
    yc.constraint(in - lv[op + OP_SHIFT] * shift_multiplier - lv[op + OP_CARRY]);
}

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& /*nv*/, YC& yc) {
    rotate_right<P, V, YC>(lv, W_I_MINUS_15, W_I_MINUS_15_RR_7, 7, yc);
    rotate_right<P, V, YC>(lv, W_I_MINUS_15, W_I_MINUS_15_RR_18, 18, yc);
    rotate_right<P, V, YC>(lv, W_I_MINUS_2, W_I_MINUS_2_RR_17, 17, yc);
    rotate_right<P, V, YC>(lv, W_I_MINUS_2, W_I_MINUS_2_RR_19, 19, yc);
    shift_right<P, V, YC>(lv, W_I_MINUS_15, W_I_MINUS_15_RS_3, 3, yc);
    shift_right<P, V, YC>(lv, W_I_MINUS_2, W_I_MINUS_2_RS_10, 10, yc);
    // wrapping_add_4(s_1, w_i_minus_7, s_0, w_i_minus_16) -> w_i, every constraint scaled by is_real_round
    const P real = lv[IS_REAL_ROUND];
    const int a = S_1, b = W_I_MINUS_7, c = S_0, d = W_I_MINUS_16;
    P result = from_bytes<P>(lv, W_I_VALUE);
    for (int i = 0; i < 4; i++) { P cy = lv[W_I_CARRY + i]; yc.constraint(cy * (P(1) - cy) * real); }
    yc.constraint((lv[W_I_CARRY] + lv[W_I_CARRY + 1] + lv[W_I_CARRY + 2] + lv[W_I_CARRY + 3] - P(1)) * real);
    P carry = lv[W_I_CARRY + 1] * P(1) + lv[W_I_CARRY + 2] * P(2) + lv[W_I_CARRY + 3] * P(3);
    P overflowed = (lv[a] + lv[b] + lv[c] + lv[d]) + (lv[a + 1] + lv[b + 1] + lv[c + 1] + lv[d + 1]) * P(1u << 8) +
                   (lv[a + 2] + lv[b + 2] + lv[c + 2] + lv[d + 2]) * P(1u << 16) + (lv[a + 3] + lv[b + 3] + lv[c + 3] + lv[d + 3]) * P(1u << 24);
    yc.constraint((overflowed - carry * P((uint64_t)1 << 32) - result) * real);
}

inline std::vector<Column> ctl_data_inputs() {
    std::vector<int> c = range(W_I_MINUS_15, W_I_MINUS_15 + 4);
    for (int s : {W_I_MINUS_2, W_I_MINUS_16, W_I_MINUS_7}) for (int i = 0; i < 4; i++) c.push_back(s + i);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_data_outputs() {
    std::vector<int> c = range(W_I_VALUE, W_I_VALUE + 4);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> xor_logic(int in0, int in1, int out) {
    return {Column::constant_(0b100110 * (1 << 6)), Column::le_bytes(range(in0, in0 + 4)), Column::le_bytes(range(in1, in1 + 4)),
            Column::le_bytes(range(out, out + 4))};
}
inline std::vector<Column> ctl_s_0_inter_looking_logic() { return xor_logic(W_I_MINUS_15_RR_7, W_I_MINUS_15_RR_18, S_0_INTER); }
inline std::vector<Column> ctl_s_0_looking_logic() { return xor_logic(S_0_INTER, W_I_MINUS_15_RS_3, S_0); }
inline std::vector<Column> ctl_s_1_inter_looking_logic() { return xor_logic(W_I_MINUS_2_RR_17, W_I_MINUS_2_RR_19, S_1_INTER); }
inline std::vector<Column> ctl_s_1_looking_logic() { return xor_logic(S_1_INTER, W_I_MINUS_2_RS_10, S_1); }
inline Filter ctl_filter() { return Filter::new_simple(Column::single(IS_REAL_ROUND)); }

}  // namespace sha_extend
}  // namespace tables
}  // namespace zkm
