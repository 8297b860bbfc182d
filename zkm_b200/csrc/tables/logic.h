// Logic table (69 columns): AND / OR / XOR / NOR on bit-decomposed 32-bit inputs.
// Column map: reference prover/src/logic.rs:26-50; constraints: logic.rs:186-240 (64 bit-booleans,
// then one result check); CTL selectors: logic.rs:52-76.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace logic {

constexpr int IS_AND = 0, IS_OR = 1, IS_XOR = 2, IS_NOR = 3;
constexpr int VAL_BITS = 32;
constexpr int INPUT0 = 4, INPUT1 = INPUT0 + VAL_BITS, RESULT = INPUT1 + VAL_BITS;
constexpr int NUM_COLUMNS = RESULT + 1;

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& /*nv*/, YC& yc) {
    const P is_and = lv[IS_AND], is_or = lv[IS_OR], is_xor = lv[IS_XOR], is_nor = lv[IS_NOR];
    const P sum_coeff = is_or + is_xor - is_nor;
    const P and_coeff = is_and - is_or - is_xor * P(2) + is_nor;
    const P not_coeff = is_nor;
    for (int i = INPUT0; i < INPUT0 + 2 * VAL_BITS; i++) {
        P bit = lv[i];
        yc.constraint(bit * (bit - P(1)));
    }
    // sum_i bit_i 2^i by Horner from the top bit (same field elements as the reference's weighted sums)
    P x = P(0), y = P(0), x_land_y = P(0);
    for (int i = VAL_BITS - 1; i >= 0; i--) {
        P xb = lv[INPUT0 + i], yb = lv[INPUT1 + i];
        x = x + x + xb;
        y = y + y + yb;
        x_land_y = x_land_y + x_land_y + xb * yb;
    }
    P x_op_y = sum_coeff * (x + y) + and_coeff * x_land_y + not_coeff * P(0xFFFFFFFFull);
    yc.constraint(lv[RESULT] - x_op_y);
}

inline std::vector<Column> ctl_data() {
    std::vector<Column> res;
    res.push_back(Column::linear_combination({{IS_AND, 0b100100 * (1 << 6)}, {IS_OR, 0b100101 * (1 << 6)},
                                              {IS_XOR, 0b100110 * (1 << 6)}, {IS_NOR, 0b100111 * (1 << 6)}}));
    res.push_back(Column::le_bits(range(INPUT0, INPUT0 + VAL_BITS)));
    res.push_back(Column::le_bits(range(INPUT1, INPUT1 + VAL_BITS)));
    res.push_back(Column::single(RESULT));
    return res;
}
inline Filter ctl_filter() { return Filter::new_simple(Column::sum({IS_AND, IS_OR, IS_XOR, IS_NOR})); }

}  // namespace logic
}  // namespace tables
}  // namespace zkm
