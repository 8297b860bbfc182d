// ShaCompressSponge table (127 columns).  Column map: reference prover/src/sha_compress_sponge/columns.rs:
// 6-25 (fields in declaration order; output_hx is 8 WrappingAdd2Op = value[4] + carry[2]); constraints:
// sha_compress_sponge_stark.rs:241-280; CTL selectors :31-89.
#pragma once
#include "hd.h"
#include "dsl.h"
#include "sha_compress.h"

namespace zkm {
namespace tables {
namespace sha_compress_sponge {

constexpr int SHA_COMPRESS_SPONGE_READ_BYTES = 32;
constexpr int HX = 0, OUTPUT_STATE = 32, OUTPUT_HX = 64, HX_VIRT = 112, W_START_VIRT = 120, TIMESTAMP = 121, CONTEXT = 122, SEGMENT = 123,
              W_START_SEGMENT = 124, W_START_CONTEXT = 125, IS_REAL_ROUND = 126, NUM_COLUMNS = 127;

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& /*nv*/, YC& yc) {
    const P is_normal_round = lv[IS_REAL_ROUND];
    yc.constraint(is_normal_round * (is_normal_round - P(1)));
    for (int i = 0; i < 7; i++) yc.constraint(is_normal_round * (lv[HX_VIRT + i + 1] - lv[HX_VIRT + i] - P(4)));
    for (int i = 0; i < 8; i++) {
        // wrapping_add_2(hx[i], output_state[i]) -> output_hx[i], each constraint times is_real_round (c * filter)
        const int a = HX + 4 * i, b = OUTPUT_STATE + 4 * i, cols = OUTPUT_HX + 6 * i;
        P result = sha_extend::from_bytes<P>(lv, cols);
        for (int k = 0; k < 2; k++) { P cy = lv[cols + 4 + k]; yc.constraint(cy * (P(1) - cy) * is_normal_round); }
        yc.constraint((lv[cols + 4] + lv[cols + 5] - P(1)) * is_normal_round);
        P carry = lv[cols + 5];
        P overflowed = (lv[a] + lv[b]) + (lv[a + 1] + lv[b + 1]) * P(1u << 8) + (lv[a + 2] + lv[b + 2]) * P(1u << 16) +
                       (lv[a + 3] + lv[b + 3]) * P(1u << 24);
        yc.constraint((overflowed - carry * P((uint64_t)1 << 32) - result) * is_normal_round);
    }
}

inline std::vector<Column> ctl_looking_sha_compress_inputs() {
    std::vector<int> c = range(HX, HX + 32);
    for (int x : {TIMESTAMP, W_START_SEGMENT, W_START_CONTEXT, W_START_VIRT}) c.push_back(x);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looking_sha_compress_outputs() {
    std::vector<int> c = range(OUTPUT_STATE, OUTPUT_STATE + 32);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_looked_data() {
    std::vector<Column> res = Column::singles({CONTEXT, SEGMENT, HX_VIRT, TIMESTAMP});
    for (int i = 0; i < 8; i++) res.push_back(Column::le_bytes(range(OUTPUT_HX + 6 * i, OUTPUT_HX + 6 * i + 4)));
    return res;
}
inline std::vector<Column> ctl_looking_memory(int i) {
    int start = i / 4;
    return {Column::constant_(1), Column::single(CONTEXT), Column::single(SEGMENT), Column::single(HX_VIRT + start),
            Column::le_bytes(range(HX + 4 * start, HX + 4 * start + 4)), Column::single(TIMESTAMP)};
}
inline Filter ctl_looking_sha_compress_filter() { return Filter::new_simple(Column::single(IS_REAL_ROUND)); }
inline Filter ctl_looked_filter() { return Filter::new_simple(Column::single(IS_REAL_ROUND)); }

}  // namespace sha_compress_sponge
}  // namespace tables
}  // namespace zkm
