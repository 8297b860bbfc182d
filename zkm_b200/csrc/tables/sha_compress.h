// ShaCompress table (224 columns): one SHA-256 compression round per row, 65 rows per block.
// Column map: reference prover/src/sha_compress/columns.rs:9-55 with the gadget structs
// not_operation.rs:8-10, wrapping_add_2.rs:8-11, wrapping_add_5.rs:8-11 and sha_extend/rotate_right.rs:8-12
// (fields in declaration order); constraints: sha_compress_stark.rs:399-606 (gadgets: not_operation.rs:
// 25-38, wrapping_add_2.rs:44-78, wrapping_add_5.rs:53-98, sha_compress/logic.rs:7-33); round constants
// sha_compress_sponge/constants.rs:1-10 (K) / :12-77 (their LE bytes); CTL selectors :32-252.
#pragma once
#include "hd.h"
#include "dsl.h"
#include "sha_extend.h"

namespace zkm {
namespace tables {
namespace sha_compress {

constexpr int NUM_COMPRESS_ROWS = 65;
constexpr int STATE = 0, E_NOT = 32, W_I = 36, K_I = 40, S_1_INTER = 44, S_1 = 48, E_AND_F = 52, E_NOT_AND_G = 56, CH = 60, S_0_INTER = 64,
              S_0 = 68, A_AND_B = 72, A_AND_C = 76, B_AND_C = 80, MAJ_INTER = 84, MAJ = 88, E_RR_6 = 92, E_RR_11 = 98, E_RR_25 = 104,
              A_RR_2 = 110, A_RR_13 = 116, A_RR_22 = 122, TEMP2 = 128, D_ADD_TEMP1 = 134, TEMP1_ADD_TEMP2 = 140, TIMESTAMP = 146,
              SEGMENT = 147, CONTEXT = 148, W_I_VIRT = 149, TEMP1 = 150, ROUND = 159, NUM_COLUMNS = ROUND + NUM_COMPRESS_ROWS;
static_assert(NUM_COLUMNS == 224, "sha compress layout");
constexpr int ADD_VALUE = 0, ADD_CARRY = 4;       // WrappingAddNOp field offsets
ZKM_HD constexpr int state4(int i) { return STATE + 4 * i; }

ZKM_DEF_CONST(SHA_K, 64, {0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2})

// wrapping_add_2.rs:44-78; every constraint is scaled by `scale` by the caller in the reference
template <class P, class V, class YC>
ZKM_HD void wrapping_add_2(const V& lv, int a, int b, int cols, P scale, YC& yc) {
    P result = sha_extend::from_bytes<P>(lv, cols + ADD_VALUE);
    for (int i = 0; i < 2; i++) { P cy = lv[cols + ADD_CARRY + i]; yc.constraint(scale * (cy * (P(1) - cy))); }
    yc.constraint(scale * (lv[cols + ADD_CARRY] + lv[cols + ADD_CARRY + 1] - P(1)));
    P carry = lv[cols + ADD_CARRY + 1];
    P overflowed = (lv[a] + lv[b]) + (lv[a + 1] + lv[b + 1]) * P(1u << 8) + (lv[a + 2] + lv[b + 2]) * P(1u << 16) +
                   (lv[a + 3] + lv[b + 3]) * P(1u << 24);
    yc.constraint(scale * (overflowed - carry * P((uint64_t)1 << 32) - result));
}

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    const P is_final = lv[ROUND + NUM_COMPRESS_ROWS - 1];
    yc.constraint(is_final * (is_final - P(1)));
    const P not_final = P(1) - is_final;
    P sum_round_flags = P(0);
    for (int i = 0; i < NUM_COMPRESS_ROWS; i++) sum_round_flags = sum_round_flags + lv[ROUND + i];
    yc.constraint(sum_round_flags * (sum_round_flags - P(1)));
    for (int i = 0; i < 4; i++) {
        P bit_i = P(0);
        for (int j = 0; j < 64; j++) bit_i = bit_i + lv[ROUND + j] * P((ZKM_K(SHA_K)[j] >> (8 * i)) & 0xff);
        yc.constraint(sum_round_flags * not_final * (lv[K_I + i] - bit_i));
    }
    sha_extend::rotate_right<P, V, YC>(lv, state4(4), E_RR_6, 6, yc);
    sha_extend::rotate_right<P, V, YC>(lv, state4(4), E_RR_11, 11, yc);
    sha_extend::rotate_right<P, V, YC>(lv, state4(4), E_RR_25, 25, yc);
    sha_extend::rotate_right<P, V, YC>(lv, state4(0), A_RR_2, 2, yc);
    sha_extend::rotate_right<P, V, YC>(lv, state4(0), A_RR_13, 13, yc);
    sha_extend::rotate_right<P, V, YC>(lv, state4(0), A_RR_22, 22, yc);
    // not_operation(e) (not_operation.rs:25-38)
    for (int i = 0; i < 4; i++) yc.constraint(sum_round_flags * (lv[state4(4) + i] + lv[E_NOT + i] - P(255)));
    {   // wrapping_add_5(h, s_1, ch, k_i, w_i) -> temp1 (wrapping_add_5.rs:53-98)
        const int a = state4(7), b = S_1, c = CH, d = K_I, e = W_I;
        P result = sha_extend::from_bytes<P>(lv, TEMP1 + ADD_VALUE);
        for (int i = 0; i < 5; i++) { P cy = lv[TEMP1 + ADD_CARRY + i]; yc.constraint(sum_round_flags * (cy * (P(1) - cy))); }
        yc.constraint(sum_round_flags * (lv[TEMP1 + ADD_CARRY] + lv[TEMP1 + ADD_CARRY + 1] + lv[TEMP1 + ADD_CARRY + 2] +
                                         lv[TEMP1 + ADD_CARRY + 3] + lv[TEMP1 + ADD_CARRY + 4] - P(1)));
        P carry = lv[TEMP1 + ADD_CARRY + 1] * P(1) + lv[TEMP1 + ADD_CARRY + 2] * P(2) + lv[TEMP1 + ADD_CARRY + 3] * P(3) +
                  lv[TEMP1 + ADD_CARRY + 4] * P(4);
        P overflowed = (lv[a] + lv[b] + lv[c] + lv[d] + lv[e]) + (lv[a + 1] + lv[b + 1] + lv[c + 1] + lv[d + 1] + lv[e + 1]) * P(1u << 8) +
                       (lv[a + 2] + lv[b + 2] + lv[c + 2] + lv[d + 2] + lv[e + 2]) * P(1u << 16) +
                       (lv[a + 3] + lv[b + 3] + lv[c + 3] + lv[d + 3] + lv[e + 3]) * P(1u << 24);
        yc.constraint(sum_round_flags * (overflowed - carry * P((uint64_t)1 << 32) - result));
    }
    wrapping_add_2<P, V, YC>(lv, S_0, MAJ, TEMP2, sum_round_flags, yc);
    wrapping_add_2<P, V, YC>(lv, state4(3), TEMP1 + ADD_VALUE, D_ADD_TEMP1, sum_round_flags, yc);
    wrapping_add_2<P, V, YC>(lv, TEMP1 + ADD_VALUE, TEMP2 + ADD_VALUE, TEMP1_ADD_TEMP2, sum_round_flags, yc);
    yc.constraint(sum_round_flags * not_final * (nv[TIMESTAMP] - lv[TIMESTAMP]));
    yc.constraint(sum_round_flags * not_final * (nv[W_I_VIRT] - lv[W_I_VIRT] - P(4)));
    // state rotation a..h (logic.rs:7-16 equal_packed_constraint)
    const int src[8] = {TEMP1_ADD_TEMP2 + ADD_VALUE, state4(0), state4(1), state4(2), D_ADD_TEMP1 + ADD_VALUE, state4(4), state4(5), state4(6)};
    for (int k = 0; k < 8; k++)
        for (int i = 0; i < 4; i++) yc.constraint(sum_round_flags * not_final * (lv[src[k] + i] - nv[state4(k) + i]));
}

inline std::vector<Column> ctl_data_inputs() {
    std::vector<int> c = range(STATE, STATE + 32);
    for (int x : {TIMESTAMP, SEGMENT, CONTEXT, W_I_VIRT}) c.push_back(x);
    return Column::singles(c);
}
inline std::vector<Column> ctl_data_outputs() {
    std::vector<int> c = range(STATE, STATE + 32);
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> logic_row(u64 op, int in0, int in1, int out) {
    return {Column::constant_(op), Column::le_bytes(range(in0, in0 + 4)), Column::le_bytes(range(in1, in1 + 4)), Column::le_bytes(range(out, out + 4))};
}
constexpr u64 OP_XOR = 0b100110 * (1 << 6), OP_AND = 0b100100 * (1 << 6);
inline std::vector<Column> ctl_s_1_inter_looking_logic() { return logic_row(OP_XOR, E_RR_6, E_RR_11, S_1_INTER); }
inline std::vector<Column> ctl_s_1_looking_logic() { return logic_row(OP_XOR, S_1_INTER, E_RR_25, S_1); }
inline std::vector<Column> ctl_e_and_f_looking_logic() { return logic_row(OP_AND, state4(4), state4(5), E_AND_F); }
inline std::vector<Column> ctl_not_e_and_g_looking_logic() { return logic_row(OP_AND, E_NOT, state4(6), E_NOT_AND_G); }
inline std::vector<Column> ctl_ch_looking_logic() { return logic_row(OP_XOR, E_AND_F, E_NOT_AND_G, CH); }
inline std::vector<Column> ctl_s_0_inter_looking_logic() { return logic_row(OP_XOR, A_RR_2, A_RR_13, S_0_INTER); }
inline std::vector<Column> ctl_s_0_looking_logic() { return logic_row(OP_XOR, S_0_INTER, A_RR_22, S_0); }
inline std::vector<Column> ctl_a_and_b_looking_logic() { return logic_row(OP_AND, state4(0), state4(1), A_AND_B); }
inline std::vector<Column> ctl_a_and_c_looking_logic() { return logic_row(OP_AND, state4(0), state4(2), A_AND_C); }
inline std::vector<Column> ctl_b_and_c_looking_logic() { return logic_row(OP_AND, state4(1), state4(2), B_AND_C); }
inline std::vector<Column> ctl_maj_inter_looking_logic() { return logic_row(OP_XOR, A_AND_B, A_AND_C, MAJ_INTER); }
inline std::vector<Column> ctl_maj_looking_logic() { return logic_row(OP_XOR, MAJ_INTER, B_AND_C, MAJ); }
inline std::vector<Column> ctl_looking_memory(int) {
    return {Column::constant_(1), Column::single(CONTEXT), Column::single(SEGMENT), Column::single(W_I_VIRT), Column::le_bytes(range(W_I, W_I + 4)),
            Column::single(TIMESTAMP)};
}
inline Filter ctl_filter_inputs() { return Filter::new_simple(Column::single(ROUND)); }
inline Filter ctl_filter_outputs() { return Filter::new_simple(Column::single(ROUND + NUM_COMPRESS_ROWS - 1)); }
inline Filter ctl_logic_filter() { return Filter::new_simple(Column::sum(range(ROUND, ROUND + NUM_COMPRESS_ROWS - 1))); }

}  // namespace sha_compress
}  // namespace tables
}  // namespace zkm
