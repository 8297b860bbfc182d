// Poseidon syscall table (262 columns): one permutation per row with the S-box witnesses
// (x^3 and x^7 of every S-box input).  Column map: reference prover/src/poseidon/columns.rs:3-54;
// constraints: poseidon/poseidon_stark.rs:554-594 with the layer helpers :166-171 (constants),
// :183-190 (S-box: 2 constraints each), :244-266 (full S-box layer), :300-313 (MDS),
// :380-384,400-412 (fast partial init), :434-446 (partial S-box), :505-520 (fast partial MDS);
// CTL selectors :28-48; trace row generator (tests only) :51-95,126-145.
#pragma once
#include "hd.h"
#include "dsl.h"
#define ZKM_POSEIDON_NO_ARRAYS
#include "../poseidon_consts.h"

namespace zkm {
namespace tables {
namespace poseidon {

ZKM_DEF_CONST(PT_RC, 360, POSEIDON_ALL_ROUND_CONSTANTS_INIT)
ZKM_DEF_CONST(PT_FIRST, 12, POSEIDON_FAST_PARTIAL_FIRST_ROUND_CONSTANT_INIT)
ZKM_DEF_CONST(PT_PRC, 22, POSEIDON_FAST_PARTIAL_ROUND_CONSTANTS_INIT)
ZKM_DEF_CONST(PT_VS, 242, POSEIDON_FAST_PARTIAL_ROUND_VS_INIT)
ZKM_DEF_CONST(PT_WHAT, 242, POSEIDON_FAST_PARTIAL_ROUND_W_HATS_INIT)
ZKM_DEF_CONST(PT_INIT, 121, POSEIDON_FAST_PARTIAL_ROUND_INITIAL_MATRIX_INIT)
ZKM_DEF_CONST(PT_CIRC, 12, POSEIDON_MDS_CIRC_INIT)
ZKM_DEF_CONST(PT_DIAG, 12, POSEIDON_MDS_DIAG_INIT)

constexpr int W = 12, HALF_N_FULL_ROUNDS = 4, N_PARTIAL_ROUNDS = 22;
constexpr int FILTER = 0, START_IN = 1, START_OUT = START_IN + W, TIMESTAMP = START_OUT + W;
constexpr int START_FULL_0 = TIMESTAMP + 1;
constexpr int START_PARTIAL = START_FULL_0 + W * 2 * HALF_N_FULL_ROUNDS;
constexpr int START_FULL_1 = START_PARTIAL + N_PARTIAL_ROUNDS * 2;
constexpr int NUM_COLUMNS = START_FULL_1 + W * 2 * HALF_N_FULL_ROUNDS;
ZKM_HD constexpr int reg_in(int i) { return START_IN + i; }
ZKM_HD constexpr int reg_out(int i) { return START_OUT + i; }
ZKM_HD constexpr int reg_full0_s0(int r, int i) { return START_FULL_0 + W * 2 * r + 2 * i; }
ZKM_HD constexpr int reg_full1_s0(int r, int i) { return START_FULL_1 + W * 2 * r + 2 * i; }
ZKM_HD constexpr int reg_partial_s0(int r) { return START_PARTIAL + 2 * r; }

template <class P, class YC>
ZKM_HD void sbox(const P& input, const P& inter, const P& output, YC& yc) {
    yc.constraint(input * input * input - inter);
    yc.constraint(input * inter * inter - output);
}

template <class P>
ZKM_HD void mds_layer(P* state) {
    P res[W];
    for (int i = 0; i < W; i++) {
        P acc = P(0);
        for (int j = 0; j < W; j++) acc = acc + state[(j + i) % W] * P(ZKM_K(PT_CIRC)[j]);
        acc = acc + state[i] * P(ZKM_K(PT_DIAG)[i]);
        res[i] = acc;
    }
    for (int i = 0; i < W; i++) state[i] = res[i];
}

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& /*nv*/, YC& yc) {
    P state[W];
    for (int i = 0; i < W; i++) state[i] = lv[reg_in(i)];
    int round_ctr = 0;
    for (int r = 0; r < HALF_N_FULL_ROUNDS; r++) {
        for (int i = 0; i < W; i++) state[i] = state[i] + P(ZKM_K(PT_RC)[i + W * round_ctr]);
        for (int i = 0; i < W; i++) {
            P tmp = lv[reg_full0_s0(r, i)], out = lv[reg_full0_s0(r, i) + 1];
            sbox<P>(state[i], tmp, out, yc);
            state[i] = out;
        }
        mds_layer<P>(state);
        round_ctr++;
    }
    // partial rounds, fast schedule
    for (int i = 0; i < W; i++) state[i] = state[i] + P(ZKM_K(PT_FIRST)[i]);
    {
        P res[W];
        res[0] = state[0];
        for (int c = 1; c < W; c++) res[c] = P(0);
        for (int r = 1; r < W; r++)
            for (int c = 1; c < W; c++) res[c] = res[c] + state[r] * P(ZKM_K(PT_INIT)[(r - 1) * 11 + (c - 1)]);
        for (int i = 0; i < W; i++) state[i] = res[i];
    }
    for (int r = 0; r < N_PARTIAL_ROUNDS; r++) {
        P inter = lv[reg_partial_s0(r)], out = lv[reg_partial_s0(r) + 1];
        sbox<P>(state[0], inter, out, yc);
        state[0] = out;
        if (r < N_PARTIAL_ROUNDS - 1) state[0] = state[0] + P(ZKM_K(PT_PRC)[r]);
        // mds_partial_layer_fast_field
        P s0 = state[0];
        P d = s0 * P(ZKM_K(PT_CIRC)[0] + ZKM_K(PT_DIAG)[0]);
        for (int i = 1; i < W; i++) d = d + state[i] * P(ZKM_K(PT_WHAT)[r * 11 + i - 1]);
        for (int i = 1; i < W; i++) state[i] = s0 * P(ZKM_K(PT_VS)[r * 11 + i - 1]) + state[i];
        state[0] = d;
    }
    round_ctr += N_PARTIAL_ROUNDS;
    for (int r = 0; r < HALF_N_FULL_ROUNDS; r++) {
        for (int i = 0; i < W; i++) state[i] = state[i] + P(ZKM_K(PT_RC)[i + W * round_ctr]);
        for (int i = 0; i < W; i++) {
            P tmp = lv[reg_full1_s0(r, i)], out = lv[reg_full1_s0(r, i) + 1];
            sbox<P>(state[i], tmp, out, yc);
            state[i] = out;
        }
        mds_layer<P>(state);
        round_ctr++;
    }
    for (int i = 0; i < W; i++) yc.constraint(state[i] - lv[reg_out(i)]);
}

inline std::vector<Column> ctl_data_inputs() {
    std::vector<int> c;
    for (int i = 0; i < W; i++) c.push_back(reg_in(i));
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_data_outputs() {
    std::vector<int> c;
    for (int i = 0; i < W; i++) c.push_back(reg_out(i));
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline Filter ctl_filter_inputs() { return Filter::new_simple(Column::single(FILTER)); }
inline Filter ctl_filter_outputs() { return Filter::new_simple(Column::single(FILTER)); }

}  // namespace poseidon
}  // namespace tables
}  // namespace zkm
