// Arithmetic table (54 columns, 16-bit limbs, N_LIMBS = 2).
// Column map: reference prover/src/arithmetic/columns.rs:5-127.  Constraints, in the reference's emission
// order (arithmetic_stark.rs:199-230): range counter -> mul (mul.rs:179-187, helper :143-177) ->
// mult/multu (mult.rs:158-304) -> addcy (addcy.rs:86-160) -> slt (slt.rs:50-111) -> lui (lui.rs:51-71)
// -> divu/div (div.rs:249-563) -> sll/srl (shift.rs:128-166) -> sra (sra.rs:93-160) -> lo_hi
// (lo_hi.rs:25-39); polynomial helpers from utils.rs.  logUp over the 18 shared columns against
// RANGE_COUNTER (arithmetic_stark.rs:269-276); CTL rows :61-116.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace arithmetic {

constexpr int N_LIMBS = 2, LIMB_BITS = 16;
enum {
    IS_ADD = 0, IS_ADDU, IS_ADDI, IS_ADDIU, IS_SUB, IS_SUBU, IS_MULT, IS_MULTU, IS_MUL, IS_DIV, IS_DIVU, IS_SLLV, IS_SRLV, IS_SRAV,
    IS_SLL, IS_SRL, IS_SRA, IS_SLT, IS_SLTU, IS_SLTI, IS_SLTIU, IS_LUI, IS_MFHI, IS_MTHI, IS_MFLO, IS_MTLO, START_SHARED_COLS
};
constexpr int NUM_SHARED_COLS = 9 * N_LIMBS;
constexpr int INPUT_REGISTER_0 = START_SHARED_COLS, INPUT_REGISTER_1 = INPUT_REGISTER_0 + N_LIMBS, INPUT_REGISTER_2 = INPUT_REGISTER_1 + N_LIMBS,
              OUTPUT_REGISTER = INPUT_REGISTER_2 + N_LIMBS, AUX_INPUT_REGISTER_0 = OUTPUT_REGISTER + N_LIMBS,
              AUX_INPUT_REGISTER_1 = AUX_INPUT_REGISTER_0 + N_LIMBS, AUX_INPUT_REGISTER_2 = AUX_INPUT_REGISTER_1 + N_LIMBS,
              AUX_INPUT_REGISTER_2_END = AUX_INPUT_REGISTER_2 + N_LIMBS;
constexpr int AUX_REGISTER_0 = START_SHARED_COLS, AUX_REGISTER_1 = AUX_REGISTER_0 + N_LIMBS, AUX_REGISTER_2 = AUX_REGISTER_1 + 2 * N_LIMBS,
              AUX_REGISTER_2_END = AUX_REGISTER_2 + 2 * N_LIMBS - 1;
constexpr uint64_t AUX_COEFF_ABS_MAX = 1 << 20;
constexpr int MUL_AUX_INPUT_LO = AUX_INPUT_REGISTER_0, MUL_AUX_INPUT_HI = AUX_INPUT_REGISTER_1;
constexpr int MODULAR_OUT_AUX_RED = AUX_REGISTER_0, MODULAR_MOD_IS_ZERO = AUX_REGISTER_1, MODULAR_AUX_INPUT_LO = AUX_REGISTER_1 + 1,
              MODULAR_AUX_INPUT_HI = AUX_REGISTER_2, MODULAR_DIV_DENOM_IS_ZERO = AUX_REGISTER_2_END;
constexpr int RANGE_COUNTER = START_SHARED_COLS + NUM_SHARED_COLS, RC_FREQUENCIES = RANGE_COUNTER + 1, AUX_EXTRA = RC_FREQUENCIES + 1;
constexpr int NUM_COLUMNS = START_SHARED_COLS + NUM_SHARED_COLS + 10;
constexpr int OUTPUT_REGISTER_LO = OUTPUT_REGISTER, OUTPUT_REGISTER_HI = OUTPUT_REGISTER + N_LIMBS, MULT_AUX_LO = OUTPUT_REGISTER_HI + N_LIMBS,
              MULT_AUX_HI = MULT_AUX_LO + 2 * N_LIMBS;
constexpr int QUOT_ABS = AUX_INPUT_REGISTER_2_END, REM_ABS = QUOT_ABS + N_LIMBS;      // div.rs:29-30
static_assert(NUM_COLUMNS == 54 && RANGE_COUNTER == 44 && MODULAR_DIV_DENOM_IS_ZERO == 35 && REM_ABS + N_LIMBS == RANGE_COUNTER, "arithmetic layout");

constexpr uint64_t BASE = 1ull << LIMB_BITS;
constexpr uint64_t GOLDILOCKS_INVERSE_65536 = 18446462594437939201ull;                  // addcy.rs:32
constexpr uint64_t RANGE_MAX = 1ull << 16;

// Coefficients (low to high) of sra.rs:162-171 sign_extend_poly: the interpolant through
// (0, 0), (i, sum_{k=1..i} 2^(32-k)) for i = 1..31, computed offline (tools/gen_sra_poly.py).
ZKM_DEF_CONST(SRA_POLY, 32, {0ULL, 6556208684488608402ULL, 10411650145091517360ULL, 9575747500968261281ULL, 3218865349166832205ULL,
    6003928160650219699ULL, 4659039088580492741ULL, 14725394300870097084ULL, 1391456848276765763ULL, 6047636245641262293ULL,
    13529437519242381204ULL, 13442593302542460522ULL, 1593346132779232121ULL, 10787744126197692619ULL, 7621994036651826023ULL,
    5105723605058418965ULL, 14800830534546910133ULL, 11713039910610883359ULL, 16483774255748568464ULL, 8218921582491356248ULL,
    16893122587496637607ULL, 18353552016915605209ULL, 17525415630691714770ULL, 13351797609243744194ULL, 4946066446722770778ULL,
    2414861715068685227ULL, 8529652317273235355ULL, 12214710769812389200ULL, 14209687019612721744ULL, 15034969890298596498ULL,
    11759614642361326477ULL, 12473867207094203560ULL})

// utils.rs pol_adjoin_root: (x - root) * a(x), same length
template <class P, int N>
ZKM_HD void pol_adjoin_root(const P* a, P root, P* res) {
    res[0] = -(root * a[0]);
    for (int d = 1; d < N; d++) res[d] = a[d - 1] - root * a[d];
}
// utils.rs pol_mul_lo
template <class P, int N>
ZKM_HD void pol_mul_lo(const P* a, const P* b, P* res) {
    for (int deg = 0; deg < N; deg++) {
        P acc = P(0);
        for (int i = 0; i <= deg; i++) acc = acc + a[i] * b[deg - i];
        res[deg] = acc;
    }
}

// mul.rs:143-177 eval_packed_generic_mul
template <class P, class V, class YC>
ZKM_HD void eval_mul(const V& lv, P filter, const P* left, const P* right, YC& yc) {
    const P base = P(BASE), offset = P(AUX_COEFF_ABS_MAX);
    P aux[N_LIMBS], constr[N_LIMBS], adj[N_LIMBS];
    for (int i = 0; i < N_LIMBS; i++) aux[i] = lv[MUL_AUX_INPUT_LO + i] + (lv[MUL_AUX_INPUT_HI + i] * base - offset);
    pol_mul_lo<P, N_LIMBS>(left, right, constr);
    for (int i = 0; i < N_LIMBS; i++) constr[i] = constr[i] - lv[OUTPUT_REGISTER + i];
    pol_adjoin_root<P, N_LIMBS>(aux, base, adj);
    for (int i = 0; i < N_LIMBS; i++) constr[i] = constr[i] - adj[i];
    for (int i = 0; i < N_LIMBS; i++) yc.constraint(filter * constr[i]);
}

// mult.rs:267-304 eval_packed_generic_mult_helper
template <class P, class V, class YC>
ZKM_HD void eval_mult_helper(const V& lv, P filter, const P* left, const P* right, const P* output, YC& yc) {
    constexpr int N = 2 * N_LIMBS;
    const P base = P(BASE), offset = P(AUX_COEFF_ABS_MAX);
    P aux[N], constr[N], adj[N];
    for (int i = 0; i < N; i++) aux[i] = lv[MULT_AUX_LO + i] + (lv[MULT_AUX_HI + i] * base - offset);
    pol_mul_lo<P, N>(left, right, constr);
    for (int i = 0; i < N; i++) constr[i] = constr[i] - output[i];
    pol_adjoin_root<P, N>(aux, base, adj);
    for (int i = 0; i < N; i++) constr[i] = constr[i] - adj[i];
    for (int i = 0; i < N; i++) yc.constraint(filter * constr[i]);
}

// mult.rs:158-265
template <class P, class V, class YC>
ZKM_HD void eval_mult(const V& lv, YC& yc) {
    P in0[N_LIMBS], in1[N_LIMBS], out[2 * N_LIMBS];
    for (int i = 0; i < N_LIMBS; i++) { in0[i] = lv[INPUT_REGISTER_0 + i]; in1[i] = lv[INPUT_REGISTER_1 + i]; }
    for (int i = 0; i < N_LIMBS; i++) { out[i] = lv[OUTPUT_REGISTER_LO + i]; out[N_LIMBS + i] = lv[OUTPUT_REGISTER_HI + i]; }
    {   // signed: sign-extend both inputs
        const P filter = lv[IS_MULT], base = P(BASE), add = P(1ull << (LIMB_BITS - 1));
        P l[2 * N_LIMBS], r[2 * N_LIMBS];
        for (int side = 0; side < 2; side++) {
            const P* input = side == 0 ? in0 : in1;
            P* result = side == 0 ? l : r;
            P is_neg = lv[AUX_EXTRA + side];
            yc.constraint(filter * is_neg * (P(1) - is_neg));
            P sum = lv[INPUT_REGISTER_2 + side];
            P input_hi = input[N_LIMBS - 1];
            yc.constraint(filter * (input_hi + add - sum - is_neg * base));
            P pad = is_neg * P(65535);
            for (int i = 0; i < N_LIMBS; i++) { result[i] = input[i]; result[N_LIMBS + i] = pad; }
        }
        eval_mult_helper<P, V, YC>(lv, filter, l, r, out, yc);
    }
    {   // unsigned
        P l[2 * N_LIMBS], r[2 * N_LIMBS];
        for (int i = 0; i < N_LIMBS; i++) { l[i] = in0[i]; r[i] = in1[i]; l[N_LIMBS + i] = P(0); r[N_LIMBS + i] = P(0); }
        eval_mult_helper<P, V, YC>(lv, lv[IS_MULTU], l, r, out, yc);
    }
}

// addcy.rs:86-140 eval_packed_generic_addcy
template <class P, class YC>
ZKM_HD void eval_addcy(YC& yc, P filter, const P* x, const P* y, const P* z, const P* given_cy, bool is_two_row_op) {
    const P overflow = P(BASE), overflow_inv = P(GOLDILOCKS_INVERSE_65536);
    P cy = P(0);
    for (int i = 0; i < N_LIMBS; i++) {
        P t = cy + x[i] + y[i] - z[i];
        if (is_two_row_op) yc.constraint_transition(filter * t * (overflow - t));
        else yc.constraint(filter * t * (overflow - t));
        cy = t * overflow_inv;
    }
    if (is_two_row_op) {
        yc.constraint_transition(filter * (cy - given_cy[0]));
        for (int i = 1; i < N_LIMBS; i++) yc.constraint_transition(filter * given_cy[i]);
    } else {
        yc.constraint(filter * given_cy[0] * (given_cy[0] - P(1)));
        yc.constraint(filter * (cy - given_cy[0]));
        for (int i = 1; i < N_LIMBS; i++) yc.constraint(filter * given_cy[i]);
    }
}

template <class P, class V>
ZKM_HD void read2(const V& lv, int start, P* out) { for (int i = 0; i < N_LIMBS; i++) out[i] = lv[start + i]; }

// slt.rs:50-111
template <class P, class V, class YC>
ZKM_HD void eval_slt(const V& lv, YC& yc) {
    P is_lt = lv[IS_SLT] + lv[IS_SLTU];
    P is_lti = lv[IS_SLTI] + lv[IS_SLTIU];
    const P filter = is_lt + is_lti;
    const P sign = lv[IS_SLT] + lv[IS_SLTI];
    P x[N_LIMBS], y[N_LIMBS], z[N_LIMBS], given_cy[N_LIMBS], rd[N_LIMBS];
    read2<P>(lv, INPUT_REGISTER_1, x);          // right
    read2<P>(lv, AUX_INPUT_REGISTER_0, y);      // diff
    read2<P>(lv, INPUT_REGISTER_0, z);          // left
    read2<P>(lv, AUX_INPUT_REGISTER_1, given_cy);
    read2<P>(lv, OUTPUT_REGISTER, rd);
    const P overflow = P(BASE), overflow_inv = P(GOLDILOCKS_INVERSE_65536);
    P cy = P(0);
    for (int i = 0; i < N_LIMBS; i++) {
        P t = cy + x[i] + y[i] - z[i];
        yc.constraint(filter * t * (overflow - t));
        cy = t * overflow_inv;
    }
    yc.constraint(filter * given_cy[0] * (given_cy[0] - P(1)));
    yc.constraint(filter * (cy - given_cy[0]) * (P(1) - sign));
    yc.constraint(filter * given_cy[1] * (P(1) - cy - given_cy[0]));
    yc.constraint_transition(filter * (rd[0] - given_cy[0]));
    for (int i = 1; i < N_LIMBS; i++) {
        yc.constraint(filter * given_cy[i] * (P(1) - sign));
        yc.constraint_transition(filter * rd[i]);
    }
}

template <class P, class V>
ZKM_HD P div_shift_flags(const V& lv) {
    return lv[IS_DIV] + lv[IS_DIVU] + lv[IS_SRL] + lv[IS_SRLV] + lv[IS_SRA] + lv[IS_SRAV];
}

// div.rs:249-330 modular_constr_poly (+ check_reduced :222-247); result: 2*N_LIMBS coefficients
template <class P, class V, class YC>
ZKM_HD void modular_constr_poly(const V& lv, const V& nv, YC& yc, P filter, P* output, P* modulus, const P* quot, P* constr_poly) {
    constexpr int N2 = 2 * N_LIMBS;
    const P mod_is_zero = nv[MODULAR_MOD_IS_ZERO];
    yc.constraint_transition(filter * (mod_is_zero * mod_is_zero - mod_is_zero));
    P limb_sum = P(0);
    for (int i = 0; i < N_LIMBS; i++) limb_sum = limb_sum + modulus[i];
    yc.constraint_transition(filter * limb_sum * mod_is_zero);
    modulus[0] = modulus[0] + mod_is_zero;
    const P div_denom_is_zero = nv[MODULAR_DIV_DENOM_IS_ZERO];
    const P flags = div_shift_flags<P>(lv);
    yc.constraint_transition(filter * (mod_is_zero * flags - div_denom_is_zero));
    output[0] = output[0] + div_denom_is_zero;
    {   // check_reduced
        P out_aux_red[N_LIMBS], is_less_than[N_LIMBS];
        for (int i = 0; i < N_LIMBS; i++) { out_aux_red[i] = nv[MODULAR_OUT_AUX_RED + i]; is_less_than[i] = P(0); }
        is_less_than[0] = P(1) - mod_is_zero * flags;
        eval_addcy<P, YC>(yc, filter, modulus, out_aux_red, output, is_less_than, true);
    }
    output[0] = output[0] - div_denom_is_zero;
    // pol_mul_wide2(quot[2N], modulus[N]) -> 3N-1 coefficients
    P prod[3 * N_LIMBS - 1];
    for (int i = 0; i < 3 * N_LIMBS - 1; i++) prod[i] = P(0);
    for (int i = 0; i < N2; i++)
        for (int j = 0; j < N_LIMBS; j++) prod[i + j] = prod[i + j] + quot[i] * modulus[j];
    for (int i = N2; i < 3 * N_LIMBS - 1; i++) yc.constraint_transition(filter * prod[i]);
    for (int i = 0; i < N2; i++) constr_poly[i] = prod[i];
    for (int i = 0; i < N_LIMBS; i++) constr_poly[i] = constr_poly[i] + output[i];
    const P base = P(BASE), offset = P(AUX_COEFF_ABS_MAX);
    P aux[N2], adj[N2];
    for (int i = 0; i < N2; i++) aux[i] = P(0);
    for (int i = 0; i < 2 * N_LIMBS - 1; i++) aux[i] = nv[MODULAR_AUX_INPUT_LO + i] - offset;
    for (int i = 0; i < 2 * N_LIMBS - 1; i++) aux[i] = aux[i] + base * nv[MODULAR_AUX_INPUT_HI + i];
    pol_adjoin_root<P, N2>(aux, base, adj);
    for (int i = 0; i < N2; i++) constr_poly[i] = constr_poly[i] + adj[i];
}

// div.rs:521-563 eval_packed_div_helper
template <class P, class V, class YC>
ZKM_HD void eval_div_helper(const V& lv, const V& nv, YC& yc, P filter, int num_range, int den_range, int quo_range, int rem_range) {
    constexpr int N2 = 2 * N_LIMBS;
    yc.constraint_last_row(filter);
    P den[N_LIMBS], quo[N2], rem[N_LIMBS], constr_poly[N2];
    read2<P>(lv, den_range, den);
    for (int i = 0; i < N2; i++) quo[i] = i < N_LIMBS ? lv[quo_range + i] : P(0);
    read2<P>(lv, rem_range, rem);
    modular_constr_poly<P, V, YC>(lv, nv, yc, filter, rem, den, quo, constr_poly);
    for (int i = 0; i < N_LIMBS; i++) constr_poly[i] = constr_poly[i] - lv[num_range + i];
    for (int i = 0; i < N2; i++) yc.constraint_transition(filter * constr_poly[i]);
}

// div.rs:403-519 eval_packed_div (signed)
template <class P, class V, class YC>
ZKM_HD void eval_div_signed(const V& lv, const V& nv, YC& yc) {
    const P filter = lv[IS_DIV];
    const P over_flow = P(BASE), add = P(1ull << (LIMB_BITS - 1));
    const int input_idx[4] = {INPUT_REGISTER_0, INPUT_REGISTER_1, OUTPUT_REGISTER_LO, OUTPUT_REGISTER_HI};
    const int abs_idx[4] = {INPUT_REGISTER_2, AUX_INPUT_REGISTER_2, QUOT_ABS, REM_ABS};
    const int sum_idx[4] = {MODULAR_DIV_DENOM_IS_ZERO + 1, MODULAR_DIV_DENOM_IS_ZERO + 2, MODULAR_DIV_DENOM_IS_ZERO + 3, MODULAR_DIV_DENOM_IS_ZERO + 4};
    const int is_neg_idx[4] = {MODULAR_DIV_DENOM_IS_ZERO + 5, MODULAR_DIV_DENOM_IS_ZERO + 7, RC_FREQUENCIES + 1, RC_FREQUENCIES + 3};
    const int lo_borrow_idx[4] = {MODULAR_DIV_DENOM_IS_ZERO + 6, MODULAR_DIV_DENOM_IS_ZERO + 8, RC_FREQUENCIES + 2, RC_FREQUENCIES + 4};
    P negs[4];
    for (int k = 0; k < 4; k++) {      // check_abs
        P is_neg = nv[is_neg_idx[k]];
        yc.constraint_transition(filter * is_neg * (P(1) - is_neg));
        P sum = nv[sum_idx[k]];
        P input_hi = lv[input_idx[k] + N_LIMBS - 1];
        yc.constraint_transition(filter * (input_hi + add - sum - is_neg * over_flow));
        P input_lo_borrow = nv[lo_borrow_idx[k]];
        yc.constraint_transition(filter * input_lo_borrow * (P(1) - input_lo_borrow));
        P neg_inputs[2] = {input_lo_borrow * over_flow - lv[input_idx[k]], over_flow - lv[input_idx[k] + 1] - input_lo_borrow};
        for (int i = 0; i < N_LIMBS; i++)
            yc.constraint_transition(filter * (is_neg * neg_inputs[i] + (P(1) - is_neg) * lv[input_idx[k] + i] - lv[abs_idx[k] + i]));
        negs[k] = is_neg;
    }
    const P is_input0_neg = negs[0], is_input1_neg = negs[1], is_quot_neg = negs[2], is_rem_neg = negs[3];
    const P is_same_sign = nv[RC_FREQUENCIES + 5];
    yc.constraint_transition(filter * (is_input0_neg + is_input1_neg - P(2) * is_input0_neg * is_input1_neg - is_same_sign));
    P quot_limbs_sum = P(0), rem_limbs_sum = P(0);
    for (int i = 0; i < N_LIMBS; i++) { quot_limbs_sum = quot_limbs_sum + lv[OUTPUT_REGISTER_LO + i]; rem_limbs_sum = rem_limbs_sum + lv[OUTPUT_REGISTER_HI + i]; }
    yc.constraint_transition(filter * (is_quot_neg - is_same_sign) * quot_limbs_sum);
    yc.constraint_transition(filter * (is_rem_neg - is_input0_neg) * rem_limbs_sum);
    eval_div_helper<P, V, YC>(lv, nv, yc, filter, INPUT_REGISTER_2, AUX_INPUT_REGISTER_2, QUOT_ABS, REM_ABS);
}

// sra.rs:93-160
template <class P, class V, class YC>
ZKM_HD void eval_sra(const V& lv, const V& nv, YC& yc) {
    const P filter = lv[IS_SRA] + lv[IS_SRAV];
    for (int i = 1; i < N_LIMBS; i++) yc.constraint_transition(filter * lv[INPUT_REGISTER_0 + i]);
    const P shift0 = lv[INPUT_REGISTER_0];
    const P is_neg = lv[AUX_INPUT_REGISTER_2_END + 1];
    yc.constraint_transition(filter * is_neg * (P(1) - is_neg));
    const P over_flow = P(BASE), add = P(1ull << (LIMB_BITS - 1));
    const P sum = lv[AUX_INPUT_REGISTER_2_END];
    const P input_hi = lv[INPUT_REGISTER_1 + N_LIMBS - 1];
    yc.constraint_transition(filter * (input_hi + add - sum - is_neg * over_flow));
    const P shift_sq = nv[AUX_INPUT_REGISTER_2_END];
    yc.constraint_transition(filter * (shift_sq - shift0 * shift0));
    // Horner in x^2 over the reversed coefficient list, 2 coefficients per step; witnesses: the 8
    // AUX_EXTRA cells of this row then the 8 of the next row
    P acc = P(0);
    for (int k = 0; k < 16; k++) {
        P w = k < 8 ? lv[AUX_EXTRA + k] : nv[AUX_EXTRA + k - 8];
        P j0 = P(ZKM_K(SRA_POLY)[31 - 2 * k]), j1 = P(ZKM_K(SRA_POLY)[30 - 2 * k]);
        yc.constraint_transition(filter * (acc * shift_sq + j0 * shift0 + j1 - w));
        acc = w;
    }
    const P acc_lo = nv[AUX_INPUT_REGISTER_2], acc_hi = nv[AUX_INPUT_REGISTER_2 + 1];
    yc.constraint_transition(filter * (acc_hi * over_flow + acc_lo - acc));
    eval_div_helper<P, V, YC>(lv, nv, yc, filter, INPUT_REGISTER_1, INPUT_REGISTER_2, AUX_INPUT_REGISTER_2, AUX_INPUT_REGISTER_0);
    const P accs[2] = {acc_lo, acc_hi};
    for (int i = 0; i < N_LIMBS; i++)
        yc.constraint_transition(filter * (lv[AUX_INPUT_REGISTER_2 + i] + accs[i] * is_neg - lv[OUTPUT_REGISTER + i]));
}

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    // range counter: starts at 0, increments by 0 or 1, ends at 2^16 - 1 (arithmetic_stark.rs:205-219)
    const P rc1 = lv[RANGE_COUNTER], rc2 = nv[RANGE_COUNTER];
    yc.constraint_first_row(rc1);
    const P incr = rc2 - rc1;
    yc.constraint_transition(incr * incr - incr);
    yc.constraint_last_row(rc1 - P(RANGE_MAX - 1));

    P in0[N_LIMBS], in1[N_LIMBS], in2[N_LIMBS], out[N_LIMBS], aux0[N_LIMBS];
    read2<P>(lv, INPUT_REGISTER_0, in0); read2<P>(lv, INPUT_REGISTER_1, in1); read2<P>(lv, INPUT_REGISTER_2, in2);
    read2<P>(lv, OUTPUT_REGISTER, out); read2<P>(lv, AUX_INPUT_REGISTER_0, aux0);
    // mul
    eval_mul<P, V, YC>(lv, lv[IS_MUL], in0, in1, yc);
    yc.checkpoint();
    // mult / multu
    eval_mult<P, V, YC>(lv, yc);
    yc.checkpoint();
    // addcy (addcy.rs:142-160)
    eval_addcy<P, YC>(yc, lv[IS_ADD], in0, in1, out, aux0, false);
    eval_addcy<P, YC>(yc, lv[IS_SUB], in1, out, in0, aux0, false);
    eval_addcy<P, YC>(yc, lv[IS_ADDI], in0, in1, out, aux0, false);
    eval_addcy<P, YC>(yc, lv[IS_ADDIU], in0, in1, out, aux0, false);
    yc.checkpoint();
    // slt
    eval_slt<P, V, YC>(lv, yc);
    // lui (lui.rs:59-71): a multiplication
    eval_mul<P, V, YC>(lv, lv[IS_LUI], in0, in1, yc);
    // divu then div (div.rs:381-401)
    eval_div_helper<P, V, YC>(lv, nv, yc, lv[IS_DIVU], INPUT_REGISTER_0, INPUT_REGISTER_1, OUTPUT_REGISTER, AUX_INPUT_REGISTER_0);
    yc.checkpoint();
    eval_div_signed<P, V, YC>(lv, nv, yc);
    yc.checkpoint();
    // sll (shift.rs:128-139) then srl (:141-157)
    eval_mul<P, V, YC>(lv, lv[IS_SLL] + lv[IS_SLLV], in1, in2, yc);
    eval_div_helper<P, V, YC>(lv, nv, yc, lv[IS_SRL] + lv[IS_SRLV], INPUT_REGISTER_1, INPUT_REGISTER_2, OUTPUT_REGISTER, AUX_INPUT_REGISTER_0);
    yc.checkpoint();
    // sra
    eval_sra<P, V, YC>(lv, nv, yc);
    yc.checkpoint();
    // lo_hi (lo_hi.rs:25-39)
    const P f = lv[IS_MFHI] + lv[IS_MTHI] + lv[IS_MFLO] + lv[IS_MTLO];
    for (int i = 0; i < N_LIMBS; i++) yc.constraint(f * (in0[i] - out[i]));
}

// arithmetic_stark.rs:33-116
inline TableWithColumns ctl_arithmetic_rows(int table) {
    const std::pair<int, u64> COMBINED_OPS[26] = {
        {IS_ADD, 0b100000 * (1 << 6)}, {IS_ADDU, 0b100001 * (1 << 6)}, {IS_ADDI, 0b001000}, {IS_ADDIU, 0b001001},
        {IS_SUB, 0b100010 * (1 << 6)}, {IS_SUBU, 0b100011 * (1 << 6)}, {IS_MULT, 0b011000 * (1 << 6)}, {IS_MULTU, 0b011001 * (1 << 6)},
        {IS_MUL, 0b011100 + 0b000010 * (1 << 6)}, {IS_DIV, 0b011010 * (1 << 6)}, {IS_DIVU, 0b011011 * (1 << 6)},
        {IS_SLLV, 0b000100 * (1 << 6)}, {IS_SRLV, 0b000110 * (1 << 6)}, {IS_SRAV, 0b000111 * (1 << 6)}, {IS_SLL, 0b000000 * (1 << 6)},
        {IS_SRL, 0b000010 * (1 << 6)}, {IS_SRA, 0b000011 * (1 << 6)}, {IS_SLT, 0b101010 * (1 << 6)}, {IS_SLTU, 0b101011 * (1 << 6)},
        {IS_SLTI, 0b001010}, {IS_SLTIU, 0b001011}, {IS_LUI, 0b001111}, {IS_MFHI, 0b010000 * (1 << 6)}, {IS_MTHI, 0b010001 * (1 << 6)},
        {IS_MFLO, 0b010010 * (1 << 6)}, {IS_MTLO, 0b010011 * (1 << 6)}};
    std::vector<std::pair<int, u64>> ops(COMBINED_OPS, COMBINED_OPS + 26);
    std::vector<Column> res;
    res.push_back(Column::linear_combination(ops));
    const int regs[3] = {INPUT_REGISTER_0, INPUT_REGISTER_1, OUTPUT_REGISTER};
    for (int r : regs)
        for (int i = 0; i < N_LIMBS / 2; i++) res.push_back(Column::linear_combination({{r + 2 * i, 1}, {r + 2 * i + 1, BASE}}));
    std::vector<int> flags;
    for (auto& o : ops) flags.push_back(o.first);
    return TableWithColumns(table, res, Filter::new_simple(Column::sum(flags)));
}

inline std::vector<Lookup> lookups() {
    Lookup l;
    l.columns = Column::singles(range(START_SHARED_COLS, START_SHARED_COLS + NUM_SHARED_COLS));
    l.table_column = Column::single(RANGE_COUNTER);
    l.frequencies_column = Column::single(RC_FREQUENCIES);
    l.filter_columns.assign(NUM_SHARED_COLS, Filter::none());
    return {l};
}

}  // namespace arithmetic
}  // namespace tables
}  // namespace zkm
