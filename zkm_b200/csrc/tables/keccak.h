// Keccak-f table (2431 columns, 24 rows per permutation).  Column map: reference
// prover/src/keccak/columns.rs:5-134; constraints: keccak/keccak_stark.rs:248-415 (eval_round_flags is
// commented out in the reference, :256); bit helpers keccak/logic.rs:17-42; round constants
// keccak/constants.rs:1-26; CTL selectors keccak_stark.rs:36-58.
#pragma once
#include "hd.h"
#include "dsl.h"

namespace zkm {
namespace tables {
namespace keccak {

constexpr int NUM_ROUNDS = 24, NUM_INPUTS = 25;
constexpr int TIMESTAMP = NUM_ROUNDS, START_A = TIMESTAMP + 1, START_C = START_A + 5 * 5 * 2, START_C_PRIME = START_C + 5 * 64,
              START_A_PRIME = START_C_PRIME + 5 * 64, START_A_PRIME_PRIME = START_A_PRIME + 5 * 5 * 64,
              START_A_PRIME_PRIME_0_0_BITS = START_A_PRIME_PRIME + 5 * 5 * 2, REG_A_PRIME_PRIME_PRIME_0_0_LO = START_A_PRIME_PRIME_0_0_BITS + 64,
              REG_A_PRIME_PRIME_PRIME_0_0_HI = REG_A_PRIME_PRIME_PRIME_0_0_LO + 1, NUM_COLUMNS = REG_A_PRIME_PRIME_PRIME_0_0_HI + 1;
static_assert(NUM_COLUMNS == 2431, "keccak layout");

ZKM_DEF_CONST(KECCAK_R, 25, {0, 36, 3, 41, 18, 1, 44, 10, 45, 2, 62, 6, 43, 15, 61, 28, 55, 25, 21, 56, 27, 20, 39, 8, 14})
ZKM_DEF_CONST(KECCAK_RC, 24, {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL,
    0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL,
    0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL,
    0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL})

ZKM_HD constexpr int reg_step(int i) { return i; }
ZKM_HD constexpr int reg_a(int x, int y) { return START_A + (x * 5 + y) * 2; }
ZKM_HD constexpr int reg_c(int x, int z) { return START_C + x * 64 + z; }
ZKM_HD constexpr int reg_c_prime(int x, int z) { return START_C_PRIME + x * 64 + z; }
ZKM_HD constexpr int reg_a_prime(int x, int y, int z) { return START_A_PRIME + x * 64 * 5 + y * 64 + z; }
ZKM_HD int reg_b(int x, int y, int z) {
    int a = (x + 3 * y) % 5, b = x;
    int rot = (int)ZKM_K(KECCAK_R)[a * 5 + b];
    return reg_a_prime(a, b, (z + 64 - rot) % 64);
}
ZKM_HD constexpr int reg_a_prime_prime(int x, int y) { return START_A_PRIME_PRIME + x * 2 * 5 + y * 2; }
ZKM_HD constexpr int reg_a_prime_prime_0_0_bit(int i) { return START_A_PRIME_PRIME_0_0_BITS + i; }
ZKM_HD constexpr int reg_a_prime_prime_prime(int x, int y) { return (x == 0 && y == 0) ? REG_A_PRIME_PRIME_PRIME_0_0_LO : reg_a_prime_prime(x, y); }
inline int reg_input_limb(int i) { int u = i / 2, y = u / 5, x = u % 5; return reg_a(x, y) + (i % 2); }
inline int reg_output_limb(int i) { int u = i / 2, y = u / 5, x = u % 5; return reg_a_prime_prime_prime(x, y) + (i % 2); }

template <class P> ZKM_HD P xor_gen(P x, P y) { return x + y - x * (y + y); }
template <class P> ZKM_HD P xor3_gen(P x, P y, P z) { return xor_gen<P>(x, xor_gen<P>(y, z)); }
template <class P> ZKM_HD P andn_gen(P x, P y) { return (P(1) - x) * y; }

template <class P, class V, class YC>
ZKM_HD void eval(const V& lv, const V& nv, YC& yc) {
    const P filter = lv[reg_step(NUM_ROUNDS - 1)];
    yc.constraint(filter * (filter - P(1)));
    const P final_step = lv[reg_step(NUM_ROUNDS - 1)];
    const P not_final_step = P(1) - final_step;
    yc.constraint(not_final_step * filter);
    P sum_round_flags = P(0);
    for (int i = 0; i < NUM_ROUNDS; i++) sum_round_flags = sum_round_flags + lv[reg_step(i)];
    yc.constraint(sum_round_flags * not_final_step * (nv[TIMESTAMP] - lv[TIMESTAMP]));
    // C'[x, z] = xor(C[x, z], C[x - 1, z], C[x + 1, z - 1])
    for (int x = 0; x < 5; x++)
        for (int z = 0; z < 64; z++) {
            P xr = xor3_gen<P>(lv[reg_c(x, z)], lv[reg_c((x + 4) % 5, z)], lv[reg_c((x + 1) % 5, (z + 63) % 64)]);
            yc.constraint(lv[reg_c_prime(x, z)] - xr);
        }
    // A[x, y] limbs recomposed from A'[x, y, z] ^ C[x, z] ^ C'[x, z]
    for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) {
            P lo = P(0), hi = P(0);
            for (int z = 31; z >= 0; z--) lo = lo + lo + xor3_gen<P>(lv[reg_a_prime(x, y, z)], lv[reg_c(x, z)], lv[reg_c_prime(x, z)]);
            for (int z = 63; z >= 32; z--) hi = hi + hi + xor3_gen<P>(lv[reg_a_prime(x, y, z)], lv[reg_c(x, z)], lv[reg_c_prime(x, z)]);
            yc.constraint(lo - lv[reg_a(x, y)]);
            yc.constraint(hi - lv[reg_a(x, y) + 1]);
        }
    // xor_{i} A'[x, i, z] == C'[x, z]: diff in {0, 2, 4}
    for (int x = 0; x < 5; x++)
        for (int z = 0; z < 64; z++) {
            P sum = P(0);
            for (int i = 0; i < 5; i++) sum = sum + lv[reg_a_prime(x, i, z)];
            P diff = sum - lv[reg_c_prime(x, z)];
            yc.constraint(diff * (diff - P(2)) * (diff - P(4)));
        }
    // A''[x, y] = xor(B[x, y], andn(B[x + 1, y], B[x + 2, y]))
    for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) {
            P lo = P(0), hi = P(0);
            for (int z = 31; z >= 0; z--)
                lo = lo + lo + xor_gen<P>(lv[reg_b(x, y, z)], andn_gen<P>(lv[reg_b((x + 1) % 5, y, z)], lv[reg_b((x + 2) % 5, y, z)]));
            for (int z = 63; z >= 32; z--)
                hi = hi + hi + xor_gen<P>(lv[reg_b(x, y, z)], andn_gen<P>(lv[reg_b((x + 1) % 5, y, z)], lv[reg_b((x + 2) % 5, y, z)]));
            yc.constraint(lo - lv[reg_a_prime_prime(x, y)]);
            yc.constraint(hi - lv[reg_a_prime_prime(x, y) + 1]);
        }
    // A''[0, 0] bit decomposition
    {
        P lo = P(0), hi = P(0);
        for (int z = 31; z >= 0; z--) lo = lo + lo + lv[reg_a_prime_prime_0_0_bit(z)];
        for (int z = 63; z >= 32; z--) hi = hi + hi + lv[reg_a_prime_prime_0_0_bit(z)];
        yc.constraint(lo - lv[reg_a_prime_prime(0, 0)]);
        yc.constraint(hi - lv[reg_a_prime_prime(0, 0) + 1]);
    }
    // A'''[0, 0] = A''[0, 0] XOR RC
    {
        P lo = P(0), hi = P(0);
        for (int half = 0; half < 2; half++) {
            P acc = P(0);
            for (int z = 32 * half + 31; z >= 32 * half; z--) {
                P rc_bit = P(0);
                for (int r = 0; r < NUM_ROUNDS; r++) rc_bit = rc_bit + lv[reg_step(r)] * P((ZKM_K(KECCAK_RC)[r] >> z) & 1);
                acc = acc + acc + xor_gen<P>(lv[reg_a_prime_prime_0_0_bit(z)], rc_bit);
            }
            if (half == 0) lo = acc; else hi = acc;
        }
        yc.constraint(lo - lv[reg_a_prime_prime_prime(0, 0)]);
        yc.constraint(hi - lv[reg_a_prime_prime_prime(0, 0) + 1]);
    }
    // output of a non-final round feeds the next row's input
    for (int x = 0; x < 5; x++)
        for (int y = 0; y < 5; y++) {
            P output_lo = lv[reg_a_prime_prime_prime(x, y)], output_hi = lv[reg_a_prime_prime_prime(x, y) + 1];
            P input_lo = nv[reg_a(x, y)], input_hi = nv[reg_a(x, y) + 1];
            P not_last_round = P(1) - lv[reg_step(NUM_ROUNDS - 1)];
            yc.constraint_transition(not_last_round * (output_lo - input_lo));
            yc.constraint_transition(not_last_round * (output_hi - input_hi));
        }
}

inline std::vector<Column> ctl_data_inputs() {
    std::vector<int> c;
    for (int i = 0; i < 2 * NUM_INPUTS; i++) c.push_back(reg_input_limb(i));
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline std::vector<Column> ctl_data_outputs() {
    std::vector<int> c;
    for (int i = 0; i < 2 * NUM_INPUTS; i++) c.push_back(reg_output_limb(i));
    c.push_back(TIMESTAMP);
    return Column::singles(c);
}
inline Filter ctl_filter_inputs() { return Filter::new_simple(Column::single(reg_step(0))); }
inline Filter ctl_filter_outputs() { return Filter::new_simple(Column::single(reg_step(NUM_ROUNDS - 1))); }

}  // namespace keccak
}  // namespace tables
}  // namespace zkm
