// Device-side table generators from operation logs: the tables whose rows are a pure function of a
// compact log of operations are built directly in HBM, column-major, so that only the log crosses PCIe.
//   Logic     reference logic.rs:108-183  (Operation::into_row, generate_trace_rows: one row per operation, zero padding)
//             log entry = 3 words: operator (0 AND, 1 OR, 2 XOR, 3 NOR), input0, input1 (u32 each) -> 69 columns (23x smaller)
//   Poseidon  reference poseidon/poseidon_stark.rs:51-95 (poseidon_with_witness), :105-145 (generate_trace_rows[_for_perm]:
//             one permutation per row with the x^3 / x^7 witness of every S-box; padding rows = the permutation of zero with
//             FILTER = 0)  log entry = 13 words: the 12 input elements (canonical), timestamp -> 262 columns (20x smaller)
// The Memory table's generator lives in memtrace.cu.  One thread per row; every store is coalesced across the warp.
#include "dev.cuh"
#include "tables/logic.h"
#include "tables/poseidon.h"

namespace zkm {

static size_t padded_rows(size_t n_ops, size_t min_rows) {
    size_t n = n_ops > min_rows ? n_ops : min_rows, p = 1;
    while (p < n) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------------------------ Logic
__global__ void logic_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols, unsigned* bad) {
    namespace lg = tables::logic;
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u64 op = 4, a = 0, b = 0;
    if (r < n_ops) {
        op = ops[3 * r]; a = ops[3 * r + 1]; b = ops[3 * r + 2];
        if (op > 3 || a >> 32 || b >> 32) { atomicExch(bad, 1u); op = 4; a = b = 0; }
    }
    const u32 x = (u32)a, y = (u32)b;
    const u32 res = op == 0 ? (x & y) : op == 1 ? (x | y) : op == 2 ? (x ^ y) : op == 3 ? ~(x | y) : 0u;
    cols[(size_t)lg::IS_AND * n + r] = op == 0; cols[(size_t)lg::IS_OR * n + r] = op == 1;
    cols[(size_t)lg::IS_XOR * n + r] = op == 2; cols[(size_t)lg::IS_NOR * n + r] = op == 3;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        cols[(size_t)(lg::INPUT0 + i) * n + r] = (x >> i) & 1;
        cols[(size_t)(lg::INPUT1 + i) * n + r] = (y >> i) & 1;
    }
    cols[(size_t)lg::RESULT * n + r] = res;
}

size_t logic_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    namespace lg = tables::logic;
    static_assert(lg::NUM_COLUMNS == 69, "logic layout");
    const size_t n = padded_rows(n_ops, min_rows);
    DevBuf ops(3 * n_ops + 1, s), flag(1, s);
    if (n_ops) ops.upload(h_ops, 3 * n_ops);
    flag.zero();
    cols.alloc((size_t)lg::NUM_COLUMNS * n, s);
    ProfScope ps("logic_trace", s, 24.0 * (double)n_ops + 8.0 * lg::NUM_COLUMNS * (double)n);
    logic_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ops.p, n_ops, n, cols.p, (unsigned*)flag.p);
    ZKM_LAUNCHED();
    u64 bad = 0;
    flag.download(&bad, 1);
    ZKM_CHECK((unsigned)bad == 0, "logic operation out of range (operator 0..3, 32-bit inputs)");
    return n;
}

// --------------------------------------------------------------------------------------------------------- Poseidon
__global__ void __launch_bounds__(128) poseidon_rows_kernel(const u64* __restrict__ ops, size_t n_ops, size_t n, u64* __restrict__ cols,
                                                            unsigned* bad) {
    namespace pz = tables::poseidon;
    using namespace tables::poseidon;                  // ZKM_K(name) pastes the bare table name
    size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    auto put = [&](int c, gl v) { cols[(size_t)c * n + row] = v.v; };
    gl st[12];
    const bool real = row < n_ops;
    u64 ts = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 v = real ? ops[13 * row + i] : 0;
        if (v >= GL_P) { atomicExch(bad, 1u); v = 0; }
        st[i] = gl(v);
        put(pz::reg_in(i), st[i]);
    }
    if (real) ts = ops[13 * row + 12];
    put(pz::FILTER, gl(real ? 1 : 0));
    put(pz::TIMESTAMP, gl(ts));
    int round = 0;
    auto full = [&](int r, bool second) {
#pragma unroll 1
        for (int i = 0; i < 12; i++) {
            gl x = st[i] + gl(ZKM_K(PT_RC)[12 * round + i]);
            gl x3 = x * x * x, x7 = x3 * x3 * x;
            const int base = second ? pz::reg_full1_s0(r, i) : pz::reg_full0_s0(r, i);
            put(base, x3); put(base + 1, x7);
            st[i] = x7;
        }
        pz::mds_layer<gl>(st);
        round++;
    };
#pragma unroll 1
    for (int r = 0; r < 4; r++) full(r, false);
    // partial rounds in the reference's fast formulation (poseidon_stark.rs:76-95,389-400,463-501)
#pragma unroll 1
    for (int i = 0; i < 12; i++) st[i] = st[i] + gl(ZKM_K(PT_FIRST)[i]);
    {
        gl o[12];
        o[0] = st[0];
#pragma unroll 1
        for (int c = 0; c < 11; c++) {
            gl acc = gl::zero();
            for (int r = 0; r < 11; r++) acc = acc + st[r + 1] * gl(ZKM_K(PT_INIT)[r * 11 + c]);
            o[c + 1] = acc;
        }
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = o[i];
    }
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        gl x3 = st[0] * st[0] * st[0], x7 = x3 * x3 * st[0];
        put(pz::reg_partial_s0(r), x3); put(pz::reg_partial_s0(r) + 1, x7);
        st[0] = x7;
        if (r < 21) st[0] = st[0] + gl(ZKM_K(PT_PRC)[r]);
        gl d = st[0] * gl(ZKM_K(PT_CIRC)[0] + ZKM_K(PT_DIAG)[0]);
        for (int j = 1; j < 12; j++) d = d + st[j] * gl(ZKM_K(PT_WHAT)[r * 11 + j - 1]);
        for (int j = 1; j < 12; j++) st[j] = st[j] + st[0] * gl(ZKM_K(PT_VS)[r * 11 + j - 1]);
        st[0] = d;
    }
    round += 22;
#pragma unroll 1
    for (int r = 0; r < 4; r++) full(r, true);
#pragma unroll
    for (int i = 0; i < 12; i++) put(pz::reg_out(i), st[i]);
}

size_t poseidon_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    namespace pz = tables::poseidon;
    static_assert(pz::NUM_COLUMNS == 262, "poseidon layout");
    const size_t n = padded_rows(n_ops, min_rows);
    DevBuf ops(13 * n_ops + 1, s), flag(1, s);
    if (n_ops) ops.upload(h_ops, 13 * n_ops);
    flag.zero();
    cols.alloc((size_t)pz::NUM_COLUMNS * n, s);
    ProfScope ps("poseidon_trace", s, 104.0 * (double)n_ops + 8.0 * pz::NUM_COLUMNS * (double)n);
    poseidon_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(ops.p, n_ops, n, cols.p, (unsigned*)flag.p);
    ZKM_LAUNCHED();
    u64 bad = 0;
    flag.download(&bad, 1);
    ZKM_CHECK((unsigned)bad == 0, "poseidon input is not a canonical field element");
    return n;
}

}  // namespace zkm

// ----------------------------------------------------------------------------------------------------------- Keccak
// reference keccak/keccak_stark.rs:62-237 KeccakStark::generate_trace_rows: 24 rows per permutation (one per round) holding the
// round input lanes A as 32-bit limbs, the column parities C and C' and the theta output A' as bits, the chi output A'' as
// limbs, the bits of A''[0][0] and the limbs of A''[0][0] ^ RC[round]; zero rows up to the next power of two of
// max(24 * perms, min_rows).  Log entry = 26 words: the 25 input lanes (input[y * 5 + x] is lane (x, y)), timestamp.
// A round needs the previous round's output, so one thread walks the 24 rounds of a permutation and writes 24 x 2431 cells
// (a warp's stores to one column are 24 rows apart; the table is small in every real segment: 2^6..2^10 rows).
#include "tables/keccak.h"
namespace zkm {
__global__ void __launch_bounds__(64) keccak_rows_kernel(const u64* __restrict__ ops, size_t n_perms, size_t n, u64* __restrict__ cols) {
    namespace kc = tables::keccak;
    using namespace tables::keccak;                      // ZKM_K(KECCAK_R), ZKM_K(KECCAK_RC)
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_perms) return;
    u64 A[5][5];
#pragma unroll
    for (int x = 0; x < 5; x++)
#pragma unroll
        for (int y = 0; y < 5; y++) A[x][y] = ops[26 * p + y * 5 + x];
    const u64 ts = ops[26 * p + 25];
    auto rotl = [](u64 v, int r) { r &= 63; return r ? (v << r) | (v >> (64 - r)) : v; };
#pragma unroll 1
    for (int rnd = 0; rnd < kc::NUM_ROUNDS; rnd++) {
        const size_t row = p * kc::NUM_ROUNDS + rnd;
        auto put = [&](int c, u64 v) { cols[(size_t)c * n + row] = v; };
        put(kc::reg_step(rnd), 1);
        put(kc::TIMESTAMP, ts);
        u64 C[5], Cp[5];
#pragma unroll
        for (int x = 0; x < 5; x++) {
            C[x] = A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4];
#pragma unroll
            for (int y = 0; y < 5; y++) { put(kc::reg_a(x, y), A[x][y] & 0xFFFFFFFFull); put(kc::reg_a(x, y) + 1, A[x][y] >> 32); }
        }
#pragma unroll
        for (int x = 0; x < 5; x++) Cp[x] = C[x] ^ C[(x + 4) % 5] ^ rotl(C[(x + 1) % 5], 1);
        u64 Ap[5][5];
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) Ap[x][y] = A[x][y] ^ C[x] ^ Cp[x];
#pragma unroll 1
        for (int z = 0; z < 64; z++) {
#pragma unroll
            for (int x = 0; x < 5; x++) {
                put(kc::reg_c(x, z), (C[x] >> z) & 1);
                put(kc::reg_c_prime(x, z), (Cp[x] >> z) & 1);
#pragma unroll
                for (int y = 0; y < 5; y++) put(kc::reg_a_prime(x, y, z), (Ap[x][y] >> z) & 1);
            }
        }
        // B[x, y] = rot(A'[(x + 3y) % 5, x], R[(x + 3y) % 5][x])   (columns.rs:83-92);  A''[x, y] = B[x, y] ^ (~B[x+1, y] & B[x+2, y])
        u64 B[5][5];
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) { const int a = (x + 3 * y) % 5; B[x][y] = rotl(Ap[a][x], (int)ZKM_K(KECCAK_R)[a * 5 + x]); }
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) {
                A[x][y] = B[x][y] ^ (~B[(x + 1) % 5][y] & B[(x + 2) % 5][y]);
                put(kc::reg_a_prime_prime(x, y), A[x][y] & 0xFFFFFFFFull); put(kc::reg_a_prime_prime(x, y) + 1, A[x][y] >> 32);
            }
#pragma unroll 1
        for (int z = 0; z < 64; z++) put(kc::reg_a_prime_prime_0_0_bit(z), (A[0][0] >> z) & 1);
        A[0][0] ^= ZKM_K(KECCAK_RC)[rnd];
        put(kc::REG_A_PRIME_PRIME_PRIME_0_0_LO, A[0][0] & 0xFFFFFFFFull);
        put(kc::REG_A_PRIME_PRIME_PRIME_0_0_HI, A[0][0] >> 32);
    }
}

size_t keccak_generate_trace_dev(const u64* h_ops, size_t n_ops, size_t min_rows, DevBuf& cols, cudaStream_t s) {
    namespace kc = tables::keccak;
    const size_t n = padded_rows(n_ops * kc::NUM_ROUNDS, min_rows);
    ZKM_CHECK(n <= ((size_t)1 << 22), "too many Keccak permutations");
    DevBuf ops(26 * n_ops + 1, s);
    if (n_ops) ops.upload(h_ops, 26 * n_ops);
    cols.alloc((size_t)kc::NUM_COLUMNS * n, s);
    cols.zero();
    ProfScope ps("keccak_trace", s, 208.0 * (double)n_ops + 8.0 * kc::NUM_COLUMNS * (double)n);
    if (n_ops) {
        keccak_rows_kernel<<<(unsigned)((n_ops + 63) / 64), 64, 0, s>>>(ops.p, n_ops, n, cols.p);
        ZKM_LAUNCHED();
    }
    return n;
}
}  // namespace zkm

// ------------------------------------------------------------------------------------------------------- Arithmetic
// reference arithmetic/arithmetic_stark.rs:155-192 ArithmeticStark::generate_trace: every operation becomes one row, or two for
// DIV / DIVU / SRL(V) / SRA(V) (mod.rs:237-312 binary_op_to_rows), rows are zero-padded to a power of two >= 2^16, then
// generate_range_checks (:127-153) fills RANGE_COUNTER / RC_FREQUENCIES.  Row generators: addcy.rs:30-56, mul.rs:70-124,
// mult.rs:54-156, slt.rs:14-46, lui.rs:34-48, lo_hi.rs:15-24, div.rs:22-309 (generate_div, generate_divu_helper,
// generate_modular_op), shift.rs:42-91, sra.rs:31-91 (+ :273-305 sign-extend interpolant), utils.rs pol_* helpers -- all in
// the reference's i64 limb arithmetic.  Log entry = 3 words: operator (the IS_* column index 0..25 of columns.rs:5-35, i.e.
// BinaryOperator::row_filter), input0, input1 exactly as Operation::binary receives them (u32); results are recomputed
// (mod.rs:48-133 BinaryOperator::result).  tests/arith_gen.py is the Python restatement this is checked against.
#include "tables/arithmetic.h"

namespace zkm {
namespace {
namespace ar = tables::arithmetic;
typedef long long i64;
constexpr int NCOL = ar::NUM_COLUMNS;
constexpr u64 M32 = 0xFFFFFFFFull, MASK16 = 0xFFFF;
constexpr int IN0 = ar::INPUT_REGISTER_0, IN1 = ar::INPUT_REGISTER_1, IN2 = ar::INPUT_REGISTER_2, OUT = ar::OUTPUT_REGISTER,
              AUXIN0 = ar::AUX_INPUT_REGISTER_0, AUXIN1 = ar::AUX_INPUT_REGISTER_1, AUXIN2 = ar::AUX_INPUT_REGISTER_2,
              AUXIN2_END = ar::AUX_INPUT_REGISTER_2_END, OUT_LO = ar::OUTPUT_REGISTER_LO, OUT_HI = ar::OUTPUT_REGISTER_HI;
constexpr i64 AUX_MAX = (i64)ar::AUX_COEFF_ABS_MAX;

__device__ __forceinline__ bool arith_two_rows(int op) {
    return op == ar::IS_DIV || op == ar::IS_DIVU || op == ar::IS_SRL || op == ar::IS_SRLV || op == ar::IS_SRA || op == ar::IS_SRAV;
}
__device__ __forceinline__ i64 s32(u64 x) { return (i64)(int)(u32)x; }
__device__ __forceinline__ u64 sign_extend16(u64 v) { return (v >> 15) != 0 ? ((v & 0xFFFF) | 0xFFFF0000ull) : (v & 0xFFFF); }
__device__ __forceinline__ u64 fe(i64 x) { return x < 0 ? GL_P - (u64)(-x) : (u64)x; }             // F::from_noncanonical_i64, |x| small
__device__ __forceinline__ void put_u32(u64* row, int at, u64 x) { row[at] = x & MASK16; row[at + 1] = (x >> 16) & MASK16; }

// mod.rs:48-133 BinaryOperator::result -> (result0, result1)
__device__ void arith_result(int op, u64 a, u64 b, u64& r0, u64& r1) {
    r1 = 0;
    switch (op) {
        case ar::IS_ADD: case ar::IS_ADDU: r0 = (a + b) & M32; break;
        case ar::IS_ADDI: case ar::IS_ADDIU: r0 = (a + sign_extend16(b)) & M32; break;
        case ar::IS_SUB: case ar::IS_SUBU: r0 = (a - b) & M32; break;
        case ar::IS_SLL: r0 = b > 31 ? 0 : (a << b) & M32; break;
        case ar::IS_SRL: r0 = b > 31 ? 0 : a >> b; break;
        case ar::IS_SRA: r0 = b > 31 ? 0 : (u64)(s32(a) >> b) & M32; break;
        case ar::IS_SLLV: r0 = (a << (b & 0x1F)) & M32; break;
        case ar::IS_SRLV: r0 = a >> (b & 0x1F); break;
        case ar::IS_SRAV: r0 = (u64)(s32(a) >> (b & 0x1F)) & M32; break;
        case ar::IS_MUL: r0 = (a * b) & M32; break;
        case ar::IS_SLTU: r0 = a < b; break;
        case ar::IS_SLT: r0 = s32(a) < s32(b); break;
        case ar::IS_SLTIU: r0 = a < sign_extend16(b); break;
        case ar::IS_SLTI: r0 = s32(a) < s32(sign_extend16(b)); break;
        case ar::IS_LUI: r0 = (sign_extend16(a) << 16) & M32; break;
        case ar::IS_MULT: { u64 o = (u64)(s32(a) * s32(b)); r0 = o & M32; r1 = o >> 32; break; }
        case ar::IS_MULTU: { u64 o = a * b; r0 = o & M32; r1 = o >> 32; break; }
        case ar::IS_DIV: { i64 x = s32(a), y = s32(b); i64 q = x / y; r0 = (u64)q & M32; r1 = (u64)(x - q * y) & M32; break; }   // truncating, like Rust
        case ar::IS_DIVU: r0 = a / b; r1 = a % b; break;
        default: r0 = a; break;                                  // MFHI / MTHI / MFLO / MTLO
    }
}
// utils.rs:281 pol_remove_root_2exp (the last element stays zero)
template <int N>
__device__ __forceinline__ void remove_root_2exp(const i64* a, i64* q) {
    q[0] = -(a[0] >> 16);
#pragma unroll
    for (int d = 1; d < N - 1; d++) q[d] = (q[d - 1] - a[d]) >> 16;
    q[N - 1] = 0;
}
// mul.rs:70-111 generate_mul on N-limb operands writing N output limbs at out_at and the aux columns at (aux_lo, aux_hi)
template <int N>
__device__ __forceinline__ void gen_mul_limbs(u64* row, const i64* l, const i64* r, int out_at, int out_hi_at, int aux_lo, int aux_hi) {
    i64 un[N], out[N], aux[N];
#pragma unroll
    for (int d = 0; d < N; d++) { i64 s = 0; for (int i = 0; i <= d; i++) s += l[i] * r[d - i]; un[d] = s; }     // pol_mul_lo
    i64 cy = 0;
#pragma unroll
    for (int c = 0; c < N; c++) { i64 t = un[c] + cy; cy = t >> 16; out[c] = t & 0xFFFF; }
#pragma unroll
    for (int i = 0; i < N; i++) {
        // mult.rs writes the low half at OUT_LO and the high half at OUT_HI; mul.rs has N = N_LIMBS and only OUT
        if (i < ar::N_LIMBS) row[out_at + i] = fe(out[i]); else row[out_hi_at + i - ar::N_LIMBS] = fe(out[i]);
        un[i] -= out[i];
    }
    remove_root_2exp<N>(un, aux);
    aux[N - 1] = -cy;
#pragma unroll
    for (int i = 0; i < N; i++) { i64 c = aux[i] + AUX_MAX; row[aux_lo + i] = (u64)c & MASK16; row[aux_hi + i] = ((u64)c >> 16) & MASK16; }
}
// div.rs:192-309 generate_modular_op; returns out limbs (2) and quotient limbs (4)
__device__ void gen_modular_op(u64* lv, u64* nv, int op, const i64* pol_input /*3 limbs*/, int modulus_at, i64* out_limbs, i64* quot_limbs) {
    i64 ml[2] = {(i64)lv[modulus_at], (i64)lv[modulus_at + 1]};
    i64 modulus = ml[0] + (ml[1] << 16);
    i64 constr[4] = {pol_input[0], pol_input[1], pol_input[2], 0};
    i64 mod_is_zero = 0;
    const bool divlike = op == ar::IS_DIV || op == ar::IS_DIVU || op == ar::IS_SRL || op == ar::IS_SRLV;
    if (modulus == 0) {
        if (divlike) modulus = (i64)1 << 32; else { modulus = 1; ml[0] = 1; }
        mod_is_zero = 1;
    }
    const i64 inp = constr[0] + (constr[1] << 16) + (constr[2] << 32);
    const i64 output = inp % modulus;                       // inp >= 0
    out_limbs[0] = output & 0xFFFF; out_limbs[1] = (output >> 16) & 0xFFFF;
    const i64 quot = (inp - output) / modulus;
#pragma unroll
    for (int i = 0; i < 4; i++) quot_limbs[i] = (quot >> (16 * i)) & 0xFFFF;
    const i64 red = ((i64)1 << 32) - modulus + output;
    constr[0] -= out_limbs[0]; constr[1] -= out_limbs[1];
    // pol_mul_wide2(quot_limbs (4), modulus_limbs (2)): only the low 4 coefficients can be non-zero
#pragma unroll
    for (int i = 0; i < 4; i++) {
        i64 p = quot_limbs[i] * ml[0] + (i > 0 ? quot_limbs[i - 1] * ml[1] : 0);
        constr[i] -= p;
    }
    i64 aux[4];
    remove_root_2exp<4>(constr, aux);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        i64 c = aux[i] + AUX_MAX;
        nv[ar::MODULAR_AUX_INPUT_LO + i] = (u64)c & MASK16;
        nv[ar::MODULAR_AUX_INPUT_HI + i] = ((u64)c >> 16) & MASK16;
    }
    nv[ar::MODULAR_MOD_IS_ZERO] = (u64)mod_is_zero;
    nv[ar::MODULAR_OUT_AUX_RED] = (u64)(red & 0xFFFF); nv[ar::MODULAR_OUT_AUX_RED + 1] = (u64)((red >> 16) & 0xFFFF);
    nv[ar::MODULAR_DIV_DENOM_IS_ZERO] = (u64)(mod_is_zero * (i64)(lv[ar::IS_DIV] + lv[ar::IS_DIVU] + lv[ar::IS_SRL] + lv[ar::IS_SRLV]));
}
// div.rs:142-186 generate_divu_helper
__device__ void gen_divu_helper(u64* lv, u64* nv, int op, int input_at, int modulus_at, bool write_rem_aux) {
    i64 pol[3] = {(i64)lv[input_at], (i64)lv[input_at + 1], 0};
    i64 out[2], quo[4];
    gen_modular_op(lv, nv, op, pol, modulus_at, out, quo);
    if (write_rem_aux) { lv[AUXIN0] = (u64)out[0]; lv[AUXIN0 + 1] = (u64)out[1]; }
}
__device__ __forceinline__ bool div_fill(u64* lv, u64* nv, u64 x, int abs_at, int sum_idx, int neg_idx, int borrow_idx) {
    const bool neg = s32(x) < 0;
    nv[neg_idx] = neg;
    nv[sum_idx] = (x >> 16) ^ 0x8000;
    nv[borrow_idx] = (x & 0xFFFF) > 0;
    i64 v = s32(x);
    put_u32(lv, abs_at, (u64)(v < 0 ? -v : v));
    return neg;
}

__global__ void arith_rows_per_op_kernel(const u64* __restrict__ ops, size_t n_ops, u64* __restrict__ rows, unsigned* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ops) return;
    const u64 op = ops[3 * i], a = ops[3 * i + 1], b = ops[3 * i + 2];
    bool ok = op < 26 && !(a >> 32) && !(b >> 32);
    // what the reference panics on: division by zero, i32::MIN / -1
    if (ok && (op == ar::IS_DIV || op == ar::IS_DIVU) && b == 0) ok = false;
    if (ok && op == ar::IS_DIV && a == 0x80000000ull && b == 0xFFFFFFFFull) ok = false;
    if (!ok) atomicExch(bad, 1u);
    rows[i] = ok ? (arith_two_rows((int)op) ? 2 : 1) : 1;
}
// exclusive scan of `v` (n values) by one CTA; out[n] = total
__global__ void arith_scan_kernel(const u64* v, size_t n, u64* out) {
    __shared__ u64 warp_sums[32];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < n; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        u64 x0 = i < n ? v[i] : 0, x = x0;
        for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            u64 w = threadIdx.x < (blockDim.x >> 5) ? warp_sums[threadIdx.x] : 0;
            for (int o = 1; o < 32; o <<= 1) { u64 y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        u64 before = carry + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + x - x0;
        if (i < n) out[i] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + x0;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(128) arith_rows_kernel(const u64* __restrict__ ops, const u64* __restrict__ offs, size_t n_ops, size_t n,
                                                         u64* __restrict__ cols) {
    using namespace tables::arithmetic;                 // ZKM_K(SRA_POLY)
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ops) return;
    const int op = (int)ops[3 * k];
    const u64 a = ops[3 * k + 1], b = ops[3 * k + 2];
    u64 lv[NCOL], nv[NCOL];
#pragma unroll 1
    for (int c = 0; c < NCOL; c++) { lv[c] = 0; nv[c] = 0; }
    u64 r0, r1;
    arith_result(op, a, b, r0, r1);
    lv[op] = 1;
    switch (op) {
        case IS_ADD: case IS_SUB: case IS_ADDI: case IS_ADDIU: case IS_ADDU: case IS_SUBU: {     // addcy.rs:30-56
            put_u32(lv, IN0, a); put_u32(lv, IN1, b); put_u32(lv, IN2, 0);
            u64 res, cy;
            if (op == IS_SUB || op == IS_SUBU) { res = (a - b) & M32; cy = a < b; } else { res = (a + b) & M32; cy = a + b > M32; }
            put_u32(lv, AUXIN0, cy); put_u32(lv, OUT, res);
            break;
        }
        case IS_MUL: {                                                                            // mul.rs:113-124
            put_u32(lv, IN0, a); put_u32(lv, IN1, b); put_u32(lv, IN2, 0);
            i64 l[2] = {(i64)lv[IN0], (i64)lv[IN0 + 1]}, r[2] = {(i64)lv[IN1], (i64)lv[IN1 + 1]};
            gen_mul_limbs<2>(lv, l, r, OUT, OUT, MUL_AUX_INPUT_LO, MUL_AUX_INPUT_HI);
            break;
        }
        case IS_SLT: case IS_SLTI: case IS_SLTU: case IS_SLTIU: {                                 // slt.rs:14-46
            put_u32(lv, IN0, a); put_u32(lv, IN1, b); put_u32(lv, IN2, 0);
            const u64 diff = (a - b) & M32, cy = a < b;
            u64 cy_val = cy;
            if ((op == IS_SLT || op == IS_SLTI) && ((a & 0x80000000ull) != (b & 0x80000000ull))) cy_val = (1ull << 16) | (1 - cy);
            put_u32(lv, AUXIN0, diff); put_u32(lv, AUXIN1, cy_val); put_u32(lv, OUT, r0);
            break;
        }
        case IS_MULT: case IS_MULTU: {                                                            // mult.rs:54-156
            put_u32(lv, IN0, a); put_u32(lv, IN1, b);
            i64 l[4] = {(i64)lv[IN0], (i64)lv[IN0 + 1], 0, 0}, r[4] = {(i64)lv[IN1], (i64)lv[IN1 + 1], 0, 0};
            if (op == IS_MULT) {
                const bool an = s32(a) < 0, bn = s32(b) < 0;
                lv[AUX_EXTRA] = an; lv[AUX_EXTRA + 1] = bn;
                lv[IN2] = (a >> 16) ^ 0x8000; lv[IN2 + 1] = (b >> 16) ^ 0x8000;
                l[2] = l[3] = an ? 0xFFFF : 0; r[2] = r[3] = bn ? 0xFFFF : 0;
            }
            gen_mul_limbs<4>(lv, l, r, OUT_LO, OUT_HI, MULT_AUX_LO, MULT_AUX_HI);
            break;
        }
        case IS_DIV: case IS_DIVU: {                                                              // div.rs:22-140
            put_u32(lv, IN0, a); put_u32(lv, IN1, b); put_u32(lv, OUT_LO, r0); put_u32(lv, OUT_HI, r1);
            if (op == IS_DIVU) { gen_divu_helper(lv, nv, IS_DIVU, IN0, IN1, false); break; }
            const int Dz = MODULAR_DIV_DENOM_IS_ZERO;
            const bool n0 = div_fill(lv, nv, a, IN2, Dz + 1, Dz + 5, Dz + 6);
            const bool n1 = div_fill(lv, nv, b, AUXIN2, Dz + 2, Dz + 7, Dz + 8);
            nv[RC_FREQUENCIES + 5] = n0 != n1;
            div_fill(lv, nv, r0, QUOT_ABS, Dz + 3, RC_FREQUENCIES + 1, RC_FREQUENCIES + 2);
            div_fill(lv, nv, r1, REM_ABS, Dz + 4, RC_FREQUENCIES + 3, RC_FREQUENCIES + 4);
            gen_divu_helper(lv, nv, IS_DIV, IN2, AUXIN2, false);
            break;
        }
        case IS_LUI: {                                                                            // lui.rs:34-48
            put_u32(lv, IN0, a); put_u32(lv, IN1, 1ull << 16); put_u32(lv, OUT, r0);
            i64 l[2] = {(i64)lv[IN0], (i64)lv[IN0 + 1]}, r[2] = {(i64)lv[IN1], (i64)lv[IN1 + 1]};
            gen_mul_limbs<2>(lv, l, r, OUT, OUT, MUL_AUX_INPUT_LO, MUL_AUX_INPUT_HI);
            break;
        }
        case IS_SLL: case IS_SLLV: case IS_SRL: case IS_SRLV: {                                   // shift.rs:42-91 (shift = b, input = a)
            put_u32(lv, IN0, b); put_u32(lv, IN1, a); put_u32(lv, OUT, r0);
            put_u32(lv, IN2, 1ull << (b & 0x1F));
            if (op == IS_SLL || op == IS_SLLV) {
                i64 l[2] = {(i64)lv[IN1], (i64)lv[IN1 + 1]}, r[2] = {(i64)lv[IN2], (i64)lv[IN2 + 1]};
                gen_mul_limbs<2>(lv, l, r, OUT, OUT, MUL_AUX_INPUT_LO, MUL_AUX_INPUT_HI);
            } else gen_divu_helper(lv, nv, op, IN1, IN2, true);
            break;
        }
        case IS_SRA: case IS_SRAV: {                                                              // sra.rs:31-91
            const u64 shift = b, inp = a;
            put_u32(lv, IN0, shift); put_u32(lv, IN1, inp); put_u32(lv, OUT, r0);
            put_u32(lv, IN2, 1ull << (shift & 0x1F));
            put_u32(lv, AUXIN2, inp >> shift);
            lv[AUXIN2_END] = (inp >> 16) ^ 0x8000;
            lv[AUXIN2_END + 1] = inp >> 31;
            {   // sra.rs:284-305 eval_poly of the sign-extend interpolant at `shift`: 16 Horner partial results
                gl x(shift), x2 = x * x, acc = gl::zero();
                int w = 0;
#pragma unroll 1
                for (int t = 15; t >= 0; t--) {
                    acc = gl(ZKM_K(SRA_POLY)[2 * t]) + gl(ZKM_K(SRA_POLY)[2 * t + 1]) * x + acc * x2;
                    if (w < 8) lv[AUX_EXTRA + w] = acc.v; else nv[AUX_EXTRA + w - 8] = acc.v;
                    w++;
                }
            }
            put_u32(nv, AUXIN2, (((1ull << shift) - 1) << ((32 - shift) % 32)) & M32);
            nv[AUXIN2_END] = shift * shift;
            gen_divu_helper(lv, nv, op, IN1, IN2, true);
            break;
        }
        default:                                                                                  // lo_hi.rs:15-24
            put_u32(lv, IN0, a); put_u32(lv, OUT, r0);
            break;
    }
    const size_t r = offs[k];
#pragma unroll 1
    for (int c = 0; c < NCOL; c++) cols[(size_t)c * n + r] = lv[c];
    if (arith_two_rows(op)) {
#pragma unroll 1
        for (int c = 0; c < NCOL; c++) cols[(size_t)c * n + r + 1] = nv[c];
    }
}
}  // namespace

size_t arithmetic_generate_trace_dev(const u64* h_ops, size_t n_ops, DevBuf& cols, cudaStream_t s) {
    ZKM_CHECK(n_ops < ((size_t)1 << 26), "too many arithmetic operations");
    DevBuf ops(3 * n_ops + 1, s), per(n_ops + 1, s), offs(n_ops + 1, s), flag(1, s);
    if (n_ops) ops.upload(h_ops, 3 * n_ops);
    flag.zero();
    u64 total = 0;
    if (n_ops) {
        arith_rows_per_op_kernel<<<(unsigned)((n_ops + 255) / 256), 256, 0, s>>>(ops.p, n_ops, per.p, (unsigned*)flag.p);
        ZKM_LAUNCHED();
        arith_scan_kernel<<<1, 1024, 0, s>>>(per.p, n_ops, offs.p);
        ZKM_LAUNCHED();
        offs.download(&total, 1, n_ops);
    }
    u64 bad = 0;
    flag.download(&bad, 1);
    ZKM_CHECK((unsigned)bad == 0, "arithmetic operation out of range (operator 0..25, 32-bit inputs, non-zero divisor, no i32::MIN / -1)");
    size_t n = 1;
    while (n < (size_t)total) n <<= 1;
    if (n < ar::RANGE_MAX) n = ar::RANGE_MAX;
    cols.alloc((size_t)NCOL * n, s);
    cols.zero();
    ProfScope ps("arithmetic_trace", s, 24.0 * (double)n_ops + 8.0 * NCOL * (double)n);
    if (n_ops) {
        arith_rows_kernel<<<(unsigned)((n_ops + 127) / 128), 128, 0, s>>>(ops.p, offs.p, n_ops, n, cols.p);
        ZKM_LAUNCHED();
    }
    DevBuf range_bad(1, s);
    arith_generate_range_checks(cols.p, n, ar::START_SHARED_COLS, ar::NUM_SHARED_COLS, ar::RANGE_COUNTER, ar::RC_FREQUENCIES,
                                (unsigned*)range_bad.p, s);
    range_bad.download(&bad, 1);
    ZKM_CHECK((unsigned)bad == 0, "column value exceeds the max range value 65536");
    return n;
}

}  // namespace zkm
