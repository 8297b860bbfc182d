"""Valid-trace generator for the Arithmetic table, restated from the reference's witness generators (test
infrastructure, like oracle/): arithmetic/mod.rs:48-312 (BinaryOperator::result, binary_op_to_rows),
addcy.rs:30-56, mul.rs:70-124, mult.rs:54-156, slt.rs:14-46, lui.rs:34-48, lo_hi.rs:15-24, shift.rs:42-91,
sra.rs:31-91 + :273-305 (sign-extend interpolant), div.rs:22-309 (generate_div, generate_divu_helper,
generate_modular_op), utils.rs (pol_* helpers), arithmetic_stark.rs:127-195 (range counter, frequencies,
padding).  Used by the generate => constraints-vanish tests the reference keeps in every arithmetic module
(`generate_eval_consistency`) and by the prove -> verify parity tests of the Arithmetic System."""
import numpy as np

P = 0xFFFFFFFF00000001
N_LIMBS, LIMB_BITS = 2, 16
MASK16 = 0xFFFF
(IS_ADD, IS_ADDU, IS_ADDI, IS_ADDIU, IS_SUB, IS_SUBU, IS_MULT, IS_MULTU, IS_MUL, IS_DIV, IS_DIVU, IS_SLLV, IS_SRLV, IS_SRAV,
 IS_SLL, IS_SRL, IS_SRA, IS_SLT, IS_SLTU, IS_SLTI, IS_SLTIU, IS_LUI, IS_MFHI, IS_MTHI, IS_MFLO, IS_MTLO) = range(26)
START_SHARED_COLS = 26
NUM_SHARED_COLS = 9 * N_LIMBS
IN0, IN1, IN2, OUT, AUXIN0, AUXIN1, AUXIN2 = (START_SHARED_COLS + 2 * k for k in range(7))
AUXIN2_END = AUXIN2 + N_LIMBS
AUX_REG0 = START_SHARED_COLS
AUX_REG1 = AUX_REG0 + N_LIMBS
AUX_REG2 = AUX_REG1 + 2 * N_LIMBS
AUX_REG2_END = AUX_REG2 + 2 * N_LIMBS - 1
AUX_COEFF_ABS_MAX = 1 << 20
MUL_AUX_LO, MUL_AUX_HI = AUXIN0, AUXIN1
MODULAR_OUT_AUX_RED, MODULAR_MOD_IS_ZERO = AUX_REG0, AUX_REG1
MODULAR_AUX_INPUT_LO, MODULAR_AUX_INPUT_HI, MODULAR_DIV_DENOM_IS_ZERO = AUX_REG1 + 1, AUX_REG2, AUX_REG2_END
RANGE_COUNTER = START_SHARED_COLS + NUM_SHARED_COLS
RC_FREQUENCIES = RANGE_COUNTER + 1
AUX_EXTRA = RC_FREQUENCIES + 1
NUM_COLUMNS = START_SHARED_COLS + NUM_SHARED_COLS + 10
OUT_LO, OUT_HI = OUT, OUT + N_LIMBS
MULT_AUX_LO = OUT_HI + N_LIMBS
MULT_AUX_HI = MULT_AUX_LO + 2 * N_LIMBS
QUOT_ABS = AUXIN2_END
REM_ABS = QUOT_ABS + N_LIMBS
RANGE_MAX = 1 << 16
assert NUM_COLUMNS == 54 and MODULAR_DIV_DENOM_IS_ZERO == 35 and REM_ABS + N_LIMBS == RANGE_COUNTER

M32 = 0xFFFFFFFF


def s32(x):
    return x - (1 << 32) if x & 0x80000000 else x


def sign_extend16(v):          # witness/util.rs:97-106 sign_extend::<16>
    return (v & 0xFFFF) | 0xFFFF0000 if (v >> 15) != 0 else v & 0xFFFF


def fe(x):                      # F::from_canonical_i64 / from_noncanonical_i64
    return x % P


def put_u32(row, at, x):        # utils.rs:325 u32_to_array
    row[at] = x & MASK16
    row[at + 1] = (x >> 16) & MASK16


def limbs(row, at, n=N_LIMBS):  # utils.rs:316 read_value_i64_limbs
    return [int(row[at + i]) for i in range(n)]


def pol_mul_lo(a, b):           # utils.rs:186
    n = len(a)
    return [sum(a[i] * b[d - i] for i in range(d + 1)) for d in range(n)]


def pol_mul_wide2(a, b):        # utils.rs:157   a: 2N, b: N -> 3N - 1
    res = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            res[i + j] += x * y
    return res


def pol_remove_root_2exp(a, exp=LIMB_BITS):   # utils.rs:281 (last element deliberately zero)
    n = len(a)
    q = [0] * n
    q[0] = -(a[0] >> exp)
    for d in range(1, n - 1):
        q[d] = (q[d - 1] - a[d]) >> exp
    return q


# ---------------------------------------------------------------------------------------------- result
def result(op, a, b):           # mod.rs:48-133
    if op in (IS_ADD, IS_ADDU):
        return (a + b) & M32, 0
    if op in (IS_ADDI, IS_ADDIU):
        return (a + sign_extend16(b)) & M32, 0
    if op in (IS_SUB, IS_SUBU):
        return (a - b) & M32, 0
    if op == IS_SLL:
        return (0 if b > 31 else (a << b) & M32), 0
    if op == IS_SRL:
        return (0 if b > 31 else a >> b), 0
    if op == IS_SRA:
        return (0 if b > 31 else (s32(a) >> b) & M32), 0
    if op == IS_SLLV:
        return (a << (b & 0x1F)) & M32, 0
    if op == IS_SRLV:
        return a >> (b & 0x1F), 0
    if op == IS_SRAV:
        return (s32(a) >> (b & 0x1F)) & M32, 0
    if op == IS_MUL:
        return (a * b) & M32, 0
    if op == IS_SLTU:
        return int(a < b), 0
    if op == IS_SLT:
        return int(s32(a) < s32(b)), 0
    if op == IS_SLTIU:
        return int(a < sign_extend16(b)), 0
    if op == IS_SLTI:
        return int(s32(a) < s32(sign_extend16(b))), 0
    if op == IS_LUI:
        return (sign_extend16(a) << 16) & M32, 0
    if op == IS_MULT:
        out = (s32(a) * s32(b)) & 0xFFFFFFFFFFFFFFFF
        return out & M32, out >> 32
    if op == IS_MULTU:
        out = a * b
        return out & M32, out >> 32
    if op == IS_DIV:            # Rust i32 division truncates toward zero
        x, y = s32(a), s32(b)
        q = abs(x) // abs(y)
        q = -q if (x < 0) != (y < 0) else q
        return q & M32, (x - q * y) & M32
    if op == IS_DIVU:
        return a // b, a % b
    if op in (IS_MFHI, IS_MTHI, IS_MFLO, IS_MTLO):
        return a, 0
    raise ValueError(op)


# ------------------------------------------------------------------------------------------- generators
def gen_addcy(row, op, a, b):   # addcy.rs:30-56
    put_u32(row, IN0, a); put_u32(row, IN1, b); put_u32(row, IN2, 0)
    if op in (IS_SUB, IS_SUBU):
        res, cy = (a - b) & M32, int(a < b)
    else:
        res, cy = (a + b) & M32, int(a + b > M32)
    put_u32(row, AUXIN0, cy)
    put_u32(row, OUT, res)


def gen_mul_limbs(row, left, right):   # mul.rs:70-111 generate_mul
    unreduced = pol_mul_lo(left, right)
    out, cy = [0] * N_LIMBS, 0
    for c in range(N_LIMBS):
        t = unreduced[c] + cy
        cy = t >> LIMB_BITS
        out[c] = t & MASK16
    for i in range(N_LIMBS):
        row[OUT + i] = fe(out[i])
        unreduced[i] -= out[i]
    aux = pol_remove_root_2exp(unreduced)
    aux[N_LIMBS - 1] = -cy
    aux = [c + AUX_COEFF_ABS_MAX for c in aux]
    assert all(abs(c) <= 2 * AUX_COEFF_ABS_MAX for c in aux)
    for i, c in enumerate(aux):
        row[MUL_AUX_LO + i] = c & MASK16
        row[MUL_AUX_HI + i] = (c >> 16) & MASK16


def gen_mul(row, a, b):         # mul.rs:113-124
    put_u32(row, IN0, a); put_u32(row, IN1, b); put_u32(row, IN2, 0)
    gen_mul_limbs(row, limbs(row, IN0), limbs(row, IN1))


def gen_mult_helper(row, left, right):   # mult.rs:109-156
    unreduced = pol_mul_lo(left, right)
    out, cy = [0] * (2 * N_LIMBS), 0
    for c in range(2 * N_LIMBS):
        t = unreduced[c] + cy
        cy = t >> LIMB_BITS
        out[c] = t & MASK16
    for i in range(N_LIMBS):
        row[OUT_LO + i] = fe(out[i])
        row[OUT_HI + i] = fe(out[N_LIMBS + i])
    for i in range(2 * N_LIMBS):
        unreduced[i] -= out[i]
    aux = pol_remove_root_2exp(unreduced)
    aux[2 * N_LIMBS - 1] = -cy
    aux = [c + AUX_COEFF_ABS_MAX for c in aux]
    assert all(abs(c) <= 2 * AUX_COEFF_ABS_MAX for c in aux)
    for i, c in enumerate(aux):
        row[MULT_AUX_LO + i] = c & MASK16
        row[MULT_AUX_HI + i] = (c >> 16) & MASK16


def gen_mult(row, op, a, b):    # mult.rs:54-107
    put_u32(row, IN0, a); put_u32(row, IN1, b)
    l, r = limbs(row, IN0), limbs(row, IN1)
    if op == IS_MULT:
        an, bn = s32(a) < 0, s32(b) < 0
        row[AUX_EXTRA] = int(an)
        row[AUX_EXTRA + 1] = int(bn)
        row[IN2] = (a >> LIMB_BITS) ^ 0x8000
        row[IN2 + 1] = (b >> LIMB_BITS) ^ 0x8000
        l = l + [0xFFFF if an else 0] * N_LIMBS
        r = r + [0xFFFF if bn else 0] * N_LIMBS
    else:
        l = l + [0] * N_LIMBS
        r = r + [0] * N_LIMBS
    gen_mult_helper(row, l, r)


def gen_slt(row, op, a, b, rd):  # slt.rs:14-46
    put_u32(row, IN0, a); put_u32(row, IN1, b); put_u32(row, IN2, 0)
    diff, cy = (a - b) & M32, int(a < b)
    cy_val = cy
    if op in (IS_SLT, IS_SLTI) and (a & 0x80000000) != (b & 0x80000000):
        cy_val = (1 << 16) | (1 - cy)          # `!cy as u32` on a bool
    put_u32(row, AUXIN0, diff)
    put_u32(row, AUXIN1, cy_val)
    put_u32(row, OUT, rd)


def gen_lui(row, imm, rt):      # lui.rs:34-48
    put_u32(row, IN0, imm); put_u32(row, IN1, 1 << 16); put_u32(row, OUT, rt)
    gen_mul_limbs(row, limbs(row, IN0), limbs(row, IN1))


def gen_lo_hi(row, a, res):     # lo_hi.rs:15-24
    put_u32(row, IN0, a); put_u32(row, OUT, res)


def gen_modular_op(lv, nv, op, pol_input, modulus_at):   # div.rs:192-309
    modulus_limbs = limbs(lv, modulus_at)
    modulus = modulus_limbs[0] + (modulus_limbs[1] << 16)
    constr = list(pol_input) + [0]
    mod_is_zero = 0
    if modulus == 0:
        if op in (IS_DIV, IS_DIVU, IS_SRL, IS_SRLV):
            modulus = 1 << 32
        else:
            modulus = 1
            modulus_limbs[0] = 1
        mod_is_zero = 1
    inp = sum(c << (16 * i) for i, c in enumerate(constr))
    output = inp % modulus                      # Python % is already non-negative for modulus > 0
    out_limbs = [output & MASK16, (output >> 16) & MASK16]
    quot = (inp - output) // modulus
    assert quot >= 0
    quot_limbs = [(quot >> (16 * i)) & MASK16 for i in range(2 * N_LIMBS)]
    red = (1 << 32) - modulus + output
    out_aux_red = [red & MASK16, (red >> 16) & MASK16]
    assert red >> 32 == 0
    for i in range(N_LIMBS):
        constr[i] -= out_limbs[i]
    prod = pol_mul_wide2(quot_limbs, modulus_limbs)
    for i in range(2 * N_LIMBS):
        constr[i] -= prod[i]
    assert all(x == 0 for x in prod[2 * N_LIMBS:])
    aux = [c + AUX_COEFF_ABS_MAX for c in pol_remove_root_2exp(constr)]
    assert all(abs(c) <= 2 * AUX_COEFF_ABS_MAX for c in aux)
    for i in range(2 * N_LIMBS - 1):
        nv[MODULAR_AUX_INPUT_LO + i] = aux[i] & MASK16
        nv[MODULAR_AUX_INPUT_HI + i] = (aux[i] >> 16) & MASK16
    nv[MODULAR_MOD_IS_ZERO] = mod_is_zero
    for i in range(N_LIMBS):
        nv[MODULAR_OUT_AUX_RED + i] = fe(out_aux_red[i])
    nv[MODULAR_DIV_DENOM_IS_ZERO] = mod_is_zero * (int(lv[IS_DIV]) + int(lv[IS_DIVU]) + int(lv[IS_SRL]) + int(lv[IS_SRLV]))
    return out_limbs, quot_limbs


def gen_divu_helper(lv, nv, op, input_at, modulus_at, output_at, rem_at):   # div.rs:142-186
    pol_input = limbs(lv, input_at) + [0] * (N_LIMBS - 1)
    out, quo = gen_modular_op(lv, nv, op, pol_input, modulus_at)
    assert all(x == 0 for x in quo[N_LIMBS:])
    assert limbs(lv, output_at) == quo[:N_LIMBS], "computed output doesn't match expected"
    if rem_at is not None:
        assert limbs(lv, rem_at) == out, "computed rem doesn't match expected"
    else:
        for i in range(N_LIMBS):
            lv[AUXIN0 + i] = out[i]


def gen_div(lv, nv, op, a, b, quot, rem):   # div.rs:22-140
    put_u32(lv, IN0, a); put_u32(lv, IN1, b); put_u32(lv, OUT_LO, quot); put_u32(lv, OUT_HI, rem)
    if op == IS_DIVU:
        gen_divu_helper(lv, nv, IS_DIVU, IN0, IN1, OUT_LO, OUT_HI)
        return

    def fill(x, abs_at, sum_idx, neg_idx, borrow_idx):
        neg = s32(x) < 0
        nv[neg_idx] = int(neg)
        nv[sum_idx] = (x >> LIMB_BITS) ^ 0x8000
        nv[borrow_idx] = int((x & 0xFFFF) > 0)
        put_u32(lv, abs_at, abs(s32(x)))
        return neg
    D = MODULAR_DIV_DENOM_IS_ZERO
    n0 = fill(a, IN2, D + 1, D + 5, D + 6)
    n1 = fill(b, AUXIN2, D + 2, D + 7, D + 8)
    nv[RC_FREQUENCIES + 5] = int(n0 ^ n1)
    fill(quot, QUOT_ABS, D + 3, RC_FREQUENCIES + 1, RC_FREQUENCIES + 2)
    fill(rem, REM_ABS, D + 4, RC_FREQUENCIES + 3, RC_FREQUENCIES + 4)
    gen_divu_helper(lv, nv, IS_DIV, IN2, AUXIN2, QUOT_ABS, REM_ABS)


def gen_shift(lv, nv, op, shift, inp, res):   # shift.rs:42-91
    put_u32(lv, IN0, shift); put_u32(lv, IN1, inp); put_u32(lv, OUT, res)
    put_u32(lv, IN2, 1 << (shift & 0x1F))
    if op in (IS_SLL, IS_SLLV):
        gen_mul_limbs(lv, limbs(lv, IN1), limbs(lv, IN2))
    else:
        gen_divu_helper(lv, nv, op, IN1, IN2, OUT, None)


def _inv(x):
    return pow(x, P - 2, P)


def sign_extend_poly():         # sra.rs:273-282: interpolant through (0,0), (i, sum_{k<=i} 2^(32-k))
    pts, s = [(0, 0)], 0
    for i in range(1, 32):
        s += 1 << (32 - i)
        pts.append((i, s))
    n = len(pts)
    coeffs = [0] * n
    for j, (xj, yj) in enumerate(pts):
        # basis polynomial prod_{m != j} (x - x_m) / (x_j - x_m)
        basis, denom = [1], 1
        for m, (xm, _) in enumerate(pts):
            if m == j:
                continue
            nb = [0] * (len(basis) + 1)
            for k, c in enumerate(basis):
                nb[k] = (nb[k] - c * xm) % P
                nb[k + 1] = (nb[k + 1] + c) % P
            basis = nb
            denom = denom * (xj - xm) % P
        scale = yj * _inv(denom) % P
        for k, c in enumerate(basis):
            coeffs[k] = (coeffs[k] + c * scale) % P
    return coeffs


_SRA_POLY = None


def eval_aux_sign_extend(x):    # sra.rs:284-305 eval_poly on sign_extend_poly
    global _SRA_POLY
    if _SRA_POLY is None:
        _SRA_POLY = sign_extend_poly()
    poly = _SRA_POLY
    results, acc = [], 0
    for k in range(len(poly) // 2 - 1, -1, -1):
        acc = (poly[2 * k] + poly[2 * k + 1] * x + acc * x * x) % P
        results.append(acc)
    return results


def gen_sra(lv, nv, op, shift, inp, res):   # sra.rs:31-91
    put_u32(lv, IN0, shift); put_u32(lv, IN1, inp); put_u32(lv, OUT, res)
    put_u32(lv, IN2, 1 << (shift & 0x1F))
    put_u32(lv, AUXIN2, inp >> shift)
    lv[AUXIN2_END] = (inp >> 16) ^ 0x8000
    lv[AUXIN2_END + 1] = inp >> 31
    aux = eval_aux_sign_extend(shift)
    for i in range(8):
        lv[AUX_EXTRA + i] = aux[i]
        nv[AUX_EXTRA + i] = aux[8 + i]
    put_u32(nv, AUXIN2, (((1 << shift) - 1) << ((32 - shift) % 32)) & M32)
    nv[AUXIN2_END] = shift * shift
    gen_divu_helper(lv, nv, op, IN1, IN2, AUXIN2, None)


def op_to_rows(op, a, b):       # mod.rs:237-312 binary_op_to_rows (inputs as Operation::binary receives them)
    r0, r1 = result(op, a, b)
    lv = [0] * NUM_COLUMNS
    nv = [0] * NUM_COLUMNS
    lv[op] = 1
    if op in (IS_ADD, IS_SUB, IS_ADDI, IS_ADDIU, IS_ADDU, IS_SUBU):
        gen_addcy(lv, op, a, b); return [lv]
    if op == IS_MUL:
        gen_mul(lv, a, b); return [lv]
    if op in (IS_SLT, IS_SLTI, IS_SLTU, IS_SLTIU):
        gen_slt(lv, op, a, b, r0); return [lv]
    if op in (IS_MULT, IS_MULTU):
        gen_mult(lv, op, a, b); return [lv]
    if op in (IS_DIV, IS_DIVU):
        gen_div(lv, nv, op, a, b, r0, r1); return [lv, nv]
    if op == IS_LUI:
        gen_lui(lv, a, r0); return [lv]
    if op in (IS_SLL, IS_SLLV):
        gen_shift(lv, nv, op, b, a, r0); return [lv]
    if op in (IS_SRL, IS_SRLV):
        gen_shift(lv, nv, op, b, a, r0); return [lv, nv]
    if op in (IS_SRA, IS_SRAV):
        gen_sra(lv, nv, op, b, a, r0); return [lv, nv]
    gen_lo_hi(lv, a, r0); return [lv]


def random_ops(count, seed=11):
    """Operations as witness/operation.rs issues them: immediates arrive sign-extended (operation.rs:390,416),
    shift amounts are < 32, divisors are non-zero (Rust's `/` panics on zero)."""
    rng = np.random.default_rng(seed)
    edge = [0, 1, 2, 0x7FFF, 0x8000, 0xFFFF, 0x10000, 0x7FFFFFFF, 0x80000000, 0x80000001, 0xFFFFFFFE, 0xFFFFFFFF]
    ops = []
    for i in range(count):
        op = int(rng.integers(0, 26))

        def pick():
            return edge[int(rng.integers(0, len(edge)))] if rng.random() < 0.25 else int(rng.integers(0, 1 << 32))
        a, b = pick(), pick()
        if op in (IS_SLL, IS_SRL, IS_SRA, IS_SLLV, IS_SRLV, IS_SRAV):
            b = int(rng.integers(0, 32))
        if op in (IS_DIV, IS_DIVU) and b == 0:
            b = 3
        if op == IS_DIV and a == 0x80000000 and b == 0xFFFFFFFF:
            b = 7                                   # i32::MIN / -1 overflows (Rust panics)
        if op == IS_LUI:
            a = sign_extend16(a & 0xFFFF)
        if op in (IS_ADDI, IS_ADDIU, IS_SLTI, IS_SLTIU):
            b = sign_extend16(b & 0xFFFF)
        ops.append((op, a, b))
    return ops


def arithmetic_trace(ops, log_n=16):
    """arithmetic_stark.rs:155-195 generate_trace + :127-153 generate_range_checks -> (54, n) uint64."""
    rows = []
    for op, a, b in ops:
        rows.extend(op_to_rows(op, a, b))
    n = max(1 << max(0, (len(rows) - 1).bit_length()), RANGE_MAX)
    assert n == 1 << log_n, (len(rows), log_n)
    t = np.zeros((NUM_COLUMNS, n), dtype=np.uint64)
    if rows:
        t[:, :len(rows)] = np.array(rows, dtype=np.uint64).T
    t[RANGE_COUNTER, :RANGE_MAX] = np.arange(RANGE_MAX, dtype=np.uint64)
    t[RANGE_COUNTER, RANGE_MAX:] = RANGE_MAX - 1
    shared = t[START_SHARED_COLS:START_SHARED_COLS + NUM_SHARED_COLS]
    assert int(shared.max()) < RANGE_MAX, "column value exceeds the max range value"
    t[RC_FREQUENCIES, :RANGE_MAX] += np.bincount(shared.ravel().astype(np.int64), minlength=RANGE_MAX).astype(np.uint64)
    return t
