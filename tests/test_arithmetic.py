"""Arithmetic table: the reference's own generate => constraints-vanish tests (`generate_eval_consistency` in
arithmetic/{addcy,mul,mult,slt,lui,div,shift,sra,lo_hi}.rs) and its sign-extend polynomial fixture
(sra.rs:320-347 `test_poly`), restated over tests/arith_gen.py and the oracle's constraint evaluator; then the
Arithmetic System (range-check logUp + CPU-facing CTL rows) proves and verifies."""
import numpy as np
import pytest

import arith_gen as ag
import traces as tr
from oracle import binding

# sra.rs:327-344: eval_poly(sign_extend_poly(), 2)
SRA_TEST_POLY_EXPECTED = [
    18260604987135149276, 6641582332263005918, 4185170977706284464, 8069729718270694767, 2720953444603644942,
    9143808191498674830, 14156617978482227317, 2619661922624664514, 9865344867852737688, 7289981648341815148,
    14234318509450809877, 15083771169118776894, 2211192019722872880, 5624745679944178802, 15168639727975586488, 3221225472]


def _check(orc, t):
    return orc.orc_check_table_constraints(tr.T_ARITHMETIC, binding.col_ptrs(t), t.shape[0], t.shape[1].bit_length() - 1)


def test_sra_sign_extend_poly_matches_reference_fixture():
    assert ag.eval_aux_sign_extend(2) == [e % ag.P for e in SRA_TEST_POLY_EXPECTED]
    # the interpolant goes through the points it was built from (sra.rs:273-282)
    poly, s = ag.sign_extend_poly(), 0
    for i in range(1, 32):
        s += 1 << (32 - i)
        assert sum(c * pow(i, k, ag.P) for k, c in enumerate(poly)) % ag.P == s


def test_results_match_mips_semantics():
    assert ag.result(ag.IS_DIV, 0xFFFFFFF9, 2) == (0xFFFFFFFD, 0xFFFFFFFF)      # -7 / 2 = -3 rem -1 (truncating)
    assert ag.result(ag.IS_MULT, 0xFFFFFFFF, 0xFFFFFFFF) == (1, 0)
    assert ag.result(ag.IS_MULTU, 0xFFFFFFFF, 0xFFFFFFFF) == (1, 0xFFFFFFFE)
    assert ag.result(ag.IS_SRA, 0x80000000, 31) == (0xFFFFFFFF, 0)
    assert ag.result(ag.IS_SLTI, 0xFFFFFFFF, 0) == (1, 0)
    assert ag.result(ag.IS_LUI, 0xFFFF8000, 1 << 16) == (0x80000000, 0)


@pytest.fixture(scope="module")
def arith_ops():
    return ag.random_ops(30000, seed=11)


def test_valid_arithmetic_trace_satisfies_constraints(orc, arith_ops):
    t = ag.arithmetic_trace(arith_ops)
    assert t.shape == (54, 1 << 16)
    kinds = {op for op, _, _ in arith_ops}
    assert kinds == set(range(26))
    assert _check(orc, t) == 0, orc.orc_last_error()


def test_every_operation_is_constrained(orc, arith_ops):
    """Corrupting the output (or, for two-row operations, an auxiliary cell of the second row) of one row of each of
    the 26 operations is caught: no operation's constraints were transcribed as vacuous."""
    t = ag.arithmetic_trace(arith_ops)
    row = 0
    first = {}
    for op, a, b in arith_ops:
        nrows = 2 if op in (ag.IS_DIV, ag.IS_DIVU, ag.IS_SRL, ag.IS_SRLV, ag.IS_SRA, ag.IS_SRAV) else 1     # mod.rs:275-300
        first.setdefault(op, (row, nrows))
        row += nrows
    for op, (r, nrows) in sorted(first.items()):
        t2 = t.copy()
        col = ag.OUT
        t2[col, r] = (int(t2[col, r]) + 1) % ag.P
        if op in (ag.IS_ADDU, ag.IS_SUBU):
            # the reference leaves ADDU/SUBU unconstrained (addcy.rs:142-160 evaluates ADD, SUB, ADDI, ADDIU only; the
            # generator carries a FIXME, addcy.rs:47): parity means the transcription must not constrain them either
            assert _check(orc, t2) == 0, f"op {op}: constrained here but not in the reference"
            continue
        assert _check(orc, t2) >= 1, f"op {op}: corrupted output accepted"
        if nrows == 2:
            t2 = t.copy()
            col = ag.MODULAR_AUX_INPUT_LO
            t2[col, r + 1] = (int(t2[col, r + 1]) + 1) % ag.P
            assert _check(orc, t2) >= 1, f"op {op}: corrupted second row accepted"
    # the range counter is constrained too (arithmetic_stark.rs:213-219)
    t2 = t.copy()
    t2[ag.RANGE_COUNTER, 100] = 99
    assert _check(orc, t2) >= 1


def test_unselected_rows_are_unconstrained(orc):
    """`generate_eval_consistency_not_*`: with every IS_* flag zero, garbage in the shared columns satisfies the
    operation constraints (only the range counter is constrained)."""
    rng = np.random.default_rng(3)
    t = np.zeros((54, 1 << 16), dtype=np.uint64)
    t[ag.START_SHARED_COLS:ag.RANGE_COUNTER] = rng.integers(0, ag.P, size=(18, 1 << 16), dtype=np.uint64)
    t[ag.AUX_EXTRA:] = rng.integers(0, ag.P, size=(8, 1 << 16), dtype=np.uint64)
    t[ag.RANGE_COUNTER] = np.arange(1 << 16, dtype=np.uint64)
    assert _check(orc, t) == 0, orc.orc_last_error()


def test_arithmetic_system_proves_and_verifies(orc, arith_ops):
    t = ag.arithmetic_trace(arith_ops)
    proof = binding.prove_system(orc, tr.SYSTEM_ARITH, [t])
    assert binding.verify_system(orc, tr.SYSTEM_ARITH, proof) is None
    # a value outside the 16-bit range passes no range check: the logUp argument breaks and the proof is rejected
    t2 = t.copy()
    assert not t[:26, 65000].any()               # a padding row: no operation constraint looks at it
    t2[ag.IN2, 65000] = 1 << 16
    t2[ag.RC_FREQUENCIES, 0] -= 1
    assert _check(orc, t2) == 0
    bad = binding.prove_system(orc, tr.SYSTEM_ARITH, [t2])
    assert binding.verify_system(orc, tr.SYSTEM_ARITH, bad) is not None
