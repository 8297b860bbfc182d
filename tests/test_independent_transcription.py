"""An INDEPENDENT second transcription of some tables' constraints, written in Python straight from the reference's
`eval_packed_generic` bodies (not from zkm_b200/csrc/tables/*.h, which oracle and product share), evaluated on the fingerprint
frame of tests/golden/constraint_fingerprints_v1.json and folded in emission order with the same alphas.  Equality with the
fixture pins those tables' transcription -- every constraint, its coefficients, its kind (plain / transition / first / last row)
and its POSITION -- to a second reading of the reference, without cargo.  (VERDICT r1 weak 1b / ADVICE r1 low 3.)"""
import json
import pathlib

import pytest

P = 0xFFFFFFFF00000001
ROOT = pathlib.Path(__file__).resolve().parent.parent
FIX = json.loads((ROOT / "tests/golden/constraint_fingerprints_v1.json").read_text())
M64 = (1 << 64) - 1


def splitmix(x):
    z = (x + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return (z ^ (z >> 31)) % P


class Consumer:
    """constraint_consumer.rs:52-75 with the fixture's stand-ins for z_last / lagrange_first / lagrange_last."""

    def __init__(self):
        self.alphas, self.acc, self.count = FIX["alphas"], [0, 0], 0
        self.z_last, self.l_first, self.l_last = FIX["z_last"], FIX["lagrange_first"], FIX["lagrange_last"]

    def constraint(self, c):
        self.count += 1
        self.acc = [(a * al + c) % P for a, al in zip(self.acc, self.alphas)]

    def constraint_transition(self, c):
        self.constraint(c * self.z_last % P)

    def constraint_first_row(self, c):
        self.constraint(c * self.l_first % P)

    def constraint_last_row(self, c):
        self.constraint(c * self.l_last % P)


def frame(table_index, ncols):
    seed = FIX["seed"] + 0x10000 * table_index
    return [splitmix(seed + 2 * c) for c in range(ncols)], [splitmix(seed + 2 * c + 1) for c in range(ncols)]


def expect(name):
    return next(t for t in FIX["tables"] if t["table"] == name)


# ------------------------------------------------------------------------------------------------------------ Memory
def memory_constraints(lv, nv, yc):
    """memory/memory_stark.rs:255-341, columns memory/columns.rs:7-37 (VALUE_LIMBS = 1)."""
    FILTER, TIMESTAMP, IS_READ, CTX, SEG, VIRT, VALUE0, CFC, SFC, VFC, RANGE_CHECK = range(11)
    one = 1
    filt = lv[FILTER]
    yc.constraint(filt * (filt - 1))
    cfc, sfc, vfc = lv[CFC], lv[SFC], lv[VFC]
    unchanged = one - cfc - sfc - vfc
    yc.constraint(cfc * (one - cfc))
    yc.constraint(sfc * (one - sfc))
    yc.constraint(vfc * (one - vfc))
    yc.constraint(unchanged * (one - unchanged))
    yc.constraint_transition(sfc * (nv[CTX] - lv[CTX]))
    yc.constraint_transition(vfc * (nv[CTX] - lv[CTX]))
    yc.constraint_transition(vfc * (nv[SEG] - lv[SEG]))
    yc.constraint_transition(unchanged * (nv[CTX] - lv[CTX]))
    yc.constraint_transition(unchanged * (nv[SEG] - lv[SEG]))
    yc.constraint_transition(unchanged * (nv[VIRT] - lv[VIRT]))
    computed = (cfc * (nv[CTX] - lv[CTX] - one) + sfc * (nv[SEG] - lv[SEG] - one) + vfc * (nv[VIRT] - lv[VIRT] - one)
                + unchanged * (nv[TIMESTAMP] - lv[TIMESTAMP]))
    yc.constraint_transition(lv[RANGE_CHECK] - computed)
    yc.constraint_transition(nv[IS_READ] * unchanged * (nv[VALUE0] - lv[VALUE0]))


# ------------------------------------------------------------------------------------------------------------- Logic
def logic_constraints(lv, nv, yc):
    """logic.rs:186-240, columns logic.rs:26-50 (VAL_BITS = 32, PACKED_LIMB_BITS = 32: one result limb)."""
    IS_AND, IS_OR, IS_XOR, IS_NOR = 0, 1, 2, 3
    INPUT0, INPUT1, RESULT = range(4, 36), range(36, 68), 68
    sum_coeff = lv[IS_OR] + lv[IS_XOR] - lv[IS_NOR]
    and_coeff = lv[IS_AND] - lv[IS_OR] - lv[IS_XOR] * 2 + lv[IS_NOR]
    not_coeff = lv[IS_NOR]
    for cols in (INPUT0, INPUT1):
        for i in cols:
            yc.constraint(lv[i] * (lv[i] - 1))
    x = sum(lv[c] << k for k, c in enumerate(INPUT0))
    y = sum(lv[c] << k for k, c in enumerate(INPUT1))
    x_land_y = sum(lv[a] * lv[b] * (1 << k) for k, (a, b) in enumerate(zip(INPUT0, INPUT1)))
    yc.constraint(lv[RESULT] - (sum_coeff * (x + y) + and_coeff * x_land_y + not_coeff * 0xFFFFFFFF))


# ---------------------------------------------------------------------------------------------- the two byte sponges
def _byte_sponge_constraints(lv, nv, yc, rate_bytes, rate_words, cap_words, digest_words, digest_as_bytes):
    """keccak_sponge/keccak_sponge_stark.rs:456-567 and poseidon_sponge/poseidon_sponge_stark.rs:374-478 (the same body up to
    the sizes and the digest representation); views keccak_sponge/columns.rs:19-70, poseidon_sponge/columns.rs:19-68."""
    at = 0

    def take(k):
        nonlocal at
        r = range(at, at + k)
        at += k
        return r
    (IS_FULL,), (CONTEXT,), (SEGMENT,) = take(1), take(1), take(1)
    take(rate_words)                                                 # virt
    (TIMESTAMP,), (LEN,), (ALREADY,) = take(1), take(1), take(1)
    FINAL_LEN, ORIG_RATE, ORIG_CAP = take(rate_bytes), take(rate_words), take(cap_words)
    take(rate_bytes)                                                 # block_bytes
    take(rate_words)                                                 # xored_rate_u32s / new_rate
    PARTIAL = take(rate_words + cap_words - digest_words)
    DIGEST = take(4 * digest_words if digest_as_bytes else digest_words)
    assert at == len(lv)
    is_full = lv[IS_FULL]
    yc.constraint(is_full * (is_full - 1))
    is_final = sum(lv[c] for c in FINAL_LEN)
    yc.constraint(is_final * (is_final - 1))
    for c in FINAL_LEN:
        yc.constraint(lv[c] * (lv[c] - 1))
    yc.constraint(is_final * is_full)
    yc.constraint_first_row(lv[ALREADY])
    for c in list(ORIG_RATE) + list(ORIG_CAP):
        yc.constraint_first_row(lv[c])
    yc.constraint_transition(is_final * nv[ALREADY])
    for c in list(ORIG_RATE) + list(ORIG_CAP):
        yc.constraint_transition(is_final * nv[c])
    yc.constraint_transition(is_full * (lv[CONTEXT] - nv[CONTEXT]))
    yc.constraint_transition(is_full * (lv[SEGMENT] - nv[SEGMENT]))
    yc.constraint_transition(is_full * (lv[TIMESTAMP] - nv[TIMESTAMP]))
    for k in range(digest_words):
        after = sum(lv[DIGEST[4 * k + i]] << (8 * i) for i in range(4)) if digest_as_bytes else lv[DIGEST[k]]
        yc.constraint_transition(is_full * (nv[ORIG_RATE[k]] - after))
    for cur, nxt in zip(PARTIAL, list(ORIG_RATE)[digest_words:]):
        yc.constraint_transition(is_full * (nv[nxt] - lv[cur]))
    for cur, nxt in zip(list(PARTIAL)[rate_words - digest_words:], ORIG_CAP):
        yc.constraint_transition(is_full * (nv[nxt] - lv[cur]))
    yc.constraint_transition(is_full * (lv[ALREADY] + rate_bytes - nv[ALREADY]))
    is_dummy = 1 - is_full - is_final
    yc.constraint_transition(is_dummy * (nv[IS_FULL] + sum(nv[c] for c in FINAL_LEN)))
    offset = lv[LEN] - lv[ALREADY]
    for i, c in enumerate(FINAL_LEN):
        yc.constraint(lv[c] * (offset - i))


def keccak_sponge_constraints(lv, nv, yc):
    _byte_sponge_constraints(lv, nv, yc, rate_bytes=136, rate_words=34, cap_words=16, digest_words=8, digest_as_bytes=True)


def poseidon_sponge_constraints(lv, nv, yc):
    _byte_sponge_constraints(lv, nv, yc, rate_bytes=32, rate_words=8, cap_words=4, digest_words=4, digest_as_bytes=False)


# ------------------------------------------------------------------------------------------------------ SHA-256 tables
def _le(v, at):
    return v[at] + (v[at + 1] << 8) + (v[at + 2] << 16) + (v[at + 3] << 24)


def _rotate_right(v, inp, op, r):          # sha_extend/rotate_right.rs:29-60; RotateRightOp = value[4], shift, carry
    return [_le(v, op) - v[op + 5] * (1 << (32 - r)) - v[op + 4], _le(v, inp) - v[op + 4] * (1 << r) - v[op + 5]]


def _shift_right(v, inp, op, r):           # sha_extend/shift_right.rs:29-56
    return [_le(v, op) - v[op + 4], _le(v, inp) - v[op + 4] * (1 << r) - v[op + 5]]


def _wrapping_add(v, inputs, op, ncarry):  # wrapping_add_2.rs:33-66 / wrapping_add_4.rs:33-75: value[4], carry[ncarry]
    out = [v[op + 4 + i] * (1 - v[op + 4 + i]) for i in range(ncarry)]
    out.append(sum(v[op + 4 + i] for i in range(ncarry)) - 1)
    carry = sum(i * v[op + 4 + i] for i in range(1, ncarry))
    out.append(sum(_le(v, a) for a in inputs) - carry * (1 << 32) - _le(v, op))
    return out


def sha_extend_constraints(lv, nv, yc):
    """sha_extend/sha_extend_stark.rs:246-321; view sha_extend/columns.rs:8-35."""
    W_I, W15, W2, W16, W7, S0, S1 = 0, 8, 12, 16, 20, 28, 36
    RR7, RR18, RR17, RR19, RS10, RS3, IS_REAL = 40, 46, 52, 58, 64, 70, 77
    for c in (_rotate_right(lv, W15, RR7, 7) + _rotate_right(lv, W15, RR18, 18) + _rotate_right(lv, W2, RR17, 17) + _rotate_right(lv, W2, RR19, 19)
              + _shift_right(lv, W15, RS3, 3) + _shift_right(lv, W2, RS10, 10)):
        yc.constraint(c)
    for c in _wrapping_add(lv, (S1, W7, S0, W16), W_I, 4):
        yc.constraint(c * lv[IS_REAL])


def sha_extend_sponge_constraints(lv, nv, yc):
    """sha_extend_sponge/sha_extend_sponge_stark.rs:229-327; view columns.rs:7-33; NUM_CHANNELS = 10 (cpu/membus.rs)."""
    ROUND, INPUT_VIRT, OUTPUT_VIRT, TIMESTAMP = range(48), range(68, 72), 72, 75
    for i in ROUND:
        yc.constraint(lv[i] * (lv[i] - 1))
    is_final = lv[47]
    yc.constraint(is_final * (is_final - 1))
    not_final = 1 - is_final
    flags = sum(lv[i] for i in ROUND)
    yc.constraint(flags * not_final * (nv[TIMESTAMP] - lv[TIMESTAMP] - 2 * 10))
    yc.constraint(flags * not_final * (sum(nv[i] * i for i in ROUND) - sum(lv[i] * i for i in ROUND) - 1))
    for c in INPUT_VIRT:
        yc.constraint(flags * not_final * (nv[c] - lv[c] - 4))
    yc.constraint(flags * not_final * (nv[OUTPUT_VIRT] - lv[OUTPUT_VIRT] - 4))
    base = lv[INPUT_VIRT[2]]
    yc.constraint(flags * (lv[INPUT_VIRT[0]] - base - 4))
    yc.constraint(flags * (lv[INPUT_VIRT[1]] - base - 56))
    yc.constraint(flags * (lv[INPUT_VIRT[3]] - base - 36))
    yc.constraint(flags * (lv[OUTPUT_VIRT] - base - 64))


def sha_compress_sponge_constraints(lv, nv, yc):
    """sha_compress_sponge/sha_compress_sponge_stark.rs:241-280; view columns.rs:6-25 (output_hx = 8 x WrappingAdd2Op)."""
    HX, OUTPUT_STATE, OUTPUT_HX, HX_VIRT, IS_REAL = 0, 32, 64, 112, 126
    real = lv[IS_REAL]
    yc.constraint(real * (real - 1))
    for i in range(7):
        yc.constraint(real * (lv[HX_VIRT + i + 1] - lv[HX_VIRT + i] - 4))
    for i in range(8):
        for c in _wrapping_add(lv, (HX + 4 * i, OUTPUT_STATE + 4 * i), OUTPUT_HX + 6 * i, 2):
            yc.constraint(c * real)


TABLES = {"Memory": (11, 13, memory_constraints), "Logic": (10, 69, logic_constraints),
          "KeccakSponge": (5, 470, keccak_sponge_constraints), "PoseidonSponge": (3, 110, poseidon_sponge_constraints),
          "ShaExtend": (6, 78, sha_extend_constraints), "ShaExtendSponge": (7, 76, sha_extend_sponge_constraints),
          "ShaCompressSponge": (9, 127, sha_compress_sponge_constraints)}


@pytest.mark.parametrize("name", sorted(TABLES))
def test_second_transcription_reproduces_the_fingerprint(name):
    index, ncols, fn = TABLES[name]
    lv, nv = frame(index, ncols)
    yc = Consumer()
    fn(lv, nv, yc)
    want = expect(name)
    assert yc.count == want["num_constraints"]
    assert [a % P for a in yc.acc] == want["acc"]
