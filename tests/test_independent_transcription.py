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


SHA_K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]      # FIPS 180-4


def sha_compress_constraints(lv, nv, yc):
    """sha_compress/sha_compress_stark.rs:399-606; view sha_compress/columns.rs:9-55; gadgets not_operation.rs:23-33,
    wrapping_add_5.rs:37-82, wrapping_add_2.rs, sha_extend/rotate_right.rs, sha_compress/logic.rs:7-16; round constants
    sha_compress_sponge/constants.rs (their little-endian bytes)."""
    STATE, E_NOT, W_I, K_I, S_1, CH, S_0, MAJ = 0, 32, 36, 40, 48, 60, 68, 88
    E_RR_6, E_RR_11, E_RR_25, A_RR_2, A_RR_13, A_RR_22 = 92, 98, 104, 110, 116, 122
    TEMP2, D_ADD_TEMP1, TEMP1_ADD_TEMP2, TIMESTAMP, W_I_VIRT, TEMP1, ROUND = 128, 134, 140, 146, 149, 150, 159
    st = lambda i: STATE + 4 * i
    is_final = lv[ROUND + 64]
    yc.constraint(is_final * (is_final - 1))
    not_final = 1 - is_final
    flags = sum(lv[ROUND + i] for i in range(65))
    yc.constraint(flags * (flags - 1))
    for i in range(4):
        byte_i = sum(lv[ROUND + j] * ((SHA_K[j] >> (8 * i)) & 0xFF) for j in range(64))
        yc.constraint(flags * not_final * (lv[K_I + i] - byte_i))
    for op, inp, r in ((E_RR_6, st(4), 6), (E_RR_11, st(4), 11), (E_RR_25, st(4), 25), (A_RR_2, st(0), 2), (A_RR_13, st(0), 13), (A_RR_22, st(0), 22)):
        for c in _rotate_right(lv, inp, op, r):
            yc.constraint(c)
    for i in range(4):
        yc.constraint(flags * (lv[st(4) + i] + lv[E_NOT + i] - 255))
    for c in _wrapping_add(lv, (st(7), S_1, CH, K_I, W_I), TEMP1, 5):
        yc.constraint(flags * c)
    for inputs, op in (((S_0, MAJ), TEMP2), ((st(3), TEMP1), D_ADD_TEMP1), ((TEMP1, TEMP2), TEMP1_ADD_TEMP2)):
        for c in _wrapping_add(lv, inputs, op, 2):
            yc.constraint(flags * c)
    yc.constraint(flags * not_final * (nv[TIMESTAMP] - lv[TIMESTAMP]))
    yc.constraint(flags * not_final * (nv[W_I_VIRT] - lv[W_I_VIRT] - 4))
    for cur, nxt in ((TEMP1_ADD_TEMP2, st(0)), (st(0), st(1)), (st(1), st(2)), (st(2), st(3)), (D_ADD_TEMP1, st(4)), (st(4), st(5)), (st(5), st(6)),
                     (st(6), st(7))):
        for i in range(4):
            yc.constraint(flags * not_final * (lv[cur + i] - nv[nxt + i]))


# ---------------------------------------------------------------------------------------------------------- Poseidon
def _poseidon_constants():
    """The reference's poseidon/constants.rs tables as extracted into oracle/poseidon_consts.h by tools/gen_poseidon_consts.py
    (data, pinned by the permutation's known answers -- not part of the constraint transcription being cross-checked)."""
    import re
    text = (ROOT / "oracle/poseidon_consts.h").read_text().replace("\\\n", " ")
    out = {}
    for name, body in re.findall(r"#define POSEIDON_(\w+)_INIT \{([^}]*)\}", text):
        out[name] = [int(x.rstrip("ULL"), 0) for x in re.findall(r"0x[0-9a-fA-F]+(?:ULL)?|\b\d+(?:ULL)?", body)]
    return out


def poseidon_constraints(lv, nv, yc):
    """poseidon/poseidon_stark.rs:554-594 with the layer helpers :165-170 (constants), :184-192,245-267 (S-box witness),
    :294-308 (MDS), :371-375,402-414 (first partial-round constants, initial matrix), :436-448,503-519 (partial rounds);
    columns poseidon/columns.rs:3-60."""
    K = _poseidon_constants()
    RC, CIRC, DIAG = K["ALL_ROUND_CONSTANTS"], K["MDS_CIRC"], K["MDS_DIAG"]
    FIRST, PRC, VS, WH, INIT = (K["FAST_PARTIAL_FIRST_ROUND_CONSTANT"], K["FAST_PARTIAL_ROUND_CONSTANTS"], K["FAST_PARTIAL_ROUND_VS"],
                                K["FAST_PARTIAL_ROUND_W_HATS"], K["FAST_PARTIAL_ROUND_INITIAL_MATRIX"])
    assert (len(RC), len(VS), len(WH), len(INIT)) == (360, 242, 242, 121)
    W, HALF, NP = 12, 4, 22
    reg_in = lambda i: 1 + i
    reg_out = lambda i: 1 + W + i
    start_full0 = 1 + 2 * W + 1
    start_partial = start_full0 + 2 * W * HALF
    start_full1 = start_partial + 2 * NP
    state = [lv[reg_in(i)] for i in range(W)]

    def sbox(x, inter, out):
        yc.constraint((x * x * x - inter) % P)
        yc.constraint((x * inter * inter - out) % P)

    def full(r, first, ctr):
        nonlocal state
        state = [(state[i] + RC[i + W * ctr]) % P for i in range(W)]
        base = start_full0 if first else start_full1
        for i in range(W):
            inter, out = lv[base + 2 * W * r + 2 * i], lv[base + 2 * W * r + 2 * i + 1]
            sbox(state[i], inter, out)
            state[i] = out
        state = [(sum(state[(j + i) % W] * CIRC[j] for j in range(W)) + state[i] * DIAG[i]) % P for i in range(W)]
    ctr = 0
    for r in range(HALF):
        full(r, True, ctr)
        ctr += 1
    state = [(state[i] + FIRST[i]) % P for i in range(W)]
    res = [state[0]] + [0] * (W - 1)
    for r in range(1, W):
        for c in range(1, W):
            res[c] = (res[c] + state[r] * INIT[(r - 1) * 11 + (c - 1)]) % P
    state = res

    def partial_mds(r):
        nonlocal state
        d = state[0] * (CIRC[0] + DIAG[0])
        for i in range(1, W):
            d += state[i] * WH[r * 11 + i - 1]
        state = [d % P] + [(state[0] * VS[r * 11 + i - 1] + state[i]) % P for i in range(1, W)]
    for r in range(NP):
        inter, out = lv[start_partial + 2 * r], lv[start_partial + 2 * r + 1]
        sbox(state[0], inter, out)
        state[0] = out
        if r < NP - 1:
            state[0] = (state[0] + PRC[r]) % P
        partial_mds(r)
    ctr += NP
    for r in range(HALF):
        full(r, False, ctr)
        ctr += 1
    for i in range(W):
        yc.constraint(state[i] - lv[reg_out(i)])


# ------------------------------------------------------------------------------------------------------------ Keccak
KECCAK_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
             0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
             0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
             0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]   # FIPS 202
KECCAK_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]                      # keccak/columns.rs:41-47


def keccak_constraints(lv, nv, yc):
    """keccak/keccak_stark.rs:248-415; columns keccak/columns.rs:8-131; xor_gen / xor3_gen / andn_gen keccak/logic.rs:16-23,54-56;
    round-constant bits keccak/constants.rs (bit i of RC[round], least significant first)."""
    ROUNDS, TIMESTAMP = 24, 24
    START_A = 25
    START_C = START_A + 50
    START_CP = START_C + 320
    START_AP = START_CP + 320
    START_APP = START_AP + 1600
    START_APP00 = START_APP + 50
    APPP00_LO = START_APP00 + 64
    assert APPP00_LO + 2 == len(lv) == 2431
    reg_a = lambda x, y: START_A + (x * 5 + y) * 2
    reg_c = lambda x, z: START_C + x * 64 + z
    reg_cp = lambda x, z: START_CP + x * 64 + z
    reg_ap = lambda x, y, z: START_AP + x * 320 + y * 64 + z
    reg_app = lambda x, y: START_APP + x * 10 + y * 2
    reg_appp = lambda x, y: APPP00_LO if (x, y) == (0, 0) else reg_app(x, y)

    def reg_b(x, y, z):
        a, b = (x + 3 * y) % 5, x
        return reg_ap(a, b, (z + 64 - KECCAK_ROT[a][b]) % 64)
    xor = lambda x, y: (x + y - 2 * x * y) % P
    xor3 = lambda x, y, z: xor(x, xor(y, z))
    andn = lambda x, y: (1 - x) * y % P
    limb = lambda bits: sum(b << k for k, b in enumerate(bits))
    filt = lv[ROUNDS - 1]
    yc.constraint(filt * (filt - 1))
    not_final = 1 - lv[ROUNDS - 1]
    yc.constraint(not_final * filt)
    flags = sum(lv[i] for i in range(ROUNDS))
    yc.constraint(flags * not_final * (nv[TIMESTAMP] - lv[TIMESTAMP]))
    for x in range(5):
        for z in range(64):
            yc.constraint(lv[reg_cp(x, z)] - xor3(lv[reg_c(x, z)], lv[reg_c((x + 4) % 5, z)], lv[reg_c((x + 1) % 5, (z + 63) % 64)]))
    for x in range(5):
        for y in range(5):
            bit = lambda z: xor3(lv[reg_ap(x, y, z)], lv[reg_c(x, z)], lv[reg_cp(x, z)])
            yc.constraint(limb([bit(z) for z in range(32)]) - lv[reg_a(x, y)])
            yc.constraint(limb([bit(z) for z in range(32, 64)]) - lv[reg_a(x, y) + 1])
    for x in range(5):
        for z in range(64):
            diff = sum(lv[reg_ap(x, i, z)] for i in range(5)) - lv[reg_cp(x, z)]
            yc.constraint(diff * (diff - 2) * (diff - 4))
    for x in range(5):
        for y in range(5):
            bit = lambda z: xor(lv[reg_b(x, y, z)], andn(lv[reg_b((x + 1) % 5, y, z)], lv[reg_b((x + 2) % 5, y, z)]))
            yc.constraint(limb([bit(z) for z in range(32)]) - lv[reg_app(x, y)])
            yc.constraint(limb([bit(z) for z in range(32, 64)]) - lv[reg_app(x, y) + 1])
    bits00 = [lv[START_APP00 + i] for i in range(64)]
    yc.constraint(limb(bits00[:32]) - lv[reg_app(0, 0)])
    yc.constraint(limb(bits00[32:]) - lv[reg_app(0, 0) + 1])
    xored = lambda i: xor(bits00[i], sum(lv[r] * ((KECCAK_RC[r] >> i) & 1) for r in range(ROUNDS)))
    yc.constraint(limb([xored(z) for z in range(32)]) - lv[APPP00_LO])
    yc.constraint(limb([xored(z) for z in range(32, 64)]) - lv[APPP00_LO + 1])
    not_last = 1 - lv[ROUNDS - 1]
    for x in range(5):
        for y in range(5):
            yc.constraint_transition(not_last * (lv[reg_appp(x, y)] - nv[reg_a(x, y)]))
            yc.constraint_transition(not_last * (lv[reg_appp(x, y) + 1] - nv[reg_a(x, y) + 1]))


# -------------------------------------------------------------------------------------------------------- Arithmetic
def _sign_extend_poly():
    """arithmetic/sra.rs:271-282: the interpolant through (0, 0), (i, 2^31 + .. + 2^(32-i)) for i = 1..31 (plonky2's
    `interpolant` = the unique polynomial of degree < 32), coefficients low to high."""
    pts, acc = [(0, 0)], 0
    for i in range(1, 32):
        acc += 1 << (32 - i)
        pts.append((i, acc))
    coeffs = [0] * 32
    for k, (xk, yk) in enumerate(pts):
        num, den = [1], 1                                   # prod_{j != k} (x - xj), low to high
        for j, (xj, _) in enumerate(pts):
            if j == k:
                continue
            num = [(a - xj * b) % P for a, b in zip([0] + num, num + [0])]
            den = den * (xk - xj) % P
        scale = yk * pow(den, P - 2, P) % P
        for d in range(32):
            coeffs[d] = (coeffs[d] + num[d] * scale) % P
    return coeffs


def arithmetic_constraints(lv, nv, yc):
    """arithmetic/arithmetic_stark.rs:198-229 and the modules it calls, in its order: mul.rs:126-188, mult.rs:158-318,
    addcy.rs:87-160, slt.rs:50-112, lui.rs:51-71, div.rs:361-628, shift.rs:93-134, sra.rs:93-155, lo_hi.rs:25-36; column map
    arithmetic/columns.rs (N_LIMBS = 2), polynomial helpers arithmetic/utils.rs."""
    (IS_ADD, IS_ADDU, IS_ADDI, IS_ADDIU, IS_SUB, IS_SUBU, IS_MULT, IS_MULTU, IS_MUL, IS_DIV, IS_DIVU, IS_SLLV, IS_SRLV, IS_SRAV, IS_SLL, IS_SRL,
     IS_SRA, IS_SLT, IS_SLTU, IS_SLTI, IS_SLTIU, IS_LUI, IS_MFHI, IS_MTHI, IS_MFLO, IS_MTLO) = range(26)
    S = 26
    IN0, IN1, IN2, OUT, AUXIN0, AUXIN1, AUXIN2 = (range(S + 2 * k, S + 2 * k + 2) for k in range(7))
    AUX_REG_0, AUX_REG_1, AUX_REG_2 = range(S, S + 2), range(S + 2, S + 6), range(S + 6, S + 9)
    MOD_OUT_AUX_RED, MOD_IS_ZERO, MOD_AUX_LO, MOD_AUX_HI, DENOM_IS_ZERO = AUX_REG_0, AUX_REG_1[0], range(AUX_REG_1[0] + 1, AUX_REG_1[-1] + 1), AUX_REG_2, AUX_REG_2[-1] + 1
    RANGE_COUNTER = S + 18
    RC_FREQ = RANGE_COUNTER + 1
    AUX_EXTRA = range(RC_FREQ + 1, RC_FREQ + 9)
    OUT_LO, OUT_HI = OUT, range(OUT[-1] + 1, OUT[-1] + 3)
    MULT_AUX_LO = range(OUT_HI[-1] + 1, OUT_HI[-1] + 5)
    MULT_AUX_HI = range(MULT_AUX_LO[-1] + 1, MULT_AUX_LO[-1] + 5)
    QUOT_ABS = range(AUXIN2[-1] + 1, AUXIN2[-1] + 3)
    REM_ABS = range(QUOT_ABS[-1] + 1, QUOT_ABS[-1] + 3)
    assert len(lv) == S + 18 + 10 == 54
    BASE, OFFSET, INV = 1 << 16, 1 << 20, 18446462594437939201
    assert BASE * INV % P == 1
    rd = lambda v, rng: [v[i] for i in rng]

    def mul_lo(a, b):
        return [sum(a[i] * b[d - i] for i in range(d + 1)) for d in range(len(a))]

    def adjoin_root(a, root):
        return [-root * a[0]] + [a[d - 1] - root * a[d] for d in range(1, len(a))]

    def mul_like(filt, left, right, out, aux_lo, aux_hi):          # mul.rs:126-177 / mult.rs:275-318
        aux = [lv[lo] + lv[hi] * BASE - OFFSET for lo, hi in zip(aux_lo, aux_hi)]
        poly = [c - o - r for c, o, r in zip(mul_lo(left, right), out, adjoin_root(aux, BASE))]
        for c in poly:
            yc.constraint(filt * c)

    def addcy(filt, x, y, z, given_cy, two_row):                   # addcy.rs:87-140
        emit = yc.constraint_transition if two_row else yc.constraint
        cy = 0
        for xi, yi, zi in zip(x, y, z):
            t = cy + xi + yi - zi
            emit(filt * t * (BASE - t))
            cy = t * INV
        if not two_row:
            emit(filt * given_cy[0] * (given_cy[0] - 1))
        emit(filt * (cy - given_cy[0]))
        for i in range(1, 2):
            emit(filt * given_cy[i])

    div_like_flags = lv[IS_DIV] + lv[IS_DIVU] + lv[IS_SRL] + lv[IS_SRLV] + lv[IS_SRA] + lv[IS_SRAV]

    def div_helper(filt, num_rng, den_rng, quo_rng, rem_rng):      # div.rs:605-633 with modular_constr_poly :395-469, check_reduced :361-393
        yc.constraint_last_row(filt)
        num, modulus, quo, output = rd(lv, num_rng), rd(lv, den_rng), rd(lv, quo_rng) + [0, 0], rd(lv, rem_rng)
        mod_is_zero = nv[MOD_IS_ZERO]
        yc.constraint_transition(filt * (mod_is_zero * mod_is_zero - mod_is_zero))
        yc.constraint_transition(filt * sum(modulus) * mod_is_zero)
        modulus[0] += mod_is_zero
        denom_is_zero = nv[DENOM_IS_ZERO]
        yc.constraint_transition(filt * (mod_is_zero * div_like_flags - denom_is_zero))
        output[0] += denom_is_zero
        addcy(filt, modulus, rd(nv, MOD_OUT_AUX_RED), output, [1 - mod_is_zero * div_like_flags, 0], True)
        output[0] -= denom_is_zero
        prod = [0] * 5
        for i, qi in enumerate(quo):
            for j, mj in enumerate(modulus):
                prod[i + j] += qi * mj
        for x in prod[4:]:
            yc.constraint_transition(filt * x)
        poly = prod[:4]
        for i, o in enumerate(output):
            poly[i] += o
        aux = [nv[i] - OFFSET for i in MOD_AUX_LO] + [0]
        for k, j in enumerate(MOD_AUX_HI):
            aux[k] += BASE * nv[j]
        poly = [c + r for c, r in zip(poly, adjoin_root(aux, BASE))]
        for i, x in enumerate(num):
            poly[i] -= x
        for c in poly:
            yc.constraint_transition(filt * c)

    # range counter (arithmetic_stark.rs:209-217)
    rc1, rc2 = lv[RANGE_COUNTER], nv[RANGE_COUNTER]
    yc.constraint_first_row(rc1)
    yc.constraint_transition((rc2 - rc1) * (rc2 - rc1) - (rc2 - rc1))
    yc.constraint_last_row(rc1 - ((1 << 16) - 1))
    # mul
    mul_like(lv[IS_MUL], rd(lv, IN0), rd(lv, IN1), rd(lv, OUT), AUXIN0, AUXIN1)
    # mult / multu
    out4 = rd(lv, OUT_LO) + rd(lv, OUT_HI)
    filt = lv[IS_MULT]

    def sign_extend(is_neg_idx, sum_idx, inp):
        is_neg = lv[is_neg_idx]
        yc.constraint(filt * is_neg * (1 - is_neg))
        yc.constraint(filt * (inp[1] + (1 << 15) - lv[sum_idx] - is_neg * BASE))
        return inp + [is_neg * 0xFFFF] * 2
    left = sign_extend(AUX_EXTRA[0], IN2[0], rd(lv, IN0))
    right = sign_extend(AUX_EXTRA[0] + 1, IN2[0] + 1, rd(lv, IN1))
    mul_like(filt, left, right, out4, MULT_AUX_LO, MULT_AUX_HI)
    mul_like(lv[IS_MULTU], rd(lv, IN0) + [0, 0], rd(lv, IN1) + [0, 0], out4, MULT_AUX_LO, MULT_AUX_HI)
    # addcy
    in0, in1, out, aux = rd(lv, IN0), rd(lv, IN1), rd(lv, OUT), rd(lv, AUXIN0)
    addcy(lv[IS_ADD], in0, in1, out, aux, False)
    addcy(lv[IS_SUB], in1, out, in0, aux, False)
    addcy(lv[IS_ADDI], in0, in1, out, aux, False)
    addcy(lv[IS_ADDIU], in0, in1, out, aux, False)
    # slt (slt.rs:50-112: x = in1, y = aux, z = in0, given_cy = AUX_INPUT_REGISTER_1, rd = out)
    filt = lv[IS_SLT] + lv[IS_SLTU] + lv[IS_SLTI] + lv[IS_SLTIU]
    sign = lv[IS_SLT] + lv[IS_SLTI]
    given_cy, rdv = rd(lv, AUXIN1), out
    cy = 0
    for xi, yi, zi in zip(in1, aux, in0):
        t = cy + xi + yi - zi
        yc.constraint(filt * t * (BASE - t))
        cy = t * INV
    yc.constraint(filt * given_cy[0] * (given_cy[0] - 1))
    yc.constraint(filt * (cy - given_cy[0]) * (1 - sign))
    yc.constraint(filt * given_cy[1] * (1 - cy - given_cy[0]))
    yc.constraint_transition(filt * (rdv[0] - given_cy[0]))
    yc.constraint(filt * given_cy[1] * (1 - sign))
    yc.constraint_transition(filt * rdv[1])
    # lui
    mul_like(lv[IS_LUI], rd(lv, IN0), rd(lv, IN1), rd(lv, OUT), AUXIN0, AUXIN1)
    # divu, div (div.rs:481-603)
    div_helper(lv[IS_DIVU], IN0, IN1, OUT, AUXIN0)
    filt = lv[IS_DIV]

    def check_abs(inp, abs_rng, sum_idx, is_neg_idx, borrow_idx):
        is_neg = nv[is_neg_idx]
        yc.constraint_transition(filt * is_neg * (1 - is_neg))
        yc.constraint_transition(filt * (lv[inp[-1]] + (1 << 15) - nv[sum_idx] - is_neg * BASE))
        borrow = nv[borrow_idx]
        yc.constraint_transition(filt * borrow * (1 - borrow))
        neg_inputs = [borrow * BASE - lv[inp[0]], BASE - lv[inp[0] + 1] - borrow]
        for i, j, neg in zip(inp, abs_rng, neg_inputs):
            yc.constraint_transition(filt * (is_neg * neg + (1 - is_neg) * lv[i] - lv[j]))
        return is_neg
    D = DENOM_IS_ZERO
    n0 = check_abs(IN0, IN2, D + 1, D + 5, D + 6)
    n1 = check_abs(IN1, AUXIN2, D + 2, D + 7, D + 8)
    nq = check_abs(OUT_LO, QUOT_ABS, D + 3, RC_FREQ + 1, RC_FREQ + 2)
    nr = check_abs(OUT_HI, REM_ABS, D + 4, RC_FREQ + 3, RC_FREQ + 4)
    same = nv[RC_FREQ + 5]
    yc.constraint_transition(filt * (n0 + n1 - 2 * n0 * n1 - same))
    yc.constraint_transition(filt * (nq - same) * sum(rd(lv, OUT_LO)))
    yc.constraint_transition(filt * (nr - n0) * sum(rd(lv, OUT_HI)))
    div_helper(filt, IN2, AUXIN2, QUOT_ABS, REM_ABS)
    # shift: sll as a multiplication, srl as a division (shift.rs:93-134)
    mul_like(lv[IS_SLL] + lv[IS_SLLV], rd(lv, IN1), rd(lv, IN2), rd(lv, OUT), AUXIN0, AUXIN1)
    div_helper(lv[IS_SRL] + lv[IS_SRLV], IN1, IN2, OUT, AUXIN0)
    # sra (sra.rs:93-155)
    filt = lv[IS_SRA] + lv[IS_SRAV]
    shift = rd(lv, IN0)
    yc.constraint_transition(filt * shift[1])
    is_neg = lv[AUXIN2[-1] + 2]
    yc.constraint_transition(filt * is_neg * (1 - is_neg))
    yc.constraint_transition(filt * (lv[IN1[-1]] + (1 << 15) - lv[AUXIN2[-1] + 1] - is_neg * BASE))
    shift_sq = nv[AUXIN2[-1] + 1]
    yc.constraint_transition(filt * (shift_sq - shift[0] * shift[0]))
    coeffs = _sign_extend_poly()[::-1]
    acc = 0
    for w, k in zip(rd(lv, AUX_EXTRA) + rd(nv, AUX_EXTRA), range(0, 32, 2)):
        yc.constraint_transition(filt * (acc * shift_sq + coeffs[k] * shift[0] + coeffs[k + 1] - w))
        acc = w
    acc_lo, acc_hi = nv[AUXIN2[0]], nv[AUXIN2[0] + 1]
    yc.constraint_transition(filt * (acc_hi * BASE + acc_lo - acc))
    div_helper(filt, IN1, IN2, AUXIN2, AUXIN0)
    for x, y, z in zip(rd(lv, AUXIN2), (acc_lo, acc_hi), rd(lv, OUT)):
        yc.constraint_transition(filt * (x + y * is_neg - z))
    # lo_hi
    filt = lv[IS_MFHI] + lv[IS_MTHI] + lv[IS_MFLO] + lv[IS_MTLO]
    for i, o in zip(rd(lv, IN0), rd(lv, OUT)):
        yc.constraint(filt * (i - o))


def cpu_table_constraints(lv, nv, yc):
    """cpu/cpu_stark.rs:259-284 and the ten modules it calls: tests/cpu_constraints_ref.py."""
    import cpu_constraints_ref
    cpu_constraints_ref.cpu_constraints(lv, nv, yc)


TABLES = {"Memory": (11, 13, memory_constraints), "Logic": (10, 69, logic_constraints), "ShaCompress": (8, 224, sha_compress_constraints),
          "Arithmetic": (0, 54, arithmetic_constraints), "Cpu": (1, 259, cpu_table_constraints),
          "Keccak": (4, 2431, keccak_constraints),
          "Poseidon": (2, 262, poseidon_constraints),
          "KeccakSponge": (5, 470, keccak_sponge_constraints), "PoseidonSponge": (3, 110, poseidon_sponge_constraints),
          "ShaExtend": (6, 78, sha_extend_constraints), "ShaExtendSponge": (7, 76, sha_extend_sponge_constraints),
          "ShaCompressSponge": (9, 127, sha_compress_sponge_constraints)}


def test_all_twelve_tables_are_covered():
    assert sorted(TABLES) == sorted(t["table"] for t in FIX["tables"]) and len(TABLES) == 12


@pytest.mark.parametrize("name", sorted(TABLES))
def test_second_transcription_matches_constraint_by_constraint(orc, name):
    """Stronger than the fold: every single constraint value of the shared C++ transcription (oracle = product) equals the Python
    one at the same position, on the fingerprint frame and on a second frame."""
    import numpy as np
    from oracle import binding
    index, ncols, fn = TABLES[name]
    for extra in (0, 0x777):
        seed = FIX["seed"] + 0x10000 * index + extra
        lv, nv = [splitmix(seed + 2 * c) for c in range(ncols)], [splitmix(seed + 2 * c + 1) for c in range(ncols)]

        class Listing(Consumer):
            def __init__(self):
                super().__init__()
                self.values = []

            def constraint(self, c):
                self.values.append(c % P)
                super().constraint(c)
        yc = Listing()
        fn(lv, nv, yc)
        out = np.zeros(1024, dtype=np.uint64)
        n = orc.orc_table_constraint_values(index, seed, binding.u64ptr(out), out.size)
        assert n == len(yc.values) == expect(name)["num_constraints"]
        assert [int(x) for x in out[:n]] == yc.values


@pytest.mark.parametrize("name", sorted(TABLES))
def test_second_transcription_reproduces_the_fingerprint(name):
    index, ncols, fn = TABLES[name]
    lv, nv = frame(index, ncols)
    yc = Consumer()
    fn(lv, nv, yc)
    want = expect(name)
    assert yc.count == want["num_constraints"]
    assert [a % P for a in yc.acc] == want["acc"]


# ------------------------------------------------------------------------------------- cross-table and in-table lookups
def _fold(values):
    acc = values[0] % P
    for v in values[1:]:
        acc = (acc * FIX["alphas"][0] + v) % P
    return acc


def test_second_transcription_of_the_fifteen_ctls():
    """all_cross_table_lookups() (all_stark.rs:136-542) re-read in tests/ctl_defs_ref.py: the same tables in the same order, the same
    number of columns, and the fold of (filter, columns...) on each table's fingerprint frame equals the fixture entry by entry --
    309 entries (294 looking tables + 15 looked tables), 230 of them in the memory CTL."""
    import ctl_defs_ref as cd
    ctls = cd.all_cross_table_lookups()
    assert len(ctls) == len(FIX["ctls"]) == 15
    checked = 0
    for (looking, looked), want in zip(ctls, FIX["ctls"]):
        assert len(looking) == want["num_looking"], want["index"]
        for e, w in zip(looking + [looked], want["entries"]):
            assert (e["table"], len(e["columns"])) == (w["table"], w["num_columns"]), (want["index"], w)
            lv, nv = frame(cd.TABLES.index(e["table"]), cd.NCOLS[e["table"]])
            assert _fold([e["filter"](lv, nv)] + [c(lv, nv) for c in e["columns"]]) == w["fp"], (want["index"], w)
            checked += 1
    assert checked == 309


def test_second_transcription_of_the_in_table_lookups():
    import ctl_defs_ref as cd
    got = cd.lookups()
    assert [(t, len(cols)) for t, cols, _t, _f in got] == [(w["table"], w["num_columns"]) for w in FIX["lookups"]]
    for (table, cols, table_col, freq_col), w in zip(got, FIX["lookups"]):
        lv, nv = frame(cd.TABLES.index(table), cd.NCOLS[table])
        # filter_columns = [None; n]: an absent filter counts as 1 (cross_table_lookup.rs get_helper_cols / eval_helper_columns)
        values = [0] + [c(lv, nv) for c in cols] + [table_col(lv, nv), freq_col(lv, nv)] + [1] * len(cols)
        assert _fold(values) == w["fp"], table
