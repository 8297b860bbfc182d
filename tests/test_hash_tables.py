"""Keccak and KeccakSponge tables: the reference's generator tests restated (`keccak_correctness_test`,
keccak_stark.rs:655-687; `test_generation`, keccak_sponge_stark.rs:761-790), generate => constraints-vanish, and the
Keccak slice of AllStark (Keccak, KeccakSponge, Logic, Memory with their real cross-table lookups: 34 XOR lookers,
136 memory-read lookers) proving and verifying."""
import hashlib

import numpy as np
import pytest

import hash_gen as hg
import traces as tr
from oracle import binding


def _check(orc, kind, t):
    return orc.orc_check_table_constraints(kind, binding.col_ptrs(t), t.shape[0], t.shape[1].bit_length() - 1)


def _sha3_256(data):            # SHA3-256 over the restated Keccak-f (same permutation, padding 0x06)
    st, msg = [0] * 25, bytearray(data) + b"\x06"
    msg += b"\x00" * (-len(msg) % 136)
    msg[-1] |= 0x80
    for o in range(0, len(msg), 136):
        for i in range(17):
            st[i] ^= int.from_bytes(msg[o + 8 * i:o + 8 * i + 8], "little")
        st = hg.keccakf(st)
    return b"".join(s.to_bytes(8, "little") for s in st[:4])


def test_keccak_rows_compute_keccak_f():
    for d in (b"", b"abc", bytes(range(200)), b"x" * 136):
        assert _sha3_256(d) == hashlib.sha3_256(d).digest()


def test_sponge_rows_compute_keccak256():
    assert hg.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    # 135-byte input: both padding bits share one byte (keccak_sponge_stark.rs:334-340)
    rows, _ = hg.keccak_sponge_rows_for_op(([0] * 40, 1, b"\x07" * 135, 0, 0))
    assert len(rows) == 1 and int(rows[0][hg.KS_BLOCK_BYTES + 135]) == 0x81


def test_keccak_trace_satisfies_constraints(orc):
    t = hg.random_keccak_trace(7, perms=4)          # 96 rounds + 32 padding rows
    assert _check(orc, tr.T_KECCAK, t) == 0, orc.orc_last_error()
    rng = np.random.default_rng(1)
    for col in (hg.K_START_A + 3, hg.K_START_C + 70, hg.K_START_C_PRIME + 5, hg.K_START_A_PRIME + 777, hg.K_START_A_PP + 12,
                hg.K_START_A_PP_00_BITS + 9, hg.K_A_PPP_00_LO, hg.K_TIMESTAMP, 5):
        t2 = t.copy()
        r = int(rng.integers(1, 90))
        t2[col, r] = (int(t2[col, r]) + 1) % tr.P
        assert _check(orc, tr.T_KECCAK, t2) >= 1, f"column {col} row {r}: corruption accepted"


def test_keccak_sponge_trace_satisfies_constraints(orc):
    ops = hg.random_sponge_ops(7, lens=[0, 4, 132, 136, 272, 140, 408])
    t, perms = hg.keccak_sponge_trace(ops, 4)
    assert len(perms) == 14
    assert _check(orc, tr.T_KECCAK_SPONGE, t) == 0, orc.orc_last_error()
    for col, r in ((hg.KS_ALREADY, 4), (hg.KS_LEN, 4), (hg.KS_IS_FULL, 2), (hg.KS_ORIG_RATE + 3, 5), (hg.KS_ORIG_CAP + 1, 5),
                   (hg.KS_TIMESTAMP, 4), (hg.KS_UPDATED_DIGEST_BYTES + 2, 5), (hg.KS_PARTIAL_UPDATED + 40, 5)):
        t2 = t.copy()
        t2[col, r] = (int(t2[col, r]) + 1) % tr.P
        assert _check(orc, tr.T_KECCAK_SPONGE, t2) >= 1, f"column {col} row {r}: corruption accepted"
    # block bytes, xored rate and the permutation outputs of final rows are bound only through the cross-table lookups
    # (Logic, Memory, Keccak): the table alone accepts a change there, as in the reference (keccak_sponge_stark.rs:456-567)
    t2 = t.copy()
    t2[hg.KS_BLOCK_BYTES + 8, 1] += 1
    assert _check(orc, tr.T_KECCAK_SPONGE, t2) == 0


@pytest.fixture(scope="module")
def keccak_traces():
    return tr.keccak_system_traces()


def test_keccak_system_proves_and_verifies(orc, keccak_traces):
    for kind, t in zip((tr.T_KECCAK, tr.T_KECCAK_SPONGE, tr.T_LOGIC, tr.T_MEMORY), keccak_traces):
        assert _check(orc, kind, t) == 0
    proof = binding.prove_system(orc, tr.SYSTEM_KECCAK, keccak_traces)
    assert binding.verify_system(orc, tr.SYSTEM_KECCAK, proof) is None


@pytest.mark.parametrize("which", ["xor", "read", "perm_output"])
def test_keccak_system_rejects_broken_lookups(orc, keccak_traces, which):
    """Each table stays valid on its own; only the multiset equality across tables breaks."""
    ts = [t.copy() for t in keccak_traces]
    if which == "xor":            # a XOR row the sponge never asked for
        lg = ts[2]
        r = 3
        lg[4, r] ^= 1             # flip bit 0 of input 0 and fix the result
        lg[68, r] ^= 1
        kind, t = tr.T_LOGIC, lg
    elif which == "read":         # a memory read returning another value
        m = ts[3]
        vals = m[6, :5].copy()
        m[6, :4] = (int(vals[0]) + 1)   # the 4 byte-reads of the first word stay consistent with each other
        kind, t = tr.T_MEMORY, m
    else:                         # the permutation table's last-round output changed consistently in the next input? no next: last perm
        k = ts[0]
        last = 24 * 10 - 1        # final round of the last permutation; the next row is padding
        k[hg.reg_a_pp(1, 1), last] ^= 1
        kind, t = tr.T_KECCAK, k
    if which != "perm_output":
        assert _check(orc, kind, t) == 0
    proof = binding.prove_system(orc, tr.SYSTEM_KECCAK, ts)
    assert binding.verify_system(orc, tr.SYSTEM_KECCAK, proof) is not None


# ------------------------------------------------------------------------------------------ Poseidon slice
@pytest.fixture(scope="module")
def poseidon_traces(orc):
    return tr.poseidon_system_traces(orc)


def test_poseidon_sponge_trace_satisfies_constraints(orc, poseidon_traces):
    t = poseidon_traces[1]
    assert t.shape[0] == hg.POSEIDON_SPONGE_COLUMNS
    assert _check(orc, tr.T_POSEIDON_SPONGE, t) == 0, orc.orc_last_error()
    # rows: lens (0, 4, 28, 32, 64, 36, 100) -> final rows 0, 1, 2, 4, 7, 9, 13; full rows 3, 5, 6, 8, 10, 11, 12
    for col, r in ((hg.PS_ALREADY, 4), (hg.PS_LEN, 4), (hg.PS_IS_FULL, 1), (hg.PS_ORIG_RATE + 3, 5), (hg.PS_ORIG_CAP + 1, 6),
                   (hg.PS_TIMESTAMP, 4), (hg.PS_UPDATED_DIGEST + 2, 5), (hg.PS_PARTIAL_UPDATED + 7, 5)):
        t2 = t.copy()
        t2[col, r] = (int(t2[col, r]) + 1) % tr.P
        assert _check(orc, tr.T_POSEIDON_SPONGE, t2) >= 1, f"column {col} row {r}: corruption accepted"


def test_poseidon_system_proves_and_verifies(orc, poseidon_traces):
    for kind, t in zip((tr.T_POSEIDON, tr.T_POSEIDON_SPONGE, tr.T_MEMORY), poseidon_traces):
        assert _check(orc, kind, t) == 0
    proof = binding.prove_system(orc, tr.SYSTEM_POSEIDON_SPONGE, poseidon_traces)
    assert binding.verify_system(orc, tr.SYSTEM_POSEIDON_SPONGE, proof) is None
    # a block byte the memory table never served
    ts = [t.copy() for t in poseidon_traces]
    ts[1][hg.PS_BLOCK_BYTES + 1, 1] ^= 1
    ts[1][hg.PS_NEW_RATE, 1] ^= 1 << 8
    bad = binding.prove_system(orc, tr.SYSTEM_POSEIDON_SPONGE, ts)
    assert binding.verify_system(orc, tr.SYSTEM_POSEIDON_SPONGE, bad) is not None
