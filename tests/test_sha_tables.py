"""SHA-256 tables (ShaExtend, ShaExtendSponge, ShaCompress, ShaCompressSponge): the reference's generator fixtures
(`test_correction`: inputs 0,1,2,3 -> w_i = 40965, sha_extend_stark.rs:442-475 and sha_extend_sponge_stark.rs:471-521;
`test_generation`, sha_compress_stark.rs:935-969), generate => constraints-vanish, and the two SHA slices of AllStark with
their real cross-table lookups (Logic XOR/AND, Memory reads) proving and verifying."""
import hashlib
import struct

import numpy as np
import pytest

import hash_gen as hg
import traces as tr
from oracle import binding

IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]


def _check(orc, kind, t):
    return orc.orc_check_table_constraints(kind, binding.col_ptrs(t), t.shape[0], t.shape[1].bit_length() - 1)


def test_sha_extend_reference_fixture():
    row, w_i, _ = hg.sha_extend_row(0, 1, 2, 3, 0)
    assert w_i == 40965
    assert [int(b) for b in row[hg.SE_W_I:hg.SE_W_I + 4]] == list((40965).to_bytes(4, "little"))


def test_rows_compute_sha256():
    """Message schedule rows + compression rows + the sponge's final additions give SHA-256 (pins K, rotations, order)."""
    msg = b"zkm_b200 sha rows"
    block = msg + b"\x80" + b"\x00" * (55 - len(msg)) + struct.pack(">Q", 8 * len(msg))
    w = list(struct.unpack(">16I", block))
    for i in range(16, 64):
        w.append(hg.sha_extend_row(w[i - 15], w[i - 2], w[i - 16], w[i - 7], 0)[1])
    st = list(IV)
    for i in range(64):
        st = hg.sha_compress_row(st, w[i], hg.SHA_K[i], i, 0, 0)[1]
    digest = b"".join(struct.pack(">I", (a + b) & 0xFFFFFFFF) for a, b in zip(IV, st))
    assert digest == hashlib.sha256(msg).digest()


@pytest.fixture(scope="module")
def extend_traces():
    return tr.sha_extend_system_traces()


@pytest.fixture(scope="module")
def compress_traces():
    return tr.sha_compress_system_traces()


def test_sha_extend_traces_satisfy_constraints(orc, extend_traces):
    ext, sp = extend_traces[0], extend_traces[1]
    assert _check(orc, tr.T_SHA_EXTEND, ext) == 0, orc.orc_last_error()
    assert _check(orc, tr.T_SHA_EXTEND_SPONGE, sp) == 0, orc.orc_last_error()
    for col in (hg.SE_W_I, hg.SE_W_I_CARRY + 1, hg.SE_RR7 + 4, hg.SE_RR18 + 5, hg.SE_RS10 + 1, hg.SE_RS3 + 4, hg.SE_W_M7 + 2):
        t2 = ext.copy()
        t2[col, 9] = (int(t2[col, 9]) + 1) % tr.P
        assert _check(orc, tr.T_SHA_EXTEND, t2) >= 1, f"ShaExtend column {col}: corruption accepted"
    for col, r in ((hg.SES_TIMESTAMP, 5), (hg.SES_INPUT_VIRT + 1, 5), (hg.SES_OUTPUT_VIRT, 47), (hg.SES_ROUND + 7, 5)):
        t2 = sp.copy()
        t2[col, r] = (int(t2[col, r]) + 1) % tr.P
        assert _check(orc, tr.T_SHA_EXTEND_SPONGE, t2) >= 1, f"ShaExtendSponge column {col}: corruption accepted"


def test_sha_compress_traces_satisfy_constraints(orc, compress_traces):
    c, sp = compress_traces[0], compress_traces[1]
    assert _check(orc, tr.T_SHA_COMPRESS, c) == 0, orc.orc_last_error()
    assert _check(orc, tr.T_SHA_COMPRESS_SPONGE, sp) == 0, orc.orc_last_error()
    for col in (hg.SC_K_I, hg.SC_E_NOT + 1, hg.SC_E_RR6 + 4, hg.SC_A_RR22 + 5, hg.SC_TEMP1 + 2, hg.SC_TEMP1 + 5, hg.SC_TEMP2,
                hg.SC_D_ADD_TEMP1 + 3, hg.SC_TEMP1_ADD_TEMP2 + 4, hg.SC_STATE + 13, hg.SC_TIMESTAMP, hg.SC_W_I_VIRT):
        t2 = c.copy()
        t2[col, 20] = (int(t2[col, 20]) + 1) % tr.P
        assert _check(orc, tr.T_SHA_COMPRESS, t2) >= 1, f"ShaCompress column {col}: corruption accepted"
    for col in (hg.SCS_HX_VIRT + 3, hg.SCS_OUTPUT_HX + 6 * 2, hg.SCS_OUTPUT_HX + 6 * 5 + 4, hg.SCS_IS_REAL):
        t2 = sp.copy()
        t2[col, 1] = (int(t2[col, 1]) + 1) % tr.P
        assert _check(orc, tr.T_SHA_COMPRESS_SPONGE, t2) >= 1, f"ShaCompressSponge column {col}: corruption accepted"


@pytest.mark.parametrize("sid,which", [(tr.SYSTEM_SHA_EXTEND, "extend"), (tr.SYSTEM_SHA_COMPRESS, "compress")])
def test_sha_systems_prove_and_verify(orc, extend_traces, compress_traces, sid, which):
    traces = extend_traces if which == "extend" else compress_traces
    proof = binding.prove_system(orc, sid, traces)
    assert binding.verify_system(orc, sid, proof) is None
    # a logic row the SHA table never asked for: every table still satisfies its own constraints, the lookup breaks
    ts = [t.copy() for t in traces]
    ts[2][4, 3] ^= 1
    ts[2][68, 3] = int(ts[2][68, 3]) ^ (1 if int(ts[2][2, 3]) else int(ts[2][36, 3]))    # keep the XOR / AND result right
    assert _check(orc, tr.T_LOGIC, ts[2]) == 0
    bad = binding.prove_system(orc, sid, ts)
    assert binding.verify_system(orc, sid, bad) is not None
    # a memory read served with another value
    ts = [t.copy() for t in traces]
    ts[3][6, :4] = int(ts[3][6, 0]) + 1
    assert _check(orc, tr.T_MEMORY, ts[3]) == 0
    bad = binding.prove_system(orc, sid, ts)
    assert binding.verify_system(orc, sid, bad) is not None
