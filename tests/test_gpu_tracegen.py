"""Device-side generators of the six hash-precompile tables (zkm_b200/csrc/tracegen_hash.cu) against the Python
restatements of the reference generators in tests/hash_gen.py, and the drop-in prove call with EVERY table but Cpu entering as
its operation log (zkm_b200_prove_with_ops) against the proof over the host-built tables."""
import numpy as np
import pytest

import hash_gen as hg
import traces as tr
from oracle import binding
from zkm_b200 import lib as zl

pytestmark = pytest.mark.gpu


def _height(rows, min_rows=64):
    return max(min_rows, 1 << max(0, (rows - 1).bit_length()))


def _same(got, want):
    assert got.shape == want.shape, (got.shape, want.shape)
    bad = np.argwhere(got != want)
    assert bad.size == 0, [(int(c), int(r), int(got[c, r]), int(want[c, r])) for c, r in bad[:8]]


def _word(row, at):
    return int(row[at]) | int(row[at + 1]) << 8 | int(row[at + 2]) << 16 | int(row[at + 3]) << 24


def test_sha_extend_tables_generated_on_the_device(zkm, orc):
    """ShaExtend (78 columns) and ShaExtendSponge (76 columns), sha_extend_stark.rs:122-237 / sha_extend_sponge_stark.rs:128-227:
    0, 1 and 3 message schedules (48 rows each), zero padding, minimum height; constraints vanish on the generated tables."""
    for count in (0, 1, 3):
        seqs = [(0x4000 + 0x400 * k, 1000 + 2000 * k) for k in range(count)]
        ext_rows, sp_rows, _x, _m = hg.sha_extend_sequences(seqs, seed=31 + count)
        ins = [[_word(r, hg.SE_W_M15), _word(r, hg.SE_W_M2), _word(r, hg.SE_W_M16), _word(r, hg.SE_W_M7)] for r in ext_rows]
        ops6 = np.array([i + [int(r[hg.SE_TIMESTAMP])] for i, r in zip(ins, ext_rows)], dtype=np.uint64).reshape(len(ins), 5)
        ops7 = np.array([[int(np.argmax(s[:48]))] + i + [int(v) for v in s[hg.SES_INPUT_VIRT:hg.SES_INPUT_VIRT + 4]] +
                         [int(s[hg.SES_OUTPUT_VIRT]), int(s[hg.SES_CONTEXT]), int(s[hg.SES_SEGMENT]), int(s[hg.SES_TIMESTAMP])]
                         for i, s in zip(ins, sp_rows)], dtype=np.uint64).reshape(len(ins), 13)
        lg = _height(len(ins)).bit_length() - 1
        got6, got7 = zl.table_from_ops(zkm, 6, ops6), zl.table_from_ops(zkm, 7, ops7)
        _same(got6, hg.rows_to_trace(ext_rows, hg.SHA_EXTEND_COLUMNS, lg))
        _same(got7, hg.rows_to_trace(sp_rows, hg.SHA_EXTEND_SPONGE_COLUMNS, lg))
    assert orc.orc_check_table_constraints(6, binding.col_ptrs(got6), 78, lg) == 0
    assert orc.orc_check_table_constraints(7, binding.col_ptrs(got7), 76, lg) == 0
    with pytest.raises(zl.ZkmError, match="32-bit"):
        zl.table_from_ops(zkm, 6, np.array([[1 << 32, 0, 0, 0, 0]], dtype=np.uint64))
    with pytest.raises(zl.ZkmError, match="out of range"):
        zl.table_from_ops(zkm, 7, np.array([[48] + [0] * 12], dtype=np.uint64))


def test_sha_compress_tables_generated_on_the_device(zkm, orc):
    """ShaCompress (224 columns, one row per round as the reference's generator takes them: ([u8; 41], address, timestamp)) and
    ShaCompressSponge (127 columns, the 64-round compression recomputed on the device), sha_compress_stark.rs:234-391 /
    sha_compress_sponge_stark.rs:120-237."""
    for count in (0, 1, 2):
        calls = [(0x8000 + 0x40 * k, 0x9000 + 0x200 * k, 500 + 100 * k) for k in range(count)]
        c_rows, s_rows, _l, _m, io = hg.sha_compressions(calls, seed=41 + count)
        ops8 = np.array([[_word(r, hg.SC_STATE + 4 * i) for i in range(8)] + [_word(r, hg.SC_W_I), _word(r, hg.SC_K_I), int(np.argmax(r[hg.SC_ROUND:])),
                         int(r[hg.SC_W_I_VIRT]), int(r[hg.SC_SEGMENT]), int(r[hg.SC_CONTEXT]), int(r[hg.SC_TIMESTAMP])] for r in c_rows],
                        dtype=np.uint64).reshape(len(c_rows), 15)
        ops9 = np.array([list(hx) + list(w) + [h + 4 * j for j in range(8)] + [wp, 0, 0, 0, 0, ts] for (hx, w, _o), (h, wp, ts) in zip(io, calls)],
                        dtype=np.uint64).reshape(count, 86)
        got8, got9 = zl.table_from_ops(zkm, 8, ops8), zl.table_from_ops(zkm, 9, ops9)
        lg8, lg9 = _height(len(c_rows)).bit_length() - 1, _height(count).bit_length() - 1
        _same(got8, hg.rows_to_trace(c_rows, hg.SHA_COMPRESS_COLUMNS, lg8))
        _same(got9, hg.rows_to_trace(s_rows, hg.SHA_COMPRESS_SPONGE_COLUMNS, lg9))
    assert orc.orc_check_table_constraints(8, binding.col_ptrs(got8), 224, lg8) == 0
    assert orc.orc_check_table_constraints(9, binding.col_ptrs(got9), 127, lg9) == 0
    with pytest.raises(zl.ZkmError, match="out of range"):
        zl.table_from_ops(zkm, 8, np.array([[0] * 10 + [65, 0, 0, 0, 0]], dtype=np.uint64))


def test_byte_sponge_tables_generated_on_the_device(zkm, orc):
    """KeccakSponge (470 columns, rate 136 bytes, xor absorb) and PoseidonSponge (110 columns, rate 32 bytes, overwrite absorb)
    from the variable-width log of (addresses, timestamp, input bytes): keccak_sponge_stark.rs:222-447 /
    poseidon_sponge_stark.rs:187-365.  Lengths cover the empty input, rate - 1 (both pad bits in one byte), rate, rate + 1 and
    several blocks; the final digest rows equal keccak256 (the reference's test_generation)."""
    for table, rate, ncols in ((5, 136, 470), (3, 32, 110)):
        for lens in ([], [0], [rate - 1], [rate], [rate + 1, 3], [4, 8, 2 * rate - 1, 3 * rate, 0, 5 * rate + 17, rate - 2]):
            ops = hg.random_sponge_ops(len(lens), seed=22 + len(lens), lens=lens)
            rows = sum(ln // rate + 1 for ln in lens)
            lg = _height(rows).bit_length() - 1
            want = hg.keccak_sponge_trace(ops, lg)[0] if table == 5 else hg.poseidon_sponge_trace(orc, ops, lg)[0]
            log, n_ops = zl.sponge_log(ops)
            got = zl.table_from_ops(zkm, table, log, n_ops=n_ops)
            assert got.shape == (ncols, 1 << lg)
            _same(got, want)
        assert orc.orc_check_table_constraints(table, binding.col_ptrs(got), ncols, lg) == 0
    # a truncated log is an error, not an out-of-bounds read
    log, n_ops = zl.sponge_log(hg.random_sponge_ops(2, seed=5, lens=[40, 300]))
    bad = log.copy()
    bad[0] -= 3
    with pytest.raises(zl.ZkmError, match="truncated|does not match"):
        zl.table_from_ops(zkm, 5, bad, n_ops=n_ops)
    # keccak256 through the generated rows (keccak_sponge_stark.rs:761-790 test_generation)
    data = bytes(range(200))
    t = zl.table_from_ops(zkm, 5, zl.sponge_log([([4 * k for k in range(51)], 7, data, 0, 0)])[0], n_ops=1)
    digest = bytes(int(b) for b in t[hg.KS_UPDATED_DIGEST_BYTES:hg.KS_UPDATED_DIGEST_BYTES + 32, 1])
    assert digest == hg.keccak256(data)


def test_prove_with_every_table_but_cpu_from_its_log(zkm, orc):
    """zkm_b200_prove_with_ops with all eleven operation logs: Arithmetic, Poseidon, PoseidonSponge, Keccak, KeccakSponge, the
    four SHA tables, Logic and Memory are generated on the device inside the prove call and only the Cpu table crosses PCIe as
    a table.  The proof equals the proof over the host-built tables of the same valid 12-table trace, and verifies."""
    traces, ops = tr.all_stark_valid_traces(orc, return_ops=True)
    for t in (3, 5):
        ops[t] = zl.sponge_log(ops[t])
    for t, log in ops.items():
        got = zl.table_from_ops(zkm, t, log[0], n_ops=log[1]) if isinstance(log, tuple) else zl.table_from_ops(zkm, t, log)
        _same(got, traces[t])
    ref = zl.prove_with_traces(zkm, traces)
    got = zl.prove_with_ops(zkm, traces, ops)
    assert got.size == ref.size and (got == ref).all()
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, got) is None
