"""GPU parity tests (through the C ABI) for the commitment slice: Poseidon permutation, NTT, coset
LDE and Merkle cap — SURVEY §8 rows a1/a2 — against the CPU oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import P, random_columns
from oracle.binding import u64ptr, col_ptrs
from zkm_b200 import lib as zl

pytestmark = pytest.mark.gpu


def _chk(zkm, rc, err):
    zl.check(zkm, rc, err)


def test_poseidon_permutation_matches_oracle(zkm, orc):
    st = random_columns(1000, 12, seed=99)
    st[0] = 0
    st[1] = np.arange(12)
    st[2] = P - 1
    a = st.copy(); b = st.copy()
    err = C.c_void_p()
    _chk(zkm, zkm.zkm_b200_poseidon_permute(u64ptr(a), a.shape[0], C.byref(err)), err)
    orc.orc_poseidon_permute_many(u64ptr(b), b.shape[0])
    assert (a == b).all()
    assert int(a[0][0]) == 0x3c18a9786cb0b359 and int(a[1][0]) == 0xd64e1e3efc5b8e9e    # Appendix D


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 6, 9, 12, 13, 14, 16, 17])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_ntt_matches_oracle(zkm, orc, log_n, kind):
    ncols = 3 if log_n > 12 else 21
    cols = random_columns(ncols, 1 << log_n, seed=1000 + log_n)
    a = cols.copy(); b = cols.copy()
    err = C.c_void_p()
    _chk(zkm, zkm.zkm_b200_ntt(u64ptr(a), ncols, log_n, kind, C.byref(err)), err)
    orc.orc_ntt(u64ptr(b), ncols, log_n, kind)
    assert (a == b).all()


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_large_roundtrip_and_linearity(zkm, log_n):
    """Full-size properties (no CPU oracle needed): ifft(fft(x)) == x and fft(x + y) == fft(x) + fft(y)."""
    n = 1 << log_n
    x = random_columns(2, n, seed=31)
    s = ((x[0].astype(object) + x[1].astype(object)) % P).astype(np.uint64)
    buf = np.stack([x[0], x[1], s])
    err = C.c_void_p()
    _chk(zkm, zkm.zkm_b200_ntt(u64ptr(buf), 3, log_n, 0, C.byref(err)), err)
    lhs = buf[2].astype(object)
    rhs = (buf[0].astype(object) + buf[1].astype(object)) % P
    assert (lhs == rhs).all()
    _chk(zkm, zkm.zkm_b200_ntt(u64ptr(buf), 3, log_n, 1, C.byref(err)), err)
    assert (buf[0] == x[0]).all() and (buf[1] == x[1]).all()


@pytest.mark.parametrize("ncols,log_n,from_values", [(1, 6, 1), (3, 6, 1), (4, 6, 0), (5, 6, 1), (8, 7, 1), (9, 8, 1),
                                                     (13, 10, 1), (54, 12, 1), (69, 13, 1), (259, 14, 1), (4, 14, 0),
                                                     (16, 4, 1), (2431, 6, 1)])
def test_commit_matches_oracle(zkm, orc, ncols, log_n, from_values):
    n = 1 << log_n
    cols = random_columns(ncols, n, seed=7 + ncols)
    cap_g = np.zeros(64, dtype=np.uint64); cap_o = np.zeros(64, dtype=np.uint64)
    t, keep = zl.make_table(cols)
    h = C.c_void_p(); err = C.c_void_p()
    fn = zkm.zkm_b200_commit_values if from_values else zkm.zkm_b200_commit_coeffs
    _chk(zkm, fn(C.byref(t), 2, 4, C.byref(h), u64ptr(cap_g), C.byref(err)), err)
    ho = orc.orc_commit(col_ptrs(cols), ncols, log_n, 2, 4, from_values, u64ptr(cap_o))
    assert ho
    try:
        assert (cap_g == cap_o).all()
        for c in sorted({0, ncols // 2, ncols - 1}):
            cg = np.zeros(n, dtype=np.uint64); co = np.zeros(n, dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_get_coeffs(h, c, u64ptr(cg), C.byref(err)), err)
            orc.orc_batch_get_coeffs(ho, c, u64ptr(co))
            assert (cg == co).all()
            lg = np.zeros(4 * n, dtype=np.uint64); lo = np.zeros(4 * n, dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_get_lde(h, c, u64ptr(lg), C.byref(err)), err)
            orc.orc_batch_get_lde(ho, c, u64ptr(lo))
            assert (lg == lo).all()
        plen = log_n + 2 - 4
        for leaf in (0, 1, 4 * n - 1, (4 * n) // 3):
            rg = np.zeros(ncols, dtype=np.uint64); ro = np.zeros(ncols, dtype=np.uint64)
            sg = np.zeros(max(1, plen * 4), dtype=np.uint64); so = np.zeros(max(1, plen * 4), dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_open(h, leaf, u64ptr(rg), u64ptr(sg), C.byref(err)), err)
            orc.orc_batch_open(ho, leaf, u64ptr(ro), u64ptr(so))
            assert (rg == ro).all() and (sg == so).all()
    finally:
        zkm.zkm_b200_batch_free(h)
        orc.orc_batch_free(ho)


@pytest.mark.gpu
@pytest.mark.parametrize("ncols,log_n", [(3, 5), (135, 12), (20, 14)])
def test_commit_at_the_recursion_blow_up(zkm, orc, ncols, log_n):
    """PolynomialBatch::from_values at rate_bits = 3, cap_height = 4: the shape plonky2's PLONK prover commits with under
    `standard_recursion_config` (135 wire columns; SURVEY §8 f1 -- the recursion prover reuses these kernels at blow-up 8).
    Cap, coefficients, the whole 8n-point LDE of three columns, opened rows and Merkle paths against the oracle."""
    n, rb = 1 << log_n, 3
    cols = random_columns(ncols, n, seed=77 + ncols)
    cap_g = np.zeros(64, dtype=np.uint64); cap_o = np.zeros(64, dtype=np.uint64)
    t, keep = zl.make_table(cols)
    h = C.c_void_p(); err = C.c_void_p()
    _chk(zkm, zkm.zkm_b200_commit_values(C.byref(t), rb, 4, C.byref(h), u64ptr(cap_g), C.byref(err)), err)
    ho = orc.orc_commit(col_ptrs(cols), ncols, log_n, rb, 4, 1, u64ptr(cap_o))
    assert ho
    try:
        assert (cap_g == cap_o).all()
        for c in sorted({0, ncols // 2, ncols - 1}):
            lg = np.zeros(n << rb, dtype=np.uint64); lo = np.zeros(n << rb, dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_get_lde(h, c, u64ptr(lg), C.byref(err)), err)
            orc.orc_batch_get_lde(ho, c, u64ptr(lo))
            assert (lg == lo).all()
        plen = log_n + rb - 4
        for leaf in (0, 1, (n << rb) - 1, (n << rb) // 3):
            rg = np.zeros(ncols, dtype=np.uint64); ro = np.zeros(ncols, dtype=np.uint64)
            sg = np.zeros(max(1, plen * 4), dtype=np.uint64); so = np.zeros(max(1, plen * 4), dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_open(h, leaf, u64ptr(rg), u64ptr(sg), C.byref(err)), err)
            orc.orc_batch_open(ho, leaf, u64ptr(ro), u64ptr(so))
            assert (rg == ro).all() and (sg == so).all()
    finally:
        zkm.zkm_b200_batch_free(h)
        orc.orc_batch_free(ho)


def test_error_reporting(zkm):
    err = C.c_void_p()
    buf = np.zeros(8, dtype=np.uint64)
    rc = zkm.zkm_b200_ntt(u64ptr(buf), 1, 3, 9, C.byref(err))
    assert rc == -1 and err.value
    zkm.zkm_b200_free_string(err)


@pytest.mark.parametrize("log_n", [20, 22])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_ntt_full_size_matches_oracle(zkm, orc, log_n, kind):
    """BASELINE config #2 / #3 row counts (2^20, 2^22): fft, ifft and coset_ifft(7) of two columns, every output word
    compared with the oracle's transform (round trips and linearity alone would survive a consistent index permutation)."""
    cols = random_columns(2, 1 << log_n, seed=2000 + log_n)
    a = cols.copy(); b = cols.copy()
    err = C.c_void_p()
    _chk(zkm, zkm.zkm_b200_ntt(u64ptr(a), 2, log_n, kind, C.byref(err)), err)
    orc.orc_ntt(u64ptr(b), 2, log_n, kind)
    assert (a == b).all()


@pytest.mark.parametrize("ncols,log_n", [(9, 20), (3, 22)])
def test_commit_full_size_matches_oracle(zkm, orc, ncols, log_n):
    """PolynomialBatch::from_values at the benchmark row counts: Merkle cap, coefficients, the whole 4n-point LDE of the
    first and last column, leaf rows and authentication paths, all against the oracle."""
    n = 1 << log_n
    cols = random_columns(ncols, n, seed=77 + log_n)
    cap_g = np.zeros(64, dtype=np.uint64); cap_o = np.zeros(64, dtype=np.uint64)
    t, keep = zl.make_table(cols)
    h = C.c_void_p(); err = C.c_void_p()
    _chk(zkm, zkm.zkm_b200_commit_values(C.byref(t), 2, 4, C.byref(h), u64ptr(cap_g), C.byref(err)), err)
    ho = orc.orc_commit(col_ptrs(cols), ncols, log_n, 2, 4, 1, u64ptr(cap_o))
    assert ho
    try:
        assert (cap_g == cap_o).all()
        for c in (0, ncols - 1):
            cg = np.zeros(n, dtype=np.uint64); co = np.zeros(n, dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_get_coeffs(h, c, u64ptr(cg), C.byref(err)), err)
            orc.orc_batch_get_coeffs(ho, c, u64ptr(co))
            assert (cg == co).all()
            lg = np.zeros(4 * n, dtype=np.uint64); lo = np.zeros(4 * n, dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_get_lde(h, c, u64ptr(lg), C.byref(err)), err)
            orc.orc_batch_get_lde(ho, c, u64ptr(lo))
            assert (lg == lo).all()
        plen = log_n + 2 - 4
        for leaf in (0, 1, 4 * n - 1, (4 * n) // 3, 2 * n + 12345):
            rg = np.zeros(ncols, dtype=np.uint64); ro = np.zeros(ncols, dtype=np.uint64)
            sg = np.zeros(plen * 4, dtype=np.uint64); so = np.zeros(plen * 4, dtype=np.uint64)
            _chk(zkm, zkm.zkm_b200_batch_open(h, leaf, u64ptr(rg), u64ptr(sg), C.byref(err)), err)
            orc.orc_batch_open(ho, leaf, u64ptr(ro), u64ptr(so))
            assert (rg == ro).all() and (sg == so).all()
    finally:
        zkm.zkm_b200_batch_free(h)
        orc.orc_batch_free(ho)
