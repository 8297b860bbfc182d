"""CPU tests: the oracle's field / Poseidon / NTT / Merkle layer against the reference's in-tree
pins: Poseidon known answers (Appendix D), GOLDILOCKS_INVERSE_2EXP32
(reference prover/src/cpu/jumps.rs:15), naive == fast partial-round schedule, transform round trips."""
import ctypes as C

import numpy as np

from conftest import P, random_columns
from oracle.binding import u64ptr, col_ptrs

KAT_ZERO = [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca,
            0xd7709673896996dc, 0x46a84e87642f44ed, 0xd032648251ee0b3c, 0x1c687363b207df62,
            0xdf8565563e8045fe, 0x40f5b37ff4254dae, 0xd070f637b431067c, 0x1792b1c4342109d7]
KAT_IOTA = [0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2,
            0x414754bfebd051f0, 0xcb1f8980294a023f, 0x6eb2a9e4d54a9d0f, 0x1902bc3af467e056,
            0xf045d5eafdc6021f, 0xe4150f77caaa3be5, 0xc9bfd01d39b50cce, 0x5c0a27fcb0e1459b]


def test_field_pins(orc):
    # 2^-32 mod p, reference prover/src/cpu/jumps.rs:15
    assert orc.orc_inv(1 << 32) == 18446744065119617026
    assert orc.orc_mul(18446744065119617026, 1 << 32) == 1
    # POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32)
    assert pow(7, (P - 1) >> 32, P) == 1753635133440165772
    assert orc.orc_root_of_unity(32) == 1753635133440165772
    assert orc.orc_root_of_unity(1) == P - 1
    rng = np.random.default_rng(1)
    for _ in range(200):
        a, b = (int(x) % P for x in rng.integers(0, 2**63, 2, dtype=np.uint64) * 2 + 1)
        assert orc.orc_mul(a, b) == a * b % P
        assert orc.orc_add(a, b) == (a + b) % P
        assert orc.orc_sub(a, b) == (a - b) % P
    for a, b in [(P - 1, P - 1), (0, P - 1), (P - 1, 1), (0xFFFFFFFF, 0xFFFFFFFF00000000)]:
        assert orc.orc_mul(a, b) == a * b % P
        assert orc.orc_add(a, b) == (a + b) % P
        assert orc.orc_sub(a, b) == (a - b) % P


def test_ext_field(orc):
    # X^2 = 7
    a = np.array([0, 1], dtype=np.uint64); out = np.zeros(2, dtype=np.uint64)
    orc.orc_ext_mul(u64ptr(a), u64ptr(a), u64ptr(out))
    assert list(out) == [7, 0]
    x = np.array([123456789, 987654321], dtype=np.uint64); inv = np.zeros(2, dtype=np.uint64)
    orc.orc_ext_inv(u64ptr(x), u64ptr(inv))
    orc.orc_ext_mul(u64ptr(x), u64ptr(inv), u64ptr(out))
    assert list(out) == [1, 0]


def test_poseidon_known_answers(orc):
    for fast in (0, 1):
        s = np.zeros(12, dtype=np.uint64)
        orc.orc_poseidon_permute(u64ptr(s), fast)
        assert [int(x) for x in s] == KAT_ZERO
        s = np.arange(12, dtype=np.uint64)
        orc.orc_poseidon_permute(u64ptr(s), fast)
        assert [int(x) for x in s] == KAT_IOTA


def test_poseidon_naive_equals_fast(orc):
    st = random_columns(50, 12, seed=77)
    for row in st:
        a = row.copy(); b = row.copy()
        orc.orc_poseidon_permute(u64ptr(a), 0)
        orc.orc_poseidon_permute(u64ptr(b), 1)
        assert (a == b).all()


def test_hashing_modes(orc):
    out = np.zeros(4, dtype=np.uint64)
    v = np.array([5, 6, 7], dtype=np.uint64)
    orc.orc_hash_or_noop(u64ptr(v), 3, u64ptr(out))
    assert list(out) == [5, 6, 7, 0]                       # <= 4 elements: no permutation
    v = np.arange(8, dtype=np.uint64)
    orc.orc_hash_or_noop(u64ptr(v), 8, u64ptr(out))
    s = np.zeros(12, dtype=np.uint64); s[:8] = v
    orc.orc_poseidon_permute(u64ptr(s), 0)
    assert (out == s[:4]).all()
    # 11 elements: two permutations, second chunk overwrites only 3 words
    v = np.arange(100, 111, dtype=np.uint64)
    orc.orc_hash_or_noop(u64ptr(v), 11, u64ptr(out))
    s = np.zeros(12, dtype=np.uint64); s[:8] = v[:8]
    orc.orc_poseidon_permute(u64ptr(s), 0)
    s[:3] = v[8:]
    orc.orc_poseidon_permute(u64ptr(s), 0)
    assert (out == s[:4]).all()
    l = np.arange(4, dtype=np.uint64); r = np.arange(4, 8, dtype=np.uint64)
    orc.orc_two_to_one(u64ptr(l), u64ptr(r), u64ptr(out))
    s = np.zeros(12, dtype=np.uint64); s[:8] = np.arange(8)
    orc.orc_poseidon_permute(u64ptr(s), 0)
    assert (out == s[:4]).all()


def _naive_dft(v, w):
    n = len(v)
    return [sum(int(v[i]) * pow(w, i * k, P) for i in range(n)) % P for k in range(n)]


def test_ntt_matches_naive_dft_and_roundtrips(orc):
    for log_n in (0, 1, 3, 5):
        n = 1 << log_n
        cols = random_columns(2, n, seed=11 + log_n)
        d = cols.copy()
        orc.orc_ntt(u64ptr(d), 2, log_n, 0)
        w = orc.orc_root_of_unity(log_n)
        for c in range(2):
            assert [int(x) for x in d[c]] == _naive_dft(cols[c], w)
        orc.orc_ntt(u64ptr(d), 2, log_n, 1)
        assert (d == cols).all()
    cols = random_columns(3, 1 << 10, seed=5)
    d = cols.copy()
    orc.orc_ntt(u64ptr(d), 3, 10, 3)      # coset_fft(7)
    orc.orc_ntt(u64ptr(d), 3, 10, 2)      # coset_ifft(7)
    assert (d == cols).all()


def test_commit_lde_is_low_degree_extension(orc):
    log_n, ncols = 5, 6
    n = 1 << log_n
    cols = random_columns(ncols, n, seed=3)
    cap = np.zeros(16 * 4, dtype=np.uint64)
    ptrs = col_ptrs(cols)
    h = orc.orc_commit(ptrs, ncols, log_n, 2, 4, 1, u64ptr(cap))
    assert h
    coeffs = np.zeros(n, dtype=np.uint64); lde = np.zeros(4 * n, dtype=np.uint64)
    orc.orc_batch_get_coeffs(h, 2, u64ptr(coeffs))
    orc.orc_batch_get_lde(h, 2, u64ptr(lde))
    w4 = orc.orc_root_of_unity(log_n + 2)
    wn = orc.orc_root_of_unity(log_n)
    ev = lambda x: sum(int(coeffs[k]) * pow(x, k, P) for k in range(n)) % P
    for i in (0, 1, 7, n - 1):
        assert ev(pow(wn, i, P)) == int(cols[2][i])                 # from_values interpolates the trace
    for m in (0, 1, 2, 3, 50, 4 * n - 1):
        assert ev(7 * pow(w4, m, P) % P) == int(lde[m])             # values on the coset 7*H_4n
    # Merkle: leaf j = LDE row bitrev(j); path verifies up to the cap
    leaf = np.zeros(ncols, dtype=np.uint64); sib = np.zeros((log_n + 2 - 4) * 4, dtype=np.uint64)
    j = 37
    orc.orc_batch_open(h, j, u64ptr(leaf), u64ptr(sib))
    rev = int(format(j, "0%db" % (log_n + 2))[::-1], 2)
    full = np.zeros((ncols, 4 * n), dtype=np.uint64)
    for c in range(ncols):
        orc.orc_batch_get_lde(h, c, u64ptr(full[c]))
    assert (leaf == full[:, rev]).all()
    cur = np.zeros(4, dtype=np.uint64)
    orc.orc_hash_or_noop(u64ptr(leaf), ncols, u64ptr(cur))
    idx = j
    for l in range(log_n + 2 - 4):
        s = sib[4 * l:4 * l + 4].copy(); nxt = np.zeros(4, dtype=np.uint64)
        if idx & 1:
            orc.orc_two_to_one(u64ptr(s), u64ptr(cur), u64ptr(nxt))
        else:
            orc.orc_two_to_one(u64ptr(cur), u64ptr(s), u64ptr(nxt))
        cur = nxt; idx >>= 1
    assert (cur == cap[4 * idx:4 * idx + 4]).all()
    orc.orc_batch_free(h)
