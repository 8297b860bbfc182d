"""Valid trace generators for the small test Systems (zkm_b200/csrc/tables/systems.h).  Written from
the reference's witness generators: logic.rs:140-183 (LogicStark::generate_trace_row),
memory/memory_stark.rs:44-244 (into_row, first-change flags, range check, counter, frequencies, padding),
poseidon/poseidon_stark.rs:105-145 (via the oracle's orc_gen_poseidon_rows)."""
import numpy as np

P = 0xFFFFFFFF00000001
SYSTEM_ALL_STARK, SYSTEM_LOGIC, SYSTEM_MINI3, SYSTEM_POSEIDON, SYSTEM_MEMORY, SYSTEM_ARITH = 0, 1, 2, 3, 4, 5
T_ARITHMETIC, T_POSEIDON, T_LOGIC, T_MEMORY = 0, 2, 10, 11


def logic_trace(log_n: int, seed: int = 1, used_frac: float = 0.8) -> np.ndarray:
    """(69, n) uint64. Columns: IS_AND, IS_OR, IS_XOR, IS_NOR, 32 bits of x, 32 bits of y, result."""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    t = np.zeros((69, n), dtype=np.uint64)
    used = int(n * used_frac)
    op = rng.integers(0, 4, size=n)
    x = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    y = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    res = np.where(op == 0, x & y, np.where(op == 1, x | y, np.where(op == 2, x ^ y, (~(x | y)) & np.uint64(0xFFFFFFFF))))
    for k in range(4):
        t[k, :used] = (op[:used] == k)
    for i in range(32):
        t[4 + i, :used] = (x[:used] >> np.uint64(i)) & np.uint64(1)
        t[36 + i, :used] = (y[:used] >> np.uint64(i)) & np.uint64(1)
    t[68, :used] = res[:used]
    return t


def memory_trace(log_n: int, seed: int = 2, used_frac: float = 0.7) -> np.ndarray:
    """(13, n) uint64: FILTER, TIMESTAMP, IS_READ, CTX, SEG, VIRT, VALUE, CTX_FC, SEG_FC, VIRT_FC, RANGE_CHECK, COUNTER, FREQ."""
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    used = max(2, int(n * used_frac))
    # ops over a small address space so every range-check value stays below n
    ctx = rng.integers(0, 3, size=used)
    seg = rng.integers(0, 4, size=used)
    virt = rng.integers(0, max(2, n // 16), size=used)
    ts = 1 + rng.permutation(used)                 # distinct timestamps below n: every delta is range-checkable
    order = np.lexsort((ts, virt, seg, ctx))
    ctx, seg, virt, ts = ctx[order], seg[order], virt[order], ts[order]
    is_read = rng.integers(0, 2, size=used)
    value = np.zeros(used, dtype=np.uint64)
    cur = {}
    for i in range(used):
        key = (ctx[i], seg[i], virt[i])
        if is_read[i] and key in cur:
            value[i] = cur[key]
        else:
            value[i] = rng.integers(0, 1 << 32)
            cur[key] = value[i]
    t = np.zeros((13, n), dtype=np.uint64)
    t[0, :used] = 1
    t[1, :used] = ts; t[2, :used] = is_read; t[3, :used] = ctx; t[4, :used] = seg; t[5, :used] = virt; t[6, :used] = value
    # padding: repeat last op as a dummy read (filter 0)
    for c in (1, 3, 4, 5, 6):
        t[c, used:] = t[c, used - 1]
    t[2, used:] = 1
    a = t.astype(object)
    for i in range(n - 1):
        cfc = a[3, i] != a[3, i + 1]
        sfc = (a[4, i] != a[4, i + 1]) and not cfc
        vfc = (a[5, i] != a[5, i + 1]) and not sfc and not cfc
        t[7, i], t[8, i], t[9, i] = int(cfc), int(sfc), int(vfc)
        if cfc:
            rc = a[3, i + 1] - a[3, i] - 1
        elif sfc:
            rc = a[4, i + 1] - a[4, i] - 1
        elif vfc:
            rc = a[5, i + 1] - a[5, i] - 1
        else:
            rc = a[1, i + 1] - a[1, i]
        assert 0 <= rc < n, rc
        t[10, i] = rc
    t[11] = np.arange(n, dtype=np.uint64)
    freq = np.bincount(t[10].astype(np.int64), minlength=n)
    t[12] = freq.astype(np.uint64)
    return t


def poseidon_trace(orc, log_n: int, seed: int = 3, used_frac: float = 0.75) -> np.ndarray:
    """(262, n) uint64; unused rows are all-zero inputs with FILTER = 0 (reference generate_trace pads with
    the permutation of the zero input, filter 0: poseidon_stark.rs:126-145)."""
    from oracle.binding import u64ptr
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    used = int(n * used_frac)
    inputs = np.zeros((n, 12), dtype=np.uint64)
    inputs[:used] = rng.integers(0, P, size=(used, 12), dtype=np.uint64)
    tsv = np.zeros(n, dtype=np.uint64)
    tsv[:used] = np.arange(1, used + 1)
    rows = np.zeros((n, 262), dtype=np.uint64)
    orc.orc_gen_poseidon_rows(u64ptr(inputs), u64ptr(tsv), n, u64ptr(rows))
    rows[used:, 0] = 0
    return np.ascontiguousarray(rows.T)


def arithmetic_trace(count: int = 30000, seed: int = 11, log_n: int = 16) -> np.ndarray:
    """(54, 2^16) uint64: `count` random operations of all 26 kinds (tests/arith_gen.py)."""
    import arith_gen as ag
    return ag.arithmetic_trace(ag.random_ops(count, seed), log_n)
