"""Valid trace generators for the small test Systems (zkm_b200/csrc/tables/systems.h).  Written from
the reference's witness generators: logic.rs:140-183 (LogicStark::generate_trace_row),
memory/memory_stark.rs:44-244 (into_row, first-change flags, range check, counter, frequencies, padding),
poseidon/poseidon_stark.rs:105-145 (via the oracle's orc_gen_poseidon_rows)."""
import numpy as np

P = 0xFFFFFFFF00000001
SYSTEM_ALL_STARK, SYSTEM_LOGIC, SYSTEM_MINI3, SYSTEM_POSEIDON, SYSTEM_MEMORY, SYSTEM_ARITH, SYSTEM_KECCAK, SYSTEM_POSEIDON_SPONGE, SYSTEM_SHA_EXTEND, SYSTEM_SHA_COMPRESS, SYSTEM_CPU = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
T_CPU = 1
T_SHA_EXTEND, T_SHA_EXTEND_SPONGE, T_SHA_COMPRESS, T_SHA_COMPRESS_SPONGE = 6, 7, 8, 9
T_ARITHMETIC, T_POSEIDON, T_POSEIDON_SPONGE, T_KECCAK, T_KECCAK_SPONGE, T_LOGIC, T_MEMORY = 0, 2, 3, 4, 5, 10, 11


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


def logic_trace_from_ops(ops, log_n: int) -> np.ndarray:
    """ops = [(kind 0..3 = AND/OR/XOR/NOR, x, y)] -> (69, n); unused rows are all-zero (logic.rs:140-183)."""
    n = 1 << log_n
    assert len(ops) <= n
    t = np.zeros((69, n), dtype=np.uint64)
    for r, (k, x, y) in enumerate(ops):
        t[k, r] = 1
        for i in range(32):
            t[4 + i, r] = (x >> i) & 1
            t[36 + i, r] = (y >> i) & 1
        t[68, r] = [x & y, x | y, x ^ y, (~(x | y)) & 0xFFFFFFFF][k]
    return t


def memory_trace_from_ops(ops, log_n: int) -> np.ndarray:
    """ops = [(ctx, seg, virt, timestamp, is_read, value[, filter = 1])], consistent (reads return the last written value or the
    first value seen) -> (13, n): sorted by (ctx, seg, virt, timestamp), first-change flags, range check, counter,
    frequencies, padded by repeating the last operation as a filtered-off read (memory_stark.rs:44-244)."""
    n = 1 << log_n
    ops = sorted(ops, key=lambda o: (o[0], o[1], o[2], o[3]))
    used = len(ops)
    assert 1 <= used <= n
    t = np.zeros((13, n), dtype=np.uint64)
    t[0, :used] = [o[6] if len(o) > 6 else 1 for o in ops]
    a = np.array([o[:6] for o in ops], dtype=np.uint64).T
    t[3, :used], t[4, :used], t[5, :used], t[1, :used], t[2, :used], t[6, :used] = a[0], a[1], a[2], a[3], a[4], a[5]
    for c in (1, 3, 4, 5, 6):
        t[c, used:] = t[c, used - 1]
    t[2, used:] = 1
    o = t.astype(object)
    for i in range(n - 1):
        cfc = o[3, i] != o[3, i + 1]
        sfc = (o[4, i] != o[4, i + 1]) and not cfc
        vfc = (o[5, i] != o[5, i + 1]) and not sfc and not cfc
        t[7, i], t[8, i], t[9, i] = int(cfc), int(sfc), int(vfc)
        rc = (o[3, i + 1] - o[3, i] - 1) if cfc else (o[4, i + 1] - o[4, i] - 1) if sfc else (o[5, i + 1] - o[5, i] - 1) if vfc \
            else (o[1, i + 1] - o[1, i])
        assert 0 <= rc < n, rc
        t[10, i] = rc
    t[11] = np.arange(n, dtype=np.uint64)
    t[12] = np.bincount(t[10].astype(np.int64), minlength=n).astype(np.uint64)
    return t


def keccak_system_traces(lens=(0, 4, 132, 136, 272, 140), seed: int = 22):
    """Keccak slice of AllStark: sponge operations over inputs of the given byte lengths; the permutation, XOR and
    memory-read rows the sponge rows refer to are derived from them, as witness generation does upstream
    (witness/operation.rs keccak_general -> KeccakSpongeOp, logic ops, memory reads)."""
    import hash_gen as hg
    rng = np.random.default_rng(seed)
    ops, base = [], 64
    for i, ln in enumerate(lens):
        data = bytes(int(b) for b in rng.integers(0, 256, size=ln))
        nwords = ln // 4 + 1
        ops.append(([base + 4 * k for k in range(nwords)], 100 + i, data, 0, 0))
        base += 4 * nwords + 8
    sponge_rows = sum(len(o[2]) // hg.RATE_BYTES + 1 for o in ops)
    sponge, perms = hg.keccak_sponge_trace(ops, max(6, (sponge_rows - 1).bit_length()))
    keccak_in = []
    for pre, _post, ts in perms:
        lanes = [pre[2 * k] | (pre[2 * k + 1] << 32) for k in range(25)]
        keccak_in.append((lanes, ts))
    keccak = hg.keccak_trace(keccak_in, max(6, (len(keccak_in) * 24 - 1).bit_length()))
    xors, reads = [], []
    for r in range(sponge_rows):
        row = sponge[:, r]
        ts = int(row[hg.KS_TIMESTAMP])
        blk = [int(b) for b in row[hg.KS_BLOCK_BYTES:hg.KS_BLOCK_BYTES + hg.RATE_BYTES]]
        words = [blk[4 * k] | (blk[4 * k + 1] << 8) | (blk[4 * k + 2] << 16) | (blk[4 * k + 3] << 24) for k in range(hg.RATE_U32S)]
        for k in range(hg.RATE_U32S):
            xors.append((2, int(row[hg.KS_ORIG_RATE + k]), words[k]))
        final_len = [int(v) for v in row[hg.KS_IS_FINAL_LEN:hg.KS_IS_FINAL_LEN + hg.RATE_BYTES]]
        nbytes = hg.RATE_BYTES if int(row[hg.KS_IS_FULL]) else final_len.index(1)
        for i in range(nbytes):
            w = i // 4       # MIPS memory words are big-endian: ctl_looking_memory packs bytes [3, 2, 1, 0] little-endian
            be = (blk[4 * w] << 24) | (blk[4 * w + 1] << 16) | (blk[4 * w + 2] << 8) | blk[4 * w + 3]
            reads.append((0, 0, int(row[hg.KS_VIRT + w]), ts, 1, be))
    logic = logic_trace_from_ops(xors, max(6, (len(xors) - 1).bit_length()))
    memory = memory_trace_from_ops(reads, max(6, (len(reads) - 1).bit_length()))
    return [keccak, sponge, logic, memory]


def poseidon_system_traces(orc, lens=(0, 4, 28, 32, 64, 36, 100), seed: int = 23):
    """Poseidon slice of AllStark: sponge operations, the permutation rows and memory reads they refer to."""
    import hash_gen as hg
    from oracle.binding import u64ptr
    rng = np.random.default_rng(seed)
    ops, base = [], 64
    for i, ln in enumerate(lens):
        data = bytes(int(b) for b in rng.integers(0, 256, size=ln))
        nwords = ln // 4 + 1
        ops.append(([base + 4 * k for k in range(nwords)], 200 + i, data, 0, 0))
        base += 4 * nwords + 8
    nrows = sum(len(o[2]) // hg.PS_RATE_BYTES + 1 for o in ops)
    sponge, perms = hg.poseidon_sponge_trace(orc, ops, max(6, (nrows - 1).bit_length()))
    n_p = 1 << max(6, (len(perms) - 1).bit_length())
    inputs = np.zeros((n_p, 12), dtype=np.uint64)
    tsv = np.zeros(n_p, dtype=np.uint64)
    for k, (pre, _post, ts) in enumerate(perms):
        inputs[k] = pre
        tsv[k] = ts
    rows = np.zeros((n_p, 262), dtype=np.uint64)
    orc.orc_gen_poseidon_rows(u64ptr(inputs), u64ptr(tsv), n_p, u64ptr(rows))
    rows[len(perms):, 0] = 0
    reads = []
    for r in range(nrows):
        row = sponge[:, r]
        blk = [int(b) for b in row[hg.PS_BLOCK_BYTES:hg.PS_BLOCK_BYTES + hg.PS_RATE_BYTES]]
        final_len = [int(v) for v in row[hg.PS_IS_FINAL_LEN:hg.PS_IS_FINAL_LEN + hg.PS_RATE_BYTES]]
        nbytes = hg.PS_RATE_BYTES if int(row[hg.PS_IS_FULL]) else final_len.index(1)
        for i in range(nbytes):
            w = i // 4
            be = (blk[4 * w] << 24) | (blk[4 * w + 1] << 16) | (blk[4 * w + 2] << 8) | blk[4 * w + 3]
            reads.append((0, 0, int(row[hg.PS_VIRT + w]), int(row[hg.PS_TIMESTAMP]), 1, be))
    memory = memory_trace_from_ops(reads, max(6, (len(reads) - 1).bit_length()))
    return [np.ascontiguousarray(rows.T), sponge, memory]


def sha_extend_system_traces(seqs=((64, 1000), (1024, 5000)), seed: int = 31):
    """SHA-256 message-schedule slice of AllStark: 48 rounds per sequence, their XORs and memory traffic."""
    import hash_gen as hg
    ext, sp, xors, mem = hg.sha_extend_sequences(seqs, seed)
    lg = lambda k: max(6, (k - 1).bit_length())
    mem = [m + ((1,) if m[4] else (0,)) for m in mem]        # the CPU's writes carry filter 0 here: no looker in this slice
    return [hg.rows_to_trace(ext, hg.SHA_EXTEND_COLUMNS, lg(len(ext))), hg.rows_to_trace(sp, hg.SHA_EXTEND_SPONGE_COLUMNS, lg(len(sp))),
            logic_trace_from_ops(xors, lg(len(xors))), memory_trace_from_ops(mem, lg(len(mem)))]


def sha_compress_system_traces(calls=((64, 128, 3000), (512, 576, 7000)), seed: int = 41):
    """SHA-256 compression slice of AllStark: 65 rows per compression, 12 logic operations and one message-word read
    per round, the chaining-value reads of the sponge row."""
    import hash_gen as hg
    c, sp, logic, mem, _ = hg.sha_compressions(calls, seed)
    lg = lambda k: max(6, (k - 1).bit_length())
    return [hg.rows_to_trace(c, hg.SHA_COMPRESS_COLUMNS, lg(len(c))), hg.rows_to_trace(sp, hg.SHA_COMPRESS_SPONGE_COLUMNS, lg(len(sp))),
            logic_trace_from_ops(logic, lg(len(logic))), memory_trace_from_ops(mem, lg(len(mem)))]


def cpu_system_traces():
    """Instruction-execution slice of AllStark: the test program of tests/cpu_program.py run by the interpreter of
    tests/cpu_gen.py; the arithmetic, logic and memory tables are generated from the operations it logged."""
    import arith_gen as ag
    import cpu_gen as cg
    import cpu_program as cp
    image, end = cp.build()
    cpu = cg.MiniCpu(image, cp.ENTRY)
    while cpu.pc != end:
        cpu.step()
        assert cpu.clock() < 250
    lg = lambda k: max(6, (k - 1).bit_length())
    return [cpu.cpu_trace(8), ag.arithmetic_trace(cpu.arith_ops, 16), logic_trace_from_ops(cpu.logic_ops, lg(len(cpu.logic_ops))),
            cg.memory_generate_trace(cpu.mem_ops)]


def all_stark_valid_traces(orc, sha_blocks=0, return_cpu=False, return_ops=False):
    """A valid trace of all 12 AllStark tables: the test program with its syscalls and the Keccak / SHA-256 precompiles
    (tests/cpu_program.py, with_syscalls), the image-id Poseidon hash of the bootstrap, and every table generated from
    the operations the interpreter logged -- what Traces::into_tables does upstream (witness/traces.rs:230-318)."""
    import arith_gen as ag
    import cpu_gen as cg
    import cpu_program as cp
    import hash_gen as hg
    from oracle.binding import u64ptr
    image, end = cp.build(with_syscalls=True, sha_blocks=sha_blocks)
    cpu = cg.MiniCpu(image, cp.ENTRY, image_id_words=(cp.IMAGE_ID, [0x01020304 * (k + 1) & 0xFFFFFFFF for k in range(9)]))
    while cpu.pc != end:
        cpu.step()
        assert cpu.clock() < 1000 + 200 * sha_blocks
    lg = lambda k: max(6, (max(k, 1) - 1).bit_length())
    # Poseidon sponge + permutations; the digest goes back into the CPU's image-id row
    ps_rows, ps_perms = [], []
    for addrs, ts, data, ctx, seg, cpu_row in cpu.poseidon_ops:
        r, perms = hg.poseidon_sponge_rows_for_op(orc, (addrs, ts, data, ctx, seg))
        ps_rows += r
        ps_perms += [(pre, ts) for pre, _ in perms]
        for k in range(4):
            cpu_row[cg.GENERAL + k] = perms[-1][1][k]
    n_p = 1 << lg(len(ps_perms))
    p_in, p_ts = np.zeros((n_p, 12), dtype=np.uint64), np.zeros(n_p, dtype=np.uint64)
    for k, (pre, ts) in enumerate(ps_perms):
        p_in[k], p_ts[k] = pre, ts
    p_rows = np.zeros((n_p, 262), dtype=np.uint64)
    orc.orc_gen_poseidon_rows(u64ptr(p_in), u64ptr(p_ts), n_p, u64ptr(p_rows))
    p_rows[len(ps_perms):, 0] = 0
    # Keccak sponge + permutations + XORs
    ks_rows, k_in, logic = [], [], list(cpu.logic_ops)
    for op in cpu.keccak_ops:
        r, perms = hg.keccak_sponge_rows_for_op(op)
        for row, (pre, _post) in zip(r, perms):
            blk = [int(b) for b in row[hg.KS_BLOCK_BYTES:hg.KS_BLOCK_BYTES + hg.RATE_BYTES]]
            for k in range(hg.RATE_U32S):
                logic.append((2, int(row[hg.KS_ORIG_RATE + k]), blk[4 * k] | (blk[4 * k + 1] << 8) | (blk[4 * k + 2] << 16) | (blk[4 * k + 3] << 24)))
            k_in.append(([pre[2 * k] | (pre[2 * k + 1] << 32) for k in range(25)], op[1]))
        ks_rows += r
    # SHA extend / compress tables from the logged calls
    se_rows, ses_rows = [], []
    for ins, virts, out_virt, ts, rnd, w_i in cpu.sha_extend_ops:
        se_rows.append(hg.sha_extend_row(*ins, ts)[0])
        sp = np.zeros(hg.SHA_EXTEND_SPONGE_COLUMNS, dtype=np.uint64)
        sp[hg.SES_ROUND + rnd] = 1
        for at, v in zip((hg.SES_W_M15, hg.SES_W_M2, hg.SES_W_M16, hg.SES_W_M7), ins):
            sp[at:at + 4] = hg._le4(v)
        sp[hg.SES_W_I:hg.SES_W_I + 4] = hg._le4(w_i)
        sp[hg.SES_INPUT_VIRT:hg.SES_INPUT_VIRT + 4], sp[hg.SES_OUTPUT_VIRT], sp[hg.SES_TIMESTAMP] = virts, out_virt, ts
        ses_rows.append(sp)
    sc_rows, scs_rows = [], []
    for hx, w, h_ptr, w_ptr, ts in cpu.sha_compress_ops:
        st = list(hx)
        for i in range(64):
            row, st, _ = hg.sha_compress_row(st, w[i], hg.SHA_K[i], i, w_ptr + 4 * i, ts)
            sc_rows.append(row)
        sc_rows.append(hg.sha_compress_row(st, 0, 0, 64, w_ptr + 4 * 64, ts)[0])
        sp = np.zeros(hg.SHA_COMPRESS_SPONGE_COLUMNS, dtype=np.uint64)
        sp[hg.SCS_TIMESTAMP], sp[hg.SCS_IS_REAL], sp[hg.SCS_W_START_VIRT] = ts, 1, w_ptr
        sp[hg.SCS_HX_VIRT:hg.SCS_HX_VIRT + 8] = [h_ptr + 4 * j for j in range(8)]
        for j in range(8):
            sp[hg.SCS_HX + 4 * j:hg.SCS_HX + 4 * j + 4] = hg._le4(hx[j])
            sp[hg.SCS_OUTPUT_STATE + 4 * j:hg.SCS_OUTPUT_STATE + 4 * j + 4] = hg._le4(st[j])
            hg._wadd(sp, hg.SCS_OUTPUT_HX + 6 * j, 2, hx[j], st[j])
        scs_rows.append(sp)
    out = [ag.arithmetic_trace(cpu.arith_ops, 16),
            cpu.cpu_trace(lg(cpu.clock() + 1)),
            np.ascontiguousarray(p_rows.T),
            hg.rows_to_trace(ps_rows, hg.POSEIDON_SPONGE_COLUMNS, lg(len(ps_rows))),
            hg.keccak_trace(k_in, lg(len(k_in) * 24)),
            hg.rows_to_trace(ks_rows, hg.KECCAK_SPONGE_COLUMNS, lg(len(ks_rows))),
            hg.rows_to_trace(se_rows, hg.SHA_EXTEND_COLUMNS, lg(len(se_rows))),
            hg.rows_to_trace(ses_rows, hg.SHA_EXTEND_SPONGE_COLUMNS, lg(len(ses_rows))),
            hg.rows_to_trace(sc_rows, hg.SHA_COMPRESS_COLUMNS, lg(len(sc_rows))),
            hg.rows_to_trace(scs_rows, hg.SHA_COMPRESS_SPONGE_COLUMNS, lg(len(scs_rows))),
            logic_trace_from_ops(logic, lg(len(logic))),
            cg.memory_generate_trace(cpu.mem_ops)]
    if return_ops:
        # the operation logs zkm_b200_prove_with_ops takes for every table but Cpu (include/zkm_b200.h): the same operations the
        # tables above were generated from
        ses_ops = [[rnd, *ins, *virts, out_virt, 0, 0, ts] for ins, virts, out_virt, ts, rnd, _w in cpu.sha_extend_ops]
        sc_ops, scs_ops = [], []
        for hx, w, h_ptr, w_ptr, ts in cpu.sha_compress_ops:
            st = list(hx)
            for i in range(65):
                w_i, k_i = (w[i], hg.SHA_K[i]) if i < 64 else (0, 0)
                sc_ops.append([*st, w_i, k_i, i, w_ptr + 4 * i, 0, 0, ts])
                st = hg.sha_compress_row(st, w_i, k_i, i, 0, 0)[1]
            scs_ops.append([*hx, *w, *[h_ptr + 4 * j for j in range(8)], w_ptr, 0, 0, 0, 0, ts])
        arr = lambda rows, k: np.array(rows, dtype=np.uint64).reshape(len(rows), k)
        ops = {0: arr(cpu.arith_ops, 3),
               2: np.concatenate([p_in[:len(ps_perms)], p_ts[:len(ps_perms), None]], axis=1),
               3: [(addrs, ts, data, ctx, seg) for addrs, ts, data, ctx, seg, _row in cpu.poseidon_ops],
               4: arr([list(lanes) + [ts] for lanes, ts in k_in], 26),
               5: list(cpu.keccak_ops),
               6: arr([[*ins, ts] for ins, _v, _o, ts, _r, _w in cpu.sha_extend_ops], 5),
               7: arr(ses_ops, 13), 8: arr(sc_ops, 15), 9: arr(scs_ops, 86),
               10: arr(logic, 3), 11: arr(cpu.mem_ops, 7)}
        return out, ops
    return (out, cpu) if return_cpu else out
