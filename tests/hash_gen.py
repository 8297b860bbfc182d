"""Valid-trace generators for the hash tables, restated from the reference's witness generators (test
infrastructure): keccak/keccak_stark.rs:62-237 + keccak/columns.rs (Keccak-f round rows),
keccak_sponge/keccak_sponge_stark.rs:222-447 + keccak_sponge/columns.rs (absorb rows, pad10*1).
The reference pins these generators with `keccak_correctness_test` (keccak_stark.rs:655-687: last-round output ==
tiny-keccak's keccakf) and `test_generation` (keccak_sponge_stark.rs:761-790: digest bytes == keccak256(input)); the
tests restate both (independent permutation checked against hashlib's SHA3, which shares Keccak-f[1600])."""
import numpy as np

P = 0xFFFFFFFF00000001
M64 = (1 << 64) - 1

# ------------------------------------------------------------------------------------------- Keccak-f
RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
      0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
      0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
      0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
R = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]   # keccak/columns.rs:41-47
NUM_ROUNDS, NUM_INPUTS = 24, 25
K_TIMESTAMP = NUM_ROUNDS
K_START_A = K_TIMESTAMP + 1
K_START_C = K_START_A + 50
K_START_C_PRIME = K_START_C + 320
K_START_A_PRIME = K_START_C_PRIME + 320
K_START_A_PP = K_START_A_PRIME + 1600
K_START_A_PP_00_BITS = K_START_A_PP + 50
K_A_PPP_00_LO = K_START_A_PP_00_BITS + 64
KECCAK_COLUMNS = K_A_PPP_00_LO + 2
assert KECCAK_COLUMNS == 2431


def rotl(v, r):
    r %= 64
    return ((v << r) | (v >> (64 - r))) & M64 if r else v


def reg_a(x, y):
    return K_START_A + (x * 5 + y) * 2


def reg_a_pp(x, y):
    return K_START_A_PP + x * 10 + y * 2


def reg_a_ppp(x, y):
    return K_A_PPP_00_LO if (x == 0 and y == 0) else reg_a_pp(x, y)


def bits64(v):
    return [(v >> z) & 1 for z in range(64)]


def keccak_round_row(row, A, rnd):
    """keccak_stark.rs:131-226: fills one round's row from the 5x5 lanes A[x][y]; returns the output lanes."""
    row[rnd] = 1
    for x in range(5):
        for y in range(5):
            row[reg_a(x, y)] = A[x][y] & 0xFFFFFFFF
            row[reg_a(x, y) + 1] = A[x][y] >> 32
    C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
    Cp = [C[x] ^ C[(x + 4) % 5] ^ rotl(C[(x + 1) % 5], 1) for x in range(5)]
    Ap = [[A[x][y] ^ C[x] ^ Cp[x] for y in range(5)] for x in range(5)]
    for x in range(5):
        row[K_START_C + x * 64:K_START_C + x * 64 + 64] = bits64(C[x])
        row[K_START_C_PRIME + x * 64:K_START_C_PRIME + x * 64 + 64] = bits64(Cp[x])
        for y in range(5):
            at = K_START_A_PRIME + x * 320 + y * 64
            row[at:at + 64] = bits64(Ap[x][y])
    # B[x, y, z] = A'[(x + 3y) % 5, x, z - R[(x + 3y) % 5][x]]   (columns.rs:83-92)
    B = [[rotl(Ap[(x + 3 * y) % 5][x], R[(x + 3 * y) % 5][x]) for y in range(5)] for x in range(5)]
    App = [[B[x][y] ^ (~B[(x + 1) % 5][y] & B[(x + 2) % 5][y] & M64) for y in range(5)] for x in range(5)]
    for x in range(5):
        for y in range(5):
            row[reg_a_pp(x, y)] = App[x][y] & 0xFFFFFFFF
            row[reg_a_pp(x, y) + 1] = App[x][y] >> 32
    row[K_START_A_PP_00_BITS:K_START_A_PP_00_BITS + 64] = bits64(App[0][0])
    out00 = App[0][0] ^ RC[rnd]
    row[K_A_PPP_00_LO] = out00 & 0xFFFFFFFF
    row[K_A_PPP_00_LO + 1] = out00 >> 32
    App[0][0] = out00
    return App


def keccak_rows_for_perm(inp, timestamp):
    """keccak_stark.rs:83-117 (input[y * 5 + x] is lane (x, y))."""
    rows = np.zeros((NUM_ROUNDS, KECCAK_COLUMNS), dtype=np.uint64)
    A = [[int(inp[y * 5 + x]) for y in range(5)] for x in range(5)]
    for rnd in range(NUM_ROUNDS):
        rows[rnd, K_TIMESTAMP] = timestamp
        A = keccak_round_row(rows[rnd], A, rnd)
    return rows, [A[i % 5][i // 5] for i in range(25)]


def keccakf(lanes):
    return keccak_rows_for_perm(lanes, 0)[1]


def keccak_trace(inputs_and_timestamps, log_n):
    """keccak_stark.rs:62-81,228-237 -> (2431, n) uint64; padding rows are all-zero."""
    n = 1 << log_n
    assert len(inputs_and_timestamps) * NUM_ROUNDS <= n
    t = np.zeros((n, KECCAK_COLUMNS), dtype=np.uint64)
    for i, (inp, ts) in enumerate(inputs_and_timestamps):
        t[i * NUM_ROUNDS:(i + 1) * NUM_ROUNDS] = keccak_rows_for_perm(inp, ts)[0]
    return np.ascontiguousarray(t.T)


def random_keccak_trace(log_n, seed=21, perms=None):
    rng = np.random.default_rng(seed)
    perms = ((1 << log_n) // NUM_ROUNDS) if perms is None else perms
    ins = [([int(v) for v in rng.integers(0, 1 << 64, size=25, dtype=np.uint64)], 10 + i) for i in range(perms)]
    return keccak_trace(ins, log_n)


# --------------------------------------------------------------------------------------- KeccakSponge
RATE_BYTES, RATE_U32S, CAP_U32S, DIGEST_U32S, WIDTH_U32S = 136, 34, 16, 8, 50
KS_IS_FULL, KS_CONTEXT, KS_SEGMENT, KS_VIRT = 0, 1, 2, 3
KS_TIMESTAMP = KS_VIRT + RATE_U32S
KS_LEN = KS_TIMESTAMP + 1
KS_ALREADY = KS_LEN + 1
KS_IS_FINAL_LEN = KS_ALREADY + 1
KS_ORIG_RATE = KS_IS_FINAL_LEN + RATE_BYTES
KS_ORIG_CAP = KS_ORIG_RATE + RATE_U32S
KS_BLOCK_BYTES = KS_ORIG_CAP + CAP_U32S
KS_XORED_RATE = KS_BLOCK_BYTES + RATE_BYTES
KS_PARTIAL_UPDATED = KS_XORED_RATE + RATE_U32S
KS_UPDATED_DIGEST_BYTES = KS_PARTIAL_UPDATED + (WIDTH_U32S - DIGEST_U32S)
KECCAK_SPONGE_COLUMNS = KS_UPDATED_DIGEST_BYTES + 32
assert KECCAK_SPONGE_COLUMNS == 470


def keccakf_u32s(state):        # cpu/kernel/keccak_util.rs:6-18
    lanes = [state[2 * i] | (state[2 * i + 1] << 32) for i in range(25)]
    lanes = keccakf(lanes)
    return [(lanes[i // 2] >> (32 * (i % 2))) & 0xFFFFFFFF for i in range(50)]


def _sponge_common(row, op, already, state):   # keccak_sponge_stark.rs:351-438
    virts, ts, data, ctx, seg = op
    idx = already // 4
    end = min((already + RATE_BYTES) // 4, len(virts))
    v = list(virts[idx:end]) + [0] * (RATE_U32S - max(0, end - idx))
    row[KS_CONTEXT], row[KS_SEGMENT] = ctx, seg
    row[KS_VIRT:KS_VIRT + RATE_U32S] = v[:RATE_U32S]
    row[KS_TIMESTAMP], row[KS_LEN], row[KS_ALREADY] = ts, len(data), already
    row[KS_ORIG_RATE:KS_ORIG_RATE + RATE_U32S] = state[:RATE_U32S]
    row[KS_ORIG_CAP:KS_ORIG_CAP + CAP_U32S] = state[RATE_U32S:]
    blk = [int(b) for b in row[KS_BLOCK_BYTES:KS_BLOCK_BYTES + RATE_BYTES]]
    state = list(state)
    for i in range(RATE_U32S):
        state[i] ^= blk[4 * i] | (blk[4 * i + 1] << 8) | (blk[4 * i + 2] << 16) | (blk[4 * i + 3] << 24)
    row[KS_XORED_RATE:KS_XORED_RATE + RATE_U32S] = state[:RATE_U32S]
    state = keccakf_u32s(state)
    row[KS_PARTIAL_UPDATED:KS_PARTIAL_UPDATED + WIDTH_U32S - DIGEST_U32S] = state[DIGEST_U32S:]
    for l in range(DIGEST_U32S):
        for i in range(4):
            row[KS_UPDATED_DIGEST_BYTES + 4 * l + i] = (state[l] >> (8 * i)) & 0xFF
    return state


def keccak_sponge_rows_for_op(op):
    """op = (virt address of every 32-bit input word, timestamp, input bytes, context, segment);
    keccak_sponge_stark.rs:253-349.  Returns the rows and the (pre-permutation, post-permutation) u32 states of each row."""
    virts, ts, data, ctx, seg = op
    rows, perms = [], []
    state = [0] * WIDTH_U32S
    already = 0
    nfull = len(data) // RATE_BYTES
    for b in range(nfull + 1):
        row = np.zeros(KECCAK_SPONGE_COLUMNS, dtype=np.uint64)
        chunk = data[b * RATE_BYTES:(b + 1) * RATE_BYTES]
        if b < nfull:
            row[KS_IS_FULL] = 1
            row[KS_BLOCK_BYTES:KS_BLOCK_BYTES + RATE_BYTES] = list(chunk)
        else:
            row[KS_BLOCK_BYTES:KS_BLOCK_BYTES + len(chunk)] = list(chunk)
            if len(chunk) == RATE_BYTES - 1:
                row[KS_BLOCK_BYTES + len(chunk)] = 0b10000001
            else:
                row[KS_BLOCK_BYTES + len(chunk)] = 1
                row[KS_BLOCK_BYTES + RATE_BYTES - 1] = 0b10000000
            row[KS_IS_FINAL_LEN + len(chunk)] = 1
        state = _sponge_common(row, op, already, state)
        xored = [int(x) for x in row[KS_XORED_RATE:KS_XORED_RATE + RATE_U32S]] + [int(x) for x in row[KS_ORIG_CAP:KS_ORIG_CAP + CAP_U32S]]
        perms.append((xored, list(state)))
        rows.append(row)
        already += RATE_BYTES
    return rows, perms


def keccak_sponge_trace(ops, log_n):
    rows, perms = [], []
    for op in ops:
        r, p = keccak_sponge_rows_for_op(op)
        rows += r
        perms += [(pre, post, op[1]) for pre, post in p]
    n = 1 << log_n
    assert len(rows) <= n
    t = np.zeros((n, KECCAK_SPONGE_COLUMNS), dtype=np.uint64)
    if rows:
        t[:len(rows)] = np.array(rows)
    return np.ascontiguousarray(t.T), perms


def random_sponge_ops(count, seed=22, lens=None):
    rng = np.random.default_rng(seed)
    ops = []
    for i in range(count):
        ln = int(lens[i]) if lens is not None else int(rng.integers(0, 80)) * 4
        data = bytes(int(b) for b in rng.integers(0, 256, size=ln))
        base = int(rng.integers(1, 1 << 20)) * 4
        virts = [base + 4 * k for k in range((ln + 3) // 4 + 1)]
        ops.append((virts, 100 + i, data, 0, 0))
    return ops


def keccak256(data):
    """The digest the sponge rows compute (pad10*1 with 0x01, rate 136), for the reference's test_generation check."""
    _, perms = keccak_sponge_rows_for_op(([0] * (len(data) // 4 + 2), 0, data, 0, 0))
    post = perms[-1][1]
    return b"".join(int(w).to_bytes(4, "little") for w in post[:8])


# ------------------------------------------------------------------------------------- PoseidonSponge
# poseidon_sponge/poseidon_sponge_stark.rs:187-365 + poseidon_sponge/columns.rs:19-68.  State elements are field
# elements; every block OVERWRITES the rate (new_rate = the block's 8 little-endian u32s), no xor.
PS_RATE, PS_CAP, PS_WIDTH, PS_DIGEST, PS_RATE_BYTES = 8, 4, 12, 4, 32
PS_IS_FULL, PS_CONTEXT, PS_SEGMENT, PS_VIRT = 0, 1, 2, 3
PS_TIMESTAMP = PS_VIRT + PS_RATE
PS_LEN = PS_TIMESTAMP + 1
PS_ALREADY = PS_LEN + 1
PS_IS_FINAL_LEN = PS_ALREADY + 1
PS_ORIG_RATE = PS_IS_FINAL_LEN + PS_RATE_BYTES
PS_ORIG_CAP = PS_ORIG_RATE + PS_RATE
PS_BLOCK_BYTES = PS_ORIG_CAP + PS_CAP
PS_NEW_RATE = PS_BLOCK_BYTES + PS_RATE_BYTES
PS_PARTIAL_UPDATED = PS_NEW_RATE + PS_RATE
PS_UPDATED_DIGEST = PS_PARTIAL_UPDATED + (PS_WIDTH - PS_DIGEST)
POSEIDON_SPONGE_COLUMNS = PS_UPDATED_DIGEST + PS_DIGEST
assert POSEIDON_SPONGE_COLUMNS == 110


def poseidon_sponge_rows_for_op(orc, op):
    """op = (virt of every input word, timestamp, input bytes, context, segment).  Returns rows and, per row, the
    permutation (input state, output state)."""
    from oracle.binding import u64ptr
    virts, ts, data, ctx, seg = op
    rows, perms = [], []
    state = [0] * PS_WIDTH
    nfull = len(data) // PS_RATE_BYTES
    for b in range(nfull + 1):
        already = b * PS_RATE_BYTES
        row = np.zeros(POSEIDON_SPONGE_COLUMNS, dtype=np.uint64)
        chunk = data[already:already + PS_RATE_BYTES]
        if b < nfull:
            row[PS_IS_FULL] = 1
            row[PS_BLOCK_BYTES:PS_BLOCK_BYTES + PS_RATE_BYTES] = list(chunk)
        else:
            row[PS_BLOCK_BYTES:PS_BLOCK_BYTES + len(chunk)] = list(chunk)
            if len(chunk) == PS_RATE_BYTES - 1:
                row[PS_BLOCK_BYTES + len(chunk)] = 0b10000001
            else:
                row[PS_BLOCK_BYTES + len(chunk)] = 1
                row[PS_BLOCK_BYTES + PS_RATE_BYTES - 1] = 0b10000000
            row[PS_IS_FINAL_LEN + len(chunk)] = 1
        idx = already // 4
        end = min((already + PS_RATE_BYTES) // 4, len(virts))
        v = list(virts[idx:end]) + [0] * (PS_RATE - max(0, end - idx))
        row[PS_CONTEXT], row[PS_SEGMENT] = ctx, seg
        row[PS_VIRT:PS_VIRT + PS_RATE] = v[:PS_RATE]
        row[PS_TIMESTAMP], row[PS_LEN], row[PS_ALREADY] = ts, len(data), already
        row[PS_ORIG_RATE:PS_ORIG_RATE + PS_RATE] = state[:PS_RATE]
        row[PS_ORIG_CAP:PS_ORIG_CAP + PS_CAP] = state[PS_RATE:]
        blk = [int(x) for x in row[PS_BLOCK_BYTES:PS_BLOCK_BYTES + PS_RATE_BYTES]]
        words = [blk[4 * i] | (blk[4 * i + 1] << 8) | (blk[4 * i + 2] << 16) | (blk[4 * i + 3] << 24) for i in range(PS_RATE)]
        row[PS_NEW_RATE:PS_NEW_RATE + PS_RATE] = words
        pre = words + state[PS_RATE:]
        st = np.array(pre, dtype=np.uint64)
        orc.orc_poseidon_permute(u64ptr(st), 0)
        state = [int(x) for x in st]
        row[PS_PARTIAL_UPDATED:PS_PARTIAL_UPDATED + PS_WIDTH - PS_DIGEST] = state[PS_DIGEST:]
        row[PS_UPDATED_DIGEST:PS_UPDATED_DIGEST + PS_DIGEST] = state[:PS_DIGEST]
        perms.append((pre, list(state)))
        rows.append(row)
    return rows, perms


def poseidon_sponge_trace(orc, ops, log_n):
    rows, perms = [], []
    for op in ops:
        r, p = poseidon_sponge_rows_for_op(orc, op)
        rows += r
        perms += [(pre, post, op[1]) for pre, post in p]
    n = 1 << log_n
    assert len(rows) <= n
    t = np.zeros((n, POSEIDON_SPONGE_COLUMNS), dtype=np.uint64)
    if rows:
        t[:len(rows)] = np.array(rows)
    return np.ascontiguousarray(t.T), perms


# ------------------------------------------------------------------------------- ShaExtend / ShaExtendSponge
# sha_extend/sha_extend_stark.rs:122-237 + columns.rs:8-37 + rotate_right.rs:14-27,108-117, shift_right.rs:15-28,
# wrapping_add_4.rs:16-31; sha_extend_sponge/sha_extend_sponge_stark.rs:128-227 + columns.rs:7-33.
SE_W_I, SE_W_I_CARRY, SE_W_M15, SE_W_M2, SE_W_M16, SE_W_M7 = 0, 4, 8, 12, 16, 20
SE_S0_INTER, SE_S0, SE_S1_INTER, SE_S1 = 24, 28, 32, 36
SE_RR7, SE_RR18, SE_RR17, SE_RR19, SE_RS10, SE_RS3 = 40, 46, 52, 58, 64, 70      # each: value[4], shift, carry
SE_TIMESTAMP, SE_IS_REAL = 76, 77
SHA_EXTEND_COLUMNS = 78
M32 = 0xFFFFFFFF


def _le4(v):
    return [(v >> (8 * i)) & 0xFF for i in range(4)]


def rotr32(v, r):
    r %= 32
    return ((v >> r) | (v << (32 - r))) & M32 if r else v


def shr_carry(v, r):            # rotate_right.rs:108-117
    r %= 32
    return (v >> r, v & ((1 << r) - 1)) if r else (v, 0)


def _shift_op(row, at, v, r, rotate):
    out = rotr32(v, r) if rotate else v >> (r % 32)
    row[at:at + 4] = _le4(out)
    row[at + 4], row[at + 5] = shr_carry(v, r)
    return out


def sha_extend_row(w15, w2, w16, w7, timestamp):
    row = np.zeros(SHA_EXTEND_COLUMNS, dtype=np.uint64)
    row[SE_TIMESTAMP], row[SE_IS_REAL] = timestamp, 1
    row[SE_W_M15:SE_W_M15 + 4], row[SE_W_M2:SE_W_M2 + 4] = _le4(w15), _le4(w2)
    row[SE_W_M16:SE_W_M16 + 4], row[SE_W_M7:SE_W_M7 + 4] = _le4(w16), _le4(w7)
    rr7, rr18, rs3 = _shift_op(row, SE_RR7, w15, 7, True), _shift_op(row, SE_RR18, w15, 18, True), _shift_op(row, SE_RS3, w15, 3, False)
    s0i = rr7 ^ rr18
    s0 = s0i ^ rs3
    rr17, rr19, rs10 = _shift_op(row, SE_RR17, w2, 17, True), _shift_op(row, SE_RR19, w2, 19, True), _shift_op(row, SE_RS10, w2, 10, False)
    s1i = rr17 ^ rr19
    s1 = s1i ^ rs10
    row[SE_S0_INTER:SE_S0_INTER + 4], row[SE_S0:SE_S0 + 4] = _le4(s0i), _le4(s0)
    row[SE_S1_INTER:SE_S1_INTER + 4], row[SE_S1:SE_S1 + 4] = _le4(s1i), _le4(s1)
    total = s1 + w7 + s0 + w16
    row[SE_W_I:SE_W_I + 4] = _le4(total & M32)
    row[SE_W_I_CARRY + (total >> 32)] = 1
    xors = [(2, rr7, rr18), (2, s0i, rs3), (2, rr17, rr19), (2, s1i, rs10)]
    return row, total & M32, xors


SES_ROUND, SES_W_M15, SES_W_M2, SES_W_M16, SES_W_M7, SES_W_I = 0, 48, 52, 56, 60, 64
SES_INPUT_VIRT, SES_OUTPUT_VIRT, SES_CONTEXT, SES_SEGMENT, SES_TIMESTAMP = 68, 72, 73, 74, 75
SHA_EXTEND_SPONGE_COLUMNS = 76
NUM_CHANNELS = 10                # cpu/membus.rs


def sha_extend_sequences(seqs, seed=31):
    """seqs = [(base address of w[0], first timestamp)]: 48 extension rounds each over random w[0..16].
    Returns (ShaExtend rows, ShaExtendSponge rows, XOR ops, memory ops) as witness generation produces them upstream
    (witness/operation.rs sha_extend: one ShaExtendSpongeOp per round, output written back by the CPU)."""
    rng = np.random.default_rng(seed)
    ext_rows, sp_rows, xors, mem = [], [], [], []
    for base, ts0 in seqs:
        w = [int(v) for v in rng.integers(0, 1 << 32, size=16)]
        for i in range(16, 64):
            r = i - 16
            ts = ts0 + 2 * NUM_CHANNELS * r
            ins = (w[i - 15], w[i - 2], w[i - 16], w[i - 7])
            row, wi, x = sha_extend_row(*ins, ts)
            w.append(wi)
            ext_rows.append(row)
            xors += x
            sp = np.zeros(SHA_EXTEND_SPONGE_COLUMNS, dtype=np.uint64)
            sp[SES_ROUND + r] = 1
            for at, v in zip((SES_W_M15, SES_W_M2, SES_W_M16, SES_W_M7), ins):
                sp[at:at + 4] = _le4(v)
            sp[SES_W_I:SES_W_I + 4] = _le4(wi)
            virts = [base + 4 * (i - 15), base + 4 * (i - 2), base + 4 * (i - 16), base + 4 * (i - 7)]
            sp[SES_INPUT_VIRT:SES_INPUT_VIRT + 4] = virts
            sp[SES_OUTPUT_VIRT] = base + 4 * i
            sp[SES_TIMESTAMP] = ts
            sp_rows.append(sp)
            for k in range(4):          # one lookup per byte: every word is read four times (ctl_looking_memory(i), i / 4)
                mem += [(0, 0, virts[k], ts, 1, ins[k])] * 4
            mem.append((0, 0, base + 4 * i, ts + 1, 0, wi))     # the CPU's write of w[i] (not looked up from this table)
    return ext_rows, sp_rows, xors, mem


def rows_to_trace(rows, ncols, log_n):
    n = 1 << log_n
    assert len(rows) <= n
    t = np.zeros((n, ncols), dtype=np.uint64)
    if rows:
        t[:len(rows)] = np.array(rows)
    return np.ascontiguousarray(t.T)


# ----------------------------------------------------------------------------- ShaCompress / ShaCompressSponge
# sha_compress/sha_compress_stark.rs:234-391 + columns.rs:9-55 (one row per round, 64 rounds + the output row, round
# flag 64), witness/util.rs:605-694 (sha_compress_sponge_log: what rows and memory reads one compression produces),
# sha_compress_sponge/sha_compress_sponge_stark.rs:120-237 + columns.rs:6-25.
SHA_K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
SC_STATE, SC_E_NOT, SC_W_I, SC_K_I, SC_S1_INTER, SC_S1, SC_E_AND_F, SC_E_NOT_AND_G, SC_CH = 0, 32, 36, 40, 44, 48, 52, 56, 60
SC_S0_INTER, SC_S0, SC_A_AND_B, SC_A_AND_C, SC_B_AND_C, SC_MAJ_INTER, SC_MAJ = 64, 68, 72, 76, 80, 84, 88
SC_E_RR6, SC_E_RR11, SC_E_RR25, SC_A_RR2, SC_A_RR13, SC_A_RR22 = 92, 98, 104, 110, 116, 122
SC_TEMP2, SC_D_ADD_TEMP1, SC_TEMP1_ADD_TEMP2 = 128, 134, 140          # value[4], carry[2]
SC_TIMESTAMP, SC_SEGMENT, SC_CONTEXT, SC_W_I_VIRT, SC_TEMP1, SC_ROUND = 146, 147, 148, 149, 150, 159   # temp1: value[4], carry[5]
SHA_COMPRESS_COLUMNS = SC_ROUND + 65
assert SHA_COMPRESS_COLUMNS == 224


def _wadd(row, at, ncarry, *vals):
    total = sum(vals)
    row[at:at + 4] = _le4(total & M32)
    assert (total >> 32) < ncarry
    row[at + 4 + (total >> 32)] = 1
    return total & M32


def sha_compress_row(st, w_i, k_i, rnd, w_i_virt, timestamp, ctx=0, seg=0):
    """One ShaCompress row (sha_compress_stark.rs:266-391); returns (row, next state, logic ops)."""
    a, b, c, d, e, f, g, h = st
    row = np.zeros(SHA_COMPRESS_COLUMNS, dtype=np.uint64)
    row[SC_TIMESTAMP], row[SC_SEGMENT], row[SC_CONTEXT], row[SC_W_I_VIRT] = timestamp, seg, ctx, w_i_virt
    row[SC_ROUND + rnd] = 1
    for i, v in enumerate(st):
        row[SC_STATE + 4 * i:SC_STATE + 4 * i + 4] = _le4(v)
    row[SC_W_I:SC_W_I + 4], row[SC_K_I:SC_K_I + 4] = _le4(w_i), _le4(k_i)
    rr6, rr11, rr25 = (_shift_op(row, at, e, r, True) for at, r in ((SC_E_RR6, 6), (SC_E_RR11, 11), (SC_E_RR25, 25)))
    s1i = rr6 ^ rr11
    s1 = s1i ^ rr25
    e_and_f = e & f
    e_not = (~e) & M32
    enag = e_not & g
    ch = e_and_f ^ enag
    for at, v in ((SC_S1_INTER, s1i), (SC_S1, s1), (SC_E_AND_F, e_and_f), (SC_E_NOT, e_not), (SC_E_NOT_AND_G, enag), (SC_CH, ch)):
        row[at:at + 4] = _le4(v)
    temp1 = _wadd(row, SC_TEMP1, 5, h, s1, ch, k_i, w_i)
    rr2, rr13, rr22 = (_shift_op(row, at, a, r, True) for at, r in ((SC_A_RR2, 2), (SC_A_RR13, 13), (SC_A_RR22, 22)))
    s0i = rr2 ^ rr13
    s0 = s0i ^ rr22
    ab, ac, bc = a & b, a & c, b & c
    mi = ab ^ ac
    maj = mi ^ bc
    for at, v in ((SC_S0_INTER, s0i), (SC_S0, s0), (SC_A_AND_B, ab), (SC_A_AND_C, ac), (SC_B_AND_C, bc), (SC_MAJ_INTER, mi), (SC_MAJ, maj)):
        row[at:at + 4] = _le4(v)
    temp2 = _wadd(row, SC_TEMP2, 2, s0, maj)
    new_e = _wadd(row, SC_D_ADD_TEMP1, 2, d, temp1)
    new_a = _wadd(row, SC_TEMP1_ADD_TEMP2, 2, temp1, temp2)
    ops = [(2, rr6, rr11), (2, s1i, rr25), (0, e, f), (0, e_not, g), (2, e_and_f, enag), (2, rr2, rr13), (2, s0i, rr22),
           (0, a, b), (0, a, c), (0, b, c), (2, ab, ac), (2, mi, bc)]
    return row, [new_a, a, b, c, new_e, e, f, g], ops


SCS_HX, SCS_OUTPUT_STATE, SCS_OUTPUT_HX, SCS_HX_VIRT, SCS_W_START_VIRT, SCS_TIMESTAMP = 0, 32, 64, 112, 120, 121
SCS_CONTEXT, SCS_SEGMENT, SCS_W_START_SEGMENT, SCS_W_START_CONTEXT, SCS_IS_REAL = 122, 123, 124, 125, 126
SHA_COMPRESS_SPONGE_COLUMNS = 127


def sha_compressions(calls, seed=41):
    """calls = [(h_ptr, w_ptr, timestamp)] with random chaining values and message schedules.  Returns (ShaCompress rows,
    ShaCompressSponge rows, logic ops, memory reads, [(hx, w, output_hx)])."""
    rng = np.random.default_rng(seed)
    c_rows, s_rows, logic, mem, io = [], [], [], [], []
    for h_ptr, w_ptr, ts in calls:
        hx = [int(v) for v in rng.integers(0, 1 << 32, size=8)]
        w = [int(v) for v in rng.integers(0, 1 << 32, size=64)]
        for j in range(8):
            mem += [(0, 0, h_ptr + 4 * j, ts, 1, hx[j])] * 4
        st = list(hx)
        for i in range(64):
            mem += [(0, 0, w_ptr + 4 * i, ts, 1, w[i])] * 4
            row, st, ops = sha_compress_row(st, w[i], SHA_K[i], i, w_ptr + 4 * i, ts)
            c_rows.append(row)
            logic += ops
        row, _, _ = sha_compress_row(st, 0, 0, 64, w_ptr + 4 * 64, ts)      # the 65th row carries the output state
        c_rows.append(row)
        sp = np.zeros(SHA_COMPRESS_SPONGE_COLUMNS, dtype=np.uint64)
        sp[SCS_TIMESTAMP], sp[SCS_IS_REAL], sp[SCS_W_START_VIRT] = ts, 1, w_ptr
        sp[SCS_HX_VIRT:SCS_HX_VIRT + 8] = [h_ptr + 4 * j for j in range(8)]
        out = []
        for j in range(8):
            sp[SCS_HX + 4 * j:SCS_HX + 4 * j + 4] = _le4(hx[j])
            sp[SCS_OUTPUT_STATE + 4 * j:SCS_OUTPUT_STATE + 4 * j + 4] = _le4(st[j])
            out.append(_wadd(sp, SCS_OUTPUT_HX + 6 * j, 2, hx[j], st[j]))
        s_rows.append(sp)
        io.append((hx, w, out))
    return c_rows, s_rows, logic, mem, io
