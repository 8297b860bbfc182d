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
