"""The Cpu table's constraints, transcribed a SECOND time -- in Python, straight from the reference's cpu/*.rs `eval_packed`
bodies in the order cpu_stark.rs:273-283 calls them -- for tests/test_independent_transcription.py.  Nothing here is derived from
zkm_b200/csrc/tables/cpu.h.  Column offsets follow the #[repr(C)] declaration order of CpuColumnsView (cpu/columns/mod.rs:68-118),
OpsColumnsView (columns/ops.rs:9-44), CpuBranchView / MemoryChannelView / MemIOView (mod.rs:17-66) and the general-purpose
views (columns/general.rs:122-178)."""

P = 0xFFFFFFFF00000001
NUM_GP_CHANNELS = 9

OPS = ["binary_op", "binary_imm_op", "eq_iszero", "logic_op", "logic_imm_op", "movz_op", "movn_op", "clz_op", "clo_op", "shift", "shift_imm",
       "keccak_general", "jumps", "jumpi", "jumpdirect", "branch", "pc", "get_context", "set_context", "exit_kernel", "m_op_load", "m_op_store",
       "nop", "ext", "ins", "maddu", "rdhwr", "signext8", "signext16", "swaphalf", "teq", "ror", "syscall"]
BRANCH = ["should_jump", "gt", "lt", "eq", "is_gt", "is_lt", "is_eq", "is_ge", "is_le", "is_ne"]
MEMIO = ["is_lh", "is_lwl", "is_lw", "is_lbu", "is_lhu", "is_lwr", "is_sb", "is_sh", "is_swl", "is_sw", "is_swr", "is_ll", "is_sc", "is_sdc1", "is_lb",
         "aux_filter"]


class Row:
    """One row of the Cpu table addressed by the reference's field names."""

    def __init__(self, v):
        assert len(v) == 259
        self.v = v
        at = 0

        def take(k=1):
            nonlocal at
            r = at
            at += k
            return r
        self.is_bootstrap_kernel, self.is_exit_kernel, self.context, self.code_context = (v[take()] for _ in range(4))
        self.program_counter, self.next_program_counter, self.is_kernel_mode = (v[take()] for _ in range(3))
        self.op = {name: v[take()] for name in OPS}
        self.branch = {name: v[take()] for name in BRANCH}
        self.opcode_bits = [v[take()] for _ in range(6)]
        self.rs_bits = [v[take()] for _ in range(5)]
        self.rt_bits = [v[take()] for _ in range(5)]
        self.rd_bits = [v[take()] for _ in range(5)]
        self.shamt_bits = [v[take()] for _ in range(5)]
        self.func_bits = [v[take()] for _ in range(6)]
        self.is_poseidon_sponge, self.is_keccak_sponge, self.is_sha_extend_sponge, self.is_sha_compress_sponge = (v[take()] for _ in range(4))
        g = take(102)
        self.general = v[g:g + 102]
        self.memio = {name: v[take()] for name in MEMIO}
        self.clock = v[take()]
        self.mem_channels = []
        for _ in range(NUM_GP_CHANNELS):
            c = take(6)
            self.mem_channels.append(dict(used=v[c], is_read=v[c + 1], addr_context=v[c + 2], addr_segment=v[c + 3], addr_virtual=v[c + 4], value=v[c + 5]))
        assert at == 259

    # general-purpose views (columns/general.rs)
    def syscall(self):
        g = self.general
        return dict(cond=g[0:12], sysnum=g[12:24], a0=g[24:27], a1=g[27])

    def misc(self):
        g = self.general
        return dict(rs_bits=g[0:32], is_msb=g[32:64], is_lsb=g[64:96], auxm=g[96], auxl=g[97], auxs=g[98], rd_index=g[99], rd_index_eq_0=g[100],
                    rd_index_eq_29=g[101])

    def io(self):
        g = self.general
        return dict(rs_le=g[0:32], rt_le=g[32:64], mem_le=g[64:96], aux_rs0_mul_rs1=g[96])

    def logic(self):
        return dict(diff_pinv=self.general[0])

    def shift(self):
        return dict(high_limb_sum_inv=self.general[0])


def limb_from_bits_le(bits):        # util.rs limb_from_bits_le: sum_i bit_i 2^i
    return sum(b << i for i, b in enumerate(bits))


def bootstrap_kernel(lv, nv, yc):
    """bootstrap_kernel.rs:308-353."""
    local, nxt = lv.is_bootstrap_kernel, nv.is_bootstrap_kernel
    yc.constraint_first_row(local - 1)
    yc.constraint_last_row(local)
    delta = nxt - local
    yc.constraint_transition(delta * (delta + 1))
    for ch in lv.mem_channels:
        f = local * ch["used"]
        yc.constraint(f * ch["addr_context"])
        yc.constraint(f * (ch["addr_segment"] - 0))           # Segment::Code = 0
    for ch in lv.mem_channels:
        yc.constraint_transition(delta * ch["used"])


def decode(lv, yc):
    """decode.rs:66-100 with OPCODES :28-42 and COMBINED_OPCODES :47-55."""
    k = lv.is_kernel_mode
    yc.constraint(k * (k - 1))
    for b in lv.opcode_bits:
        yc.constraint(b * (b - 1))
    opcodes = ["eq_iszero", "keccak_general", "jumps", "branch", "pc", "get_context", "set_context", "exit_kernel"]
    combined = ["logic_op", "binary_op", "binary_imm_op", "shift", "shift_imm", "m_op_load", "m_op_store"]
    for name in opcodes + combined:
        yc.constraint(lv.op[name] * (lv.op[name] - 1))
    s = sum(lv.op[name] for name in opcodes + combined)
    yc.constraint(s * (s - 1))


def jumps(lv, nv, yc):
    """jumps.rs:17-122 (jump / jumpi / jumpdirect) and :243-408 (branch), called in that order (:675-683)."""
    INV32 = 18446744065119617026
    OVER = 1 << 32
    ch = lv.mem_channels
    is_jump, is_jumpi, is_jd = lv.op["jumps"], lv.op["jumpi"], lv.op["jumpdirect"]
    is_link, is_linki = is_jump * lv.func_bits[0], is_jumpi * lv.opcode_bits[0]
    yc.constraint(is_jump * (nv.next_program_counter - ch[0]["value"]))
    yc.constraint(is_jump * (limb_from_bits_le(lv.rs_bits) - ch[0]["addr_virtual"]))
    imm = [0, 0] + lv.func_bits + lv.shamt_bits + lv.rd_bits + lv.rt_bits + lv.rs_bits
    yc.constraint(is_jumpi * (nv.next_program_counter - (ch[2]["value"] + limb_from_bits_le(imm))))
    aux = ch[2]["value"]
    off = [0, 0] + lv.func_bits + lv.shamt_bits + lv.rd_bits + [lv.rd_bits[4]] * 14
    yc.constraint(is_jd * (aux - limb_from_bits_le(off)))
    dst = lv.program_counter + 4 + aux
    yc.constraint(is_jd * (nv.next_program_counter - dst) * (nv.next_program_counter + OVER - dst))
    yc.constraint((is_link + is_linki + is_jd) * (lv.program_counter + 8 - ch[1]["value"]))
    link_reg = ch[1]["addr_virtual"]
    yc.constraint(is_link * (link_reg - limb_from_bits_le(lv.rd_bits)))
    yc.constraint((is_linki + is_jd) * (link_reg - 31))
    # branch
    b = lv.branch
    f = lv.op["branch"]
    norm = b["is_eq"] + b["is_ne"] + b["is_le"] + b["is_gt"]
    special = b["is_ge"] + b["is_lt"]
    src1, src2, aux1, aux2, aux3, aux4 = (ch[i]["value"] for i in range(6))
    sj = b["should_jump"]
    yc.constraint(sj * (1 - sj))
    yc.constraint(sj * (1 - f))
    yc.constraint(f * (1 - (norm + special)))
    yc.constraint(f * (1 - (b["lt"] + b["gt"] + b["eq"])))
    yc.constraint(f * (aux4 - limb_from_bits_le(off)))
    bdst = lv.program_counter + 4 + aux4
    yc.constraint(sj * (nv.next_program_counter - bdst) * (nv.next_program_counter + OVER - bdst))
    yc.constraint(f * (1 - sj) * (nv.next_program_counter - (lv.program_counter + 8)))
    yc.constraint(f * (aux1 + src2 - src1) * (aux1 + src2 - src1 - OVER))
    yc.constraint(f * (aux2 + src1 - src2) * (aux2 + src1 - src2 - OVER))
    yc.constraint(f * aux1 * ((aux1 + aux2) - OVER))
    yc.constraint(f * aux3 * (1 - aux3))
    yc.constraint(f * (ch[0]["addr_virtual"] - limb_from_bits_le(lv.rs_bits)))
    rt_reg = ch[1]["addr_virtual"]
    yc.constraint(norm * (rt_reg - limb_from_bits_le(lv.rt_bits)))
    yc.constraint(special * rt_reg * (1 - rt_reg))
    ca = src2 + aux1 - src1
    yc.constraint(f * ca * (OVER - ca))
    lt = ca * INV32
    yc.constraint(b["lt"] * (1 - lt))
    cb = src1 + aux2 - src2
    yc.constraint(f * cb * (OVER - cb))
    gt = cb * INV32
    yc.constraint(b["gt"] * (1 - gt))
    ne = lt + gt
    yc.constraint(b["eq"] * ne)
    lt = b["lt"] * (1 - aux3) + (1 - b["lt"]) * aux3
    gt = b["gt"] * (1 - aux3) + (1 - b["gt"]) * aux3
    for flag, want in (("is_eq", 1 - ne), ("is_ne", ne), ("is_le", 1 - gt), ("is_ge", 1 - lt), ("is_gt", gt), ("is_lt", lt)):
        yc.constraint(b[flag] * (1 - f))
        yc.constraint(b[flag] * (sj - want))


def membus(lv, yc):
    """membus.rs:34-43."""
    yc.constraint(lv.code_context - (1 - lv.is_kernel_mode) * lv.context)
    for ch in lv.mem_channels:
        yc.constraint(ch["used"] * (ch["used"] - 1))


def _sign_extend(bits, n):          # memio.rs:42-49
    return bits[:n] + [bits[n - 1]] * (32 - n)


def _place(*pieces):
    """A 32-entry little-endian bit vector assembled from (start, bits) pieces; unset positions are zero."""
    out = [0] * 32
    for start, bits in pieces:
        out[start:start + len(bits)] = bits
    assert len(out) == 32
    return out


def memio(lv, nv, yc):
    """memio.rs:175-433 (loads) and :738-958 (stores), called in that order (:1217-1224); helpers :17-25 (load_offset), :65-77
    (enforce_half_word), :104-127 (enforce_byte)."""
    REGISTER_FILE = 4                                    # memory/segments.rs
    ch, io = lv.mem_channels, lv.io()
    rs_l, rt_l, mem_l, aux_mul = io["rs_le"], io["rt_le"], io["mem_le"], io["aux_rs0_mul_rs1"]
    offset = limb_from_bits_le(_sign_extend(_place((0, lv.func_bits), (6, lv.shamt_bits), (11, lv.rd_bits)), 16))
    rs, rt, mem = ch[0]["value"], ch[1]["value"], ch[3]["value"]
    aux_filter = lv.memio["aux_filter"]
    L = limb_from_bits_le

    def half_word(op, val_1, val_0):
        yc.constraint(op * ((rs_l[1] - 1) * (mem - val_0) + rs_l[1] * (mem - val_1)))

    def byte(op, v00, v10, v01, v11):
        yc.constraint(op * (rs_l[0] * rs_l[1] - aux_mul))
        total = ((mem - v00) * (aux_mul - rs_l[1] - rs_l[0] + 1) + (mem - v10) * (aux_mul - rs_l[0]) + (mem - v01) * (aux_mul - rs_l[1])
                 + (mem - v11) * aux_mul)
        yc.constraint(total * op)

    def common(filt):
        yc.constraint(filt * (1 - aux_filter))
        yc.constraint(filt * (ch[0]["addr_segment"] - REGISTER_FILE))
        yc.constraint(filt * (ch[1]["addr_segment"] - REGISTER_FILE))
        virt_raw = rs + offset
        yc.constraint(aux_filter * (L(rs_l) - virt_raw) * (L(rs_l) + (1 << 32) - virt_raw))
        yc.constraint(filt * (L(rt_l) - rt))
        yc.constraint(filt * (L([0, 0] + rs_l[2:]) - ch[2]["addr_virtual"]))

    # ---- loads
    filt = lv.op["m_op_load"] * lv.opcode_bits[5]
    common(filt)
    m = lv.memio
    half_word(m["is_lh"], L(_sign_extend(_place((0, mem_l[0:16])), 16)), L(_sign_extend(_place((0, mem_l[16:32])), 16)))
    byte(m["is_lwl"], L(mem_l), L(_place((0, rt_l[0:8]), (8, mem_l[0:24]))), L(_place((0, rt_l[0:16]), (16, mem_l[0:16]))),
         L(_place((0, rt_l[0:24]), (24, mem_l[0:8]))))
    yc.constraint(m["is_lw"] * (mem - L(mem_l)))
    byte(m["is_lbu"], L(_place((0, mem_l[24:32]))), L(_place((0, mem_l[16:24]))), L(_place((0, mem_l[8:16]))), L(_place((0, mem_l[0:8]))))
    half_word(m["is_lhu"], L(_place((0, mem_l[0:16]))), L(_place((0, mem_l[16:32]))))
    byte(m["is_lwr"], L(_place((0, mem_l[24:32]), (8, rt_l[8:32]))), L(_place((0, mem_l[16:32]), (16, rt_l[16:32]))),
         L(_place((0, mem_l[8:32]), (24, rt_l[24:32]))), L(mem_l))
    yc.constraint(m["is_ll"] * (mem - L(mem_l)))
    byte(m["is_lb"], *(L(_sign_extend(_place((0, mem_l[a:a + 8])), 8)) for a in (24, 16, 8, 0)))
    for c in ch[6:NUM_GP_CHANNELS - 1]:
        yc.constraint(filt * c["used"])
    # ---- stores
    filt = lv.op["m_op_store"] * lv.opcode_bits[5]
    common(filt)
    byte(m["is_sb"], L(_place((0, mem_l[0:24]), (24, rt_l[0:8]))), L(_place((0, mem_l[0:16]), (16, rt_l[0:8]), (24, mem_l[24:32]))),
         L(_place((0, mem_l[0:8]), (8, rt_l[0:8]), (16, mem_l[16:32]))), L(_place((0, rt_l[0:8]), (8, mem_l[8:32]))))
    half_word(m["is_sh"], L(_place((0, rt_l[0:16]), (16, mem_l[16:32]))), L(_place((0, mem_l[0:16]), (16, rt_l[0:16]))))
    byte(m["is_swl"], L(rt_l), L(_place((0, rt_l[8:32]), (24, mem_l[24:32]))), L(_place((0, rt_l[16:32]), (16, mem_l[16:32]))),
         L(_place((0, rt_l[24:32]), (8, mem_l[8:32]))))
    yc.constraint(m["is_sw"] * (mem - L(rt_l)))
    byte(m["is_swr"], L(_place((0, mem_l[0:24]), (24, rt_l[0:8]))), L(_place((0, mem_l[0:16]), (16, rt_l[0:16]))),
         L(_place((0, mem_l[0:8]), (8, rt_l[0:24]))), L(rt_l))
    yc.constraint(m["is_sc"] * (mem - L(rt_l)))
    yc.constraint(m["is_sdc1"] * mem)
    for c in ch[6:NUM_GP_CHANNELS - 1]:
        yc.constraint(filt * c["used"])


def shift(lv, yc):
    """shift.rs:11-117: variable then immediate displacement."""
    SHIFT_TABLE = 3                                      # memory/segments.rs
    two_exp = lv.mem_channels[3]
    for is_shift, disp in ((lv.op["shift"], lv.mem_channels[0]["value"]), (lv.op["shift_imm"], limb_from_bits_le(lv.shamt_bits))):
        yc.constraint(is_shift * two_exp["used"] * (two_exp["is_read"] - 1))
        yc.constraint(is_shift * two_exp["addr_context"])
        yc.constraint(is_shift * (two_exp["addr_segment"] - SHIFT_TABLE))
        yc.constraint(is_shift * (two_exp["addr_virtual"] - disp))


def count(lv, yc):
    """count.rs:10-70 (CLZ / CLO)."""
    clz, clo = lv.op["clz_op"], lv.op["clo_op"]
    f = clo + clz
    ch, io = lv.mem_channels, lv.io()
    yc.constraint(f * (limb_from_bits_le(lv.opcode_bits) - 0b011100))
    func = limb_from_bits_le(lv.func_bits)
    yc.constraint(clz * (func - 0b100000))
    yc.constraint(clo * (func - 0b100001))
    yc.constraint(f * (ch[0]["addr_virtual"] - limb_from_bits_le(lv.rs_bits)))
    yc.constraint(f * (ch[1]["addr_virtual"] - limb_from_bits_le(lv.rd_bits)))
    rs, bits = ch[0]["value"], io["rs_le"]
    for b in bits:
        yc.constraint(f * b * (1 - b))
    total = limb_from_bits_le(bits)
    yc.constraint(clz * (rs - total))
    yc.constraint(clo * (0xFFFFFFFF - rs - total))
    rd = ch[1]["value"]
    is_eqs, invs = iter(io["rt_le"]), iter(io["mem_le"])
    yc.constraint(f * bits[31] * rd)
    for i in range(30, -1, -1):
        partial = limb_from_bits_le(bits[i:])
        is_eq, inv = next(is_eqs), next(invs)
        diff = partial - 1
        yc.constraint(f * diff * is_eq)
        yc.constraint(f * (diff * inv + is_eq - 1))
        yc.constraint(f * is_eq * (rd - (31 - i)))
        if i == 0:
            is_eq, inv = next(is_eqs), next(invs)
            yc.constraint(f * partial * is_eq)
            yc.constraint(f * (partial * inv + is_eq - 1))
            yc.constraint(f * is_eq * (rd - 32))


def syscall(lv, yc):
    """syscall.rs:12-230."""
    MIPSEBADF = 9                                        # witness/operation.rs:98
    f = lv.op["syscall"]
    ch, sc = lv.mem_channels, lv.syscall()
    a0, a1, a2 = ch[1]["value"], ch[2]["value"], ch[3]["value"]
    v0 = v1 = 0
    result_v0, result_v1 = ch[4]["value"], ch[5]["value"]
    cond, sysnum, A0 = sc["cond"], sc["sysnum"], sc["a0"]
    is_sysmap, sz_mid_nz, sz_mid_zero, sz, sz_in_nz = sysnum[1], sc["a1"], sysnum[10], a1, sysnum[9]
    a0_zero, a0_nz = A0[0], A0[2]
    heap, result_heap = ch[6]["value"], ch[7]["value"]
    yc.constraint(f * (cond[0] - is_sysmap * a0_zero))
    yc.constraint(f * (cond[1] - cond[0] * sz_mid_nz))
    yc.constraint(f * cond[1] * (heap + sz_in_nz - result_heap))
    yc.constraint(f * (cond[2] - cond[0] * sz_mid_zero))
    yc.constraint(f * cond[2] * (heap + sz - result_heap))
    yc.constraint(f * cond[0] * (heap - result_v0))
    yc.constraint(f * (cond[3] - is_sysmap * a0_nz))
    yc.constraint(f * cond[3] * (a0 - result_v0))
    is_sysbrk, brk_gt, brk_le, initial_brk = sysnum[2], cond[10], cond[11], ch[6]["value"]
    yc.constraint(f * is_sysbrk * (1 - (brk_gt + brk_le)))
    yc.constraint(f * brk_gt * (a0 - result_v0))
    yc.constraint(f * brk_le * (initial_brk - result_v0))
    yc.constraint(f * is_sysbrk * (v1 - result_v1))
    is_sysclone = sysnum[3]
    yc.constraint(f * is_sysclone * (1 - result_v0))
    yc.constraint(f * is_sysclone * (v1 - result_v1))
    is_sysread = sysnum[5]
    yc.constraint(f * (cond[4] - is_sysread * A0[2]))
    yc.constraint(f * cond[4] * (0xFFFFFFFF - result_v0))
    yc.constraint(f * cond[4] * (MIPSEBADF - result_v1))
    yc.constraint(f * (cond[5] - is_sysread * A0[0]))
    yc.constraint(f * cond[5] * (v0 - result_v0))
    yc.constraint(f * cond[5] * (v1 - result_v1))
    is_syswrite = sysnum[6]
    yc.constraint(f * (cond[6] - is_syswrite * A0[2]))
    yc.constraint(f * cond[6] * (0xFFFFFFFF - result_v0))
    yc.constraint(f * cond[6] * (MIPSEBADF - result_v1))
    yc.constraint(f * (cond[7] - is_syswrite * A0[1]))
    yc.constraint(f * cond[7] * (a2 - result_v0))
    yc.constraint(f * cond[7] * (v1 - result_v1))
    is_sysfcntl = sysnum[7]
    yc.constraint(f * (cond[8] - is_sysfcntl * A0[0]))
    yc.constraint(f * cond[8] * (0 - result_v0))
    yc.constraint(f * cond[8] * (v1 - result_v1))
    yc.constraint(f * (cond[9] - is_sysfcntl * A0[1]))
    yc.constraint(f * cond[9] * (1 - result_v0))
    yc.constraint(f * cond[9] * (v1 - result_v1))
    yc.constraint(f * (is_sysfcntl - cond[8] - cond[9] - is_sysfcntl * A0[2]))
    yc.constraint(f * (is_sysfcntl - cond[8] - cond[9]) * (0xFFFFFFFF - result_v0))
    yc.constraint(f * (is_sysfcntl - cond[8] - cond[9]) * (MIPSEBADF - result_v1))
    yc.constraint(f * sysnum[8] * (a0 - ch[6]["value"]))


def bits(lv, yc):
    """bits.rs:9-62 (SEH / SEB / WSBH)."""
    seh, seb, wsbh = lv.op["signext16"], lv.op["signext8"], lv.op["swaphalf"]
    f = seh + seb + wsbh
    ch = lv.mem_channels
    yc.constraint(f * (ch[0]["addr_virtual"] - limb_from_bits_le(lv.rt_bits)))
    yc.constraint(f * (ch[1]["addr_virtual"] - limb_from_bits_le(lv.rd_bits)))
    rt, b = ch[0]["value"], lv.io()["rt_le"]
    for x in b:
        yc.constraint(f * x * (1 - x))
    yc.constraint(f * (rt - limb_from_bits_le(b)))
    rd = ch[1]["value"]
    yc.constraint(seb * (rd - limb_from_bits_le(b[:7] + [b[7]] * 25)))
    yc.constraint(seh * (rd - limb_from_bits_le(b[:15] + [b[15]] * 17)))
    yc.constraint(wsbh * (rd - limb_from_bits_le(b[8:16] + b[0:8] + b[24:32] + b[16:24])))


def misc(lv, yc):
    """misc.rs:821-832: rdhwr (:10-44), condmov (:108-145), teq (:197-225), extract (:266-316), ror (:563-603), insert (:397-466),
    maddu (:659-725), in that order."""
    L = limb_from_bits_le
    ch, mv = lv.mem_channels, lv.misc()
    # rdhwr
    f = lv.op["rdhwr"]
    yc.constraint(f * (ch[0]["addr_virtual"] - L(lv.rt_bits)))
    yc.constraint(f * (mv["rd_index"] - L(lv.rd_bits)))
    rt_val, local_user = ch[0]["value"], ch[1]["value"]
    yc.constraint(f * mv["rd_index_eq_0"] * mv["rd_index"])
    yc.constraint(f * mv["rd_index_eq_0"] * (rt_val - 1))
    yc.constraint(f * mv["rd_index_eq_29"] * (mv["rd_index"] - 29))
    yc.constraint(f * mv["rd_index_eq_29"] * (rt_val - local_user))
    yc.constraint(f * (1 - mv["rd_index_eq_29"] - mv["rd_index_eq_0"]) * rt_val)
    # condmov
    rs, rt, rd, out, mov = (ch[i]["value"] for i in range(5))
    movn, movz = lv.op["movn_op"], lv.op["movz_op"]
    f = movn + movz
    is_ne = lv.logic()["diff_pinv"] * rt
    yc.constraint(movn * (mov - is_ne))
    yc.constraint(movz * (mov - (1 - is_ne)))
    yc.constraint(f * mov * (1 - mov))
    yc.constraint(f * (out - (mov * rs + (1 - mov) * rd)))
    # teq
    f = lv.op["teq"]
    yc.constraint(f * (ch[1]["addr_virtual"] - L(lv.rt_bits)))
    yc.constraint(f * (ch[0]["addr_virtual"] - L(lv.rs_bits)))
    yc.constraint(f * (1 - (ch[0]["value"] - ch[1]["value"]) * lv.logic()["diff_pinv"]))
    # extract
    f = lv.op["ext"]
    yc.constraint(f * (ch[1]["addr_virtual"] - L(lv.rt_bits)))
    yc.constraint(f * (ch[0]["addr_virtual"] - L(lv.rs_bits)))
    msbd, rs_bits, lsb = L(lv.rd_bits), mv["rs_bits"], L(lv.shamt_bits)
    msb = lsb + msbd
    auxm, auxl, auxs = mv["auxm"], mv["auxl"], mv["auxs"]
    yc.constraint(f * (ch[1]["value"] * auxs + auxl - auxm))
    for i in range(32):
        is_msb, is_lsb = mv["is_msb"][i], mv["is_lsb"][i]
        yc.constraint(f * is_msb * (msb - i))
        yc.constraint(f * is_msb * (auxm - L(rs_bits[0:i + 1])))
        yc.constraint(f * is_lsb * (lsb - i))
        yc.constraint(f * is_lsb * (auxl - L(rs_bits[0:i])))
        yc.constraint(f * is_lsb * (auxs - (1 << i)))
    # ror
    f = lv.op["ror"]
    yc.constraint(f * (ch[1]["addr_virtual"] - L(lv.rd_bits)))
    yc.constraint(f * (ch[0]["addr_virtual"] - L(lv.rt_bits)))
    rt_bits, sa, rd_result = mv["rs_bits"], L(lv.shamt_bits), ch[1]["value"]
    for i in range(32):
        is_sa = mv["is_lsb"][i]
        yc.constraint(f * is_sa * (sa - i))
        yc.constraint(f * is_sa * (rd_result - L(rt_bits[i:32] + rt_bits[0:i])))
    # insert
    f = lv.op["ins"]
    yc.constraint(f * (ch[1]["addr_virtual"] - L(lv.rt_bits)))
    yc.constraint(f * (ch[2]["addr_virtual"] - L(lv.rt_bits)))
    yc.constraint(f * (ch[0]["addr_virtual"] - L(lv.rs_bits)))
    msb, lsb = L(lv.rd_bits), L(lv.shamt_bits)
    yc.constraint(f * (ch[2]["value"] - auxm - auxl * auxs))
    for i in range(32):
        is_msb, is_lsb = mv["is_msb"][i], mv["is_lsb"][i]
        yc.constraint(f * is_lsb * (lsb - i))
        yc.constraint(f * is_lsb * (auxs - (1 << i)))
        yc.constraint(f * is_msb * (msb - lsb - i))
        yc.constraint(f * is_msb * (auxl - L(rs_bits[0:i + 1])))
    # maddu
    f = lv.op["maddu"]
    yc.constraint(f * (ch[0]["addr_virtual"] - L(lv.rs_bits)))
    yc.constraint(f * (ch[1]["addr_virtual"] - L(lv.rt_bits)))
    yc.constraint(f * (ch[2]["addr_virtual"] - 33))
    yc.constraint(f * (ch[4]["addr_virtual"] - 33))
    yc.constraint(f * (ch[3]["addr_virtual"] - 32))
    yc.constraint(f * (ch[5]["addr_virtual"] - 32))
    rs, rt, hi, lo, hi_res, lo_res = (ch[i]["value"] for i in range(6))
    carry, scale = mv["auxm"], 1 << 32
    yc.constraint(f * carry * (carry - scale))
    yc.constraint(f * (rs * rt + hi * scale + lo - carry * scale - (hi_res * scale + lo_res)))


MODULES = [("bootstrap_kernel", lambda lv, nv, yc: bootstrap_kernel(lv, nv, yc)), ("decode", lambda lv, nv, yc: decode(lv, yc)),
           ("jumps", lambda lv, nv, yc: jumps(lv, nv, yc)), ("membus", lambda lv, nv, yc: membus(lv, yc)),
           ("memio", lambda lv, nv, yc: memio(lv, nv, yc)), ("shift", lambda lv, nv, yc: shift(lv, yc)), ("count", lambda lv, nv, yc: count(lv, yc)),
           ("syscall", lambda lv, nv, yc: syscall(lv, yc)), ("bits", lambda lv, nv, yc: bits(lv, yc)), ("misc", lambda lv, nv, yc: misc(lv, yc))]


def cpu_constraints(lv_raw, nv_raw, yc, upto=None):
    lv, nv = Row(lv_raw), Row(nv_raw)
    marks = []
    for name, fn in MODULES[:upto]:
        fn(lv, nv, yc)
        marks.append((name, yc.count))
    return marks
