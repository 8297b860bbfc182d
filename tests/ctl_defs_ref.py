"""The 15 cross-table lookups of AllStark and the two in-table logUp lookups, transcribed a SECOND time -- in Python, straight from
the reference: all_stark.rs:136-542 (which tables look where, in which order), cpu/cpu_stark.rs:25-244, arithmetic_stark.rs:32-112,
269-276, logic.rs:52-76, memory_stark.rs:29-39,476-483, poseidon_stark.rs:29-49, poseidon_sponge_stark.rs:28-128, keccak_stark.rs:
34-52 (+ keccak/columns.rs:14-37), keccak_sponge_stark.rs:28-200, sha_extend_stark.rs:30-113, sha_extend_sponge_stark.rs:30-101,
sha_compress_stark.rs:38-252, sha_compress_sponge_stark.rs:29-89 -- with the Column / Filter semantics of cross_table_lookup.rs:33-265.
Nothing here is derived from zkm_b200/csrc/tables/*.h.  A column is a function (local row, next row) -> value."""

TABLES = ["Arithmetic", "Cpu", "Poseidon", "PoseidonSponge", "Keccak", "KeccakSponge", "ShaExtend", "ShaExtendSponge", "ShaCompress",
          "ShaCompressSponge", "Logic", "Memory"]
NCOLS = dict(zip(TABLES, [54, 259, 262, 110, 2431, 470, 78, 76, 224, 127, 69, 13]))
NUM_CHANNELS = 10                      # cpu/membus.rs: code channel + 9 general-purpose channels
OP_XOR, OP_AND = 0b100110 << 6, 0b100100 << 6


def single(c):
    return lambda lv, nv: lv[c]


def const(k):
    return lambda lv, nv: k


def lincomb(pairs, constant=0):
    pairs = list(pairs)
    return lambda lv, nv: sum(lv[c] * f for c, f in pairs) + constant


def le_bits(cols):
    return lincomb((c, 1 << i) for i, c in enumerate(cols))


def le_bytes(cols):
    return lincomb((c, 256 ** i) for i, c in enumerate(cols))


def col_sum(cols):
    return lincomb((c, 1) for c in cols)


def simple(col):                         # Filter::new_simple
    return col


# ------------------------------------------------------------------------------------------------------ column maps
class Cpu:
    CODE_CONTEXT, PROGRAM_COUNTER = 3, 4
    OP = 7                               # OpsColumnsView: 33 flags
    BINARY_OP, BINARY_IMM_OP, LOGIC_OP, SHIFT, SHIFT_IMM = OP + 0, OP + 1, OP + 3, OP + 9, OP + 10
    OPCODE_BITS, FUNC_BITS = list(range(50, 56)), list(range(76, 82))
    IS_POSEIDON_SPONGE, IS_KECCAK_SPONGE, IS_SHA_EXTEND_SPONGE, IS_SHA_COMPRESS_SPONGE = 82, 83, 84, 85
    GENERAL, CLOCK, MEM_CHANNELS = 86, 204, 205

    @staticmethod
    def ch(i, field):
        return Cpu.MEM_CHANNELS + 6 * i + ["used", "is_read", "addr_context", "addr_segment", "addr_virtual", "value"].index(field)


def _sponge_map(rate_bytes, rate_words, cap_words, digest_words, digest_as_bytes):
    at, m = 0, {}
    for name, k in (("is_full", 1), ("context", 1), ("segment", 1), ("virt", rate_words), ("timestamp", 1), ("len", 1), ("already", 1),
                    ("final_len", rate_bytes), ("orig_rate", rate_words), ("orig_cap", cap_words), ("block_bytes", rate_bytes),
                    ("new_rate", rate_words), ("partial", rate_words + cap_words - digest_words),
                    ("digest", 4 * digest_words if digest_as_bytes else digest_words)):
        m[name] = list(range(at, at + k))
        at += k
    return m


KS = _sponge_map(136, 34, 16, 8, True)       # keccak_sponge/columns.rs ("new_rate" = xored_rate_u32s)
PS = _sponge_map(32, 8, 4, 4, False)         # poseidon_sponge/columns.rs


def cpu_binops():
    return [single(Cpu.ch(i, "value")) for i in range(3)]


def cpu_timestamp():
    return lincomb([(Cpu.CLOCK, NUM_CHANNELS)])


def entry(table, filt, columns):
    return dict(table=table, filter=filt, columns=columns)


def sponge_memory(m, i):
    start = (i // 4) * 4
    return [const(1), single(m["context"][0]), single(m["segment"][0]), single(m["virt"][i // 4]),
            le_bytes([m["block_bytes"][start + 3], m["block_bytes"][start + 2], m["block_bytes"][start + 1], m["block_bytes"][start]]),
            single(m["timestamp"][0])]


def sponge_memory_filter(m, i, rate_bytes):
    if i == rate_bytes - 1:
        return single(m["is_full"][0])
    return col_sum(m["is_full"] + m["final_len"][i + 1:])


def sponge_any_block(m):
    return col_sum(m["is_full"] + m["final_len"])


# SHA column maps (sha_extend/columns.rs, sha_extend_sponge/columns.rs, sha_compress/columns.rs, sha_compress_sponge/columns.rs)
SE = dict(w_i=0, w15=8, w2=12, w16=16, w7=20, s0_inter=24, s0=28, s1_inter=32, s1=36, rr7=40, rr18=46, rr17=52, rr19=58, rs10=64, rs3=70,
          timestamp=76, is_real=77)
SES = dict(round=0, w15=48, w2=52, w16=56, w7=60, w_i=64, input_virt=68, output_virt=72, context=73, segment=74, timestamp=75)
SC = dict(state=0, e_not=32, w_i=36, k_i=40, s1_inter=44, s1=48, e_and_f=52, e_not_and_g=56, ch=60, s0_inter=64, s0=68, a_and_b=72, a_and_c=76,
          b_and_c=80, maj_inter=84, maj=88, e_rr_6=92, e_rr_11=98, e_rr_25=104, a_rr_2=110, a_rr_13=116, a_rr_22=122, timestamp=146, segment=147,
          context=148, w_i_virt=149, round=159)
SCS = dict(hx=0, output_state=32, output_hx=64, hx_virt=112, w_start_virt=120, timestamp=121, context=122, segment=123, w_start_segment=124,
           w_start_context=125, is_real=126)
r4 = lambda at: list(range(at, at + 4))


def logic_row(op, a, b, out):
    return [const(op), le_bytes(r4(a)), le_bytes(r4(b)), le_bytes(r4(out))]


def all_cross_table_lookups():
    """Returns [(looking entries, looked entry)] in the order of all_stark.rs:136-154."""
    opfunc = Cpu.OPCODE_BITS + Cpu.FUNC_BITS
    ctls = []
    # 0 ctl_arithmetic
    arith_ops = [(0, 0b100000 << 6), (1, 0b100001 << 6), (2, 0b001000), (3, 0b001001), (4, 0b100010 << 6), (5, 0b100011 << 6), (6, 0b011000 << 6),
                 (7, 0b011001 << 6), (8, 0b011100 + (0b000010 << 6)), (9, 0b011010 << 6), (10, 0b011011 << 6), (11, 0b000100 << 6), (12, 0b000110 << 6),
                 (13, 0b000111 << 6), (14, 0), (15, 0b000010 << 6), (16, 0b000011 << 6), (17, 0b101010 << 6), (18, 0b101011 << 6), (19, 0b001010),
                 (20, 0b001011), (21, 0b001111), (22, 0b010000 << 6), (23, 0b010001 << 6), (24, 0b010010 << 6), (25, 0b010011 << 6)]
    arith_cols = [lincomb(arith_ops)] + [lincomb([(r, 1), (r + 1, 1 << 16)]) for r in (26, 28, 32)]     # IN0, IN1, OUT
    ctls.append(([entry("Cpu", col_sum([Cpu.BINARY_OP, Cpu.SHIFT, Cpu.SHIFT_IMM]), [le_bits(opfunc)] + cpu_binops()),
                  entry("Cpu", single(Cpu.BINARY_IMM_OP), [le_bits(Cpu.OPCODE_BITS)] + cpu_binops())],
                 entry("Arithmetic", col_sum([c for c, _ in arith_ops]), arith_cols)))
    # 1 ctl_poseidon_sponge
    cpu_sponge = lambda n_hash: ([single(Cpu.ch(i, "value")) for i in range(4)] + [cpu_timestamp()] + [single(Cpu.GENERAL + k) for k in range(n_hash)])
    ctls.append(([entry("Cpu", single(Cpu.IS_POSEIDON_SPONGE), cpu_sponge(4))],
                 entry("PoseidonSponge", col_sum(PS["final_len"]),
                       [single(c) for c in [PS["context"][0], PS["segment"][0], PS["virt"][0], PS["len"][0], PS["timestamp"][0]] + PS["digest"]])))
    # 2, 3 ctl_poseidon_inputs / outputs (Poseidon: FILTER 0, in 1..12, out 13..24, TIMESTAMP 25)
    ctls.append(([entry("PoseidonSponge", sponge_any_block(PS), [single(c) for c in PS["new_rate"] + PS["orig_cap"] + PS["timestamp"]])],
                 entry("Poseidon", single(0), [single(c) for c in list(range(1, 13)) + [25]])))
    ctls.append(([entry("PoseidonSponge", sponge_any_block(PS), [single(c) for c in PS["digest"] + PS["partial"] + PS["timestamp"]])],
                 entry("Poseidon", single(0), [single(c) for c in list(range(13, 25)) + [25]])))
    # 4 ctl_keccak_sponge
    looked = [single(c) for c in [KS["context"][0], KS["segment"][0], KS["virt"][0], KS["len"][0], KS["timestamp"][0]]]
    for i in range(7, -1, -1):
        looked.append(lincomb((KS["digest"][4 * i + j], 1 << (24 - 8 * j)) for j in range(4)))
    ctls.append(([entry("Cpu", single(Cpu.IS_KECCAK_SPONGE), cpu_sponge(8))], entry("KeccakSponge", col_sum(KS["final_len"]), looked)))
    # 5, 6 ctl_keccak_inputs / outputs (Keccak: step flags 0..23, TIMESTAMP 24, A from 25: reg_a(x, y) = 25 + (5x + y) 2)
    START_A = 25
    START_APP = START_A + 50 + 320 + 320 + 1600
    APPP00 = START_APP + 50 + 64
    reg_a = lambda x, y: START_A + (x * 5 + y) * 2
    reg_appp = lambda x, y: APPP00 if (x, y) == (0, 0) else START_APP + x * 10 + y * 2
    limb = lambda reg, i: reg((i // 2) % 5, (i // 2) // 5) + (i % 2)
    ctls.append(([entry("KeccakSponge", sponge_any_block(KS), [single(c) for c in KS["new_rate"] + KS["orig_cap"] + KS["timestamp"]])],
                 entry("Keccak", single(0), [single(limb(reg_a, i)) for i in range(50)] + [single(24)])))
    digest_u32s = [lincomb((KS["digest"][4 * k + i], 1 << (8 * i)) for i in range(4)) for k in range(8)]
    ctls.append(([entry("KeccakSponge", sponge_any_block(KS), digest_u32s + [single(c) for c in KS["partial"] + KS["timestamp"]])],
                 entry("Keccak", single(23), [single(limb(reg_appp, i)) for i in range(50)] + [single(24)])))
    # 7 ctl_sha_extend_sponge
    ses_filter = col_sum(range(48))
    ctls.append(([entry("Cpu", single(Cpu.IS_SHA_EXTEND_SPONGE), [single(Cpu.ch(i, "value")) for i in range(3)] + [cpu_timestamp(), single(Cpu.GENERAL)])],
                 entry("ShaExtendSponge", ses_filter, [single(SES[k]) for k in ("context", "segment", "output_virt", "timestamp")] + [le_bytes(r4(SES["w_i"]))])))
    # 8, 9 ctl_sha_extend_inputs / outputs
    ctls.append(([entry("ShaExtendSponge", ses_filter, [single(c) for k in ("w15", "w2", "w16", "w7") for c in r4(SES[k])] + [single(SES["timestamp"])])],
                 entry("ShaExtend", single(SE["is_real"]), [single(c) for k in ("w15", "w2", "w16", "w7") for c in r4(SE[k])] + [single(SE["timestamp"])])))
    ctls.append(([entry("ShaExtendSponge", ses_filter, [single(c) for c in r4(SES["w_i"])] + [single(SES["timestamp"])])],
                 entry("ShaExtend", single(SE["is_real"]), [single(c) for c in r4(SE["w_i"])] + [single(SE["timestamp"])])))
    # 10 ctl_sha_compress_sponge
    ctls.append(([entry("Cpu", single(Cpu.IS_SHA_COMPRESS_SPONGE),
                        [single(Cpu.ch(i, "value")) for i in range(3)] + [cpu_timestamp()] + [single(Cpu.GENERAL + k) for k in range(8)])],
                 entry("ShaCompressSponge", single(SCS["is_real"]),
                       [single(SCS[k]) for k in ("context", "segment", "hx_virt", "timestamp")] + [le_bytes(r4(SCS["output_hx"] + 6 * i)) for i in range(8)])))
    # 11, 12 ctl_sha_compress_inputs / outputs
    ctls.append(([entry("ShaCompressSponge", single(SCS["is_real"]),
                        [single(SCS["hx"] + i) for i in range(32)] + [single(SCS[k]) for k in ("timestamp", "w_start_segment", "w_start_context", "w_start_virt")])],
                 entry("ShaCompress", single(SC["round"]),
                       [single(SC["state"] + i) for i in range(32)] + [single(SC[k]) for k in ("timestamp", "segment", "context", "w_i_virt")])))
    ctls.append(([entry("ShaCompressSponge", single(SCS["is_real"]), [single(SCS["output_state"] + i) for i in range(32)] + [single(SCS["timestamp"])])],
                 entry("ShaCompress", single(SC["round"] + 64), [single(SC["state"] + i) for i in range(32)] + [single(SC["timestamp"])])))
    # 13 ctl_logic
    lookers = [entry("Cpu", single(Cpu.LOGIC_OP), [le_bits(opfunc)] + cpu_binops())]
    for i in range(34):                                   # num_logic_ctls = ceil(136 / 4)
        lookers.append(entry("KeccakSponge", sponge_any_block(KS),
                             [const(OP_XOR), single(KS["orig_rate"][i]), le_bytes(KS["block_bytes"][4 * i:4 * i + 4]), single(KS["new_rate"][i])]))
    se_f = single(SE["is_real"])
    lookers += [entry("ShaExtend", se_f, logic_row(OP_XOR, SE["rr7"], SE["rr18"], SE["s0_inter"])),
                entry("ShaExtend", se_f, logic_row(OP_XOR, SE["s0_inter"], SE["rs3"], SE["s0"])),
                entry("ShaExtend", se_f, logic_row(OP_XOR, SE["rr17"], SE["rr19"], SE["s1_inter"])),
                entry("ShaExtend", se_f, logic_row(OP_XOR, SE["s1_inter"], SE["rs10"], SE["s1"]))]
    sc_f = col_sum(range(SC["round"], SC["round"] + 64))
    st = lambda i: SC["state"] + 4 * i
    for op, a, b, out in ((OP_XOR, SC["e_rr_6"], SC["e_rr_11"], SC["s1_inter"]), (OP_XOR, SC["s1_inter"], SC["e_rr_25"], SC["s1"]),
                          (OP_AND, st(4), st(5), SC["e_and_f"]), (OP_AND, SC["e_not"], st(6), SC["e_not_and_g"]),
                          (OP_XOR, SC["e_and_f"], SC["e_not_and_g"], SC["ch"]), (OP_XOR, SC["a_rr_2"], SC["a_rr_13"], SC["s0_inter"]),
                          (OP_XOR, SC["s0_inter"], SC["a_rr_22"], SC["s0"]), (OP_AND, st(0), st(1), SC["a_and_b"]), (OP_AND, st(0), st(2), SC["a_and_c"]),
                          (OP_AND, st(1), st(2), SC["b_and_c"]), (OP_XOR, SC["a_and_b"], SC["a_and_c"], SC["maj_inter"]),
                          (OP_XOR, SC["maj_inter"], SC["b_and_c"], SC["maj"])):
        lookers.append(entry("ShaCompress", sc_f, logic_row(op, a, b, out)))
    logic_looked = [lincomb([(0, 0b100100 << 6), (1, 0b100101 << 6), (2, 0b100110 << 6), (3, 0b100111 << 6)]), le_bits(range(4, 36)), le_bits(range(36, 68)),
                    single(68)]
    ctls.append((lookers, entry("Logic", col_sum([0, 1, 2, 3]), logic_looked)))
    # 14 ctl_memory
    lookers = []
    for c in range(9):
        lookers.append(entry("Cpu", single(Cpu.ch(c, "used")),
                             [single(Cpu.ch(c, f)) for f in ("is_read", "addr_context", "addr_segment", "addr_virtual", "value")]
                             + [lincomb([(Cpu.CLOCK, NUM_CHANNELS)], 0)]))
    lookers += [entry("KeccakSponge", sponge_memory_filter(KS, i, 136), sponge_memory(KS, i)) for i in range(136)]
    lookers += [entry("PoseidonSponge", sponge_memory_filter(PS, i, 32), sponge_memory(PS, i)) for i in range(32)]
    for i in range(16):
        word = ("w15", "w2", "w16", "w7")[i // 4]
        lookers.append(entry("ShaExtendSponge", ses_filter, [const(1), single(SES["context"]), single(SES["segment"]), single(SES["input_virt"] + i // 4),
                                                             le_bytes(r4(SES[word])), single(SES["timestamp"])]))
    for i in range(32):
        lookers.append(entry("ShaCompressSponge", single(SCS["is_real"]),
                             [const(1), single(SCS["context"]), single(SCS["segment"]), single(SCS["hx_virt"] + i // 4), le_bytes(r4(SCS["hx"] + 4 * (i // 4))),
                              single(SCS["timestamp"])]))
    for _ in range(4):
        lookers.append(entry("ShaCompress", sc_f, [const(1), single(SC["context"]), single(SC["segment"]), single(SC["w_i_virt"]), le_bytes(r4(SC["w_i"])),
                                                   single(SC["timestamp"])]))
    # Memory: FILTER 0, TIMESTAMP 1, IS_READ 2, ADDR_CONTEXT 3, ADDR_SEGMENT 4, ADDR_VIRTUAL 5, value 6
    ctls.append((lookers, entry("Memory", single(0), [single(c) for c in (2, 3, 4, 5, 6, 1)])))
    return ctls


def lookups():
    """[(table, looked-up columns, table column, frequencies column)]: arithmetic_stark.rs:269-276, memory_stark.rs:476-483."""
    return [("Arithmetic", [single(c) for c in range(26, 44)], single(44), single(45)), ("Memory", [single(10)], single(11), single(12))]
