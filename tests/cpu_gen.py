"""A small MIPS interpreter that produces the CPU table rows and the arithmetic / logic / memory operations one step of
the reference's witness generation produces (test infrastructure).  Restated from: witness/transition.rs:21-353
(read_code_memory, decode, fill_op_flag, base_row, program-counter update :468-479), witness/operation.rs (generate_* of
every instruction used here, cited at each method), witness/util.rs:48-349 (register/memory channel helpers),
generation/mod.rs:169-186 (exit padding rows), cpu/bootstrap_kernel.rs (memory image written through the GP channels by
bootstrap rows; simplified to plain image writes, the constraints :308-351 ask no more), memory/memory_stark.rs:44-244
(sorting, fill_gaps, padding, first-change flags, range check, counter, frequencies); syscalls (operation.rs:1460-1684:
brk, mmap, clone, exit_group, read, write, fcntl, set_thread_area, unknown numbers) and the Keccak / SHA-extend / SHA-compress
precompile rows (operation.rs:1101-1458, witness/util.rs:370-694), image-id rows of the bootstrap (bootstrap_kernel.rs:70-163).
Not interpreted: the hint / verify / commit / preimage syscalls (host-side input streams) and the page-hash rows."""
import numpy as np

import arith_gen as ag

P = 0xFFFFFFFF00000001
M32 = 0xFFFFFFFF
NUM_GP_CHANNELS, NUM_CHANNELS = 9, 10
SEG_CODE, SEG_SHIFT_TABLE, SEG_REGISTER_FILE = 0, 3, 4

# cpu/columns/mod.rs:68-118, ops.rs:9-46
IS_BOOTSTRAP_KERNEL, IS_EXIT_KERNEL, CONTEXT, CODE_CONTEXT, PROGRAM_COUNTER, NEXT_PROGRAM_COUNTER, IS_KERNEL_MODE = range(7)
IS_POSEIDON_SPONGE, IS_KECCAK_SPONGE, IS_SHA_EXTEND_SPONGE, IS_SHA_COMPRESS_SPONGE = 82, 83, 84, 85
OPS = ["binary_op", "binary_imm_op", "eq_iszero", "logic_op", "logic_imm_op", "movz_op", "movn_op", "clz_op", "clo_op", "shift",
       "shift_imm", "keccak_general", "jumps", "jumpi", "jumpdirect", "branch", "pc", "get_context", "set_context", "exit_kernel",
       "m_op_load", "m_op_store", "nop", "ext", "ins", "maddu", "rdhwr", "signext8", "signext16", "swaphalf", "teq", "ror", "syscall"]
OP = {name: 7 + i for i, name in enumerate(OPS)}
BR = {name: 40 + i for i, name in enumerate(["should_jump", "gt", "lt", "eq", "is_gt", "is_lt", "is_eq", "is_ge", "is_le", "is_ne"])}
OPCODE_BITS, RS_BITS, RT_BITS, RD_BITS, SHAMT_BITS, FUNC_BITS = 50, 56, 61, 66, 71, 76
GENERAL = 86
MEMIO = 188
MEMIO_F = {name: MEMIO + i for i, name in enumerate(["lh", "lwl", "lw", "lbu", "lhu", "lwr", "sb", "sh", "swl", "sw", "swr", "ll", "sc",
                                                     "sdc1", "lb", "aux_filter"])}
CLOCK, MEM_CHANNELS, NUM_COLUMNS = 204, 205, 259
G_IO_RS_LE, G_IO_RT_LE, G_IO_MEM_LE, G_IO_AUX = GENERAL, GENERAL + 32, GENERAL + 64, GENERAL + 96
G_MISC_RS_BITS, G_MISC_IS_MSB, G_MISC_IS_LSB, G_MISC_AUXM, G_MISC_AUXL, G_MISC_AUXS = GENERAL, GENERAL + 32, GENERAL + 64, GENERAL + 96, GENERAL + 97, GENERAL + 98
G_MISC_RD_INDEX, G_MISC_RD_EQ_0, G_MISC_RD_EQ_29 = GENERAL + 99, GENERAL + 100, GENERAL + 101
G_LOGIC_DIFF_PINV = GENERAL
G_SYSCALL_COND, G_SYSCALL_SYSNUM, G_SYSCALL_A0, G_SYSCALL_A1 = GENERAL, GENERAL + 12, GENERAL + 24, GENERAL + 27


def ch(c, f):
    return MEM_CHANNELS + 6 * c + f


def sext16(v):
    return ag.sign_extend16(v)


def s32(v):
    return ag.s32(v)


def inv(x):
    x %= P
    return pow(x, P - 2, P) if x else 0


# ------------------------------------------------------------------------------------------------ assembler
def rtype(func, rs=0, rt=0, rd=0, sa=0, opcode=0):
    return (opcode << 26) | (rs << 21) | (rt << 16) | (rd << 11) | (sa << 6) | func


def itype(opcode, rs, rt, imm):
    return (opcode << 26) | (rs << 21) | (rt << 16) | (imm & 0xFFFF)


def jtype(opcode, target):
    return (opcode << 26) | (target & 0x3FFFFFF)


class MiniCpu:
    def __init__(self, image, entry, image_id_words=None):
        """image: {word address: u32} (code and data, all in segment Code as the reference keeps them);
        image_id_words = (base address, [9 words]) adds the image-id rows of check_image_id to the bootstrap."""
        self.image = dict(image)
        self.image_id_words = image_id_words
        self.mem = {}
        self.regs = [0] * 39
        self.pc, self.next_pc = entry, entry + 4
        self.rows, self.mem_ops, self.arith_ops, self.logic_ops = [], [], [], []
        self.keccak_ops, self.poseidon_ops, self.sha_extend_ops, self.sha_compress_ops = [], [], [], []
        self._post = None
        self.bootstrap()

    # ---------------------------------------------------------------- helpers (witness/util.rs)
    def clock(self):
        return len(self.rows)

    def mem_op(self, seg, virt, is_read, value, filt=True):
        self.mem_ops.append((0, seg, virt, self.clock() * NUM_CHANNELS, int(is_read), value & M32, int(filt)))

    def set_channel(self, row, c, used, is_read, seg, virt, value):
        assert row[ch(c, 0)] == 0
        row[ch(c, 0)], row[ch(c, 1)], row[ch(c, 2)], row[ch(c, 3)], row[ch(c, 4)], row[ch(c, 5)] = used, is_read, 0, seg, virt, value & M32

    def reg_read(self, row, index, c):            # util.rs:108-160
        v = self.regs[index]
        self.set_channel(row, c, 1, 1, SEG_REGISTER_FILE, index, v)
        self.mem_op(SEG_REGISTER_FILE, index, True, v)
        return v

    def reg_write(self, row, index, c, value):    # util.rs:162-224 (register 0: channel unused, operation filtered off)
        value &= M32
        if index != 0:
            self.regs[index] = value
        self.set_channel(row, c, int(index != 0), 0, SEG_REGISTER_FILE, index, value)
        self.mem_op(SEG_REGISTER_FILE, index, False, value, filt=index != 0)

    def push_no_write(self, row, c, value):       # util.rs:280-302
        self.set_channel(row, c, 0, 0, 0, 0, value)

    def mem_read(self, row, c, seg, virt):        # util.rs:304-325
        v = (1 << virt) & M32 if seg == SEG_SHIFT_TABLE else self.mem.get(virt, 0)
        self.set_channel(row, c, 1, 1, seg, virt, v)
        self.mem_op(seg, virt, True, v)
        return v

    def mem_write(self, row, c, virt, value):     # util.rs:327-349
        self.mem[virt] = value & M32
        self.set_channel(row, c, 1, 0, SEG_CODE, virt, value)
        self.mem_op(SEG_CODE, virt, False, value)

    # ---------------------------------------------------------------- bootstrap and padding
    def bootstrap(self):
        words = sorted(self.image.items())
        for k in range(0, len(words), 8):
            row = [0] * NUM_COLUMNS
            row[IS_BOOTSTRAP_KERNEL], row[CLOCK] = 1, self.clock()
            for c, (addr, w) in enumerate(words[k:k + 8]):
                self.mem_write(row, c, addr, w)
            self.rows.append(row)
        if self.image_id_words is not None:       # bootstrap_kernel.rs:70-163 check_image_id: 9 words written, then hashed
            base, vals = self.image_id_words
            addrs = [base + 4 * k for k in range(len(vals))]
            for k in range(0, len(vals), 8):
                row = [0] * NUM_COLUMNS
                row[IS_BOOTSTRAP_KERNEL], row[CLOCK] = 1, self.clock()
                for c, (addr, v) in enumerate(zip(addrs[k:k + 8], vals[k:k + 8])):
                    self.mem_write(row, c, addr, int.from_bytes(v.to_bytes(4, "little"), "big"))      # (*val).to_be()
                self.rows.append(row)
            data = b"".join(v.to_bytes(4, "little") for v in vals)
            row = [0] * NUM_COLUMNS
            row[IS_BOOTSTRAP_KERNEL], row[CLOCK], row[IS_POSEIDON_SPONGE] = 1, self.clock(), 1
            final_index = len(addrs) // 8 * 8
            row[ch(1, 5)], row[ch(2, 5)], row[ch(3, 5)] = SEG_CODE, addrs[final_index], len(data)
            self.poseidon_ops.append((addrs, self.clock() * NUM_CHANNELS, data, 0, SEG_CODE, row))    # digest filled in later
            self.sponge_reads(addrs, data, 32)
            self.rows.append(row)
            return
        row = [0] * NUM_COLUMNS                    # last bootstrap row: every channel disabled (bootstrap_kernel.rs:334-339)
        row[IS_BOOTSTRAP_KERNEL], row[CLOCK] = 1, self.clock()
        self.rows.append(row)

    def sponge_reads(self, addrs, data, rate):     # util.rs:370-441,471-531: one read per absorbed byte, of its whole word
        for i in range(len(data)):
            blk = i // rate * rate
            chunk = bytearray(data[blk:blk + rate])
            if len(chunk) < rate:                  # the final block is read through its padded form
                chunk += bytes(rate - len(chunk))
                chunk[len(data) - blk] = 1
                chunk[rate - 1] |= 0x80
            j = (i - blk) // 4 * 4
            self.mem_op(SEG_CODE, addrs[i // 4], True, int.from_bytes(chunk[j:j + 4], "big"))

    def pad(self, log_n):                          # generation/mod.rs:169-186
        n = 1 << log_n
        assert len(self.rows) < n
        while len(self.rows) < n:
            row = [0] * NUM_COLUMNS
            row[CLOCK], row[PROGRAM_COUNTER], row[NEXT_PROGRAM_COUNTER], row[IS_EXIT_KERNEL] = self.clock(), self.pc, self.next_pc, 1
            self.rows.append(row)

    # ---------------------------------------------------------------- one instruction
    def step(self):
        row = [0] * NUM_COLUMNS                    # transition.rs:508-521 base_row, util.rs:48-93 code read on the last channel
        row[CLOCK], row[PROGRAM_COUNTER], row[NEXT_PROGRAM_COUNTER] = self.clock(), self.pc, self.next_pc
        insn = self.mem.get(self.pc, 0)
        opcode, rs, rt, rd, sa, func = insn >> 26, (insn >> 21) & 31, (insn >> 16) & 31, (insn >> 11) & 31, (insn >> 6) & 31, insn & 63
        for at, v, nb in ((OPCODE_BITS, opcode, 6), (RS_BITS, rs, 5), (RT_BITS, rt, 5), (RD_BITS, rd, 5), (SHAMT_BITS, sa, 5), (FUNC_BITS, func, 6)):
            for i in range(nb):
                row[at + i] = (v >> i) & 1
        self.set_channel(row, NUM_GP_CHANNELS - 1, 1, 1, SEG_CODE, self.pc, insn)
        self.mem_op(SEG_CODE, self.pc, True, insn)
        offset, target = insn & 0xFFFF, insn & 0x3FFFFFF
        jumped = self.execute(row, opcode, func, rs, rt, rd, sa, offset, target)
        self.rows.append(row)
        if self._post is not None:                 # precompile rows follow the syscall row (operation.rs:1657-1682)
            post, self._post = self._post, None
            post()
        if not jumped:                             # transition.rs:468-479
            self.pc, self.next_pc = self.next_pc, self.next_pc + 4

    def jump_to(self, dst):                        # generation/state.rs:296-299
        self.pc, self.next_pc = self.next_pc, dst & M32

    def run(self, steps):
        for _ in range(steps):
            self.step()

    def arith(self, op, a, b):
        self.arith_ops.append((op, a & M32, b & M32))
        return ag.result(op, a & M32, b & M32)

    def execute(self, row, opcode, func, rs, rt, rd, sa, offset, target):
        R = {0b100000: ag.IS_ADD, 0b100001: ag.IS_ADDU, 0b100010: ag.IS_SUB, 0b100011: ag.IS_SUBU, 0b101010: ag.IS_SLT, 0b101011: ag.IS_SLTU}
        HILO = {0b011000: ag.IS_MULT, 0b011001: ag.IS_MULTU, 0b011010: ag.IS_DIV, 0b011011: ag.IS_DIVU}
        LOGIC = {0b100100: 0, 0b100101: 1, 0b100110: 2, 0b100111: 3}
        if opcode == 0 and func in R or (opcode == 0b011100 and func == 0b000010):
            self.binary_arith(row, ag.IS_MUL if opcode else R[func], rs, rt, rd)
        elif opcode == 0 and func in HILO:
            self.hilo(row, HILO[func], rs, rt)
        elif opcode == 0 and func in (0b010000, 0b010001, 0b010010, 0b010011):     # MFHI MTHI MFLO MTLO (transition.rs:176-199)
            op, a, d = {0b010000: (ag.IS_MFHI, 33, rd), 0b010001: (ag.IS_MTHI, rs, 33), 0b010010: (ag.IS_MFLO, 32, rd), 0b010011: (ag.IS_MTLO, rs, 32)}[func]
            self.binary_arith(row, op, a, 0, d)
        elif opcode == 0 and func in LOGIC:
            self.binary_logic(row, LOGIC[func], rs, rt, rd)
        elif opcode in (0b001100, 0b001101, 0b001110):
            self.logic_imm(row, {0b001100: 0, 0b001101: 1, 0b001110: 2}[opcode], rs, rt, offset)
        elif opcode in (0b001000, 0b001001, 0b001010, 0b001011):
            self.arith_imm(row, {0b001000: ag.IS_ADDI, 0b001001: ag.IS_ADDIU, 0b001010: ag.IS_SLTI, 0b001011: ag.IS_SLTIU}[opcode], rs, rt, offset)
        elif opcode == 0b001111:
            self.lui(row, rs, rt, offset)
        elif opcode == 0 and func == 0b000010 and rs == 1:
            self.ror(row, rd, rt, sa)
        elif opcode == 0 and func in (0b000000, 0b000010, 0b000011):
            self.shift_imm(row, {0: ag.IS_SLL, 2: ag.IS_SRL, 3: ag.IS_SRA}[func], sa, rt, rd)
        elif opcode == 0 and func in (0b000100, 0b000110, 0b000111):
            self.shift_var(row, {4: ag.IS_SLLV, 6: ag.IS_SRLV, 7: ag.IS_SRAV}[func], rs, rt, rd)
        elif opcode == 0 and func in (0b001010, 0b001011):
            self.cond_mov(row, func == 0b001010, rs, rt, rd)
        elif opcode == 0b011100 and func in (0b100000, 0b100001):
            self.count(row, func == 0b100001, rs, rd)
        elif opcode == 0 and func == 0b001111 or opcode == 0b110011:
            row[OP["nop"]] = 1
        elif opcode == 0 and func in (0x08, 0x09):
            self.jump(row, 0 if func == 0x08 else rd, rs)
            return True
        elif opcode == 0x01:
            if rt == 1:
                self.branch(row, "ge", rs, 0, offset)
            elif rt == 0:
                self.branch(row, "lt", rs, 0, offset)
            elif rt == 0x11 and rs == 0:
                self.jumpdirect(row, 31, offset)
            else:
                raise ValueError("invalid opcode")
            return True
        elif opcode in (0x02, 0x03):
            self.jumpi(row, 0 if opcode == 0x02 else 31, target)
            return True
        elif opcode in (0x04, 0x05, 0x06, 0x07):
            self.branch(row, {4: "eq", 5: "ne", 6: "le", 7: "gt"}[opcode], rs, rt if opcode < 6 else 0, offset)
            return True
        elif opcode in (0b100000, 0b100001, 0b100010, 0b100011, 0b100100, 0b100101, 0b100110, 0b110000):
            self.mload({0b100000: "lb", 0b100001: "lh", 0b100010: "lwl", 0b100011: "lw", 0b100100: "lbu", 0b100101: "lhu", 0b100110: "lwr",
                        0b110000: "ll"}[opcode], row, rs, rt, offset)
        elif opcode in (0b101000, 0b101001, 0b101010, 0b101011, 0b101110, 0b111000, 0b111101):
            self.mstore({0b101000: "sb", 0b101001: "sh", 0b101010: "swl", 0b101011: "sw", 0b101110: "swr", 0b111000: "sc", 0b111101: "sdc1"}[opcode],
                        row, rs, rt, offset)
        elif opcode == 0b011100 and func == 0b000001:
            self.maddu(row, rt, rs)
        elif opcode == 0b011111 and func == 0b000000:
            self.ext(row, rt, rs, rd, sa)
        elif opcode == 0b011111 and func == 0b000100:
            self.ins(row, rt, rs, rd, sa)
        elif opcode == 0b011111 and func == 0b111011:
            self.rdhwr(row, rt, rd)
        elif opcode == 0b011111 and func == 0b100000 and sa in (0b011000, 0b010000):
            self.signext(row, rd, rt, 16 if sa == 0b011000 else 8)
        elif opcode == 0b011111 and func == 0b100000 and sa == 0b000010:
            self.swaphalf(row, rd, rt)
        elif opcode == 0 and func == 0b110100:
            self.teq(row, rs, rt)
        elif opcode == 0 and func == 0b001100:
            self.syscall(row)
        else:
            raise ValueError(f"instruction {opcode:06b}/{func:06b} is not interpreted here")
        return False

    # ---------------------------------------------------------------- operation.rs generators
    def binary_arith(self, row, op, rs, rt, rd):   # operation.rs:286-318
        row[OP["binary_op"]] = 1
        a, b = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1)
        self.reg_write(row, rd, 2, self.arith(op, a, b)[0])

    def hilo(self, row, op, rs, rt):               # operation.rs:320-375
        row[OP["binary_op"]] = 1
        a, b = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1)
        lo, hi = self.arith(op, a, b)
        self.reg_write(row, 32, 2, lo)
        self.reg_write(row, 33, 3, hi)

    def arith_imm(self, row, op, rs, rt, imm):     # operation.rs:377-403
        row[OP["binary_imm_op"]] = 1
        a, b = self.reg_read(row, rs, 0), sext16(imm)
        self.reg_write(row, rt, 1, b)
        self.reg_write(row, rt, 2, self.arith(op, a, b)[0])

    def lui(self, row, rs, rt, imm):               # operation.rs:405-433
        row[OP["binary_imm_op"]] = 1
        a, b = sext16(imm), 1 << 16
        self.reg_write(row, rs, 0, a)
        self.push_no_write(row, 1, b)
        self.reg_write(row, rt, 1, b)
        self.reg_write(row, rt, 2, self.arith(ag.IS_LUI, a, b)[0])

    def binary_logic(self, row, kind, rs, rt, rd):  # operation.rs:233-258
        row[OP["logic_op"]] = 1
        a, b = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1)
        self.logic_ops.append((kind, a, b))
        self.reg_write(row, rd, 2, [a & b, a | b, a ^ b, ~(a | b) & M32][kind])

    def logic_imm(self, row, kind, rs, rd, imm):    # operation.rs:260-284 (no logic-table lookup: push_logic is commented out)
        row[OP["logic_imm_op"]] = 1
        a = self.reg_read(row, rs, 0)
        self.reg_write(row, rd, 2, [a & imm, a | imm, a ^ imm][kind])

    def shift_imm(self, row, op, sa, rt, rd):       # operation.rs:731-769
        row[OP["shift_imm"]] = 1
        a = self.reg_read(row, rt, 1)
        self.push_no_write(row, 0, sa)
        self.mem_read(row, 3, SEG_SHIFT_TABLE, sa)
        self.reg_write(row, rd, 2, self.arith(op, a, sa)[0])

    def shift_var(self, row, op, rs, rt, rd):       # operation.rs:771-872
        row[OP["shift"]] = 1
        s, a = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1)
        self.mem_read(row, 3, SEG_SHIFT_TABLE, s)
        self.reg_write(row, rd, 2, self.arith(op, a, s)[0])

    def pinv_diff(self, row, v0, v1):               # operation.rs:55-72
        row[G_LOGIC_DIFF_PINV] = inv(v0 - v1)

    def cond_mov(self, row, is_eq, rs, rt, rd):     # operation.rs:149-184
        row[OP["movz_op" if is_eq else "movn_op"]] = 1
        a, b, c = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1), self.reg_read(row, rd, 2)
        mov = (b == 0) if is_eq else (b != 0)
        self.pinv_diff(row, b, 0)
        self.reg_write(row, rd, 3, a if mov else c)
        self.reg_write(row, 0, 4, int(mov))

    def count(self, row, is_clo, rs, rd):           # operation.rs:186-231
        row[OP["clo_op" if is_clo else "clz_op"]] = 1
        a = self.reg_read(row, rs, 0)
        a = (~a) & M32 if is_clo else a
        self.reg_write(row, rd, 1, 32 - a.bit_length())
        for i in range(32):
            row[G_IO_RS_LE + i] = (a >> i) & 1
        k = 0
        for i in range(30, -1, -1):
            x = a >> i
            row[G_IO_RT_LE + k], row[G_IO_MEM_LE + k] = int(x == 1), inv(x - 1)
            k += 1
        row[G_IO_RT_LE + 31], row[G_IO_MEM_LE + 31] = int(a == 0), inv(a)

    def jump(self, row, link, target_reg):          # operation.rs:481-499
        row[OP["jumps"]] = 1
        dst = self.reg_read(row, target_reg, 0)
        self.reg_write(row, link, 1, self.pc + 8)
        self.jump_to(dst)

    def branch(self, row, cond, r1, r2, target):    # operation.rs:501-568
        row[OP["branch"]] = 1
        a, b = self.reg_read(row, r1, 0), self.reg_read(row, r2, 1)
        sa_, sb_ = s32(a), s32(b)
        should = {"eq": sa_ == sb_, "ne": sa_ != sb_, "ge": sa_ >= sb_, "le": sa_ <= sb_, "gt": sa_ > sb_, "lt": sa_ < sb_}[cond]
        row[BR["is_" + cond]] = 1
        row[BR["eq"]], row[BR["gt"]], row[BR["lt"]] = int(a == b), int(a > b), int(a < b)
        tgt = (sext16(target) << 2) & M32
        self.reg_write(row, 0, 2, a - b)
        self.reg_write(row, 0, 3, b - a)
        self.reg_write(row, 0, 4, int(((a ^ b) & 0x80000000) > 0))
        self.reg_write(row, 0, 5, tgt)
        row[BR["should_jump"]] = int(should)
        self.jump_to((tgt + self.pc + 4) & M32 if should else (self.pc + 8) & M32)

    def jumpi(self, row, link, target):             # operation.rs:570-596
        row[OP["jumpi"]] = 1
        pc_hi = self.pc & 0xF0000000
        self.reg_write(row, 0, 2, pc_hi)
        self.reg_write(row, link, 1, self.pc + 8)
        self.jump_to(((target << 2) + pc_hi) & M32)

    def jumpdirect(self, row, link, target):        # operation.rs:598-622
        row[OP["jumpdirect"]] = 1
        tgt = (sext16(target) << 2) & M32
        self.reg_write(row, 0, 2, tgt)
        self.reg_write(row, link, 1, self.pc + 8)
        self.jump_to((tgt + self.pc + 4) & M32)

    def ror(self, row, rd, rt, sa):                 # operation.rs:874-906
        row[OP["ror"]] = 1
        a = self.reg_read(row, rt, 0)
        for i in range(32):
            row[G_MISC_RS_BITS + i] = (a >> i) & 1
        row[G_MISC_IS_LSB + sa] = 1
        self.reg_write(row, rd, 1, ((a | (a << 32)) >> sa) & M32)

    def _io_bits(self, row, rs, rt, mem):
        for i in range(32):
            row[G_IO_RS_LE + i], row[G_IO_RT_LE + i], row[G_IO_MEM_LE + i] = (rs >> i) & 1, (rt >> i) & 1, (mem >> i) & 1

    def mload(self, op, row, base, rt_reg, offset):  # operation.rs:1686-1802
        row[OP["m_op_load"]] = 1
        rs, rt = self.reg_read(row, base, 0), self.reg_read(row, rt_reg, 1)
        raw = (rs + sext16(offset)) & M32
        mem = self.mem_read(row, 2, SEG_CODE, raw & 0xFFFFFFFC)
        self._io_bits(row, raw, rt, mem)
        row[MEMIO_F["aux_filter"]] = row[OPCODE_BITS + 5]
        row[MEMIO_F[op]] = 1
        i, aux = raw & 3, ((raw >> 1) & 1) * (raw & 1)
        if op == "lh":
            aux, val = 0, sext16((mem >> (16 - (raw & 2) * 8)) & 0xFFFF)
        elif op == "lwl":
            val = (rt & ~((M32 << (i * 8)) & M32) & M32) | ((mem << (i * 8)) & M32)
        elif op in ("lw", "ll"):
            aux, val = 0, mem
        elif op == "lbu":
            val = (mem >> (24 - i * 8)) & 0xFF
        elif op == "lhu":
            aux, val = 0, (mem >> (16 - (raw & 2) * 8)) & 0xFFFF
        elif op == "lwr":
            val = (rt & ~(M32 >> (24 - i * 8)) & M32) | (mem >> (24 - i * 8))
        else:                                        # lb: sign_extend::<8>
            b = (mem >> (24 - i * 8)) & 0xFF
            val = b | 0xFFFFFF00 if b & 0x80 else b
        row[G_IO_AUX] = aux
        self.reg_write(row, rt_reg, 3, val)

    def mstore(self, op, row, base, rt_reg, offset):  # operation.rs:1804-1928
        row[OP["m_op_store"]] = 1
        rs, rt = self.reg_read(row, base, 0), self.reg_read(row, rt_reg, 1)
        raw = (rs + sext16(offset)) & M32
        virt = raw & 0xFFFFFFFC
        mem = self.mem_read(row, 2, SEG_CODE, virt)
        self._io_bits(row, raw, rt, mem)
        row[MEMIO_F["aux_filter"]] = row[OPCODE_BITS + 5]
        row[MEMIO_F[op]] = 1
        i, aux = raw & 3, ((raw >> 1) & 1) * (raw & 1)
        if op == "sb":
            val = (mem & (M32 ^ (0xFF << (24 - i * 8)))) | ((rt & 0xFF) << (24 - i * 8))
        elif op == "sh":
            j = raw & 2
            aux, val = 0, (mem & (M32 ^ (0xFFFF << (16 - j * 8)))) | ((rt & 0xFFFF) << (16 - j * 8))
        elif op == "swl":
            val = (mem & ~(M32 >> (i * 8)) & M32) | (rt >> (i * 8))
        elif op in ("sw", "sc"):
            aux, val = 0, rt
        elif op == "swr":
            val = (mem & ~((M32 << (24 - i * 8)) & M32) & M32) | ((rt << (24 - i * 8)) & M32)
        else:                                        # sdc1
            aux, val = 0, 0
        row[G_IO_AUX] = aux
        self.mem_write(row, 3, virt, val)
        if op == "sc":
            self.reg_write(row, rt_reg, 4, 1)

    def ext(self, row, rt, rs, msbd, lsb):           # operation.rs:1943-1984
        row[OP["ext"]] = 1
        assert msbd + lsb < 32
        a = self.reg_read(row, rs, 0)
        mask_msb = (1 << (msbd + lsb + 1)) - 1
        for i in range(32):
            row[G_MISC_RS_BITS + i] = (a >> i) & 1
        row[G_MISC_IS_MSB + msbd + lsb] = 1
        row[G_MISC_IS_LSB + lsb] = 1
        row[G_MISC_AUXS], row[G_MISC_AUXM], row[G_MISC_AUXL] = 1 << lsb, a & mask_msb, a & ((1 << lsb) - 1)
        self.reg_write(row, rt, 1, (a & mask_msb) >> lsb)

    def ins(self, row, rt, rs, msb, lsb):            # operation.rs:1986-2032
        row[OP["ins"]] = 1
        assert lsb <= msb < 32
        a, b = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1)
        mask = (1 << (msb - lsb + 1)) - 1
        field = (mask << lsb) & M32
        for i in range(32):
            row[G_MISC_RS_BITS + i] = (a >> i) & 1
        row[G_MISC_IS_MSB + msb - lsb] = 1
        row[G_MISC_IS_LSB + lsb] = 1
        row[G_MISC_AUXM], row[G_MISC_AUXL], row[G_MISC_AUXS] = b & ~field & M32, a & mask, 1 << lsb
        self.reg_write(row, rt, 2, (b & ~field & M32) | ((a << lsb) & field))

    def maddu(self, row, rt, rs):                    # operation.rs:2034-2062
        row[OP["maddu"]] = 1
        a, b, hi, lo = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1), self.reg_read(row, 33, 2), self.reg_read(row, 32, 3)
        total = a * b + (hi << 32) + lo
        overflow, res = total >> 64, total & 0xFFFFFFFFFFFFFFFF
        self.reg_write(row, 33, 4, res >> 32)
        self.reg_write(row, 32, 5, res & M32)
        row[G_MISC_AUXM] = overflow << 32

    def rdhwr(self, row, rt, rd):                    # operation.rs:2063-2092
        row[OP["rdhwr"]] = 1
        row[G_MISC_RD_INDEX] = rd
        if rd == 0:
            row[G_MISC_RD_EQ_0], res = 1, 1
        elif rd == 29:
            row[G_MISC_RD_EQ_29] = 1
            res = self.reg_read(row, 38, 1)
        else:
            res = 0
        self.reg_write(row, rt, 0, res)

    def signext(self, row, rd, rt, bits):            # operation.rs:2094-2132
        row[OP["signext8" if bits == 8 else "signext16"]] = 1
        a = self.reg_read(row, rt, 0)
        for i in range(32):
            row[G_IO_RT_LE + i] = (a >> i) & 1
        mask = (1 << bits) - 1
        self.reg_write(row, rd, 1, (a & mask) | ((M32 ^ mask) if (a >> (bits - 1)) & 1 else 0))

    def swaphalf(self, row, rd, rt):                 # operation.rs:2134-2166
        row[OP["swaphalf"]] = 1
        a = self.reg_read(row, rt, 0)
        for i in range(32):
            row[G_IO_RT_LE + i] = (a >> i) & 1
        self.reg_write(row, rd, 1, (((a >> 16) & 0xFF) << 24) | (((a >> 24) & 0xFF) << 16) | ((a & 0xFF) << 8) | ((a >> 8) & 0xFF))

    def teq(self, row, rs, rt):                      # operation.rs:2168-2189
        row[OP["teq"]] = 1
        a, b = self.reg_read(row, rs, 0), self.reg_read(row, rt, 1)
        assert a != b, "trap"
        self.pinv_diff(row, a, b)

    # ---------------------------------------------------------------- syscalls and precompiles
    def syscall(self, row):                          # operation.rs:1460-1684
        row[OP["syscall"]] = 1
        num, a0, a1, a2 = (self.reg_read(row, r, c) for c, r in enumerate((2, 4, 5, 6)))
        v0 = v1 = 0
        cond, sysnum, fa0 = (lambda i: G_SYSCALL_COND + i), (lambda i: G_SYSCALL_SYSNUM + i), (lambda i: G_SYSCALL_A0 + i)
        if num in (4090, 4210):                      # SYSMMAP, SYSMMAP2
            row[sysnum(1)] = 1
            sz, unaligned = a1, False
            if sz & 0xFFF:
                row[G_SYSCALL_A1] = 1
                sz += 0x1000 - (sz & 0xFFF)
                row[sysnum(9)] = sz
                unaligned = True
            else:
                row[sysnum(10)] = 1
            if a0 == 0:
                row[cond(0)], row[fa0(0)] = 1, 1
                row[cond(1 if unaligned else 2)] = 1
                heap = self.reg_read(row, 34, 6)
                v0 = heap
                self.reg_write(row, 34, 7, heap + sz)
            else:
                row[cond(3)], row[fa0(2)] = 1, 1
                v0 = a0
        elif num == 4045:                            # SYSBRK
            row[sysnum(2)] = 1
            brk = self.reg_read(row, 37, 6)
            v0 = a0 if a0 > brk else brk
            row[cond(10 if a0 > brk else 11)] = 1
        elif num == 4120:                            # SYSCLONE
            row[sysnum(3)], v0 = 1, 1
        elif num == 4246:                            # SYSEXITGROUP
            row[sysnum(4)] = 1
        elif num == 4003:                            # SYSREAD
            row[sysnum(5)] = 1
            if a0 == 0:
                row[fa0(0)], row[cond(5)] = 1, 1
            else:
                row[fa0(2)], row[cond(4)], v0, v1 = 1, 1, M32, 0x9
        elif num == 4004:                            # SYSWRITE (fd 3, public values, is not interpreted)
            row[sysnum(6)] = 1
            if a0 in (1, 2, 4):
                row[fa0(1)], row[cond(7)], v0 = 1, 1, a2
            else:
                row[fa0(2)], row[cond(6)], v0, v1 = 1, 1, M32, 0x9
        elif num == 4055:                            # SYSFCNTL
            row[sysnum(7)] = 1
            if a0 == 0:
                row[fa0(0)], row[cond(8)] = 1, 1
            elif a0 in (1, 2):
                row[fa0(1)], row[cond(9)], v0 = 1, 1, 1
            else:
                row[fa0(2)], v0, v1 = 1, M32, 0x9
        elif num == 4283:                            # SYSSETTHREADAREA
            row[sysnum(8)] = 1
            self.reg_write(row, 38, 6, a0)
        elif num == 0x010109:                        # SYSKECCAK
            self._post = lambda: self.keccak(a0, a1, a2)
        elif num == 0x00010106:                      # SYSSHACOMPRESS
            self._post = lambda: self.sha_compress(a0, a1)
        elif num == 0x00300105:                      # SYSSHAEXTEND
            self._post = lambda: self.sha_extend(a0, a1)
        else:
            row[sysnum(11)] = 1
        self.reg_write(row, 2, 4, v0)
        self.reg_write(row, 7, 5, v1)

    def _blank(self):
        row = [0] * NUM_COLUMNS
        row[CLOCK] = self.clock()
        return row

    def keccak(self, addr, length, ptr):             # operation.rs:1101-1181, util.rs:471-564
        import hash_gen as hg
        assert length % 4 == 0
        row, addrs, data = self._blank(), [], bytearray()
        for k in range(length // 4):
            if k and k % 8 == 0:
                self.rows.append(row)
                row = self._blank()
            word = self.mem_read(row, k % 8, SEG_CODE, addr + 4 * k)
            data += word.to_bytes(4, "big")
            addrs.append(addr + 4 * k)
        self.rows.append(row)
        row = self._blank()
        row[IS_KECCAK_SPONGE] = 1
        final_idx = length // hg.RATE_BYTES * hg.RATE_U32S
        row[ch(1, 5)], row[ch(2, 5)], row[ch(3, 5)] = SEG_CODE, (addrs[final_idx] if final_idx < len(addrs) else 0), length
        digest = hg.keccak256(bytes(data))
        words_be = [int.from_bytes(digest[4 * i:4 * i + 4], "big") for i in range(8)]
        for i in range(8):
            row[GENERAL + i] = words_be[7 - i]       # khash value, reversed
        self.keccak_ops.append((addrs if addrs else [0], self.clock() * NUM_CHANNELS, bytes(data), 0, SEG_CODE))
        self.sponge_reads(addrs, bytes(data), hg.RATE_BYTES)
        self.rows.append(row)
        row = self._blank()
        for i in range(8):                           # hash_data_be[i].to_be(): the digest bytes, big-endian word by word
            self.mem_write(row, i, ptr + 4 * i, words_be[i])
        self.rows.append(row)

    def sha_extend(self, w_ptr, a1):                 # operation.rs:1183-1284, util.rs:566-603
        import hash_gen as hg
        assert a1 == 0
        for i in range(16, 64):
            row = self._blank()
            virts = [w_ptr + 4 * (i - 15), w_ptr + 4 * (i - 2), w_ptr + 4 * (i - 16), w_ptr + 4 * (i - 7)]
            ins = [self.mem_read(row, c, SEG_CODE, v) for c, v in enumerate(virts)]
            _, w_i, xors = hg.sha_extend_row(*ins, 0)
            self.logic_ops += xors
            self.mem_write(row, 4, w_ptr + 4 * i, w_i)
            self.rows.append(row)
            row = self._blank()
            row[IS_SHA_EXTEND_SPONGE] = 1
            row[ch(1, 5)], row[ch(2, 5)], row[GENERAL] = SEG_CODE, w_ptr + 4 * i, w_i
            ts = self.clock() * NUM_CHANNELS
            for v, val in zip(virts, ins):
                for _ in range(4):
                    self.mem_op(SEG_CODE, v, True, val)
            self.sha_extend_ops.append((ins, virts, w_ptr + 4 * i, ts, i - 16, w_i))
            self.rows.append(row)

    def sha_compress(self, w_ptr, h_ptr):            # operation.rs:1298-1458, util.rs:605-694
        import hash_gen as hg
        row = self._blank()
        hx = [self.mem_read(row, i, SEG_CODE, h_ptr + 4 * i) for i in range(8)]
        self.rows.append(row)
        row, w = self._blank(), []
        for i in range(64):
            if i and i % 8 == 0:
                self.rows.append(row)
                row = self._blank()
            w.append(self.mem_read(row, i % 8, SEG_CODE, w_ptr + 4 * i))
        self.rows.append(row)
        st = list(hx)
        for i in range(64):
            _, st, ops = hg.sha_compress_row(st, w[i], hg.SHA_K[i], i, 0, 0)
            self.logic_ops += ops
        out = [(a + b) & M32 for a, b in zip(hx, st)]
        row = self._blank()
        row[IS_SHA_COMPRESS_SPONGE] = 1
        row[ch(1, 5)], row[ch(2, 5)] = SEG_CODE, h_ptr
        for i in range(8):
            row[GENERAL + i] = out[i]
        ts = self.clock() * NUM_CHANNELS
        for j in range(8):
            for _ in range(4):
                self.mem_op(SEG_CODE, h_ptr + 4 * j, True, hx[j])
        for i in range(64):
            for _ in range(4):
                self.mem_op(SEG_CODE, w_ptr + 4 * i, True, w[i])
        self.sha_compress_ops.append((hx, w, h_ptr, w_ptr, ts))
        self.rows.append(row)
        row = self._blank()
        for i in range(8):
            self.mem_write(row, i, h_ptr + 4 * i, out[i])
        self.rows.append(row)

    def cpu_trace(self, log_n):
        self.pad(log_n)
        return np.ascontiguousarray(np.array(self.rows, dtype=np.uint64).T)


# ------------------------------------------------------------------------------------------------ memory table
def memory_generate_trace(ops):
    """memory_stark.rs:133-244: ops = [(ctx, seg, virt, timestamp, is_read, value, filter)] in push order."""
    key = lambda o: (o[0], o[1], o[2], o[3])
    ops = sorted(ops, key=key)
    max_rc = (1 << (len(ops) - 1).bit_length()) - 1
    extra = []
    for cur, nxt in zip(ops, ops[1:]):                       # fill_gaps :186-217
        if cur[0] != nxt[0] or cur[1] != nxt[1]:
            continue
        if cur[2] != nxt[2]:
            while nxt[2] - cur[2] - 1 > max_rc:
                cur = (cur[0], cur[1], cur[2] + max_rc + 1, 0, 1, 0, 0)
                extra.append(cur)
        else:
            while nxt[3] - cur[3] > max_rc:
                cur = (cur[0], cur[1], cur[2], cur[3] + max_rc, 1, cur[5], 0)
                extra.append(cur)
    ops += extra
    last = ops[-1]                                           # pad_memory_ops :219-237 (before the second sort, as upstream)
    n = 1 << (len(ops) - 1).bit_length()
    ops += [(last[0], last[1], last[2], last[3], 1, last[5], 0)] * (n - len(ops))
    ops = sorted(ops, key=key)
    a = np.array(ops, dtype=np.int64).T                      # ctx, seg, virt, ts, is_read, value, filter
    r0 = (a[4] == 0) & (a[0] == 0) & (a[1] == SEG_REGISTER_FILE) & (a[2] == 0)
    a[5][r0] = 0                                             # into_row :62-72: writes to R0 are recorded as 0
    t = np.zeros((13, n), dtype=np.uint64)
    t[0], t[1], t[2], t[3], t[4], t[5], t[6] = a[6], a[3], a[4], a[0], a[1], a[2], a[5]
    cfc = a[0][:-1] != a[0][1:]
    sfc = (a[1][:-1] != a[1][1:]) & ~cfc
    vfc = (a[2][:-1] != a[2][1:]) & ~sfc & ~cfc
    rc = np.where(cfc, a[0][1:] - a[0][:-1] - 1, np.where(sfc, a[1][1:] - a[1][:-1] - 1, np.where(vfc, a[2][1:] - a[2][:-1] - 1,
                                                                                                 a[3][1:] - a[3][:-1])))
    assert rc.size == 0 or (rc.min() >= 0 and rc.max() < n), "Range check is too large. Bug in fill_gaps?"
    t[7, :-1], t[8, :-1], t[9, :-1], t[10, :-1] = cfc, sfc, vfc, rc
    t[11] = np.arange(n, dtype=np.uint64)
    t[12] = np.bincount(t[10].astype(np.int64), minlength=n).astype(np.uint64)
    return t
