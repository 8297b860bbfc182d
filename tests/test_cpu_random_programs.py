"""Randomised coverage of the CPU table's constraints: straight-line MIPS programs with random instructions, registers,
immediates, shift amounts, bit-field positions and memory alignments (every load/store form at all four byte offsets),
interleaved with taken and untaken branches.  Every program's trace must satisfy the transcribed CPU constraints and the
memory log it produces must satisfy the Memory table's."""
import numpy as np
import pytest

import cpu_gen as cg
import traces as tr
from cpu_gen import itype, rtype
from oracle import binding

ENTRY, DATA = 0x1000, 0x4000
SYNC = rtype(0b001111)


def random_program(rng, count):
    """Registers 1..15 hold data, 28 the data base; $0 appears as a source and as a (discarded) destination."""
    p = [itype(0b001101, 0, 28, DATA)]                              # ori $28, $0, DATA
    for r in range(1, 16):                                          # lui/ori: random 32-bit values
        v = int(rng.integers(0, 1 << 32)) if rng.random() < 0.8 else int(rng.choice([0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF]))
        p += [itype(0b001111, 0, r, v >> 16), itype(0b001101, r, r, v & 0xFFFF)]
    reg = lambda: int(rng.integers(0, 16))
    dst = lambda: int(rng.integers(0, 16))
    for _ in range(count):
        k = int(rng.integers(0, 14))
        rs, rt, rd = reg(), reg(), dst()
        imm = int(rng.integers(0, 1 << 16))
        if k == 0:
            p.append(rtype(int(rng.choice([0b100000, 0b100001, 0b100010, 0b100011, 0b101010, 0b101011])), rs, rt, rd))
        elif k == 1:
            p.append(rtype(int(rng.choice([0b100100, 0b100101, 0b100110, 0b100111])), rs, rt, rd))
        elif k == 2:
            p.append(itype(int(rng.choice([0b001000, 0b001001, 0b001010, 0b001011, 0b001100, 0b001101, 0b001110, 0b001111])), rs, rd, imm))
        elif k == 3:
            p.append(rtype(int(rng.choice([0, 2, 3])), 0, rt, rd, int(rng.integers(0, 32))))       # sll srl sra
        elif k == 4:                                                # sllv/srlv/srav need a shift register < 32
            p.append(itype(0b001100, rs, 14, 31))                   # andi $14, rs, 31
            p.append(rtype(int(rng.choice([4, 6, 7])), 14, rt, rd))
        elif k == 5:
            f = int(rng.choice([0b011000, 0b011001, 0b011011]))
            p.append(itype(0b001101, rt, 13, 1))                    # ori $13, rt, 1: a non-zero divisor
            p.append(rtype(f, rs, 13))                              # mult multu divu
            p.append(rtype(int(rng.choice([0b010000, 0b010010])), rd=rd))                           # mfhi / mflo
        elif k == 6:
            p.append(rtype(int(rng.choice([0b001010, 0b001011])), rs, rt, rd))                      # movz movn
        elif k == 7:
            p.append(rtype(int(rng.choice([0b100000, 0b100001])), rs, 0, rd, opcode=0b011100))      # clz clo
        elif k == 8:
            lsb = int(rng.integers(0, 32))
            if rng.random() < 0.5:
                p.append(rtype(0b000000, rs, rd, int(rng.integers(0, 32 - lsb)), lsb, opcode=0b011111))     # ext
            else:
                p.append(rtype(0b000100, rs, rd, int(rng.integers(lsb, 32)), lsb, opcode=0b011111))         # ins
        elif k == 9:
            sa = int(rng.choice([0b010000, 0b011000, 0b000010]))
            p.append(rtype(0b100000, 0, rt, rd, sa, opcode=0b011111))                               # seb seh wsbh
        elif k == 10:
            p.append(rtype(0b000010, 1, rt, rd, int(rng.integers(0, 32))))                          # rotr
        elif k == 11:                                               # loads at every alignment
            op = int(rng.choice([0b100000, 0b100001, 0b100010, 0b100011, 0b100100, 0b100101, 0b100110, 0b110000]))
            off = int(rng.integers(0, 64))
            if op in (0b100011, 0b110000):
                off &= ~3
            if op in (0b100001, 0b100101):
                off &= ~1
            p.append(itype(op, 28, rd if rd else 1, off))
        elif k == 12:                                               # stores at every alignment
            op = int(rng.choice([0b101000, 0b101001, 0b101010, 0b101011, 0b101110, 0b111000, 0b111101]))
            off = int(rng.integers(0, 64))
            if op in (0b101011, 0b111000, 0b111101):
                off &= ~3
            if op == 0b101001:
                off &= ~1
            p.append(itype(op, 28, max(1, rt) if op == 0b111000 else rt, off))
        else:                                                       # a branch over one instruction, taken or not
            op = int(rng.choice([0x04, 0x05, 0x06, 0x07, 0x01]))
            b_rt = rt if op in (0x04, 0x05) else (int(rng.choice([0, 1])) if op == 0x01 else 0)
            p += [itype(op, rs, b_rt, 2), SYNC, itype(0b001001, 12, 12, 1)]
    end = ENTRY + 4 * len(p)
    p += [SYNC] * 4
    image = {ENTRY + 4 * i: w & 0xFFFFFFFF for i, w in enumerate(p)}
    for i in range(20):
        image[DATA + 4 * i] = int(rng.integers(0, 1 << 32))
    return image, end


@pytest.mark.parametrize("seed", range(6))
def test_random_programs_satisfy_cpu_and_memory_constraints(orc, seed):
    rng = np.random.default_rng(1000 + seed)
    image, end = random_program(rng, 260)
    cpu = cg.MiniCpu(image, ENTRY)
    steps = 0
    while cpu.pc != end:
        cpu.step()
        steps += 1
        assert steps < 2000
    log_n = (cpu.clock()).bit_length()
    t = cpu.cpu_trace(log_n)
    bad = orc.orc_check_table_constraints(tr.T_CPU, binding.col_ptrs(t), t.shape[0], log_n)
    assert bad == 0, orc.orc_last_error()
    m = cg.memory_generate_trace(cpu.mem_ops)
    assert orc.orc_check_table_constraints(tr.T_MEMORY, binding.col_ptrs(m), 13, m.shape[1].bit_length() - 1) == 0
    # the arithmetic operations it logged give a valid Arithmetic table too
    import arith_gen as ag
    a = ag.arithmetic_trace(cpu.arith_ops, 16)
    assert orc.orc_check_table_constraints(tr.T_ARITHMETIC, binding.col_ptrs(a), 54, 16) == 0


def test_random_program_system_proves_and_verifies(orc):
    """One random program through the Cpu + Arithmetic + Logic + Memory slice: every arithmetic/logic result and every one
    of the ~1500 memory-channel uses is matched by the cross-table lookups."""
    import arith_gen as ag
    rng = np.random.default_rng(77)
    image, end = random_program(rng, 300)
    cpu = cg.MiniCpu(image, ENTRY)
    while cpu.pc != end:
        cpu.step()
    lg = lambda k: max(6, (max(k, 1) - 1).bit_length())
    traces = [cpu.cpu_trace(cpu.clock().bit_length()), ag.arithmetic_trace(cpu.arith_ops, 16),
              tr.logic_trace_from_ops(cpu.logic_ops, lg(len(cpu.logic_ops))), cg.memory_generate_trace(cpu.mem_ops)]
    proof = binding.prove_system(orc, tr.SYSTEM_CPU, traces)
    assert binding.verify_system(orc, tr.SYSTEM_CPU, proof) is None
