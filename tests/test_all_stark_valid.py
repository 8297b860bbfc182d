"""The whole AllStark on a VALID trace: a MIPS program with arithmetic, logic, memory, control flow, 17 syscalls and the
Keccak / SHA-256 precompiles, plus the bootstrap's Poseidon image-id hash, gives valid traces of all 12 tables
(tests/traces.py all_stark_valid_traces); the proof over the reference's 12 tables and 15 cross-table lookups
(all_stark.rs:136-542) is accepted by the restated verifier, and proofs over traces whose cross-table data disagree are not."""
import numpy as np
import pytest

import cpu_gen as cg
import traces as tr
from oracle import binding


@pytest.fixture(scope="module")
def valid_traces(orc):
    return tr.all_stark_valid_traces(orc)


def _check(orc, kind, t):
    return orc.orc_check_table_constraints(kind, binding.col_ptrs(t), t.shape[0], t.shape[1].bit_length() - 1)


def test_every_table_satisfies_its_constraints(orc, valid_traces):
    assert [t.shape[0] for t in valid_traces] == [54, 259, 262, 110, 2431, 470, 78, 76, 224, 127, 69, 13]
    for kind, t in enumerate(valid_traces):
        assert _check(orc, kind, t) == 0, (kind, orc.orc_last_error())
    cpu = valid_traces[tr.T_CPU]
    assert int(cpu[cg.OP["syscall"]].sum()) == 18
    for flag in (cg.IS_POSEIDON_SPONGE, cg.IS_KECCAK_SPONGE, cg.IS_SHA_COMPRESS_SPONGE):
        assert int(cpu[flag].sum()) == 1
    assert int(cpu[cg.IS_SHA_EXTEND_SPONGE].sum()) == 48


def test_syscall_rows_are_constrained(orc, valid_traces):
    cpu = valid_traces[tr.T_CPU]
    rows = np.nonzero(cpu[cg.OP["syscall"]])[0]
    caught = 0
    for r in rows:
        t2 = cpu.copy()
        t2[cg.ch(4, 5), r] = (int(t2[cg.ch(4, 5), r]) + 1) % tr.P        # v0, the syscall's result
        caught += _check(orc, tr.T_CPU, t2) >= 1
    # results of the numbers syscall.rs knows are constrained; precompile and unknown numbers return what the prover says
    assert caught >= 12


def test_all_stark_valid_proof_verifies(orc, valid_traces):
    proof = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, valid_traces)
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, proof) is None


@pytest.mark.parametrize("what", ["keccak_digest", "sha_compress_digest", "image_id", "sha_extend_word"])
def test_all_stark_rejects_wrong_precompile_results(orc, valid_traces, what):
    """The CPU row that receives a precompile's result is tied to the sponge table only by the cross-table lookup."""
    ts = [t.copy() for t in valid_traces]
    cpu = ts[tr.T_CPU]
    flag = {"keccak_digest": cg.IS_KECCAK_SPONGE, "sha_compress_digest": cg.IS_SHA_COMPRESS_SPONGE, "image_id": cg.IS_POSEIDON_SPONGE,
            "sha_extend_word": cg.IS_SHA_EXTEND_SPONGE}[what]
    r = int(np.nonzero(cpu[flag])[0][0])
    cpu[cg.GENERAL, r] = (int(cpu[cg.GENERAL, r]) + 1) % tr.P
    assert _check(orc, tr.T_CPU, cpu) == 0
    proof = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, ts)
    assert binding.verify_system(orc, tr.SYSTEM_ALL_STARK, proof) is not None
