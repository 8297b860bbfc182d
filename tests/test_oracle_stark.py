"""CPU tests of the oracle's STARK layer (oracle/stark.h): valid traces satisfy the transcribed
constraints (the reference's generate => constraints-vanish tests), the restated prover's
proofs are accepted by the restated verifier (verifier.rs logic), and tampering is rejected."""
import numpy as np
import pytest

import traces as tr
from oracle import binding


@pytest.mark.parametrize("kind,gen", [(tr.T_LOGIC, "logic"), (tr.T_MEMORY, "memory"), (tr.T_POSEIDON, "poseidon")])
def test_valid_traces_satisfy_constraints(orc, kind, gen):
    t = {"logic": lambda: tr.logic_trace(7), "memory": lambda: tr.memory_trace(8), "poseidon": lambda: tr.poseidon_trace(orc, 6)}[gen]()
    bad = orc.orc_check_table_constraints(kind, binding.col_ptrs(t), t.shape[0], t.shape[1].bit_length() - 1)
    assert bad == 0, orc.orc_last_error()
    # and a corrupted cell is caught
    t2 = t.copy()
    col = {"logic": 68, "memory": 10, "poseidon": 20}[gen]
    t2[col, 3] = (int(t2[col, 3]) + 1) % tr.P
    bad = orc.orc_check_table_constraints(kind, binding.col_ptrs(t2), t2.shape[0], t2.shape[1].bit_length() - 1)
    assert bad >= 1


def _systems(orc):
    return {
        tr.SYSTEM_LOGIC: [tr.logic_trace(6)],
        tr.SYSTEM_POSEIDON: [tr.poseidon_trace(orc, 6)],
        tr.SYSTEM_MEMORY: [tr.memory_trace(7)],
        tr.SYSTEM_MINI3: [tr.poseidon_trace(orc, 6), tr.logic_trace(8), tr.memory_trace(7)],
    }


@pytest.mark.parametrize("sid", [tr.SYSTEM_LOGIC, tr.SYSTEM_POSEIDON, tr.SYSTEM_MEMORY, tr.SYSTEM_MINI3])
def test_oracle_prove_then_verify(orc, sid):
    traces = _systems(orc)[sid]
    proof = binding.prove_system(orc, sid, traces)
    assert binding.verify_system(orc, sid, proof) is None
    # deterministic
    assert (binding.prove_system(orc, sid, traces) == proof).all()


def test_oracle_verifier_rejects_tampering(orc):
    sid = tr.SYSTEM_MINI3
    traces = _systems(orc)[sid]
    proof = binding.prove_system(orc, sid, traces)
    assert binding.verify_system(orc, sid, proof) is None
    rng = np.random.default_rng(5)
    rejected = 0
    positions = list(rng.integers(3, proof.size, size=60)) + [proof.size - 1, 5, 6]
    for pos in positions:
        bad = proof.copy()
        bad[pos] = (int(bad[pos]) + 1) % tr.P
        if binding.verify_system(orc, sid, bad) is not None:
            rejected += 1
    # every word of the proof is bound by the transcript, a Merkle path, or a shape check, except
    # init_challenger_state (metadata the native verifier does not read: verifier.rs:27-176)
    assert rejected >= len(positions) - 3
    # wrong public values -> different challenges -> reject
    bad = proof.copy()
    bad[3 + 1 + 4 + 2] ^= 1
    assert binding.verify_system(orc, sid, bad) is not None


def test_invalid_trace_does_not_verify(orc):
    t = tr.logic_trace(6)
    t[68, 5] = (int(t[68, 5]) + 1) % tr.P
    proof = binding.prove_system(orc, tr.SYSTEM_LOGIC, [t])
    assert binding.verify_system(orc, tr.SYSTEM_LOGIC, proof) is not None


def test_non_binary_filter_is_an_error(orc):
    t = tr.logic_trace(6)
    t[0, 2] = 2        # IS_AND = 2 -> CTL filter evaluates to 2 (cross_table_lookup.rs:741)
    with pytest.raises(RuntimeError, match="Non-binary filter"):
        binding.prove_system(orc, tr.SYSTEM_LOGIC, [t])
