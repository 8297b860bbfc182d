"""Proof wire format: zkm_b200_proof_table_json must be the serde_json text of the reference's `StarkProof`
(proof.rs:177-189) -- checked here structurally against an independent walk of the flat proof buffer (a proof made by the CPU
oracle, so the test needs no GPU): field names and order, nesting, every number, compactness."""
import json
import re

import numpy as np

import traces as tr
from oracle import binding
from zkm_b200 import lib as zl

MAGIC = 0x464F4F52504D4B5A


class Walk:
    def __init__(self, w):
        self.w, self.p = [int(x) for x in w], 0

    def u(self):
        self.p += 1
        return self.w[self.p - 1]

    def words(self, n):
        self.p += n
        return self.w[self.p - n:self.p]

    def vec(self, unit):
        n = self.u()
        x = self.words(n * unit)
        return x if unit == 1 else [x[i * unit:(i + 1) * unit] for i in range(n)]


def expected_tables(proof):
    """The structure serde derives for AllProof.stark_proofs[t].proof, from the buffer layout of include/zkm_b200.h."""
    r = Walk(proof)
    assert r.u() == MAGIC and r.u() == 1
    nt = r.u()
    r.words(2 * r.u())
    rb, ra = r.words(8), r.words(8)
    ud = r.vec(1)
    H = lambda hs: [{"elements": h} for h in hs]      # noqa: E731
    out = []
    for _ in range(nt):
        r.words(12)
        d = {"trace_cap": H(r.vec(4)), "auxiliary_polys_cap": H(r.vec(4)), "quotient_polys_cap": H(r.vec(4))}
        d["openings"] = {"local_values": r.vec(2), "next_values": r.vec(2), "auxiliary_polys": r.vec(2),
                         "auxiliary_polys_next": r.vec(2), "ctl_zs_first": r.vec(1), "quotient_polys": r.vec(2)}
        ncaps = r.u()
        caps = [H(r.vec(4)) for _ in range(ncaps)]
        rounds = []
        for _ in range(r.u()):
            ev = []
            for _ in range(r.u()):
                leaf = r.vec(1)
                ev.append([leaf, {"siblings": H(r.vec(4))}])
            steps = []
            for _ in range(r.u()):
                evals = r.vec(2)
                steps.append({"evals": evals, "merkle_proof": {"siblings": H(r.vec(4))}})
            rounds.append({"initial_trees_proof": {"evals_proofs": ev}, "steps": steps})
        d["opening_proof"] = {"commit_phase_merkle_caps": caps, "query_round_proofs": rounds, "final_poly": {"coeffs": r.vec(2)},
                              "pow_witness": r.u()}
        out.append(d)
    assert r.p == len(r.w)
    return out, {"roots_before": {"root": rb}, "roots_after": {"root": ra}, "userdata": ud}


def test_stark_proof_json_matches_the_serde_layout(orc):
    lib = zl.load()
    traces = [tr.poseidon_trace(orc, 6), tr.logic_trace(8), tr.memory_trace(7)]
    proof = binding.prove_system(orc, tr.SYSTEM_MINI3, traces, roots_before=[3] * 8, userdata=bytes(range(32)))
    want, want_pv = expected_tables(proof)
    for t, w in enumerate(want):
        s = zl.proof_table_json(lib, proof, t)
        assert not re.search(r"\s", s)                                   # serde_json::to_string is compact
        got = json.loads(s)
        assert got == w
        assert list(got.keys()) == ["trace_cap", "auxiliary_polys_cap", "quotient_polys_cap", "openings", "opening_proof"]
        assert list(got["openings"].keys()) == ["local_values", "next_values", "auxiliary_polys", "auxiliary_polys_next", "ctl_zs_first",
                                                "quotient_polys"]
        assert list(got["opening_proof"].keys()) == ["commit_phase_merkle_caps", "query_round_proofs", "final_poly", "pow_witness"]
        assert json.dumps(w, separators=(",", ":")) == s                  # byte for byte: same field order, no padding
    pv = zl.public_values_json(lib, proof)
    assert json.loads(pv) == want_pv and json.dumps(want_pv, separators=(",", ":")) == pv
    assert want_pv["userdata"] == list(range(32)) and want_pv["roots_before"]["root"] == [3] * 8


def test_json_errors():
    import ctypes as C
    lib = zl.load()
    bad = np.zeros(8, dtype=np.uint64)
    out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
    assert lib.zkm_b200_proof_table_json(zl.u64ptr(bad), bad.size, 0, C.byref(out), C.byref(n), C.byref(err)) == -1
    assert b"magic" in C.cast(err, C.c_char_p).value
    lib.zkm_b200_free_string(err)


def test_hostile_proof_buffers_are_errors_not_crashes(orc):
    """Every length in the flat buffer comes from the buffer: counts chosen so that `pos + k` or `k * unit` wraps, counts far
    beyond the buffer and every truncation must come back as an error from both JSON entry points."""
    import ctypes as C
    lib = zl.load()
    proof = binding.prove_system(orc, tr.SYSTEM_LOGIC, [tr.logic_trace(6)])

    def calls(buf, words=None):
        buf = np.ascontiguousarray(buf, dtype=np.uint64)
        words = buf.size if words is None else words
        rcs = []
        for fn, args in ((lib.zkm_b200_proof_table_json, (0,)), (lib.zkm_b200_public_values_json, ())):
            out, n, err = C.c_void_p(), C.c_size_t(), C.c_void_p()
            rc = fn(zl.u64ptr(buf), C.c_size_t(words), *args, C.byref(out), C.byref(n), C.byref(err))
            assert rc in (0, -1) and (rc == 0) == (err.value is None)
            lib.zkm_b200_free_string(err if rc else out)
            rcs.append(rc)
        return rcs

    assert calls(proof) == [0, 0]
    # the length words of the buffer: walk it once to find them (header counts, then every vec length of table 0)
    r = Walk(proof)
    r.words(3)
    at = [r.p]                                   # number of challenges
    r.words(2 * r.u() + 16)
    at.append(r.p)                               # userdata length
    r.vec(1)
    r.words(12)
    for unit in (4, 4, 4, 2, 2, 2, 2, 1, 2):
        at.append(r.p)
        r.vec(unit)
    at.append(r.p)                               # number of commit-phase caps
    hostile = [(1 << 64) - 1, (1 << 63), (1 << 62), (1 << 62) + 1, (1 << 61) + 3, proof.size, 1 << 32]
    for pos in at:
        for k in hostile:
            bad = proof.copy()
            bad[pos] = k
            rcs = calls(bad)
            assert rcs[0] == -1, (pos, k)          # the table walk crosses every one of these counts
            assert rcs[1] == (-1 if pos in at[:2] else 0)
    rng = np.random.default_rng(3)
    for words in sorted({0, 1, 2, 3, 4, 30, proof.size - 1, *rng.integers(5, proof.size - 1, size=40).tolist()}):
        assert calls(proof, words)[0] == -1
    for _ in range(200):                         # random single-word corruption: any status, no crash
        bad = proof.copy()
        bad[int(rng.integers(0, proof.size))] = int(rng.integers(0, 1 << 63)) << int(rng.integers(0, 2))
        calls(bad)


def test_segment_json_is_the_serde_text_of_the_emulator_segment():
    """zkm_b200_segment_json against Python's own compact JSON of the same structure (emulator/src/state.rs:33-48: field order of
    the struct, BTreeMap<u32, u32> as an object with decimal-string keys in ascending numeric order, [u8; 32] and Vec<u8> as
    arrays of numbers), and the reference's get_input_image word order (memory.rs:524-538)."""
    lib = zl.load()
    rng = np.random.default_rng(8)
    idx = [3, 0x10, 0x7FFFD, 0x81020]
    pages = rng.integers(0, 256, size=(len(idx), 4096), dtype=np.uint8)
    pages[1] = 0
    ids = [bytes(int(b) for b in rng.integers(0, 256, size=32)) for _ in range(4)]
    streams = [b"", bytes(range(5)), bytes([255, 0, 7])]
    pvs = bytes(range(40))
    text = zl.segment_json(lib, idx, pages, 0x1234, 7, ids[0], ids[1], ids[2], ids[3], 0xFFFF0000, (1 << 40) + 5, streams, 2, pvs, 40)
    image = {}
    for k, pi in enumerate(idx):
        for i in range(1024):
            image[str((pi << 12) + 4 * i)] = int.from_bytes(pages[k, 4 * i:4 * i + 4].tobytes(), "little")
    want = {"mem_image": image, "pc": 0x1234, "segment_id": 7, "pre_image_id": list(ids[0]), "pre_hash_root": list(ids[1]), "image_id": list(ids[2]),
            "page_hash_root": list(ids[3]), "end_pc": 0xFFFF0000, "step": (1 << 40) + 5, "input_stream": [list(b) for b in streams], "input_stream_ptr": 2,
            "public_values_stream": list(pvs), "public_values_stream_ptr": 40}
    assert text.decode() == json.dumps(want, separators=(",", ":"))
    assert list(json.loads(text)["mem_image"].keys())[:2] == [str(3 << 12), str((3 << 12) + 4)]
    import pytest
    with pytest.raises(zl.ZkmError, match="ascending"):
        zl.segment_json(lib, [5, 5], pages[:2], 0, 0, ids[0], ids[1], ids[2], ids[3], 0, 0, [], 0, b"", 0)
    empty = zl.segment_json(lib, [], np.zeros((0, 4096), dtype=np.uint8), 1, 2, ids[0], ids[1], ids[2], ids[3], 3, 4, [], 0, b"", 0)
    assert json.loads(empty)["mem_image"] == {} and json.loads(empty)["input_stream"] == []
