#!/usr/bin/env python
"""Generates tests/golden/golden_v1.json from the CPU oracle (oracle/liborc.so).  The reference holds no
golden vectors for this path and cannot be run here, so these fixtures pin the ORACLE's
outputs (regression pins for both the oracle and the CUDA path); the Poseidon entries are the
known answers of SURVEY Appendix D, which come from the reference's own constants."""
import hashlib
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import binding  # noqa: E402
import traces as tr  # noqa: E402
from conftest import random_columns  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def main():
    orc = binding.load()
    g = {"poseidon_kat": {
        "zeros": ["3c18a9786cb0b359", "c4055e3364a246c3", "7953db0ab48808f4", "c71603f33a1144ca", "d7709673896996dc", "46a84e87642f44ed",
                  "d032648251ee0b3c", "1c687363b207df62", "df8565563e8045fe", "40f5b37ff4254dae", "d070f637b431067c", "1792b1c4342109d7"],
        "iota": ["d64e1e3efc5b8e9e", "53666633020aaa47", "d40285597c6a8825", "613a4f81e81231d2", "414754bfebd051f0", "cb1f8980294a023f",
                 "6eb2a9e4d54a9d0f", "1902bc3af467e056", "f045d5eafdc6021f", "e4150f77caaa3be5", "c9bfd01d39b50cce", "5c0a27fcb0e1459b"]}}
    # NTT of a fixed vector
    v = np.arange(1, 17, dtype=np.uint64).reshape(1, 16).copy()
    out = {}
    for kind, name in ((0, "fft"), (1, "ifft"), (2, "coset_ifft"), (3, "coset_fft")):
        a = v.copy()
        orc.orc_ntt(binding.u64ptr(a), 1, 4, kind)
        out[name] = [int(x) for x in a[0]]
    g["ntt_1_to_16"] = out
    # commitments
    caps = {}
    for ncols, log_n in ((3, 6), (13, 6), (54, 8)):
        cols = random_columns(ncols, 1 << log_n, seed=7 + ncols)
        cap = np.zeros(64, dtype=np.uint64)
        h = orc.orc_commit(binding.col_ptrs(cols), ncols, log_n, 2, 4, 1, binding.u64ptr(cap))
        orc.orc_batch_free(h)
        caps[f"{ncols}x2^{log_n}"] = [int(x) for x in cap[:8]] + [sha(cap)]
    g["commit_caps_first2_digests_and_sha256"] = caps
    # proofs
    proofs = {}
    for sid, name, t in ((tr.SYSTEM_LOGIC, "logic_2^6", [tr.logic_trace(6)]), (tr.SYSTEM_MEMORY, "memory_2^7", [tr.memory_trace(7)]),
                         (tr.SYSTEM_POSEIDON, "poseidon_2^6", [tr.poseidon_trace(orc, 6)]),
                         (tr.SYSTEM_MINI3, "mini3", [tr.poseidon_trace(orc, 6), tr.logic_trace(8), tr.memory_trace(7)])):
        p = binding.prove_system(orc, sid, t)
        assert binding.verify_system(orc, sid, p) is None
        proofs[name] = {"system": sid, "words": int(p.size), "sha256": sha(p), "trace_sha256": [sha(x) for x in t]}
    g["proofs"] = proofs
    (ROOT / "tests/golden/golden_v1.json").write_text(json.dumps(g, indent=1))
    print("wrote tests/golden/golden_v1.json")


if __name__ == "__main__":
    main()
