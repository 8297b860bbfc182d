#!/usr/bin/env python
"""Writes tests/golden/constraint_fingerprints_v1.json: per table the number of constraints and the alpha-fold of all of them
on one fixed pseudo-random frame, per in-table lookup and per cross-table lookup a fold of its column / filter evaluations
(oracle/oracle_capi.cpp orc_*_fingerprint).  The same numbers are printed on the reference side by the Rust test in
INTEGRATION.md section 3, so ONE cargo run pins all 12 transcribed constraint sets, their emission order and the 15 CTL
descriptions.  Here the file is a regression pin for both the oracle and (through proof equality) the CUDA kernels."""
import ctypes as C
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import binding  # noqa: E402

SEED = 0x5EEDF1A600000000
TABLES = ["Arithmetic", "Cpu", "Poseidon", "PoseidonSponge", "Keccak", "KeccakSponge", "ShaExtend", "ShaExtendSponge", "ShaCompress",
          "ShaCompressSponge", "Logic", "Memory"]


def collect():
    orc = binding.load()
    out = {"seed": SEED, "alphas": [0x9E3779B97F4A7C15 % binding_P(), 0xC2B2AE3D27D4EB4F % binding_P()], "z_last": 3, "lagrange_first": 5,
           "lagrange_last": 7, "cell": "local[c] = splitmix64(seed + 2c) % p, next[c] = splitmix64(seed + 2c + 1) % p; table t uses "
           "seed + 0x10000 * t", "tables": [], "lookups": [], "ctls": []}
    buf = (C.c_uint64 * 3)()
    for t, name in enumerate(TABLES):
        assert orc.orc_table_fingerprint(t, SEED + 0x10000 * t, buf) == 0, orc.orc_last_error()
        out["tables"].append({"table": name, "num_constraints": int(buf[0]), "acc": [int(buf[1]), int(buf[2])]})
        n = orc.orc_lookup_fingerprint(t, -1, 0, buf)
        for i in range(n):
            orc.orc_lookup_fingerprint(t, i, SEED + 0x10000 * t, buf)
            out["lookups"].append({"table": name, "index": i, "num_columns": int(buf[0]), "fp": int(buf[1])})
    for c in range(orc.orc_num_ctls(0)):
        nl = orc.orc_ctl_fingerprint(0, c, 0, SEED, buf)
        entries = []
        for e in range(nl + 1):
            assert orc.orc_ctl_fingerprint(0, c, e, SEED, buf) == nl
            entries.append({"role": "looking" if e < nl else "looked", "table": TABLES[int(buf[0])], "num_columns": int(buf[1]), "fp": int(buf[2])})
        out["ctls"].append({"index": c, "num_looking": nl, "entries": entries})
    return out


def binding_P():
    return 0xFFFFFFFF00000001


if __name__ == "__main__":
    dst = ROOT / "tests/golden/constraint_fingerprints_v1.json"
    dst.write_text(json.dumps(collect(), indent=1))
    d = json.loads(dst.read_text())
    print(f"{len(d['tables'])} tables, {len(d['lookups'])} lookups, {len(d['ctls'])} CTLs -> {dst}")
    for t in d["tables"]:
        print(f"  {t['table']:18s} {t['num_constraints']:5d} constraints")
