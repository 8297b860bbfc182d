#!/usr/bin/env python
"""Diffs the output of shim/fingerprint_test.rs (run under cargo in the reference tree) against
tests/golden/constraint_fingerprints_v1.json.   python tools/compare_fingerprints.py cargo_output.txt"""
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
want = json.loads((ROOT / "tests/golden/constraint_fingerprints_v1.json").read_text())
tables = {t["table"]: t["acc"] for t in want["tables"]}
lookups = {(l["table"], l["index"]): (l["num_columns"], l["fp"]) for l in want["lookups"]}
ctls = {(c["index"], e): (x["role"], x["table"], x["num_columns"], x["fp"]) for c in want["ctls"] for e, x in enumerate(c["entries"])}
bad = seen = 0
for line in open(sys.argv[1]):
    w = line.split()
    if not w:
        continue
    if w[0] == "table" and len(w) == 5:
        seen += 1
        if tables.get(w[1]) != [int(w[3]), int(w[4])]:
            bad += 1
            print("MISMATCH", line.strip(), "expected", tables.get(w[1]))
    elif w[0] == "lookup" and len(w) == 7:
        seen += 1
        if lookups.get((w[1], int(w[2]))) != (int(w[4]), int(w[6])):
            bad += 1
            print("MISMATCH", line.strip(), "expected", lookups.get((w[1], int(w[2]))))
    elif w[0] == "ctl" and len(w) == 10:
        seen += 1
        if ctls.get((int(w[1]), int(w[3]))) != (w[4], w[5], int(w[7]), int(w[9])):
            bad += 1
            print("MISMATCH", line.strip(), "expected", ctls.get((int(w[1]), int(w[3]))))
total = len(tables) + len(lookups) + len(ctls)
print(f"{seen} of {total} fingerprints compared, {bad} mismatches")
sys.exit(1 if bad or seen != total else 0)
