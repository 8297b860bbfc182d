#!/usr/bin/env python
"""Derives (and checks against the naive schedule) the merged partial-round constants used by
zkm_b200/csrc/poseidon_v2.cuh poseidon_permute_v8: PARTIAL_A[22] and RC26_MERGED[12].
Input: the round constants generated from reference prover/src/poseidon/constants.rs."""
import pathlib
import random
import re

P = 0xFFFFFFFF00000001
ROOT = pathlib.Path(__file__).resolve().parent.parent
src = (ROOT / "zkm_b200/csrc/poseidon_consts.h").read_text()
m = re.search(r"#define POSEIDON_ALL_ROUND_CONSTANTS_INIT \{(.*?)\}", src, re.S)
rc = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
assert len(rc) == 360
C = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]


def mds(v):
    return [(sum(v[(i + r) % 12] * C[i] for i in range(12)) + (8 * v[0] if r == 0 else 0)) % P for r in range(12)]


k, A = [0] * 12, []
for r in range(4, 26):
    d = [(k[i] + rc[12 * r + i]) % P for i in range(12)]
    A.append(d[0])
    k = mds([0] + d[1:])
rc26 = [(k[i] + rc[12 * 26 + i]) % P for i in range(12)]


def naive(s):
    for r in range(30):
        s = [(s[i] + rc[12 * r + i]) % P for i in range(12)]
        s = [pow(x, 7, P) for x in s] if (r < 4 or r >= 26) else [pow(s[0], 7, P)] + s[1:]
        s = mds(s)
    return s


def merged(s):
    for r in range(30):
        if r < 4 or r >= 26:
            c = rc26 if r == 26 else rc[12 * r:12 * r + 12]
            s = [pow((s[i] + c[i]) % P, 7, P) for i in range(12)]
        else:
            s = [pow((s[0] + A[r - 4]) % P, 7, P)] + s[1:]
        s = mds(s)
    return s


random.seed(1)
for _ in range(50):
    st = [random.randrange(P) for _ in range(12)]
    assert naive(list(st)) == merged(list(st))
assert naive([0] * 12)[0] == 0x3C18A9786CB0B359
print("PARTIAL_A =", ", ".join(f"0x{x:016x}ULL" for x in A))
print("RC26_MERGED =", ", ".join(f"0x{x:016x}ULL" for x in rc26))
