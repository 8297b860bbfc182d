#!/usr/bin/env python
"""Profiling target for `ncu --set full`: one commit of COLS columns x 2^LOGN (NTT passes, leaf hashing,
Merkle levels at production size) and, with --prove, one U<LOGP> full prove (quotient/openings/FRI kernels).
usage: ncu ... python tools/prof_target.py [--cols 54] [--logn 20] [--prove 18]"""
import argparse
import ctypes as C
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import bench  # noqa: E402
from zkm_b200 import lib as zl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cols", type=int, default=54)
ap.add_argument("--logn", type=int, default=20)
ap.add_argument("--prove", type=int, default=0)
ap.add_argument("--warm", type=int, default=0)
ap.add_argument("--trace", action="store_true")
a = ap.parse_args()
lib = zl.init(0)
err = C.c_void_p()
if a.cols:
    n = 1 << a.logn
    buf = torch.empty(a.cols * n, dtype=torch.int64, device="cuda")
    zl.check(lib, lib.zkm_b200_synth_columns_device(buf.data_ptr(), a.cols, a.logn, 0x5EED000000000000 | a.cols, C.byref(err)), err)
    cap = np.zeros(64, dtype=np.uint64)
    h = C.c_void_p()
    zl.check(lib, lib.zkm_b200_commit_values_device(buf.data_ptr(), a.cols, a.logn, 2, 4, C.byref(h), zl.u64ptr(cap), C.byref(err)), err)
    lib.zkm_b200_batch_free(h)
if a.prove:
    import os
    seg = bench.Segment(lib, f"U{a.prove}")
    for _ in range(a.warm):
        seg.step_device()
    if a.trace:
        os.environ["ZKM_TRACE"] = "1"      # host wall-clock per prover phase (steady state after the warm proofs)
    seg.step_device()
    seg.sync()
print("prof_target done")
