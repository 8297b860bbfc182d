#!/usr/bin/env python
"""Experiment: K U20 proofs on one GPU with 1 vs 2 worker contexts (two host threads, two streams)."""
import pathlib
import sys
import threading
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import bench  # noqa: E402
from zkm_b200 import lib as zl  # noqa: E402

lib = zl.init(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
segs = [bench.Segment(lib, "U20", seed_offset=i) for i in range(2)]
for s in segs:
    for _ in range(2):
        s.step_device()
    s.sync()


def run(nworkers, e2e=False):
    workers = [zl.Worker(lib) for _ in range(nworkers)]
    if e2e:
        for s in segs[:nworkers]:
            s.prepare_host()

    def body(i):
        with workers[i]:
            for _ in range(2):
                segs[i].step_device()                  # warm this context's tables and arena
            segs[i].sync()
            barrier.wait()
            for _ in range(K // nworkers):
                (segs[i].step_e2e if e2e else segs[i].step_device)()
            segs[i].sync()
    barrier = threading.Barrier(nworkers + 1)
    th = [threading.Thread(target=body, args=(i,)) for i in range(nworkers)]
    for t in th:
        t.start()
    barrier.wait()
    t0 = time.time()
    for t in th:
        t.join()
    dt = time.time() - t0
    for w in workers:
        w.close()
    return dt * 1000 / K


for e2e in (False, True):
    for n in (1, 2, 3):
        if K % n:
            continue
        print(f"e2e={e2e} workers={n}: {run(n, e2e):.1f} ms per proof", flush=True)
