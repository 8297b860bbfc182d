"""GPU test of in-segment sharding (include/zkm_b200.h "In-segment sharding"): G = 2, 4 or 8 processes, one per
GPU, prove ONE segment together through the C ABI; every rank's proof must be identical and equal to the proof the same
library computes on one GPU alone (which tests/test_gpu_prove.py compares with the oracle word for word).  Needs >= 2 GPUs
on the box (`gpurun --gpus 2`); skipped otherwise."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, heights, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hashlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from zkm_b200 import lib as zl, multi
        import traces as tr
        lib = zl.init(rank)
        traces = zl.synth_traces(lib, tr.SYSTEM_ALL_STARK, heights)       # the same segment on every rank
        single = zl.prove_with_traces(lib, traces)                        # one GPU alone
        gi, ri, g = multi.shard_group_init(lib, world)
        assert (gi, ri, g) == (0, rank, world)
        sharded = zl.prove_with_traces(lib, traces)                       # cooperative call
        again = zl.prove_with_traces(lib, traces)
        multi.shard_group_shutdown(lib)
        after = zl.prove_with_traces(lib, traces)                         # sharding off again
        assert sharded.size == single.size and (sharded == single).all(), "sharded proof differs from the single-GPU proof"
        assert (again == single).all() and (after == single).all()
        d = torch.tensor(list(hashlib.sha256(sharded.tobytes()).digest()), dtype=torch.uint8, device="cuda")
        alld = [torch.zeros_like(d) for _ in range(world)]
        dist.all_gather(alld, d)
        assert all(bool((x == alld[0]).all()) for x in alld), "ranks returned different proofs"
        if rank == 0:
            q.put(("ok", single))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_in_segment_sharded_proof_equals_single_gpu_proof(orc, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import binding
    import traces as tr
    # Arithmetic 2^16, Cpu 2^14, Logic 2^13, Memory 2^14 are sharded (>= 2^13 rows); the 2^6 / 2^7-row tables are not
    heights = [16, 14, 6, 6, 6, 6, 6, 7, 6, 6, 13, 14]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, heights, q)) for r in range(world)]
    import queue
    import time
    for p in procs:
        p.start()
    # fail fast: a rank that dies leaves its peers blocked in a collective, so poll instead of waiting for a timeout
    tag, proof, t0 = None, None, time.time()
    while tag is None and time.time() - t0 < 420:
        try:
            tag, proof = q.get(timeout=2)
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=20 if tag == "ok" else 1)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():
            p.terminate()
    assert tag == "ok" and all(c == 0 for c in codes), codes
    # and the single-GPU proof it equals is the oracle's proof of the same traces
    from zkm_b200 import lib as zl
    lib = zl.init(0)
    cpu = binding.prove_system(orc, tr.SYSTEM_ALL_STARK, zl.synth_traces(lib, tr.SYSTEM_ALL_STARK, heights))
    assert proof.size == cpu.size and (proof == cpu).all()
