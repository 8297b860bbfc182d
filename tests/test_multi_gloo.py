"""CPU test of the N > 1 path (task ⑤): world_size-2 gloo processes shard independent segments, prove them
(here with the CPU oracle standing in for the per-rank GPU prover, since this container has no GPU) and
rank 0 gathers and verifies every proof."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding
        import traces as tr
        from zkm_b200 import multi
        orc = binding.load()
        orc.orc_set_threads(2)
        nseg = 5

        def prove(i):
            # segment i = a Logic-table System with its own trace and public values
            return binding.prove_system(orc, tr.SYSTEM_LOGIC, [tr.logic_trace(6, seed=100 + i)], roots_before=[i + 1] * 8)

        proofs = multi.prove_segments(prove, nseg, rank, world)
        if rank == 0:
            assert proofs is not None and len(proofs) == nseg
            for i, p in enumerate(proofs):
                assert p is not None and binding.verify_system(orc, tr.SYSTEM_LOGIC, p) is None
                assert int(p[3 + 1 + 4]) == i + 1            # roots_before[0] of segment i: proofs are in segment order
            q.put("ok")
        else:
            assert proofs is None
    finally:
        dist.destroy_process_group()


def test_shard_segments():
    from zkm_b200 import multi
    assert multi.shard_segments(8, 0, 8) == [0]
    assert multi.shard_segments(5, 1, 2) == [1, 3]
    assert sorted(sum((multi.shard_segments(11, r, 4) for r in range(4)), [])) == list(range(11))
    with pytest.raises(ValueError):
        multi.shard_segments(4, 4, 4)


def test_two_rank_gloo_prove_and_gather():
    from oracle import binding
    binding.load()                                   # build the oracle once, before forking
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"
