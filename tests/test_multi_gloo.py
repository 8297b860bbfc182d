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


# ---- in-segment sharding (zkm_b200/csrc/shard.cuh, zkm_b200/multi.py): the ownership rules, checked with the oracle ----

def _shard_worker(rank, world, port, q):
    """Each rank plays one GPU of a shard group: from the LDE values of the cosets it OWNS (and nothing else) it hashes its
    leaves, builds its cap subtrees and answers the queries that fall into its leaf quarters; the pieces are exchanged with
    gloo all_gather (NCCL on the GPU box) and must reassemble the oracle's single-rank commitment."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding
        from oracle.binding import u64ptr, col_ptrs
        from conftest import random_columns
        from zkm_b200 import multi
        orc = binding.load()
        orc.orc_set_threads(1)
        ncols, log_n, cap_h = 9, 5, 4
        n, log_leaves = 1 << log_n, log_n + 2
        cols = random_columns(ncols, n, seed=4242)
        cap_full = np.zeros(64, dtype=np.uint64)
        h = orc.orc_commit(col_ptrs(cols), ncols, log_n, 2, cap_h, 1, u64ptr(cap_full))
        assert h
        lde = np.zeros((ncols, 4 * n), dtype=np.uint64)
        for c in range(ncols):
            orc.orc_batch_get_lde(h, c, u64ptr(lde[c]))
        # --- this rank's share: leaves of its cosets only
        mine = multi.owned_cosets(rank, world)
        pieces = []
        for j in mine:
            qtr = multi.bitrev(j, 2)
            level = []
            for leaf in range(qtr * n, (qtr + 1) * n):
                m = multi.bitrev(leaf, log_leaves)
                assert m % 4 == j                                   # the whole leaf quarter lies in coset j
                assert multi.leaf_owner(leaf, log_leaves, world) == rank
                row = np.ascontiguousarray(lde[:, m])
                d = np.zeros(4, dtype=np.uint64)
                orc.orc_hash_or_noop(u64ptr(row), ncols, u64ptr(d))
                level.append(d)
            while len(level) > (1 << cap_h) // 4:
                nxt = []
                for k in range(0, len(level), 2):
                    d = np.zeros(4, dtype=np.uint64)
                    orc.orc_two_to_one(u64ptr(level[k]), u64ptr(level[k + 1]), u64ptr(d))
                    nxt.append(d)
                level = nxt
            pieces.append(np.stack(level))
        piece = torch.from_numpy(np.concatenate(pieces).view(np.int64).reshape(-1))
        got = [torch.zeros_like(piece) for _ in range(world)]
        dist.all_gather(got, piece)                                 # the cap all-gather
        cap = multi.assemble_cap([g.numpy().view(np.uint64) for g in got], world, cap_h)
        assert (cap.reshape(-1) == cap_full).all()
        # --- queries: the owner answers, everyone ends up with the oracle's opening
        plen = log_leaves - cap_h
        for leaf in (0, 1, n + 3, 2 * n + 17, 3 * n + 5, 4 * n - 1):
            ans = torch.zeros(ncols + 4 * plen, dtype=torch.int64)
            if multi.leaf_owner(leaf, log_leaves, world) == rank:
                row = np.zeros(ncols, dtype=np.uint64); sib = np.zeros(4 * plen, dtype=np.uint64)
                orc.orc_batch_open(h, leaf, u64ptr(row), u64ptr(sib))
                ans = torch.from_numpy(np.concatenate([row, sib]).view(np.int64))
            allans = [torch.zeros_like(ans) for _ in range(world)]
            dist.all_gather(allans, ans)
            sel = allans[multi.leaf_owner(leaf, log_leaves, world)].numpy().view(np.uint64)
            row = np.zeros(ncols, dtype=np.uint64); sib = np.zeros(4 * plen, dtype=np.uint64)
            orc.orc_batch_open(h, leaf, u64ptr(row), u64ptr(sib))
            assert (sel[:ncols] == row).all() and (sel[ncols:] == sib).all()
        orc.orc_batch_free(h)
        if rank == 0:
            q.put("ok")
    finally:
        dist.destroy_process_group()


def test_shard_ownership_rules():
    from zkm_b200 import multi
    assert multi.owned_cosets(0, 2) == [0, 1] and multi.owned_cosets(1, 2) == [2, 3]
    assert [multi.owned_cosets(r, 4) for r in range(4)] == [[0], [1], [2], [3]]
    assert [multi.owned_cosets(r, 8) for r in range(8)] == [[0], [0], [1], [1], [2], [2], [3], [3]]      # two ranks per coset
    log_leaves = 9
    for g in (1, 2, 4, 8):
        owned = sorted(sum((multi.owned_cap_entries(r, g) for r in range(g)), []))
        assert owned == list(range(16))                             # every cap entry has exactly one owner
        # the two halves of the quotient domain (LDE cosets 0 and 2) go to different ranks as soon as there are two
        assert (multi.coset_owner(0, g) != multi.coset_owner(2, g)) == (g > 1)
        # a leaf's owner is the rank whose segments contain it, and its coset is one that rank computes
        nseg = 4 * multi.parts(g)
        for leaf in range(1 << log_leaves):
            r = multi.leaf_owner(leaf, log_leaves, g)
            assert (leaf * nseg) >> log_leaves in multi.owned_segments(r, g)
            assert multi.bitrev(leaf, log_leaves) % 4 in multi.owned_cosets(r, g)
        # 8 ranks: the two owners of a coset split its rows by parity (row i = natural LDE index >> 2)
        if g == 8:
            for leaf in range(1 << log_leaves):
                i = multi.bitrev(leaf, log_leaves) >> 2
                assert multi.leaf_owner(leaf, log_leaves, g) % 2 == i % 2
    with pytest.raises(ValueError):
        multi.owned_cosets(0, 3)


@pytest.mark.parametrize("world", [2, 4])
def test_in_segment_sharding_gloo(world):
    from oracle import binding
    binding.load()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"
