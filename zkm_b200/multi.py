"""Multi-GPU host logic: one process per GPU, segments ("continuations") sharded across ranks, no
data-path collective; the only exchange is the final gather of the finished proofs on rank 0.

The reference proves the segments of one execution sequentially in one process
(prover/examples/utils/src/utils.rs:58-69,107-134); segments are independent until the recursion tree
merges them, so they shard with no communication.  torch.distributed is used only
as plumbing (nccl on the GPU box, gloo in the CPU tests)."""
from typing import Callable, List, Optional, Sequence

import numpy as np


def shard_segments(num_segments: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment: segment i is proved by rank i % world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, num_segments, world_size))


def gather_proofs(local: Sequence[np.ndarray], local_ids: Sequence[int], num_segments: int, device=None) -> Optional[List[np.ndarray]]:
    """Gathers every rank's proof buffers (uint64 arrays) on rank 0, ordered by segment id.  Returns the
    list on rank 0 and None elsewhere.  Collective: all ranks must call it."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [None] * num_segments
        for i, p in zip(local_ids, local):
            out[i] = np.asarray(p, dtype=np.uint64)
        return out
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    # 1) lengths  2) padded payloads (all_gather keeps nccl happy: equal shapes on every rank)
    max_local = max(1, (num_segments + world - 1) // world)
    lens = torch.zeros(max_local, dtype=torch.int64, device=dev)
    for k, p in enumerate(local):
        lens[k] = int(np.asarray(p).size)
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens)
    width = int(max(int(l.max()) for l in all_lens))
    payload = torch.zeros((max_local, max(width, 1)), dtype=torch.int64, device=dev)
    for k, p in enumerate(local):
        a = np.ascontiguousarray(p, dtype=np.uint64).view(np.int64)
        payload[k, : a.size] = torch.from_numpy(a).to(dev)
    gathered = [torch.zeros_like(payload) for _ in range(world)] if rank == 0 else None
    if dist.get_backend() == "nccl":
        # gather is available on nccl since torch 1.11; fall back to all_gather otherwise
        try:
            dist.gather(payload, gathered, dst=0)
        except RuntimeError:
            tmp = [torch.zeros_like(payload) for _ in range(world)]
            dist.all_gather(tmp, payload)
            gathered = tmp if rank == 0 else None
    else:
        dist.gather(payload, gathered, dst=0)
    if rank != 0:
        return None
    out: List[Optional[np.ndarray]] = [None] * num_segments
    for r in range(world):
        ids = shard_segments(num_segments, r, world)
        for k, seg in enumerate(ids):
            n = int(all_lens[r][k])
            out[seg] = gathered[r][k, :n].cpu().numpy().view(np.uint64).copy()
    return out


def prove_segments(prove: Callable[[int], np.ndarray], num_segments: int, rank: int = 0, world_size: int = 1, device=None):
    """Proves segments [0, num_segments) across `world_size` ranks: `prove(i)` returns the proof buffer of
    segment i (on the GPU box: zkm_b200_prove_with_traces on this rank's GPU).  Rank 0 gets all proofs."""
    ids = shard_segments(num_segments, rank, world_size)
    local = [prove(i) for i in ids]
    return gather_proofs(local, ids, num_segments, device=device)
