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


# ---------------------------------------------------------------------------------------------------------------------
# In-segment sharding (include/zkm_b200.h "In-segment sharding"): G = 2, 4 or 8 ranks prove ONE segment together.
# The ownership rules below are the host-side mirror of zkm_b200/csrc/shard.cuh, used to form the groups and -- in the CPU
# tests (gloo, the oracle as each rank's local prover) -- to check that the exchanged pieces reassemble the single-rank result.

def bitrev(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def _check_group(group: int):
    if group not in (1, 2, 4, 8):
        raise ValueError("in-segment sharding supports groups of 1, 2, 4 or 8 ranks")


def parts(group: int) -> int:
    """Ranks that share one coset: 1 up to 4 ranks; with 8 ranks two ranks run the same coset transform and hash half of its
    leaves each."""
    return group // 4 if group > 4 else 1


def owned_cosets(rank: int, group: int) -> List[int]:
    """LDE cosets j (natural LDE index m = 4 i + j) `rank` of a `group`-rank shard group computes."""
    _check_group(group)
    if group >= 4:
        return [rank * 4 // group]
    per = 4 // group
    return list(range(rank * per, (rank + 1) * per))


def rank_of(coset: int, part: int, group: int) -> int:
    return coset * (group // 4) + part if group > 4 else coset * group // 4


def coset_owner(j: int, group: int) -> int:
    return rank_of(j, 0, group)


def leaf_owner(leaf: int, log_leaves: int, group: int) -> int:
    """Rank holding leaf `leaf` (bit-reversed LDE order) and its path: the top two leaf bits are the bit-reversed coset id,
    the next bit (8 ranks) the half of that coset's leaves."""
    _check_group(group)
    p = parts(group)
    part = (leaf >> (log_leaves - 2 - (p.bit_length() - 1))) & (p - 1)
    return rank_of(bitrev(leaf >> (log_leaves - 2), 2), part, group)


def owned_segments(rank: int, group: int) -> List[int]:
    """Leaf segments (of the 4 * parts contiguous ones) a rank hashes and builds subtrees over."""
    p = parts(group)
    return [bitrev(j, 2) * p + rank % p for j in owned_cosets(rank, group)]


def owned_cap_entries(rank: int, group: int, cap_height: int = 4) -> List[int]:
    """Cap entries (subtree roots over contiguous leaf blocks) a rank computes: those of its leaf segments."""
    per_seg = (1 << cap_height) // (4 * parts(group))
    out = []
    for sg in owned_segments(rank, group):
        out.extend(range(sg * per_seg, (sg + 1) * per_seg))
    return out


def assemble_cap(pieces: Sequence[np.ndarray], group: int, cap_height: int = 4) -> np.ndarray:
    """The all-gather of a sharded commitment: pieces[r] = rank r's owned cap entries (in owned_cap_entries order, 4 words
    each) -> the full cap."""
    cap = np.zeros(((1 << cap_height), 4), dtype=np.uint64)
    for r, piece in enumerate(pieces):
        ids = owned_cap_entries(r, group, cap_height)
        cap[ids] = np.asarray(piece, dtype=np.uint64).reshape(len(ids), 4)
    return cap


def shard_group_init(lib, group: int = 0):
    """Forms in-segment shard groups of `group` consecutive ranks (default: the whole world, up to 8) out of the torch.distributed world
    and binds this process's library context to its group: rank 0 of each group creates the NCCL unique id
    (zkm_b200_shard_unique_id) and it reaches the other members through a torch.distributed broadcast.  Returns
    (group_index, rank_in_group, group_size).  Collective over the whole world."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import lib as zl
    world, rank = dist.get_world_size(), dist.get_rank()
    g = group or min(world, 8)
    if world % g or g not in (1, 2, 4, 8):
        raise ValueError(f"cannot split {world} ranks into shard groups of {g}")
    gi, ri = rank // g, rank % g
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    ids = torch.zeros((world // g, 128), dtype=torch.uint8, device=dev)
    if ri == 0 and g > 1:
        buf = C.create_string_buffer(128)
        err = C.c_void_p()
        zl.check(lib, lib.zkm_b200_shard_unique_id(buf, C.byref(err)), err)
        ids[gi] = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
    dist.all_reduce(ids, op=dist.ReduceOp.SUM)          # every row is written by exactly one rank
    err = C.c_void_p()
    zl.check(lib, lib.zkm_b200_shard_init(ri, g, bytes(ids[gi].cpu().numpy().tobytes()), C.byref(err)), err)
    return gi, ri, g


def shard_group_shutdown(lib):
    import ctypes as C
    from . import lib as zl
    err = C.c_void_p()
    zl.check(lib, lib.zkm_b200_shard_shutdown(C.byref(err)), err)
