"""Multi-GPU plumbing: one process per GPU, no collective during the integration; the only exchange is an all-gather
of the per-member score block.  Members are sharded contiguously (shard_range) or, given a cost estimate per
member - the RHS counts of a short probe run -, sorted by cost and dealt to the ranks in a snake (balanced_deal: the same
deal lwfb_multi_set_members makes inside the library), so that every rank gets the same share of expensive members."""
from __future__ import annotations

import numpy as np


def shard_range(n_members: int, rank: int, world: int):
    """Contiguous block of members owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_members), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_scores(local_scores, n_members: int, rank: int, world: int):
    """All-gather of the [m_local, S] score blocks into [n_members, S] on every rank.
    `local_scores` is a torch tensor (CUDA with the nccl backend, CPU with gloo)."""
    import torch
    import torch.distributed as dist
    S = local_scores.shape[1]
    sizes = [shard_range(n_members, r, world) for r in range(world)]
    mx = max(b - a for a, b in sizes)
    pad = torch.full((mx, S), float("nan"), dtype=local_scores.dtype, device=local_scores.device)
    pad[: local_scores.shape[0]] = local_scores
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def balanced_deal(cost, world: int):
    """Members sorted by cost (most expensive first, ties by index) and dealt to the ranks in a snake:
    0 1 .. w-1 w-1 .. 1 0 0 1 ...  Returns one ascending array of member ids per rank; sizes differ by at most one and
    every rank holds the same share of the expensive members.  Deterministic: every rank computes the same deal."""
    cost = np.asarray(cost, dtype=np.float64)
    order = np.lexsort((np.arange(cost.size), -np.nan_to_num(cost, nan=np.inf)))
    pos = np.arange(cost.size)
    lap, k = divmod(pos, world)
    owner = np.where(lap % 2 == 0, k, world - 1 - k)
    return [np.sort(order[owner == r]) for r in range(world)]


def allgather_rows(local, rank: int, world: int):
    """All-gather of per-rank blocks [n_r, S] (any n_r) into the list of the world's blocks, padded internally."""
    import torch
    import torch.distributed as dist
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    mx = max(ns)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[:k] for o, k in zip(out, ns)]


def allgather_by_members(local_scores, member_ids, n_members: int, rank: int, world: int):
    """All-gather of the [m_local, S] score blocks of ranks that hold ARBITRARY members (member_ids: the global ids of the
    local rows, e.g. from balanced_deal) into [n_members, S] in global member order, on every rank."""
    import torch
    ids = torch.as_tensor(np.asarray(member_ids, dtype=np.int64), device=local_scores.device)
    blocks = allgather_rows(local_scores, rank, world)
    id_blocks = allgather_rows(ids.reshape(-1, 1), rank, world)
    full = torch.full((n_members, local_scores.shape[1]), float("nan"), dtype=local_scores.dtype, device=local_scores.device)
    for b, i in zip(blocks, id_blocks):
        full[i[:, 0]] = b
    return full


class DeviceArray:
    """Wraps a raw device pointer as a __cuda_array_interface__ object so torch can view it
    without a copy (used for the score block that stays on the GPU)."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}
