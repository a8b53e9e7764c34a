"""Multi-GPU plumbing: one process per GPU, members sharded contiguously, no collective during the
integration; the only exchange is an all-gather of the per-member score block."""
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


class DeviceArray:
    """Wraps a raw device pointer as a __cuda_array_interface__ object so torch can view it
    without a copy (used for the score block that stays on the GPU)."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}
