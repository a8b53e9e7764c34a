"""N>1 host logic on CPU: two gloo ranks shard an ensemble and all-gather their score blocks."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, M, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from lwfbrook90.jl_b200.parallel import allgather_scores, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo_, hi = shard_range(M, rank, world)
    idx = torch.arange(lo_, hi, dtype=torch.float64)
    local = torch.stack([idx, idx * 2, -idx, idx + 0.5], dim=1)     # stand-in for [m_local, 4] scores
    full = allgather_scores(local, M, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, full.numpy()))


@pytest.mark.parametrize("M", [7, 10])
def test_two_rank_score_allgather(M):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + M
    ps = [ctx.Process(target=_worker, args=(r, 2, port, M, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(M, dtype=float)
    expect = np.stack([idx, idx * 2, -idx, idx + 0.5], axis=1)
    for r in range(2):
        assert np.array_equal(res[r], expect)
