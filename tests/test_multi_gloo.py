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


def _worker_balanced(rank, world, port, M, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from lwfbrook90.jl_b200.parallel import allgather_by_members, allgather_rows, balanced_deal, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # probe costs of the contiguous shard, all-gathered in rank order = global member order
    rng = np.random.default_rng(5)
    cost_all = rng.pareto(1.5, M) + 1.0                      # heavy tail, the same on every rank
    lo_, hi = shard_range(M, rank, world)
    blocks = allgather_rows(torch.from_numpy(cost_all[lo_:hi]).reshape(-1, 1), rank, world)
    cost = torch.cat(blocks, dim=0)[:, 0].numpy()
    deal = balanced_deal(cost, world)
    mine = deal[rank]
    idx = torch.from_numpy(mine.astype(np.float64))
    local = torch.stack([idx, idx * 2, -idx, idx + 0.5], dim=1)      # stand-in for the scores of the dealt members
    full = allgather_by_members(local, mine, M, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, full.numpy(), cost, [d.tolist() for d in deal]))


@pytest.mark.parametrize("M", [11, 64])
def test_two_rank_cost_balanced_deal_and_gather(M):
    """SURVEY 8e in the one-process-per-GPU path: probe costs all-gathered, members dealt by cost in a snake (the same deal on
    every rank, sizes within one, cost shares within one member's cost of each other), scores gathered back into member order."""
    import torch.multiprocessing as mp
    from lwfbrook90.jl_b200.parallel import shard_range
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + M
    ps = [ctx.Process(target=_worker_balanced, args=(r, 2, port, M, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = {r: rest for r, *rest in (q.get(timeout=120) for _ in ps)}
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(M, dtype=float)
    expect = np.stack([idx, idx * 2, -idx, idx + 0.5], axis=1)
    full0, cost, deal0 = res[0]
    full1, _, deal1 = res[1]
    assert np.array_equal(full0, expect) and np.array_equal(full1, expect)
    assert deal0 == deal1 and sorted(deal0[0] + deal0[1]) == list(range(M)) and abs(len(deal0[0]) - len(deal0[1])) <= 1
    share = [cost[d].sum() for d in deal0]
    assert max(share) - min(share) <= cost.max() + 1e-12          # the snake deal's guarantee, whatever the order of arrival
    # ... and with the expensive members arriving together (a sorted parameter sweep) contiguous shards are far worse
    from lwfbrook90.jl_b200.parallel import balanced_deal
    srt = np.sort(cost)[::-1]
    sh2 = [srt[d].sum() for d in balanced_deal(srt, 2)]
    co2 = [srt[slice(*shard_range(M, r, 2))].sum() for r in range(2)]
    assert max(sh2) / min(sh2) < 1.2 < max(co2) / min(co2)
