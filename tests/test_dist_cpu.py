"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: env sharding, the flat gradient all-reduce of the
PPO update, and the (sum, sum of squares, count) merge used for advantage normalisation across ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_without_overlap():
    sys.path.insert(0, ROOT)
    from ppo_rl_satellite_b200.rollout import shard_bounds
    for n, w in ((1048576, 8), (1048576, 2), (10, 3), (7, 8), (65536, 4)):
        edges = [shard_bounds(n, w, r) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(1048576, 8, 3) == (393216, 524288)       # config 4: 131 072 envs per GPU


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "ppo-rl-satellite_b200", "dropin"))
    try:
        from ppo_continuous import allreduce_grads_
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(18, 32), torch.nn.Tanh(), torch.nn.Linear(32, 3))
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(64, 18, generator=g)
        net(x).pow(2).mean().backward()
        local = [p.grad.clone() for p in net.parameters()]
        allreduce_grads_(net)
        gathered = [[torch.zeros_like(t) for _ in range(world)] for t in local]
        for lst, t in zip(gathered, local):
            dist.all_gather(lst, t)
        ok = all(torch.allclose(p.grad, torch.stack(lst).mean(0), atol=1e-7) for p, lst in zip(net.parameters(), gathered))
        # advantage-normalisation moments: all-reduced (sum, sumsq, count) give the global mean / unbiased std
        adv = torch.randn(1000 + 10 * rank, generator=g, dtype=torch.float64) * (1 + rank) + rank
        sums = torch.tensor([adv.sum(), (adv * adv).sum(), float(adv.numel())], dtype=torch.float64)
        dist.all_reduce(sums)
        alls = [torch.zeros(1000 + 10 * r, dtype=torch.float64) for r in range(world)]
        # gather the ragged shards through an object list
        objs = [None] * world
        dist.all_gather_object(objs, adv)
        full = torch.cat(objs)
        n = sums[2]
        mean = sums[0] / n
        var = (sums[1] - n * mean * mean) / (n - 1)
        ok = ok and abs(mean - full.mean()) < 1e-12 and abs(var.sqrt() - full.std()) < 1e-10
        # per-rank batch sizes must agree before a PPO update (one gradient exchange per minibatch: unequal ceil(B / mb)
        # would dead-lock): equal sizes pass, unequal sizes raise on EVERY rank (collective check)
        from ppo_continuous import PPO_continuous
        PPO_continuous._require_equal_batches(4096, None, "cpu")
        try:
            PPO_continuous._require_equal_batches(4096 + rank, None, "cpu")
            ok = False
        except ValueError as e:
            ok = ok and "unequal per-rank batches" in str(e)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_and_moment_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
