"""World-size-2 gloo test (CPU) of the N>1 host logic: observation partition by owning rank,
per-rank partial statistics + ONE all-reduce == single-rank statistics, halo source maps.  The
arithmetic is the oracle's (no GPU here); the partition/plumbing under test is the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dataassimbench_b200.sharded import ShardPlan
from oracle import etkf as o_etkf


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, k, n_y, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)                       # same data on every rank
    X = rng.standard_normal((k, N)) + 3.0
    idx = rng.choice(N, n_y, replace=False, shuffle=False)
    y = rng.standard_normal(n_y)
    plan = ShardPlan(N, world)
    lo, hi = plan.slab(rank)
    lidx, lvals, pos = plan.local_obs(rank, idx, y)
    C, g = o_etkf.contraction(X[:, lo:hi], lvals, lidx, np.ones(len(lidx)))
    t = torch.from_numpy(np.concatenate([C.ravel(), g]))
    dist.all_reduce(t)                                    # the one collective of a cycle
    counts = torch.tensor([len(lidx)])
    dist.all_reduce(counts)
    if rank == 0:
        q.put((t.numpy(), int(counts.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_partial_stats_allreduce_equals_global():
    N, k, n_y, world = 64, 6, 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, k, n_y, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, count = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    X = rng.standard_normal((k, N)) + 3.0
    idx = rng.choice(N, n_y, replace=False, shuffle=False)
    y = rng.standard_normal(n_y)
    C, g = o_etkf.contraction(X, y, idx, np.ones(n_y))
    assert count == n_y                                   # every observation owned exactly once
    assert np.allclose(stats[:k * k].reshape(k, k), C, rtol=1e-12, atol=1e-12)
    assert np.allclose(stats[k * k:], g, rtol=1e-12, atol=1e-12)


def test_shard_plan_partition_and_halos():
    plan = ShardPlan(96, 4)
    assert plan.n_loc == 24 and plan.slab(3) == (72, 96)
    idx = np.array([95, 0, 23, 24, 50, 71, 72])
    assert list(plan.owner(idx)) == [3, 0, 0, 1, 2, 2, 3]
    l, v, pos = plan.local_obs(2, idx, np.arange(7.0))
    assert list(l) == [2, 23] and list(v) == [4.0, 5.0] and list(pos) == [4, 5]
    own, col = plan.halo_sources(0, 40, 20)              # halo wider than a slab: wraps past the neighbour
    assert len(own) == 24 + 60
    assert own[0] == 2 and col[0] == (96 - 40) % 24       # e = -40 -> global 56 -> rank 2, column 8
    assert own[40] == 0 and col[40] == 0                  # e = 0
    assert own[-1] == 1 and col[-1] == 19                 # e = 43 -> global 43 -> rank 1, column 19
    with pytest.raises(ValueError):
        ShardPlan(100, 8)
