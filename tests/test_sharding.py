"""Multi-rank host logic on CPU: index sharding, key sharding and the moment all-reduce over a
world-size-2 gloo group (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest


def test_shard_partition_and_keys():
    from phase_guided_terrain_traversal_b200 import prng, sharding
    for n, w in [(4096, 1), (4096, 8), (32768, 8), (10, 4), (3, 8)]:
        spans = [sharding.shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
        keys = np.concatenate([sharding.shard_keys(7, n, r, w) for r in range(w)])
        assert np.array_equal(keys, prng.env_keys(7, n))          # a sharded run steps the same envs as a 1-rank run
    with pytest.raises(ValueError):
        sharding.shard(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    import torch
    import torch.distributed as dist
    from phase_guided_terrain_traversal_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.from_numpy(np.random.default_rng(0).normal(2.0, 3.0, size=(n_total, 5)).astype(np.float32))
        a, b = sharding.shard(n_total, rank, world)
        cnt, mean, var = sharding.allreduce_moments(full[a:b])
        adv = sharding.normalize_advantages(full[a:b, 0])
        # max-over-ranks timing reduction as bench.py does it
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[rank] = (float(cnt), mean.numpy(), var.numpy(), adv.numpy(), float(t))
    finally:
        dist.destroy_process_group()


def test_allreduce_moments_gloo_world2():
    import torch.multiprocessing as mp
    n_total, world = 1001, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    full = np.random.default_rng(0).normal(2.0, 3.0, size=(n_total, 5)).astype(np.float32).astype(np.float64)
    advs = []
    for r in range(world):
        cnt, mean, var, adv, tmax = out[r]
        assert cnt == n_total and tmax == world
        assert np.allclose(mean, full.mean(0), rtol=1e-10) and np.allclose(var, full.var(0), rtol=1e-8)
        advs.append(adv)
    adv = np.concatenate(advs)
    ref = (full[:, 0] - full[:, 0].mean()) / (full[:, 0].std() + 1e-8)
    assert np.allclose(adv, ref, atol=1e-5)


def test_allreduce_moments_single_process():
    import torch
    from phase_guided_terrain_traversal_b200 import sharding
    x = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    cnt, mean, var = sharding.allreduce_moments(x)
    assert float(cnt) == 6 and torch.allclose(mean, x.double().mean(0)) and torch.allclose(var, x.double().var(0, unbiased=False))
