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


def _train_worker(rank, world, port, out):
    """ppo.train's host loop on two ranks (stub trainer / evaluator): only rank 0 owns an eval env and sees progress_fn's verdict."""
    import torch
    import torch.distributed as dist
    from phase_guided_terrain_traversal_b200 import ppo
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log = []

    class FakeTrainer:
        group, dev = None, torch.device("cpu")

        def __init__(self, wenv, state, cfg):
            self.env_steps, self.cfg, self.world, self.rank = 0, cfg, world, rank

        def training_step(self):
            self.env_steps += self.cfg.unroll_length * self.cfg.batch_size * self.cfg.num_minibatches
            log.append(("train", self.env_steps))
            return {}

    class FakeEvaluator:
        def __init__(self, eval_env, wrap_env_fn, randomization_fn, cfg, trainer, num_eval_envs, deterministic_eval, seed):
            assert rank == 0 and eval_env == "eval-env"
            self.trainer = trainer

        def run_evaluation(self, training_metrics=None):
            log.append(("eval", self.trainer.env_steps))
            return {"eval/episode_reward": float(self.trainer.env_steps)}

    class FakeWrapped:
        def reset(self, keys):
            return "state"

    ppo.PPOTrainer, ppo.Evaluator = FakeTrainer, FakeEvaluator
    try:
        cfg = ppo.PPOConfig(num_envs=64, batch_size=8, num_minibatches=8, unroll_length=20)
        per = 20 * 8 * 8
        cfg.num_timesteps = 9 * per                      # 3 epochs of 3 training steps at num_evals = 4
        tr = ppo.train("env", lambda e, **kw: FakeWrapped(), None, None, cfg, progress_fn=lambda n, m: n >= 6 * per,
                       eval_env="eval-env" if rank == 0 else None, num_evals=4)
        out[rank] = (list(log), bool(tr.stopped_early))
    finally:
        dist.destroy_process_group()


def test_train_loop_stops_every_rank_on_rank0_verdict_gloo_world2():
    """`ppo.train` under two ranks: rank 0 alone evaluates (brax evaluates on process 0) and alone sees the convergence verdict of
    `progress_fn` (training/train.py:224-229); the verdict is broadcast, so both ranks take the evaluation path and leave the loop after the
    same training step - rank 1 without an eval env and without ever calling the evaluator."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_train_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    per = 20 * 8 * 8
    trains = [("train", k * per) for k in range(1, 7)]
    log0, stop0 = out[0]
    log1, stop1 = out[1]
    assert stop0 and stop1
    assert [e for e in log0 if e[0] == "train"] == trains and log1 == trains                          # both stop after the 6th training step
    assert [e for e in log0 if e[0] == "eval"] == [("eval", 0), ("eval", 3 * per), ("eval", 6 * per)]
