"""Success-count evaluation - what training/evaluate.py:188-259 gets out of brax's evaluator:
`num_eval_envs` (1000) environments run one full episode with the command ranges of training/evaluate.py:127-129
(`eval_overrides`: +-[0.4, 0.4, 0.7]) and an environment counts as a success when its episode ends without a termination,
i.e. `abs(final termination reward) < 0.5` (evaluate.py:220-222). The reference evaluates through brax `ppo.train`, whose
default is `deterministic_eval=False` - SAMPLED actions - so that is the default here; `--deterministic` evaluates
`tanh(loc)` as deploy/policy_net.py:64 does. Also reports the evaluator's episode reward mean / std and episode length.

    python -m phase_guided_terrain_traversal_b200.evaluate --policy /path/to/policy177 --task_name stairs --terrain_file level07
"""
from __future__ import annotations

import argparse
import functools
from typing import Dict, Optional

import numpy as np


def evaluate(wenv, policy_net, episode_length: int = 1000, unroll_length: int = 20, seed: int = 0, collect_obs_stats: bool = False,
             deterministic: bool = True, per_metric: bool = False) -> Dict:
    """`wenv`: a freshly reset `wrapper.TrainingEnv` (its episode_length must equal `episode_length`); `policy_net`: a
    `policy.PolicyNet` with parameters. Runs ceil(episode_length / unroll_length) unrolls on the device (`deterministic`:
    tanh(loc) instead of sampled actions). `per_metric`: also sum every entry of `State.metrics` over each env's first
    episode, which is what brax's EvalWrapper keeps and `training/train.py:214-216` reads (`eval/episode_reward/tracking_lin_vel`);
    the per-step metrics only live in the env's `metrics` buffer, so this mode steps with unroll length 1."""
    import torch
    from .rollout import RolloutCollector
    if per_metric:
        unroll_length = 1
    col = RolloutCollector(wenv, policy_net, unroll_length=unroll_length, seed=seed)
    n, dev = col.abi.N, col.abi.torch_device
    alive = torch.ones(n, device=dev)            # still inside the first episode
    terminated = torch.zeros(n, device=dev)
    ep_reward = torch.zeros(n, device=dev)
    ep_len = torch.zeros(n, device=dev)
    metrics = col.abi.buf["metrics"]                  # [N, len(METRIC_KEYS)]: the step's State.metrics
    ep_metrics = torch.zeros_like(metrics) if per_metric else None
    obs_sum = obs_sq = priv_sum = priv_sq = None
    obs_cnt = 0
    steps = 0
    while steps < episode_length:
        _, ro = col.collect(deterministic=deterministic)
        done = 1.0 - ro.discount                                     # [T, N]
        term = done * (1.0 - ro.truncation)
        if per_metric:                                               # (one step per unroll: `metrics` is this step's)
            ep_metrics += alive[:, None] * metrics
        for t in range(ro.reward.shape[0]):
            ep_reward += alive * ro.reward[t]
            ep_len += alive
            terminated = torch.maximum(terminated, alive * term[t])
            alive = alive * (1.0 - done[t])
        if collect_obs_stats:
            o = ro.obs_state[:-1].reshape(-1, ro.obs_state.shape[-1]).double()
            obs_sum = o.sum(0) if obs_sum is None else obs_sum + o.sum(0)
            obs_sq = (o * o).sum(0) if obs_sq is None else obs_sq + (o * o).sum(0)
            obs_cnt += o.shape[0]
            pv = ro.obs_privileged[:-1].reshape(-1, ro.obs_privileged.shape[-1]).double()
            priv_sum = pv.sum(0) if priv_sum is None else priv_sum + pv.sum(0)
            priv_sq = (pv * pv).sum(0) if priv_sq is None else priv_sq + (pv * pv).sum(0)
        steps += ro.reward.shape[0]
    success = (terminated < 0.5)
    out = {"num_eval_envs": n, "success_count": int(success.sum().item()), "success_rate": float(success.float().mean().item()),
           "episode_reward": float(ep_reward.mean().item()), "episode_reward_std": float(ep_reward.std(unbiased=False).item()),
           "avg_episode_length": float(ep_len.mean().item())}
    if per_metric:
        from .go2.base import METRIC_KEYS
        mean, std = ep_metrics.mean(0).tolist(), ep_metrics.std(0, unbiased=False).tolist()
        out["episode_metrics"] = {k: mean[i] for i, k in enumerate(METRIC_KEYS)}
        out["episode_metrics_std"] = {k: std[i] for i, k in enumerate(METRIC_KEYS)}
    if collect_obs_stats:
        mean = obs_sum / obs_cnt
        out["obs_mean"] = mean.cpu().numpy()
        out["obs_std"] = (obs_sq / obs_cnt - mean * mean).clamp_min(0).sqrt().cpu().numpy()
        pmean = priv_sum / obs_cnt
        out["priv_mean"] = pmean.cpu().numpy()
        out["priv_std"] = (priv_sq / obs_cnt - pmean * pmean).clamp_min(0).sqrt().cpu().numpy()
    return out


def main():
    p = argparse.ArgumentParser(description="Count successful episodes of a saved policy (training/evaluate.py)")
    p.add_argument("--policy", required=True, help="brax-layout policy pickle (policy_folder/policyNNN or train --out)")
    p.add_argument("--method", default="pgtt")
    p.add_argument("--task_name", default="stairs")
    p.add_argument("--terrain_file", default="level07")
    p.add_argument("--num_eval_envs", type=int, default=1000)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--deterministic", action="store_true", help="evaluate tanh(loc) (deploy/policy_net.py) instead of sampled actions (brax default)")
    p.add_argument("--training_ranges", action="store_true", help="command ranges of training/train.py:127-129 instead of evaluate.py:127-129")
    a = p.parse_args()
    from . import policy_io, prng, terrain, wrapper
    from .go2 import joystick, joystick_pgtt, randomize, randomize_simple
    from .go2.configs import baseline_config, default_config, eval_overrides, training_overrides
    from .policy import PolicyNet
    joy, cfg_fn = (joystick_pgtt, default_config) if a.method == "pgtt" else (joystick, baseline_config)
    cfg = (training_overrides if a.training_ranges else eval_overrides)(cfg_fn())
    env = joy.Joystick(task=a.task_name, config=cfg)
    keys = prng.env_keys(a.seed, a.num_eval_envs)
    rfn = functools.partial(randomize.domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain(a.terrain_file)) if a.task_name == "stairs" \
        else functools.partial(randomize_simple.domain_randomize, rng=keys)
    wenv = wrapper.wrap_for_brax_training(env, episode_length=cfg.episode_length, randomization_fn=rfn)
    wenv.reset(keys + np.uint32(1))
    d = policy_io.load_policy(a.policy)
    net = PolicyNet((d["policy"][0][0].shape[0], *[k.shape[1] for k in d["policy"][0]]))
    net.set_params(d["policy"][0], d["policy"][1], d["mean"], d["std"])
    r = evaluate(wenv, net, episode_length=cfg.episode_length, seed=a.seed, deterministic=a.deterministic)
    print({k: v for k, v in r.items() if not (k.startswith("obs_") or k.startswith("priv_"))})


if __name__ == "__main__":
    main()
