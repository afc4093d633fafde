"""PPO learner behind `train()` - the half of training/train.py:135-161,242-263 that brax `ppo.train` provides
(formulas are [UPSTREAM-RECALL] of brax 0.12 `training/agents/ppo/{train,losses}.py`).

Per training step: `batch_size * num_minibatches // num_envs` unrolls of `unroll_length` control steps are collected by
the native rollout (tcgen05 policy kernel -> fused env step -> transition record, rollout.py), the running
observation statistics are updated, then `num_updates_per_batch` epochs x `num_minibatches` SGD steps of the clipped
PPO loss (GAE(lambda), value loss 0.25 MSE, entropy bonus) with Adam + global-norm clipping run on the collected
segments. The SGD step uses torch autograd over plain matmuls (library GEMMs - the learner is a "next" row, not the
hot path); it is captured in a CUDA graph so the 128 SGD steps of a training step are 128 graph launches.

Multi-GPU: every rank owns its env shard (sharding.py); gradients are averaged with one flat NCCL all-reduce per SGD
step, observation statistics and (optionally) advantage moments with `sharding.allreduce_moments`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import sharding
from .policy import PolicyNet
from .rollout import RolloutCollector


@dataclass
class PPOConfig:                      # names and defaults of training/train.py:135-161 + argparse defaults :285-296
    num_timesteps: int = 1
    episode_length: int = 1000
    unroll_length: int = 20
    num_minibatches: int = 32
    num_updates_per_batch: int = 4
    discounting: float = 0.97
    learning_rate: float = 3e-4
    entropy_cost: float = 1e-2
    num_envs: int = 4096
    batch_size: int = 256
    max_grad_norm: float = 1.0
    reward_scaling: float = 1.0
    gae_lambda: float = 0.95
    clipping_epsilon: float = 0.3
    normalize_observations: bool = True
    normalize_advantage: bool = True
    global_advantage_norm: bool = False   # north-star variant: moments all-reduced over every rank's minibatch
    policy_hidden_layer_sizes: Sequence[int] = (512, 256, 128)
    value_hidden_layer_sizes: Sequence[int] = (512, 256, 128)
    seed: int = 0
    use_cuda_graph: bool = True
    # learner GEMM precision: "highest" = fp32 like the reference (jax_default_matmul_precision=highest, train.py:94),
    # "high" = TF32 tensor cores (fp32 storage and accumulation, 10-bit mantissa products)
    matmul_precision: str = "highest"


# ----------------------------------------------------------------------------------------------------------------------
# pure functions (tested on CPU against numpy restatements)
# ----------------------------------------------------------------------------------------------------------------------
def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_: float, discount: float):
    """brax `compute_gae`: all inputs time-major [T, B]; returns (value targets vs, advantages), both detached."""
    import torch
    trunc_mask = 1.0 - truncation
    values_tp1 = torch.cat([values[1:], bootstrap_value[None]], 0)
    deltas = (rewards + discount * (1.0 - termination) * values_tp1 - values) * trunc_mask
    acc = torch.zeros_like(bootstrap_value)
    vs_minus_v = []
    for t in range(rewards.shape[0] - 1, -1, -1):
        acc = deltas[t] + discount * (1.0 - termination[t]) * trunc_mask[t] * lambda_ * acc
        vs_minus_v.append(acc)
    vs = torch.stack(vs_minus_v[::-1], 0) + values
    vs_tp1 = torch.cat([vs[1:], bootstrap_value[None]], 0)
    adv = (rewards + discount * (1.0 - termination) * vs_tp1 - values) * trunc_mask
    return vs.detach(), adv.detach()


def mlp(x, kernels, biases):
    import torch
    for i, (k, b) in enumerate(zip(kernels, biases)):
        x = x @ k + b
        if i + 1 < len(kernels):
            x = torch.nn.functional.silu(x)
    return x


def tanh_normal_log_prob(logits, raw_action, min_std: float = 0.001):
    """brax NormalTanhDistribution.log_prob of the PRE-tanh action, summed over action dims."""
    import torch
    loc, scale_raw = logits.chunk(2, -1)
    scale = torch.nn.functional.softplus(scale_raw) + min_std
    lp = -0.5 * ((raw_action - loc) / scale) ** 2 - torch.log(scale) - 0.5 * math.log(2 * math.pi)
    lp = lp - 2.0 * (math.log(2.0) - raw_action - torch.nn.functional.softplus(-2.0 * raw_action))
    return lp.sum(-1)


def tanh_normal_entropy(logits, eps, min_std: float = 0.001):
    """brax estimate: Normal entropy + log|det d tanh| at one sample `loc + scale * eps`, summed over action dims."""
    import torch
    loc, scale_raw = logits.chunk(2, -1)
    scale = torch.nn.functional.softplus(scale_raw) + min_std
    ent = 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(scale)
    sample = loc + scale * eps
    ent = ent + 2.0 * (math.log(2.0) - sample - torch.nn.functional.softplus(-2.0 * sample))
    return ent.sum(-1)


def ppo_loss(policy_params, value_params, batch: Dict, cfg: PPOConfig, moments_fn: Optional[Callable] = None):
    """batch: time-major [T, B, ...] tensors with NORMALISED observations `obs`, `obs_priv` ([T + 1, B, .]) plus
    raw_action, log_prob, reward, discount, truncation ([T, B]) and entropy noise `eps` [T, B, A]."""
    import torch
    pk, pb = policy_params
    vk, vb = value_params
    T = batch["reward"].shape[0]
    logits = mlp(batch["obs"][:T], pk, pb)
    baseline_all = mlp(batch["obs_priv"], vk, vb).squeeze(-1)           # [T + 1, B]
    baseline, bootstrap = baseline_all[:T], baseline_all[T]
    rewards = batch["reward"] * cfg.reward_scaling
    truncation = batch["truncation"]
    termination = (1.0 - batch["discount"]) * (1.0 - truncation)
    target_lp = tanh_normal_log_prob(logits, batch["raw_action"])
    vs, adv = compute_gae(truncation, termination, rewards, baseline.detach(), bootstrap.detach(), cfg.gae_lambda, cfg.discounting)
    if cfg.normalize_advantage:
        if moments_fn is not None:
            mean, std = moments_fn(adv)
        else:
            mean, std = adv.mean(), adv.std(unbiased=False)
        adv = (adv - mean) / (std + 1e-8)
    rho = torch.exp(target_lp - batch["log_prob"])
    s1 = rho * adv
    s2 = torch.clamp(rho, 1.0 - cfg.clipping_epsilon, 1.0 + cfg.clipping_epsilon) * adv
    policy_loss = -torch.minimum(s1, s2).mean()
    v_loss = ((vs - baseline) ** 2).mean() * 0.5 * 0.5
    entropy = tanh_normal_entropy(logits, batch["eps"]).mean()
    total = policy_loss + v_loss - cfg.entropy_cost * entropy
    return total, {"total_loss": total.detach(), "policy_loss": policy_loss.detach(), "v_loss": v_loss.detach(), "entropy": entropy.detach()}


class RunningStats:
    """brax `running_statistics`: count, mean, summed_variance -> std clipped to [1e-6, 1e6]; updated with the
    observations of every collected unroll, summed over all ranks."""

    def __init__(self, dim: int, device):
        import torch
        self.count = torch.zeros((), dtype=torch.float64, device=device)
        self.mean = torch.zeros(dim, dtype=torch.float64, device=device)
        self.summed_var = torch.zeros(dim, dtype=torch.float64, device=device)
        self.std = torch.ones(dim, dtype=torch.float64, device=device)

    def update(self, x, group=None):
        import torch
        n, bmean, bvar = sharding.allreduce_moments(x.reshape(-1, x.shape[-1]), group)
        new_count = self.count + n
        delta = bmean - self.mean
        self.mean = self.mean + delta * (n / new_count)
        # Chan et al. merge: equals brax's incremental update for a single batch
        self.summed_var = self.summed_var + bvar * n + delta * delta * (self.count * n / new_count)
        self.count = new_count
        self.std = torch.clamp(torch.sqrt(torch.clamp(self.summed_var / self.count, min=0.0)), 1e-6, 1e6)

    def normalize(self, x):
        return ((x.double() - self.mean) / self.std).to(x.dtype)


def lecun_uniform_params(sizes: Sequence[int], gen, device):
    import torch
    ks, bs = [], []
    for i, o in zip(sizes[:-1], sizes[1:]):
        lim = math.sqrt(3.0 / i)
        ks.append(((torch.rand((i, o), generator=gen, device=device) * 2 - 1) * lim).requires_grad_())
        bs.append(torch.zeros(o, device=device, requires_grad=True))
    return ks, bs


# ----------------------------------------------------------------------------------------------------------------------
# trainer
# ----------------------------------------------------------------------------------------------------------------------
class PPOTrainer:
    def __init__(self, wenv, state, cfg: PPOConfig, group=None):
        """`wenv`: a reset `wrapper.TrainingEnv` holding this rank's env shard; `state` its current State."""
        import torch
        import torch.distributed as dist
        self.torch, self.cfg, self.wenv, self.group = torch, cfg, wenv, group
        if cfg.matmul_precision not in ("highest", "high"):
            raise ValueError("matmul_precision must be 'highest' or 'high'")
        torch.set_float32_matmul_precision(cfg.matmul_precision)
        self.env = wenv.unwrapped
        self.abi = self.env._abi
        self.dev = self.abi.torch_device
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        n_local = self.abi.N
        if (cfg.batch_size * cfg.num_minibatches) % (n_local * self.world) != 0:
            raise ValueError("batch_size * num_minibatches must be a multiple of num_envs (brax constraint); "
                             f"got {cfg.batch_size} * {cfg.num_minibatches} vs {n_local * self.world}")
        self.unrolls_per_step = cfg.batch_size * cfg.num_minibatches // (n_local * self.world)
        self.segments = self.unrolls_per_step * n_local                      # local trajectory segments per training step
        self.mb = self.segments // cfg.num_minibatches                      # local segments per minibatch
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(cfg.seed)                                           # same init on every rank
        nobs, npriv = self.abi.nobs, self.abi.npriv                          # 171 / 215, or 162 / 206 for the baseline task
        self.policy_params = lecun_uniform_params((nobs, *cfg.policy_hidden_layer_sizes, 24), gen, self.dev)
        self.value_params = lecun_uniform_params((npriv, *cfg.value_hidden_layer_sizes, 1), gen, self.dev)
        self.params = [*self.policy_params[0], *self.policy_params[1], *self.value_params[0], *self.value_params[1]]
        self.opt = torch.optim.Adam(self.params, lr=cfg.learning_rate, eps=1e-8, capturable=cfg.use_cuda_graph, foreach=True)
        self.norm_state, self.norm_priv = RunningStats(nobs, self.dev), RunningStats(npriv, self.dev)
        self.net = PolicyNet((nobs, *cfg.policy_hidden_layer_sizes, 24), device=self.abi.device)
        self.collector = RolloutCollector(wenv, self.net, unroll_length=cfg.unroll_length, seed=cfg.seed * 7919 + self.rank)
        self.state = state
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(cfg.seed * 31 + 1 + self.rank)
        self.env_steps = 0
        self._graph = None
        self._static: Dict = {}
        self.metrics: Dict = {}
        self._sync_policy()

    # -- policy kernel <- learner parameters ---------------------------------------------------------------------------------
    def _sync_policy(self):
        ks, bs = self.policy_params
        if self.cfg.normalize_observations and float(self.norm_state.count) > 0:
            self.net.set_params(ks, bs, self.norm_state.mean.float(), self.norm_state.std.float())
        else:
            self.net.set_params(ks, bs)

    # -- one SGD step (optionally replayed from a CUDA graph) ----------------------------------------------------------------
    def _moments(self, adv):
        if self.cfg.global_advantage_norm and self.world > 1:
            _, mean, var = sharding.allreduce_moments(adv.reshape(-1, 1), self.group)
            return mean[0].to(adv.dtype), var[0].sqrt().to(adv.dtype)
        return adv.mean(), adv.std(unbiased=False)

    def _sgd_body(self, batch):
        torch = self.torch
        loss, m = ppo_loss(self.policy_params, self.value_params, batch, self.cfg, self._moments)
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        if self.world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in self.params])
            torch.distributed.all_reduce(flat, group=self.group)
            flat /= self.world
            off = 0
            for p in self.params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        if self.cfg.max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.cfg.max_grad_norm, foreach=True)
        self.opt.step()
        return m

    def _sgd_step(self, batch):
        torch = self.torch
        graphable = self.cfg.use_cuda_graph and self.world == 1      # NCCL inside a captured graph is avoided here
        if not graphable:
            return self._sgd_body(batch)
        if self._graph is None:
            self._static = {k: v.clone() for k, v in batch.items()}
            s = torch.cuda.Stream(self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(3):                                   # warm-up outside capture (allocator, Adam state)
                    self._sgd_body(self._static)
            torch.cuda.current_stream(self.dev).wait_stream(s)
            # no garbage collection while capturing: a collected env / policy handle would run pgtt_*_destroy
            # (cudaDeviceSynchronize + cudaFree), which is illegal inside a capture
            import gc
            gc.collect()
            torch.cuda.synchronize(self.dev)
            gc.disable()
            try:
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                    self._static_metrics = self._sgd_body(self._static)
            finally:
                gc.enable()
            # the warm-up steps changed the parameters: acceptable for training (three extra SGD steps on the first
            # minibatch), callers that need exact step counts construct the trainer with use_cuda_graph=False
        for k, v in batch.items():
            self._static[k].copy_(v)
        self._graph.replay()
        return self._static_metrics

    # -- one training step ----------------------------------------------------------------------------------------------------
    def training_step(self) -> Dict:
        torch, cfg = self.torch, self.cfg
        T = cfg.unroll_length
        segs = []
        for _ in range(self.unrolls_per_step):
            self.state, ro = self.collector.collect()
            segs.append({"obs": ro.obs_state.clone(), "obs_priv": ro.obs_privileged.clone(), "raw_action": ro.raw_action.clone(),
                         "log_prob": ro.log_prob.clone(), "reward": ro.reward.clone(), "discount": ro.discount.clone(),
                         "truncation": ro.truncation.clone()})
            self.env_steps += T * self.abi.N * self.world
        data = {k: torch.cat([s[k] for s in segs], 1) for k in segs[0]}            # [T(+1), segments, ...]
        if cfg.normalize_observations:
            self.norm_state.update(data["obs"][:T], self.group)
            self.norm_priv.update(data["obs_priv"][:T], self.group)
            data["obs"] = self.norm_state.normalize(data["obs"])
            data["obs_priv"] = self.norm_priv.normalize(data["obs_priv"])
        last = {}
        for _ in range(cfg.num_updates_per_batch):
            perm = torch.randperm(self.segments, generator=self.gen, device=self.dev)
            for i in range(cfg.num_minibatches):
                idx = perm[i * self.mb:(i + 1) * self.mb]
                batch = {k: v.index_select(1, idx) for k, v in data.items()}
                batch["eps"] = torch.randn((T, self.mb, 12), generator=self.gen, device=self.dev)
                last = self._sgd_step(batch)
        self._sync_policy()
        self.metrics = {k: float(v) for k, v in last.items()}
        self.metrics["reward_per_step"] = float(data["reward"].mean())
        self.metrics["episode_done_rate"] = float((1.0 - data["discount"]).mean())
        self.metrics["env_steps"] = self.env_steps
        return self.metrics

    # -- brax-layout export (deploy/policy_net.py:6-33 reads it) ----------------------------------------------------------------
    def save(self, path):
        from . import policy_io
        det = lambda ts: [t.detach().cpu().numpy() for t in ts]
        policy_io.save_policy(path, self.norm_state.mean.float().cpu().numpy(), self.norm_state.std.float().cpu().numpy(),
                              (det(self.policy_params[0]), det(self.policy_params[1])), (det(self.value_params[0]), det(self.value_params[1])),
                              count=float(self.norm_state.count))


def train(environment, wrap_env_fn, randomization_fn, rng_keys, cfg: PPOConfig, progress_fn: Optional[Callable] = None,
          policy_params_fn: Optional[Callable] = None, num_training_steps: Optional[int] = None):
    """Call shape of `ppo.train(environment=..., wrap_env_fn=..., randomization_fn=..., progress_fn=..., ...)` in
    training/train.py:242-263. `rng_keys`: uint32[N_local, 2] per-env keys of this rank's shard."""
    import functools
    wenv = wrap_env_fn(environment, episode_length=cfg.episode_length, action_repeat=1,
                       randomization_fn=functools.partial(randomization_fn, rng=rng_keys) if randomization_fn is not None else None)
    state = wenv.reset(rng_keys)
    trainer = PPOTrainer(wenv, state, cfg)
    per_step = cfg.unroll_length * cfg.batch_size * cfg.num_minibatches
    steps = num_training_steps if num_training_steps is not None else max(1, math.ceil(cfg.num_timesteps / per_step))
    for it in range(steps):
        m = trainer.training_step()
        if progress_fn is not None:
            progress_fn(trainer.env_steps, m)
        if policy_params_fn is not None:
            policy_params_fn(trainer.env_steps, trainer)
    return trainer
