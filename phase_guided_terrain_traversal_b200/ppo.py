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

import contextlib
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
    fused_head: bool = True               # PPO loss head + its gradients from the hand-written kernel `pgtt_ppo_head`
    native_optimizer: bool = True         # global-norm clip + Adam over one flat parameter vector from `pgtt_adam_clip` (two launches)
    parallel_nets: bool = True            # value network (forward and backward) on a second CUDA stream beside the policy network
    # learner GEMM precision: "highest" = fp32 like the reference (jax_default_matmul_precision=highest, train.py:94),
    # "high" = TF32 tensor cores (fp32 storage and accumulation, 10-bit mantissa products)
    matmul_precision: str = "highest"
    # True: both MLPs, forward and backward, from the hand-written blocked split-bf16 tcgen05 GEMMs of csrc/pgtt_mlp.cu (fp32-grade
    # accuracy, one autograd node per network); "layers": the per-layer tcgen05 GEMMs of csrc/pgtt_learner.cu; False = torch GEMMs
    # (CPU tests, comparisons)
    native_mlp: object = True
    # the SGD step as a fixed sequence of this repo's kernels over static buffers, no autograd (needs native_mlp = True, fused_head, native_optimizer)
    native_step: bool = True


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


def compute_gae_native(truncation, discount, rewards, values_all, lambda_: float, discount_factor: float, reward_scaling: float):
    """Same quantities from the hand-written kernel `pgtt_gae` (one launch instead of ~120): CUDA float32, time-major;
    `values_all` is [T + 1, B] with the bootstrap value in its last row."""
    import ctypes as C
    import torch
    from . import _native as nat
    lib = nat.load_library()
    T, B = rewards.shape
    vs, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    args = [t.contiguous() for t in (truncation, discount, rewards, values_all)]
    stream = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    rc = lib.pgtt_gae(*(a.data_ptr() for a in args), T, B, lambda_, discount_factor, reward_scaling, vs.data_ptr(), adv.data_ptr(), stream)
    if rc:
        raise nat.PgttError(rc, lib.pgtt_policy_last_error().decode())
    return vs, adv


_LINEAR = None
_ONES: Dict = {}
_SCALARS = ("log_prob", "reward", "discount", "truncation")


def _linear():
    """Dense layer with a hand-arranged backward (profiles/r01c_learner_launches.csv, tools/learner_profile.py):
    * the bias gradient is a (1 x M) @ (M x N) GEMM: torch's column reduction takes 18 us per layer on [5120, 512], the GEMM ~4 us;
    * no input gradient is formed for a layer whose input does not need one (the first layer: a [5120, 512] @ [512, 171] GEMM);
    * the input may carry zero-padded trailing columns (171 -> 172, 215 -> 216): K is then a multiple of four and the
      weight-gradient GEMM takes the aligned tensor-core path (57 -> ~10 us); the kernel is padded with zero rows on the fly."""
    global _LINEAR
    if _LINEAR is None:
        import torch

        class Linear(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, k, b, aux):
                pad = x.shape[1] - k.shape[0]
                kp = torch.nn.functional.pad(k, (0, 0, 0, pad)) if pad else k
                ctx.save_for_backward(x, kp)
                ctx.rows, ctx.aux = k.shape[0], aux
                return torch.addmm(b, x, kp)

            @staticmethod
            def backward(ctx, g):
                x, kp = ctx.saved_tensors
                key = (g.shape[0], g.device, g.dtype)
                ones = _ONES.get(key)
                if ones is None:
                    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                        ones = torch.ones((1, g.shape[0]), device=g.device, dtype=g.dtype)     # (not cached: graph-pool memory)
                    else:
                        ones = _ONES[key] = torch.ones((1, g.shape[0]), device=g.device, dtype=g.dtype)
                        if g.is_cuda:      # shared by every stream from now on: make sure it is written before any of them reads it
                            torch.cuda.current_stream(g.device).synchronize()
                aux = ctx.aux
                if aux is not None:     # parameter gradients off the critical path: only the input gradient feeds the next layer down
                    aux.wait_stream(torch.cuda.current_stream(g.device))
                    g.record_stream(aux); x.record_stream(aux)
                with torch.cuda.stream(aux) if aux is not None else contextlib.nullcontext():
                    gk, gb = (x.t() @ g)[:ctx.rows], (ones @ g).reshape(-1)
                gx = g @ kp.t() if ctx.needs_input_grad[0] else None
                return gx, gk, gb, None
        _LINEAR = Linear
    return _LINEAR


_NATIVE_LINEAR = None


def _native_linear():
    """Dense layer (+ fused SiLU) on the hand-written tcgen05 GEMMs: `pgtt_linear_forward` / `_backward_input` /
    `_backward_params` / `pgtt_silu_backward` (include/pgtt_b200.h). The parameter gradients run on `aux` (off the critical
    path of the backward chain) exactly like the torch variant above."""
    global _NATIVE_LINEAR
    if _NATIVE_LINEAR is None:
        import ctypes as C
        import torch
        from . import _native as nat

        def chk(lib, rc):
            if rc:
                raise nat.PgttError(rc, lib.pgtt_learner_last_error().decode())

        class NativeLinear(torch.autograd.Function):
            """y = x k + b, optionally followed by SiLU (then the pre-activation z is the second, non-differentiable output).
            `z_in`: pre-activation of the SiLU that produced x, or None. Convention inside `mlp`: the SiLU derivative of a hidden
            layer's output is applied by its CONSUMER (fused into the epilogue of the consumer's input-gradient GEMM), so the
            gradient a layer with `silu` receives is already dL/dz, and the gradient it returns for x is dL/dz_in."""

            @staticmethod
            def forward(ctx, x, k, b, aux, silu, z_in):
                lib = nat.load_library()
                x = x.contiguous()
                M, ldx = x.shape
                K, N = k.shape
                y = torch.empty((M, N), device=x.device, dtype=torch.float32)
                z = torch.empty((M, N) if silu else (0,), device=x.device, dtype=torch.float32)
                st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
                chk(lib, lib.pgtt_linear_forward(x.data_ptr(), ldx, k.data_ptr(), b.data_ptr(), M, K, N, int(silu), y.data_ptr(), z.data_ptr() if silu else None, st))
                ctx.save_for_backward(x, k, z_in if z_in is not None else z.new_empty(0))
                ctx.aux, ctx.has_zin = aux, z_in is not None
                ctx.mark_non_differentiable(z)
                return y, z

            @staticmethod
            def backward(ctx, g, _gz):
                lib = nat.load_library()
                x, k, z_in = ctx.saved_tensors
                M, ldx = x.shape
                K, N = k.shape
                dev = g.device
                cur = torch.cuda.current_stream(dev)
                dz = g.contiguous()
                aux = ctx.aux
                if aux is not None:
                    aux.wait_stream(cur)
                    dz.record_stream(aux); x.record_stream(aux)
                with torch.cuda.stream(aux) if aux is not None else contextlib.nullcontext():
                    sa = torch.cuda.current_stream(dev)
                    gk = torch.empty((K, N), device=dev, dtype=torch.float32)
                    gb = torch.empty((N,), device=dev, dtype=torch.float32)
                    scratch = torch.empty(int(lib.pgtt_linear_backward_params_scratch(M, K, N)), device=dev, dtype=torch.float32)
                    chk(lib, lib.pgtt_linear_backward_params(x.data_ptr(), ldx, dz.data_ptr(), M, K, N, gk.data_ptr(), gb.data_ptr(), scratch.data_ptr(),
                                                             C.c_void_p(sa.cuda_stream)))
                    if aux is not None:
                        gk.record_stream(cur); gb.record_stream(cur)
                gx = None
                if ctx.needs_input_grad[0]:
                    gx = torch.zeros((M, ldx), device=dev, dtype=torch.float32) if ldx != K else torch.empty((M, ldx), device=dev, dtype=torch.float32)
                    chk(lib, lib.pgtt_linear_backward_input(dz.data_ptr(), k.data_ptr(), M, K, N, gx.data_ptr(), ldx, z_in.data_ptr() if ctx.has_zin else None,
                                                            C.c_void_p(cur.cuda_stream)))
                return gx, gk, gb, None, None, None
        _NATIVE_LINEAR = NativeLinear
    return _NATIVE_LINEAR


@dataclass
class GatherInput:
    """A minibatch that is still a recipe: rows (t, j), t < T, j < len(idx), are rows `idx[j]` of time slices of the time-major store
    `data` [T_full, S, ld]; `mean` / `inv_std` (fp32 [width] or None) normalise on the way in. The whole-MLP node hands it to
    `pgtt_mlp_forward_gather`, whose input split kernel gathers, normalises and converts in one pass - no gathered or
    normalised copy of the observations exists."""
    data: object
    idx: object
    T: int
    width: int
    mean: object = None
    inv_std: object = None

    def first(self, T: int) -> "GatherInput":
        return GatherInput(self.data, self.idx, T, self.width, self.mean, self.inv_std)

    def materialise(self):
        x = self.data[:self.T].index_select(1, self.idx)[..., :self.width]
        return (x - self.mean) * self.inv_std if self.mean is not None else x


_NATIVE_MLP = None
_MLP_HANDLES: Dict = {}


class _MlpHandle:
    """One `pgtt_mlp` handle (device workspace for one network at one row count), destroyed with the object."""

    def __init__(self, lib, dims, rows, device):
        import ctypes as C
        from . import _native as nat
        self.lib, self.h = lib, C.c_void_p()
        rc = lib.pgtt_mlp_create(len(dims) - 1, (C.c_int * len(dims))(*dims), rows, device, C.byref(self.h))
        if rc:
            raise nat.PgttError(rc, lib.pgtt_mlp_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.pgtt_mlp_destroy(self.h)
            self.h = None


def _native_mlp():
    """The whole MLP, forward and backward, on the hand-written blocked split-bf16 tcgen05 GEMMs of csrc/pgtt_mlp.cu
    (`pgtt_mlp_forward` / `pgtt_mlp_backward`, include/pgtt_b200.h): one autograd node per network. The handle (workspace) is
    cached per (network = address of its first kernel, widths, rows); it is created outside graph capture by the trainer's
    eager warm-up steps."""
    global _NATIVE_MLP
    if _NATIVE_MLP is None:
        import ctypes as C
        import torch
        from . import _native as nat

        def chk(lib, rc):
            if rc:
                raise nat.PgttError(rc, lib.pgtt_mlp_last_error().decode())

        ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])

        class NativeMLP(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, n_layers, *params):
                lib = nat.load_library()
                ks, bs = params[:n_layers], params[n_layers:]
                dims = (ks[0].shape[0], *[k.shape[1] for k in ks])
                gather = isinstance(x, GatherInput)
                if gather:
                    dev, rows = x.data.device, x.T * x.idx.numel()
                    assert x.data.dim() == 3 and x.data.is_contiguous() and x.data.dtype == torch.float32 and x.idx.dtype == torch.int64 and x.width == dims[0]
                else:
                    x = x if x.stride(-1) == 1 and x.stride(0) >= x.shape[1] else x.contiguous()
                    dev, rows = x.device, x.shape[0]
                key = (ks[0].data_ptr(), dims, rows, dev.index)
                h = _MLP_HANDLES.get(key)
                if h is None:
                    if torch.cuda.is_current_stream_capturing():
                        raise RuntimeError("pgtt_mlp handle for a new (network, rows) requested inside CUDA graph capture: run one eager step first")
                    while len(_MLP_HANDLES) >= 8:          # bounded: the oldest workspace goes (handles hold ~100 MB of device memory each)
                        _MLP_HANDLES.pop(next(iter(_MLP_HANDLES)))
                    h = _MLP_HANDLES[key] = _MlpHandle(lib, dims, rows, dev.index)
                y = torch.empty((rows, dims[-1]), device=dev, dtype=torch.float32)
                st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                if gather:
                    chk(lib, lib.pgtt_mlp_forward_gather(h.h, x.data.data_ptr(), x.data.shape[2], x.data.shape[1], x.idx.data_ptr(), x.idx.numel(),
                                                         x.mean.data_ptr() if x.mean is not None else None, x.inv_std.data_ptr() if x.mean is not None else None,
                                                         ptrs(ks), ptrs(bs), y.data_ptr(), st))
                else:
                    chk(lib, lib.pgtt_mlp_forward(h.h, x.data_ptr(), x.stride(0), ptrs(ks), ptrs(bs), y.data_ptr(), st))
                ctx.h, ctx.n_layers = h, n_layers
                ctx.save_for_backward(*params)
                return y

            @staticmethod
            def backward(ctx, gy):
                lib = nat.load_library()
                params = ctx.saved_tensors
                L = ctx.n_layers
                gy = gy.contiguous()
                grads = [torch.empty_like(p) for p in params]
                st = C.c_void_p(torch.cuda.current_stream(gy.device).cuda_stream)
                chk(lib, lib.pgtt_mlp_backward(ctx.h.h, gy.data_ptr(), ptrs(grads[:L]), ptrs(grads[L:]), st))
                return (None, None, *grads)
        _NATIVE_MLP = NativeMLP
    return _NATIVE_MLP


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def mlp(x, kernels, biases, aux=None, native: bool = False):
    """`aux`: CUDA stream for the parameter-gradient GEMMs of the backward pass (the caller joins it before it reads the
    gradients); None = everything on the stream of the forward. `native`: the hand-written tcgen05 layers (CUDA fp32 only)."""
    import torch
    if isinstance(x, GatherInput):
        if native is True:
            y = _native_mlp().apply(x, len(kernels), *kernels, *biases)
            return y.reshape(x.T, x.idx.numel(), kernels[-1].shape[1])
        x = x.materialise()
    if native and x.is_cuda and x.dtype == torch.float32 and native != "layers":
        lead = x.shape[:-1]
        y = _native_mlp().apply(x.reshape(-1, x.shape[-1]), len(kernels), *kernels, *biases)
        return y.reshape(*lead, kernels[-1].shape[1])
    if native and x.is_cuda and x.dtype == torch.float32:
        lin = _native_linear()
        lead = x.shape[:-1]
        h = x.reshape(-1, x.shape[-1])
        z_prev = None
        for i, (k, b) in enumerate(zip(kernels, biases)):
            hidden = i + 1 < len(kernels)
            h, z = lin.apply(h, k, b, aux, hidden, z_prev)
            z_prev = z if hidden else None
        return h.reshape(*lead, kernels[-1].shape[1])
    lin = _linear()
    for i, (k, b) in enumerate(zip(kernels, biases)):
        x = lin.apply(x.reshape(-1, x.shape[-1]), k, b, aux).reshape(*x.shape[:-1], k.shape[1])
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


def _fused_head():
    """torch.autograd.Function over `pgtt_ppo_head`: loss terms and the gradients wrt logits / value predictions in one launch."""
    import ctypes as C
    import torch
    from . import _native as nat

    class PPOHead(torch.autograd.Function):
        @staticmethod
        def forward(ctx, logits, baseline, raw_action, old_lp, adv, vs, eps, moments, clip_eps, entropy_cost, min_std):
            lib = nat.load_library()
            A = raw_action.shape[-1]
            lg = logits.reshape(-1, 2 * A).contiguous()
            M = lg.shape[0]
            g_logits, g_base = torch.empty_like(lg), torch.empty(M, device=lg.device, dtype=lg.dtype)
            sums = torch.empty(4, device=lg.device, dtype=lg.dtype)
            flat = [t.reshape(-1).contiguous() for t in (baseline, raw_action, old_lp, adv, vs, eps)]
            stream = C.c_void_p(torch.cuda.current_stream(lg.device).cuda_stream)
            rc = lib.pgtt_ppo_head(lg.data_ptr(), *(t.data_ptr() for t in flat), moments.data_ptr(), M, A, clip_eps, entropy_cost, min_std,
                                   g_logits.data_ptr(), g_base.data_ptr(), sums.data_ptr(), stream)
            if rc:
                raise nat.PgttError(rc, lib.pgtt_policy_last_error().decode())
            ctx.save_for_backward(g_logits, g_base)
            ctx.shapes = (logits.shape, baseline.shape)
            return sums

        @staticmethod
        def backward(ctx, gs):
            g_logits, g_base = ctx.saved_tensors
            return (g_logits * gs[0]).reshape(ctx.shapes[0]), (g_base * gs[0]).reshape(ctx.shapes[1]), None, None, None, None, None, None, None, None, None

    return PPOHead


def ppo_loss(policy_params, value_params, batch: Dict, cfg: PPOConfig, moments_fn: Optional[Callable] = None, fused: bool = False,
             side_stream=None, aux_streams=(None, None)):
    """batch: time-major [T, B, ...] tensors with NORMALISED observations `obs`, `obs_priv` ([T + 1, B, .]) plus
    raw_action, log_prob, reward, discount, truncation ([T, B]) and entropy noise `eps` [T, B, A].
    `side_stream`: the value network runs there, concurrently with the policy network (autograd replays each backward on
    its forward's stream, so the two backward chains overlap as well); both are captured into the SGD-step graph as
    parallel branches."""
    import torch
    pk, pb = policy_params
    vk, vb = value_params
    T = batch["reward"].shape[0]
    obs_T = batch["obs"].first(T) if isinstance(batch["obs"], GatherInput) else batch["obs"][:T]
    if side_stream is not None:
        cur = torch.cuda.current_stream(batch["reward"].device)
        side_stream.wait_stream(cur)
        with torch.cuda.stream(side_stream):
            baseline_all = mlp(batch["obs_priv"], vk, vb, aux_streams[1], cfg.native_mlp).squeeze(-1)   # [T + 1, B]
        logits = mlp(obs_T, pk, pb, aux_streams[0], cfg.native_mlp)
        cur.wait_stream(side_stream)
        baseline_all.record_stream(cur)
    else:
        logits = mlp(obs_T, pk, pb, aux_streams[0], cfg.native_mlp)
        baseline_all = mlp(batch["obs_priv"], vk, vb, aux_streams[1], cfg.native_mlp).squeeze(-1)       # [T + 1, B]
    baseline, bootstrap = baseline_all[:T], baseline_all[T]
    truncation = batch["truncation"]
    if batch["reward"].is_cuda and batch["reward"].dtype == torch.float32:
        vs, adv = compute_gae_native(truncation, batch["discount"], batch["reward"], baseline_all.detach(), cfg.gae_lambda, cfg.discounting, cfg.reward_scaling)
    else:
        rewards = batch["reward"] * cfg.reward_scaling
        termination = (1.0 - batch["discount"]) * (1.0 - truncation)
        vs, adv = compute_gae(truncation, termination, rewards, baseline.detach(), bootstrap.detach(), cfg.gae_lambda, cfg.discounting)
    if cfg.normalize_advantage:
        if moments_fn is not None:
            mean, std = moments_fn(adv)
        else:
            mean, std = adv.mean(), adv.std(unbiased=False)
    else:
        mean, std = torch.zeros((), device=adv.device), torch.ones((), device=adv.device) - 1e-8
    if fused:   # hand-written head kernel: forward terms + gradients wrt logits / baseline in one launch (CUDA float32 only)
        sums = _fused_head().apply(logits, baseline, batch["raw_action"], batch["log_prob"], adv, vs, batch["eps"],
                                   torch.stack([mean, std]).to(torch.float32), cfg.clipping_epsilon, cfg.entropy_cost, 0.001)
        return sums[0], {"total_loss": sums[0].detach(), "policy_loss": sums[1].detach(), "v_loss": sums[2].detach(), "entropy": sums[3].detach()}
    adv = (adv - mean) / (std + 1e-8)
    target_lp = tanh_normal_log_prob(logits, batch["raw_action"])      # (the fused head computes it itself: not evaluated on that path)
    rho = torch.exp(target_lp - batch["log_prob"])
    s1 = rho * adv
    s2 = torch.clamp(rho, 1.0 - cfg.clipping_epsilon, 1.0 + cfg.clipping_epsilon) * adv
    policy_loss = -torch.minimum(s1, s2).mean()
    v_loss = ((vs - baseline) ** 2).mean() * 0.5 * 0.5
    entropy = tanh_normal_entropy(logits, batch["eps"]).mean()
    total = policy_loss + v_loss - cfg.entropy_cost * entropy
    return total, {"total_loss": total.detach(), "policy_loss": policy_loss.detach(), "v_loss": v_loss.detach(), "entropy": entropy.detach()}


class FlatAdam:
    """All parameters as views of ONE flat device vector; `step(flat_grad)` = `clip_grad_norm_` + `torch.optim.Adam.step` from the
    hand-written kernel pair `pgtt_adam_clip` (two launches instead of ~10 multi-tensor ones, 60 -> 8 us; deterministic norm).
    `grad_scale` folds the 1 / world averaging of an all-reduced gradient into the same pass."""

    def __init__(self, params, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        import torch
        from . import _native as nat
        self.lib = nat.load_library()
        self.params = list(params)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        with torch.no_grad():
            self.flat = torch.cat([p.detach().reshape(-1) for p in self.params]).contiguous()
            off = 0
            for p in self.params:
                p.data = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.t = torch.zeros(1, dtype=torch.float32, device=self.flat.device)
        self.scratch = torch.zeros(int(self.lib.pgtt_adam_scratch_floats()), dtype=torch.float32, device=self.flat.device)

    def flat_grad(self):
        import torch
        return torch.cat([p.grad.reshape(-1) for p in self.params])

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def step(self, flat_grad, max_norm: Optional[float], grad_scale: float = 1.0):
        import ctypes as C
        import torch
        from . import _native as nat
        assert flat_grad.is_contiguous() and flat_grad.numel() == self.flat.numel() and flat_grad.dtype == torch.float32
        stream = C.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        rc = self.lib.pgtt_adam_clip(self.flat.data_ptr(), flat_grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.t.data_ptr(),
                                     self.scratch.data_ptr(), self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                     float(max_norm) if max_norm is not None else 0.0, float(grad_scale), stream)
        if rc:
            raise nat.PgttError(rc, self.lib.pgtt_policy_last_error().decode())


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
        self.opt = torch.optim.Adam(self.params, lr=cfg.learning_rate, eps=1e-8, capturable=cfg.use_cuda_graph, fused=True)
        self.flat_opt = FlatAdam(self.params, cfg.learning_rate, eps=1e-8) if cfg.native_optimizer else None
        self.norm_state, self.norm_priv = RunningStats(nobs, self.dev), RunningStats(npriv, self.dev)
        self.net = PolicyNet((nobs, *cfg.policy_hidden_layer_sizes, 24), device=self.abi.device)
        self.collector = RolloutCollector(wenv, self.net, unroll_length=cfg.unroll_length, seed=cfg.seed * 7919 + self.rank)
        self.state = state
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(cfg.seed * 31 + 1 + self.rank)
        self.env_steps = 0
        self._graph = None
        self._side = None
        self._aux = None
        self._data: Dict = {}
        self._norm_seen = False             # host-side: the running statistics hold at least one batch
        self._nb = None
        self.metrics: Dict = {}
        self.learner_kind = {
            True: "hand-written whole-MLP forward / backward on blocked split-bf16 tcgen05 GEMMs (fp32 accumulation, minibatch gather + normalisation fused into "
                  "the input kernel; csrc/pgtt_mlp.cu) + hand-written GAE / loss-head / clip+Adam kernels",
            "layers": "hand-written per-layer tcgen05 GEMMs (split-bf16 products, fp32 accumulation; csrc/pgtt_learner.cu) + hand-written GAE / loss-head / clip+Adam kernels",
        }.get(cfg.native_mlp, "torch autograd over library GEMMs + hand-written GAE / loss-head / clip+Adam kernels")
        self._sync_policy()

    # -- policy kernel <- learner parameters ---------------------------------------------------------------------------------
    def _sync_policy(self):
        """Stream-ordered, on the device: a packing kernel reads the fp32 parameters and the running statistics (no host copy, no sync)."""
        ks, bs = [k.detach() for k in self.policy_params[0]], [b.detach() for b in self.policy_params[1]]
        if self.cfg.normalize_observations and self._norm_seen:
            self.net.set_params_device(ks, bs, self.norm_state.mean.float(), self.norm_state.std.float())
        else:
            self.net.set_params_device(ks, bs)

    # -- one SGD step (optionally replayed from a CUDA graph) ----------------------------------------------------------------
    def _moments(self, adv):
        if self.cfg.global_advantage_norm and self.world > 1:
            _, mean, var = sharding.allreduce_moments(adv.reshape(-1, 1), self.group)
            return mean[0].to(adv.dtype), var[0].sqrt().to(adv.dtype)
        return adv.mean(), adv.std(unbiased=False)

    def _sgd_body(self, batch):
        torch = self.torch
        par = self.cfg.parallel_nets
        if par and self._side is None:
            self._side = torch.cuda.Stream(self.dev)
        # auxiliary streams carry the parameter-gradient GEMMs of the per-layer paths; the whole-MLP node (native_mlp = True) issues
        # its weight-gradient GEMMs itself and never forks them (waiting on a stream that was not forked would break a capture)
        use_aux = par and self.cfg.native_mlp is not True
        if use_aux and self._aux is None:
            self._aux = (torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev))
        loss, m = ppo_loss(self.policy_params, self.value_params, batch, self.cfg, self._moments, fused=self.cfg.fused_head,
                           side_stream=self._side if par else None, aux_streams=self._aux if use_aux else (None, None))
        self.opt.zero_grad(set_to_none=True)    # backward writes fresh gradients: no fill + accumulate pair per parameter
        loss.backward()
        if use_aux:   # the parameter-gradient branches rejoin before anything reads the gradients
            cur = torch.cuda.current_stream(self.dev)
            for a in self._aux:
                cur.wait_stream(a)
        if self.flat_opt is not None:
            flat = self.flat_opt.flat_grad()
            if self.world > 1:
                torch.distributed.all_reduce(flat, group=self.group)
            self.flat_opt.step(flat, self.cfg.max_grad_norm, 1.0 / self.world)
            return m
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

    def _fused_input(self) -> bool:
        return self.cfg.native_mlp is True

    # -- the SGD step without autograd: every launch is one of this repo's kernels ------------------------------------------------------
    def _native_step_ok(self) -> bool:
        c = self.cfg
        return bool(c.native_mlp is True and c.fused_head and c.native_step and self.flat_opt is not None and c.normalize_advantage and self.mb <= 1024)

    def _sgd_body_native(self):
        """One SGD step as a fixed launch sequence over static buffers (what brax's `sgd_step` / `loss_and_pgrad` / `optimizer.update` do,
        training/train.py:135-161): minibatch index -> `pgtt_minibatch_gather` (segment ids, raw actions, the four per-transition scalars, entropy noise: one launch) -> both
        MLP forwards with the observation gather + normalisation fused in (`pgtt_mlp_forward_gather`; value network on a second stream) ->
        `pgtt_gae_moments` -> `pgtt_ppo_head` (loss terms + gradients wrt logits / values) -> both MLP backwards (`pgtt_mlp_backward`, writing
        straight into one flat gradient vector) -> [NCCL all-reduce] -> `pgtt_adam_clip`. Equals the autograd path (`_sgd_body`) to rounding
        (tests/test_ppo.py::test_native_sgd_step_equals_the_autograd_step)."""
        import ctypes as C
        from . import _native as nat
        torch, cfg, lib, dev = self.torch, self.cfg, nat.load_library(), self.dev
        T, mb = cfg.unroll_length, self.mb
        if self._nb is None:
            f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
            flat_g = torch.zeros_like(self.flat_opt.flat)
            views, off = [], 0
            for p_ in self.params:
                views.append(flat_g[off:off + p_.numel()].view_as(p_))
                off += p_.numel()
            L = len(self.policy_params[0])
            mk = lambda ks, rows: _MlpHandle(lib, (ks[0].shape[0], *[k.shape[1] for k in ks]), rows, dev.index)
            self._nb = {"logits": f(T * mb, 24), "base_all": f((T + 1) * mb), "vs": f(T, mb), "adv": f(T, mb), "mom": f(2), "sums3": torch.zeros(3, dtype=torch.float64, device=dev), "g_logits": f(T * mb, 24),
                        "g_base_all": torch.zeros((T + 1) * mb, dtype=torch.float32, device=dev), "sums": f(4), "flat_g": flat_g,
                        "g_pol": (views[:L], views[L:2 * L]), "g_val": (views[2 * L:3 * L], views[3 * L:]),
                        "h_pol": mk(self.policy_params[0], T * mb), "h_val": mk(self.value_params[0], (T + 1) * mb),
                        "idx": torch.zeros(mb, dtype=torch.int64, device=dev), "raw": f(T, mb, 12), "scal": f(len(_SCALARS), T, mb), "eps": f(T, mb, 12)}
            if self._side is None:
                self._side = torch.cuda.Stream(dev)
        nb = self._nb
        ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])

        def chk(rc, err):
            if rc:
                raise nat.PgttError(rc, err().decode())
        cur, side = torch.cuda.current_stream(dev), self._side
        st = lambda s: C.c_void_p(s.cuda_stream)
        S = self._data["obs"].shape[1]
        idx, raw, scal, eps = nb["idx"], nb["raw"], nb["scal"], nb["eps"]                 # segment ids [mb]; [T, mb, 12]; log_prob, reward, discount, truncation [4, T, mb]; [T, mb, 12]
        chk(lib.pgtt_minibatch_gather(self._perm.data_ptr(), self._mbi.data_ptr(), mb, S, T, 12, len(_SCALARS), self._data["raw_action"].data_ptr(), self._scal.data_ptr(),
                                      self._eps.data_ptr(), idx.data_ptr(), raw.data_ptr(), scal.data_ptr(), eps.data_ptr(), st(cur)), lib.pgtt_policy_last_error)
        (m_o, i_o), (m_p, i_p) = self._norm_dev
        dp = lambda t: t.data_ptr() if t is not None else None
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            chk(lib.pgtt_mlp_forward_gather(nb["h_val"].h, self._data["obs_priv"].data_ptr(), self._data["obs_priv"].shape[2], S, idx.data_ptr(), mb, dp(m_p), dp(i_p),
                                            ptrs(self.value_params[0]), ptrs(self.value_params[1]), nb["base_all"].data_ptr(), st(side)), lib.pgtt_mlp_last_error)
        chk(lib.pgtt_mlp_forward_gather(nb["h_pol"].h, self._data["obs"].data_ptr(), self._data["obs"].shape[2], S, idx.data_ptr(), mb, dp(m_o), dp(i_o),
                                        ptrs(self.policy_params[0]), ptrs(self.policy_params[1]), nb["logits"].data_ptr(), st(cur)), lib.pgtt_mlp_last_error)
        cur.wait_stream(side)
        gae_args = (scal[3].data_ptr(), scal[2].data_ptr(), scal[1].data_ptr(), nb["base_all"].data_ptr(), T, mb, cfg.gae_lambda, cfg.discounting, cfg.reward_scaling,
                    nb["vs"].data_ptr(), nb["adv"].data_ptr())
        if cfg.global_advantage_norm and self.world > 1:     # the advantage-normalisation all-reduce: three float64 sums per rank
            chk(lib.pgtt_gae_sums(*gae_args, nb["sums3"].data_ptr(), st(cur)), lib.pgtt_policy_last_error)
            torch.distributed.all_reduce(nb["sums3"], group=self.group)
            chk(lib.pgtt_moments_finalize(nb["sums3"].data_ptr(), nb["mom"].data_ptr(), st(cur)), lib.pgtt_policy_last_error)
        else:
            chk(lib.pgtt_gae_moments(*gae_args, nb["mom"].data_ptr(), st(cur)), lib.pgtt_policy_last_error)
        chk(lib.pgtt_ppo_head(nb["logits"].data_ptr(), nb["base_all"].data_ptr(), raw.data_ptr(), scal[0].data_ptr(), nb["adv"].data_ptr(), nb["vs"].data_ptr(), eps.data_ptr(),
                              nb["mom"].data_ptr(), T * mb, 12, cfg.clipping_epsilon, cfg.entropy_cost, 0.001, nb["g_logits"].data_ptr(), nb["g_base_all"].data_ptr(),
                              nb["sums"].data_ptr(), st(cur)), lib.pgtt_policy_last_error)
        side.wait_stream(cur)
        with torch.cuda.stream(side):      # (the bootstrap row of g_base_all stays zero: no gradient flows through the bootstrap value)
            chk(lib.pgtt_mlp_backward(nb["h_val"].h, nb["g_base_all"].data_ptr(), ptrs(nb["g_val"][0]), ptrs(nb["g_val"][1]), st(side)), lib.pgtt_mlp_last_error)
        chk(lib.pgtt_mlp_backward(nb["h_pol"].h, nb["g_logits"].data_ptr(), ptrs(nb["g_pol"][0]), ptrs(nb["g_pol"][1]), st(cur)), lib.pgtt_mlp_last_error)
        cur.wait_stream(side)
        flat = nb["flat_g"]
        if self.world > 1:
            torch.distributed.all_reduce(flat, group=self.group)
        self.flat_opt.step(flat, cfg.max_grad_norm, 1.0 / self.world)
        sums = nb["sums"]
        return {"total_loss": sums[0], "policy_loss": sums[1], "v_loss": sums[2], "entropy": sums[3]}

    def _minibatch(self):
        """Minibatch `self._mbi` of the current epoch, gathered ON THE DEVICE from the static full-data buffers: the index
        tensors are the only thing the host touches per SGD step, so the whole step (gather, forward, loss, backward,
        clip, Adam) is one graph replay."""
        torch = self.torch
        idx = self._perm.index_select(0, self._mbi).reshape(-1)                          # [mb] segment ids
        if self._fused_input():     # the whole-MLP node gathers, normalises and converts the observations itself (pgtt_mlp_forward_gather)
            batch = {"obs": GatherInput(self._data["obs"], idx, self.cfg.unroll_length + 1, self.abi.nobs, *self._norm_dev[0]),
                     "obs_priv": GatherInput(self._data["obs_priv"], idx, self.cfg.unroll_length + 1, self.abi.npriv, *self._norm_dev[1])}
        else:
            side = None
            if self.cfg.parallel_nets:     # the value network's input (the largest gather) is fetched on the stream that consumes it
                if self._side is None:
                    self._side = torch.cuda.Stream(self.dev)
                side = self._side
                side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
                batch = {"obs_priv": self._data["obs_priv"].index_select(1, idx)}
            batch["obs"] = self._data["obs"].index_select(1, idx)
        batch["raw_action"] = self._data["raw_action"].index_select(1, idx)
        # the four per-transition scalars live in one [4, T, S] buffer: one gather instead of four
        batch.update(zip(_SCALARS, self._scal.index_select(2, idx).unbind(0)))
        batch["eps"] = self._eps.index_select(0, self._mbi)[0]
        return batch

    def _sgd_step(self, i: int):
        torch = self.torch
        self._mbi.fill_(i)
        # multi-rank: the NCCL all-reduces (gradients, advantage moments) are captured with the rest of the step - every rank
        # captures and replays the same sequence (2.8x on 2 GPUs against issuing the step eagerly); PGTT_GRAPH_NCCL=0 opts out
        import os
        graphable = self.cfg.use_cuda_graph and (self.world == 1 or os.environ.get("PGTT_GRAPH_NCCL", "1") == "1")
        body = self._sgd_body_native if self._native_step_ok() else (lambda: self._sgd_body(self._minibatch()))
        if not graphable:
            return body()
        if self._graph is None:
            s = torch.cuda.Stream(self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                for _ in range(3):                                   # warm-up outside capture (allocator, Adam state, MLP handles)
                    body()
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
                    self._static_metrics = body()
            finally:
                gc.enable()
            # the warm-up steps changed the parameters: acceptable for training (three extra SGD steps on the first
            # minibatch), callers that need exact step counts construct the trainer with use_cuda_graph=False
        self._graph.replay()
        return self._static_metrics

    # -- one training step ----------------------------------------------------------------------------------------------------
    def training_step(self) -> Dict:
        torch, cfg = self.torch, self.cfg
        T = cfg.unroll_length
        if not self._data:       # static buffers (their addresses are baked into the captured graph)
            nobs, npriv, S = self.abi.nobs, self.abi.npriv, self.segments
            f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=self.dev)
            # observation rows are zero-padded to a multiple of four floats (aligned first-layer GEMMs, see _linear)
            z = lambda *sh: torch.zeros(sh, dtype=torch.float32, device=self.dev)
            self._data = {"obs": z(T + 1, S, pad4(nobs)), "obs_priv": z(T + 1, S, pad4(npriv)), "raw_action": f(T, S, 12)}
            self._scal = f(len(_SCALARS), T, S)
            self._data.update(zip(_SCALARS, self._scal.unbind(0)))
            self._perm = torch.zeros((cfg.num_minibatches, self.mb), dtype=torch.int64, device=self.dev)
            self._eps = f(cfg.num_minibatches, T, self.mb, 12)
            self._mbi = torch.zeros(1, dtype=torch.int64, device=self.dev)
            nrm = lambda d: (torch.zeros(d, dtype=torch.float32, device=self.dev), torch.ones(d, dtype=torch.float32, device=self.dev))
            self._norm_dev = (nrm(nobs), nrm(npriv)) if cfg.normalize_observations else ((None, None), (None, None))
        n = self.abi.N
        for u in range(self.unrolls_per_step):
            self.state, ro = self.collector.collect()
            sl = slice(u * n, (u + 1) * n)
            for k, src in (("obs", ro.obs_state), ("obs_priv", ro.obs_privileged), ("raw_action", ro.raw_action), ("log_prob", ro.log_prob),
                           ("reward", ro.reward), ("discount", ro.discount), ("truncation", ro.truncation)):
                self._data[k][:, sl, ..., :src.shape[-1]].copy_(src) if k.startswith("obs") else self._data[k][:, sl].copy_(src)
            self.env_steps += T * n * self.world
        data = self._data
        reward_mean, done_rate = data["reward"].mean(), (1.0 - data["discount"]).mean()
        if cfg.normalize_observations:
            obs, priv = data["obs"][..., :self.abi.nobs], data["obs_priv"][..., :self.abi.npriv]
            self.norm_state.update(obs[:T], self.group)
            self.norm_priv.update(priv[:T], self.group)
            self._norm_seen = True
            if self._fused_input():     # raw observations stay in the store; the statistics travel to the input kernel (static buffers: graph replays see them)
                for (m32, i32), rs in zip(self._norm_dev, (self.norm_state, self.norm_priv)):
                    m32.copy_(rs.mean)
                    i32.copy_(1.0 / rs.std)
            else:
                obs.copy_(self.norm_state.normalize(obs))
                priv.copy_(self.norm_priv.normalize(priv))
        last = {}
        for _ in range(cfg.num_updates_per_batch):
            self._perm.copy_(torch.randperm(self.segments, generator=self.gen, device=self.dev).reshape(cfg.num_minibatches, self.mb))
            self._eps.normal_(generator=self.gen)
            for i in range(cfg.num_minibatches):
                last = self._sgd_step(i)
        self._sync_policy()
        self.metrics = {k: float(v) for k, v in last.items()}
        self.metrics["reward_per_step"] = float(reward_mean)
        self.metrics["episode_done_rate"] = float(done_rate)
        self.metrics["env_steps"] = self.env_steps
        return self.metrics

    # -- brax-layout export (deploy/policy_net.py:6-33 reads it) ----------------------------------------------------------------
    def save(self, path):
        from . import policy_io
        det = lambda ts: [t.detach().cpu().numpy() for t in ts]
        policy_io.save_policy(path, self.norm_state.mean.float().cpu().numpy(), self.norm_state.std.float().cpu().numpy(),
                              (det(self.policy_params[0]), det(self.policy_params[1])), (det(self.value_params[0]), det(self.value_params[1])),
                              count=float(self.norm_state.count), value_mean=self.norm_priv.mean.float().cpu().numpy(),
                              value_std=self.norm_priv.std.float().cpu().numpy())

    def restore(self, path):
        """`restore_checkpoint_path` of brax ppo.train (training/train.py:248-256): normaliser + policy + value parameters from a
        pickle in the reference's layout (one written by `save`, or a shipped policy_folder/policyNNN); the optimiser starts fresh."""
        import torch
        from . import policy_io
        d = policy_io.load_policy(path)
        if d.get("value") is None:
            raise ValueError(f"{path} has no value-network parameters: cannot resume training from it")
        with torch.no_grad():
            for dst, src in zip(self.policy_params[0] + self.policy_params[1], list(d["policy"][0]) + list(d["policy"][1])):
                dst.copy_(torch.as_tensor(src).to(dst))
            for dst, src in zip(self.value_params[0] + self.value_params[1], list(d["value"][0]) + list(d["value"][1])):
                dst.copy_(torch.as_tensor(src).to(dst))
        cnt = float(d["count"] or 0.0)
        if cnt > 0 and d.get("value_mean") is None:
            # value weights trained on normalised inputs would otherwise meet an un-initialised normaliser
            raise ValueError(f"{path} carries value-network parameters but no privileged_state statistics: cannot resume training from it")
        for rs, mean, std in ((self.norm_state, d["mean"], d["std"]), (self.norm_priv, d.get("value_mean"), d.get("value_std"))):
            if mean is None or cnt <= 0:
                continue
            rs.count = torch.tensor(cnt, dtype=torch.float64, device=self.dev)
            rs.mean = torch.as_tensor(mean, dtype=torch.float64, device=self.dev)
            rs.std = torch.as_tensor(std, dtype=torch.float64, device=self.dev)
            rs.summed_var = rs.std * rs.std * cnt
            self._norm_seen = True
        self._sync_policy()


class Evaluator:
    """brax `acting.Evaluator` as `ppo.train` drives it (training/train.py:135-161 passes `num_evals`, `eval_env`): `num_eval_envs`
    (brax default 128) environments of `eval_env`, wrapped and randomised like the training env, run ONE episode with actions
    sampled from the current policy (`deterministic_eval=False`, the brax default) and report the first-episode sums under the
    keys `progress` reads (training/train.py:198-216): `eval/episode_reward[_std]`, `eval/episode_reward/<term>` and
    `eval/avg_episode_length`. The eval env is a second handle on the trainer's device with its own policy-kernel handle."""

    def __init__(self, eval_env, wrap_env_fn, randomization_fn, cfg: PPOConfig, trainer: "PPOTrainer", num_eval_envs: int = 128,
                 deterministic_eval: bool = False, seed: int = 0):
        import functools
        from . import prng
        self.trainer, self.cfg, self.deterministic, self.seed = trainer, cfg, bool(deterministic_eval), int(seed)
        self.keys = prng.env_keys(seed * 104729 + 17, int(num_eval_envs))
        self.wenv = wrap_env_fn(eval_env, episode_length=cfg.episode_length, action_repeat=1,
                                randomization_fn=functools.partial(randomization_fn, rng=self.keys) if randomization_fn is not None else None)
        self.net = None
        self.runs = 0

    def run_evaluation(self, training_metrics: Optional[Dict] = None) -> Dict:
        import time
        from .evaluate import evaluate
        tr = self.trainer
        t0 = time.time()
        self.runs += 1
        self.wenv.reset(self.keys + np.uint32(self.runs))        # brax: a fresh eval key per run
        if self.net is None:
            self.net = PolicyNet((tr.abi.nobs, *self.cfg.policy_hidden_layer_sizes, 24), device=tr.abi.device)
        ks, bs = tr.policy_params
        if self.cfg.normalize_observations and float(tr.norm_state.count) > 0:
            self.net.set_params(ks, bs, tr.norm_state.mean.float(), tr.norm_state.std.float())
        else:
            self.net.set_params(ks, bs)
        r = evaluate(self.wenv, self.net, episode_length=self.cfg.episode_length, seed=self.seed + self.runs, deterministic=self.deterministic,
                     per_metric=True)
        dt = time.time() - t0
        m = {"eval/episode_reward": r["episode_reward"], "eval/episode_reward_std": r["episode_reward_std"],
             "eval/avg_episode_length": r["avg_episode_length"], "eval/epoch_eval_time": dt,
             "eval/sps": r["num_eval_envs"] * self.cfg.episode_length / max(dt, 1e-9), "eval/success_rate": r["success_rate"]}
        for k, v in r["episode_metrics"].items():
            m[f"eval/episode_{k}"] = v
        for k, v in r["episode_metrics_std"].items():
            m[f"eval/episode_{k}_std"] = v
        for k, v in (training_metrics or {}).items():
            m[f"training/{k}"] = v
        return m


def _rank0_says(stop: bool, trainer: "PPOTrainer") -> bool:
    """rank 0's decision (evaluate at all? stop early?), made known to every rank: they all must take the same path through the loop."""
    if trainer.world == 1:
        return bool(stop)
    import torch
    import torch.distributed as dist
    flag = torch.tensor([1 if stop else 0], dtype=torch.int32, device=trainer.dev)
    dist.broadcast(flag, src=0, group=trainer.group)
    return bool(int(flag.item()))


def train(environment, wrap_env_fn, randomization_fn, rng_keys, cfg: PPOConfig, progress_fn: Optional[Callable] = None,
          policy_params_fn: Optional[Callable] = None, num_training_steps: Optional[int] = None, restore_checkpoint_path=None,
          eval_env=None, num_evals: int = 1, num_eval_envs: int = 128, deterministic_eval: bool = False):
    """Call shape of `ppo.train(environment=..., eval_env=..., wrap_env_fn=..., randomization_fn=..., progress_fn=..., ...)` in
    training/train.py:242-263. `rng_keys`: uint32[N_local, 2] per-env keys of this rank's shard.

    With `eval_env` the schedule is brax's: `num_evals` evaluations (the first one before any training when `num_evals > 1`),
    `ceil(num_timesteps / ((num_evals - 1) * env_steps_per_training_step))` training steps between two of them; rank 0 evaluates
    and calls `progress_fn(num_steps, eval_metrics)` then `policy_params_fn`. A truthy return value of `progress_fn` ends
    training early on every rank - the reference's `progress` returns its convergence verdict (training/train.py:224-229:
    both tracking rewards above `vel_percentage` of their maximum and the episode reward within 0.5 % of the previous
    evaluation, or within 0.1 % regardless), which a stock brax would ignore; it is honoured here.
    Without `eval_env`: `progress_fn(num_steps, training_metrics)` after every training step (no evaluation)."""
    import functools
    wenv = wrap_env_fn(environment, episode_length=cfg.episode_length, action_repeat=1,
                       randomization_fn=functools.partial(randomization_fn, rng=rng_keys) if randomization_fn is not None else None)
    state = wenv.reset(rng_keys)
    trainer = PPOTrainer(wenv, state, cfg)
    if restore_checkpoint_path is not None:
        trainer.restore(restore_checkpoint_path)
    per_step = cfg.unroll_length * cfg.batch_size * cfg.num_minibatches
    if not _rank0_says(eval_env is not None, trainer):       # (only rank 0 needs an eval env: brax evaluates on process 0)
        steps = num_training_steps if num_training_steps is not None else max(1, math.ceil(cfg.num_timesteps / per_step))
        for _ in range(steps):
            m = trainer.training_step()
            stop = progress_fn(trainer.env_steps, m) if progress_fn is not None else False
            if policy_params_fn is not None:
                policy_params_fn(trainer.env_steps, trainer)
            if _rank0_says(bool(stop) and trainer.rank == 0, trainer):
                break
        return trainer
    evaluator = Evaluator(eval_env, wrap_env_fn, randomization_fn, cfg, trainer, num_eval_envs, deterministic_eval, cfg.seed) if trainer.rank == 0 else None
    trainer.stopped_early = False
    trainer.evaluator = evaluator
    epochs = max(int(num_evals) - 1, 1)
    steps_per_epoch = num_training_steps if num_training_steps is not None else max(1, math.ceil(cfg.num_timesteps / (epochs * per_step)))
    if num_evals > 1 and evaluator is not None:
        m = evaluator.run_evaluation({})
        if progress_fn is not None:
            progress_fn(0, m)
    for _ in range(epochs):
        for _ in range(steps_per_epoch):
            tm = trainer.training_step()
        stop = False
        if evaluator is not None:
            m = evaluator.run_evaluation(tm)
            if progress_fn is not None:
                stop = bool(progress_fn(trainer.env_steps, m))
            if policy_params_fn is not None:
                policy_params_fn(trainer.env_steps, trainer)
        if _rank0_says(stop, trainer):
            trainer.stopped_early = True
            break
    return trainer
