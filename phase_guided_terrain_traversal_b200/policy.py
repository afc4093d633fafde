"""Policy network of the rollout collector: thin wrapper over the `pgtt_policy_*` C ABI.

Network spec = what training/train.py:135-161 asks brax for and deploy/policy_net.py:35-64 re-hosts:
`(obs - mean) / std` -> Dense(512) -> swish -> Dense(256) -> swish -> Dense(128) -> swish -> Dense(24) ->
`loc, scale = split(2)`; acting: `raw = loc + (softplus(scale) + 0.001) * eps`, `action = tanh(raw)`
(brax `NormalTanhDistribution`), deployment: `tanh(loc)`. Runs as ONE tcgen05 kernel per call
(csrc/pgtt_policy.cu); there is no torch / CPU fallback on the product path - `reference_forward` below is
a plain-torch fp32 statement of the same math used by the tests only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as nat

DEFAULT_SIZES = (171, 512, 256, 128, 24)


def check(lib, rc):
    if rc != 0:
        raise nat.PgttError(rc, lib.pgtt_policy_last_error().decode())


class PolicyNet:
    def __init__(self, sizes: Sequence[int] = DEFAULT_SIZES, device: int = 0, lib=None):
        import torch
        self.torch = torch
        self.lib = lib if lib is not None else nat.load_library()
        self.sizes = tuple(int(s) for s in sizes)
        self.device = device
        self.torch_device = torch.device("cuda", device)
        arr = (C.c_int * len(self.sizes))(*self.sizes)
        h = C.c_void_p()
        check(self.lib, self.lib.pgtt_policy_create(device, arr, len(self.sizes) - 1, C.byref(h)))
        self.h = h
        self.obs_dim, self.act_dim = self.sizes[0], self.sizes[-1] // 2
        self.step_counter = 0
        self.params = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.pgtt_policy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, kernels, biases, obs_mean=None, obs_std=None):
        """kernels[l]: [in, out] (flax layout), biases[l]: [out]; numpy or torch, any float dtype."""
        to_np = lambda a: np.ascontiguousarray(a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a), dtype=np.float32)
        ks, bs = [to_np(k) for k in kernels], [to_np(b) for b in biases]
        n = len(self.sizes) - 1
        assert len(ks) == n and len(bs) == n
        for l in range(n):
            assert ks[l].shape == (self.sizes[l], self.sizes[l + 1]) and bs[l].shape == (self.sizes[l + 1],), (l, ks[l].shape, bs[l].shape)
        kp = (C.c_void_p * n)(*[k.ctypes.data for k in ks])
        bp = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
        mean = to_np(obs_mean) if obs_mean is not None else None
        std = to_np(obs_std) if obs_std is not None else None
        check(self.lib, self.lib.pgtt_policy_set_params(self.h, kp, bp, mean.ctypes.data if mean is not None else None,
                                                         std.ctypes.data if std is not None else None))
        self.params = (ks, bs, mean, std)

    def set_params_device(self, kernels, biases, obs_mean=None, obs_std=None):
        """The same from CUDA float32 tensors on this handle's device, packed by a kernel on the current stream: no host copies, no
        synchronisation (`pgtt_policy_set_params_device`). `self.params` (the host copy `set_params` keeps) is dropped."""
        torch = self.torch
        n = len(self.sizes) - 1
        ts = list(kernels) + list(biases) + [t for t in (obs_mean, obs_std) if t is not None]
        assert len(kernels) == n and len(biases) == n and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in ts)
        for l in range(n):
            assert tuple(kernels[l].shape) == (self.sizes[l], self.sizes[l + 1]) and tuple(biases[l].shape) == (self.sizes[l + 1],)
        kp = (C.c_void_p * n)(*[k.data_ptr() for k in kernels])
        bp = (C.c_void_p * n)(*[b.data_ptr() for b in biases])
        stream = C.c_void_p(torch.cuda.current_stream(kernels[0].device).cuda_stream)
        check(self.lib, self.lib.pgtt_policy_set_params_device(self.h, kp, bp, obs_mean.data_ptr() if obs_mean is not None else None,
                                                                obs_std.data_ptr() if obs_std is not None else None, stream))
        self.params = None

    def init_random(self, seed: int = 0):
        """Random-init weights of this architecture (lecun-uniform like flax Dense), identity normaliser."""
        g = np.random.default_rng(seed)
        ks = [g.uniform(-1, 1, (i, o)).astype(np.float32) * np.sqrt(3.0 / i) for i, o in zip(self.sizes[:-1], self.sizes[1:])]
        bs = [np.zeros(o, np.float32) for o in self.sizes[1:]]
        self.set_params(ks, bs)
        return self

    def act(self, obs, seed: int = 0, deterministic: bool = False, eps=None, want_logits: bool = False, out=None):
        """obs: CUDA float32 [N, obs_dim]. Returns dict(action, raw_action, log_prob[, logits])."""
        torch = self.torch
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape[1] == self.obs_dim
        n = obs.shape[0]
        if out is None:
            out = {"action": torch.empty((n, self.act_dim), device=obs.device), "raw_action": torch.empty((n, self.act_dim), device=obs.device),
                   "log_prob": torch.empty((n,), device=obs.device)}
            if want_logits:
                out["logits"] = torch.empty((n, 2 * self.act_dim), device=obs.device)
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        ep = eps.data_ptr() if eps is not None else None
        lg = out["logits"].data_ptr() if "logits" in out else None
        check(self.lib, self.lib.pgtt_policy_act(self.h, obs.data_ptr(), n, seed, self.step_counter, int(deterministic), ep,
                                                  out["action"].data_ptr(), out["raw_action"].data_ptr(), out["log_prob"].data_ptr(), lg, stream))
        self.step_counter += 1
        return out

    def launch_count(self) -> int:
        return int(self.lib.pgtt_policy_launch_count(self.h))


def reference_forward(kernels, biases, obs, mean=None, std=None, eps=None, bf16_operands: bool = False):
    """TEST-ONLY plain torch fp32 statement of the acting step (the fp32 reference a floating-point
    kernel is compared with). `bf16_operands` rounds activations / weights to bf16 before each matmul
    like the tensor-core path does (accumulation stays fp32)."""
    import torch
    x = torch.as_tensor(obs, dtype=torch.float32)
    if mean is not None:
        x = (x - torch.as_tensor(mean)) * (1.0 / torch.as_tensor(std))
    rb = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if bf16_operands else (lambda t: t)
    n = len(kernels)
    for l in range(n):
        x = rb(x) @ rb(torch.as_tensor(kernels[l], dtype=torch.float32)) + torch.as_tensor(biases[l], dtype=torch.float32)
        if l < n - 1:
            x = x * torch.sigmoid(x)
    loc, sr = torch.chunk(x, 2, dim=-1)
    scale = torch.nn.functional.softplus(sr) + 0.001
    e = torch.zeros_like(loc) if eps is None else torch.as_tensor(eps, dtype=torch.float32)
    raw = loc + scale * e
    logp = (-0.5 * e * e - torch.log(scale) - 0.5 * np.log(2 * np.pi)).sum(-1) - (2 * (np.log(2.0) - raw - torch.nn.functional.softplus(-2 * raw))).sum(-1)
    return {"logits": x, "action": torch.tanh(raw), "raw_action": raw, "log_prob": logp}
