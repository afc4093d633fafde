"""PPO rollout collector: brax `generate_unroll` / `actor_step` as driven by training/train.py:135-161,242-263.

Per control step: `obs["state"]` -> fused tcgen05 policy kernel (normalise, MLP, tanh-normal sample) ->
fused env step kernel -> transition write-out. Transitions are stored time-major, `[T, N, ...]`, the
layout the learner consumes (brax `Transition(observation, action, reward, discount, next_observation,
extras{policy_extras{log_prob, raw_action}, state_extras{truncation}})`); observations are kept as
`[T + 1, N, ...]` so `next_observation[t] = observation[t + 1]` (what the auto-reset wrapper returns).
Everything stays on the device; the three kernels per step are launched on the caller's stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Optional

from . import _native as nat
from .policy import PolicyNet


@dataclass
class Rollout:
    obs_state: Any          # [T + 1, N, 171]
    obs_privileged: Any     # [T + 1, N, 215]
    action: Any             # [T, N, 12]
    raw_action: Any         # [T, N, 12]
    log_prob: Any           # [T, N]
    reward: Any             # [T, N]
    discount: Any           # [T, N] = 1 - done
    truncation: Any         # [T, N]

    @property
    def observation(self):
        return {"state": self.obs_state[:-1], "privileged_state": self.obs_privileged[:-1]}

    @property
    def next_observation(self):
        return {"state": self.obs_state[1:], "privileged_state": self.obs_privileged[1:]}


class RolloutCollector:
    def __init__(self, wenv, policy: PolicyNet, unroll_length: int = 20, seed: int = 0):
        """`wenv`: a `wrapper.TrainingEnv` that has been reset; `policy`: a `PolicyNet` with parameters set."""
        self.wenv, self.policy, self.T, self.seed = wenv, policy, int(unroll_length), int(seed)
        self.env = wenv.unwrapped
        abi = self.env._abi
        if abi is None:
            raise RuntimeError("reset the env before building a RolloutCollector")
        self.abi, self.lib, torch = abi, abi.lib, abi.torch
        self.torch = torch
        N, T, dev = abi.N, self.T, abi.torch_device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        self.buf = Rollout(f(T + 1, N, 171), f(T + 1, N, 215), f(T, N, 12), f(T, N, 12), f(T, N), f(T, N), f(T, N), f(T, N))
        self._act_out = [{"action": self.buf.action[t], "raw_action": self.buf.raw_action[t], "log_prob": self.buf.log_prob[t]} for t in range(T)]

    def _store(self, src, dst, slot):
        n = src.numel()
        stream = C.c_void_p(self.torch.cuda.current_stream(self.abi.torch_device).cuda_stream)
        rc = self.lib.pgtt_store_slot(src.data_ptr(), dst.data_ptr(), slot, n, stream)
        if rc:
            raise nat.PgttError(rc, self.lib.pgtt_policy_last_error().decode())

    def collect(self, state, deterministic: bool = False) -> tuple:
        """Runs `unroll_length` control steps from `state`; returns (final state, Rollout)."""
        b, buf = self.abi.buf, self.buf
        self._store(b["obs_state"], buf.obs_state, 0)
        self._store(b["obs_privileged"], buf.obs_privileged, 0)
        for t in range(self.T):
            self.policy.act(buf.obs_state[t], seed=self.seed, deterministic=deterministic, out=self._act_out[t])
            self.abi.step_ptr(buf.action[t].data_ptr(), wrapped=True)
            self._store(b["obs_state"], buf.obs_state, t + 1)
            self._store(b["obs_privileged"], buf.obs_privileged, t + 1)
            self._store(b["reward"], buf.reward, t)
            self._store(b["done"], buf.discount, t)
            self._store(b["truncation"], buf.truncation, t)
        buf.discount.neg_().add_(1.0)        # discount = 1 - done
        return self.env._live_state(), buf

    def launches_per_collect(self) -> int:
        return 2 + self.T * 7
