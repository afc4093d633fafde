"""PPO rollout collector: brax `generate_unroll` / `actor_step` as driven by training/train.py:135-161,242-263.

Per control step: `obs["state"]` -> fused tcgen05 policy kernel (normalise, MLP, tanh-normal sample) ->
fused env step kernel -> transition write-out. Transitions are stored time-major, `[T, N, ...]`, the
layout the learner consumes (brax `Transition(observation, action, reward, discount, next_observation,
extras{policy_extras{log_prob, raw_action}, state_extras{truncation}})`); observations are kept as
`[T + 1, N, ...]` so `next_observation[t] = observation[t + 1]` (what the auto-reset wrapper returns).
Everything stays on the device: `collect` is ONE native call (`pgtt_rollout`) that issues the three kernels
per control step (policy, env step, transition record) on the caller's stream.
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
        self.buf = Rollout(f(T + 1, N, abi.nobs), f(T + 1, N, abi.npriv), f(T, N, 12), f(T, N, 12), f(T, N), f(T, N), f(T, N), f(T, N))
        self._step = 0      # exploration-noise counter: advances by T per collect

    def collect(self, state=None, deterministic: bool = False) -> tuple:
        """Runs `unroll_length` control steps from the env's current state; returns (final state, Rollout).
        One native call (`pgtt_rollout`): 1 + T x (policy + the step's one or two) kernel launches on the current stream, nothing returns to the host."""
        buf = self.buf
        if state is not None and not (state._live and state._owner is self.env):
            self.env.set_state(state)
        rb = nat.RolloutBuffers(*(C.c_void_p(t.data_ptr()) for t in (buf.obs_state, buf.obs_privileged, buf.action, buf.raw_action, buf.log_prob,
                                                                     buf.reward, buf.discount, buf.truncation)))
        stream = C.c_void_p(self.torch.cuda.current_stream(self.abi.torch_device).cuda_stream)
        rc = self.lib.pgtt_rollout(self.abi.h, self.policy.h, self.T, C.c_uint64(self.seed), C.c_uint64(self._step), int(deterministic), C.byref(rb), stream)
        if rc:
            raise nat.PgttError(rc, (self.lib.pgtt_policy_last_error() or self.lib.pgtt_last_error()).decode())
        self._step += self.T
        return self.env._live_state(), buf

    def launches_per_collect(self) -> int:
        return 1 + self.T * (1 + self.abi.step_launches())
