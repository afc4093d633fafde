"""`wrap_for_brax_training` - the wrapper stack of training/train.py:255,262.

Reference stack (mujoco_playground + brax): Brax(DomainRandomization)VmapWrapper ->
EpisodeWrapper -> BraxAutoResetWrapper. Here all three are folded into the fused step kernel
(`pgtt_step(..., wrapped=1)`, csrc/pgtt_env.cuh:env_step): per-env model fields are read from the
buffers `pgtt_randomize` filled, `steps/truncation/episode_metrics/episode_done` are kept in `info`,
and envs that finish restore the cached first data/obs (info is NOT reset, as in the reference).
"""
from __future__ import annotations

from typing import Callable, Optional

from .go2.base import State


class TrainingEnv:
    """What `wrap_for_brax_training(env, ...)` returns: `reset(rng[N,2])`, `step(state, action[N,12])`."""

    def __init__(self, env, episode_length: int = 1000, action_repeat: int = 1, randomization_fn: Optional[Callable] = None):
        if action_repeat != 1:
            raise NotImplementedError("action_repeat != 1 is not used by the reference (go2/configs.py:13) and not implemented")
        self.env = env
        self.episode_length = int(episode_length)
        self.action_repeat = action_repeat
        env._episode_length = self.episode_length
        if env._abi is not None and int(env._abi.cfg.episode_length) != self.episode_length:
            env.close()                                # handle is rebuilt with the new episode length
        self.randomized_model = None
        if randomization_fn is not None:
            out = randomization_fn(env.mjx_model)
            batched, self.in_axes = out if isinstance(out, tuple) else (out, None)
            self.randomized_model = batched.apply(env)

    def reset(self, rng) -> State:
        return self.env.reset(rng)

    def step(self, state: State, action) -> State:
        return self.env._step(state, action, wrapped=True)

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env


def wrap_for_brax_training(env, episode_length: int = 1000, action_repeat: int = 1, randomization_fn: Optional[Callable] = None) -> TrainingEnv:
    return TrainingEnv(env, episode_length, action_repeat, randomization_fn)


wrap_for_training = wrap_for_brax_training
