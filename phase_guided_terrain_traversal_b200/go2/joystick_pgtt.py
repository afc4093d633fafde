"""`Joystick` - the phase-guided GO2 joystick task (go2/joystick_pgtt.py:35-611), batched.

Same constructor, `reset(rng)`, `step(state, action)` and `State` layout as the reference; the
arithmetic (4 physics substeps, contact flags, ray grid, gait phase, observation, 21 reward terms,
command resampling, history rolls) runs in ONE fused CUDA launch per call
(csrc/pgtt_env.cuh:env_step / env_reset) reached through the C ABI. There is no CPU fallback: without
the compiled library or without a CUDA device construction of the handle raises.

Batching: the reference env is single-env and is vmapped by the Playground wrapper; here the env is
natively batched. `reset` takes `uint32[N,2]` keys (one jax-style key per env), a single key, or an
int seed together with `num_envs`; `step` takes `action[N,12]`.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Union

import numpy as np

from .. import prng
from . import go2_constants as consts
from .base import Go2Env, State
from .configs import default_config


class Joystick(Go2Env):
    """Track a joystick command."""

    def __init__(self, task: str = "flat_terrain", config=None, config_overrides: Optional[Dict[str, Union[str, int, list]]] = None,
                 *, num_envs: Optional[int] = None, device: int = 0, rng_partitionable: bool = True):
        super().__init__(xml_path=consts.task_to_xml(task).as_posix(), config=config if config is not None else default_config(),
                         config_overrides=config_overrides, task=task, num_envs=num_envs, device=device, rng_partitionable=rng_partitionable)

    # -- go2/joystick_pgtt.py:50-131 -----------------------------------------------------------------
    def reset(self, rng) -> State:
        keys = prng.as_keys(rng, self._num_envs, self._rng_partitionable)
        abi = self._ensure_handle(keys.shape[0])
        self._check_terrain()
        abi.reset(keys)
        return self._live_state()

    # -- go2/joystick_pgtt.py:141-231 ----------------------------------------------------------------
    def step(self, state: State, action) -> State:
        return self._step(state, action, wrapped=False)

    def _step(self, state: State, action, wrapped: bool) -> State:
        abi = self._abi
        if abi is None:
            raise RuntimeError("step() before reset()")
        if state is not None and not (state._live and state._owner is self):
            self.set_state(state)
        torch = abi.torch
        if not isinstance(action, torch.Tensor):
            action = torch.as_tensor(np.asarray(action, dtype=np.float32))
        if action.device != abi.torch_device or action.dtype != torch.float32 or not action.is_contiguous():
            action = action.to(device=abi.torch_device, dtype=torch.float32, non_blocking=True).contiguous()
        if tuple(action.shape) != (abi.N, 12):
            raise ValueError(f"action must have shape ({abi.N}, 12), got {tuple(action.shape)}")
        abi._keep = [action]
        abi.step_ptr(action.data_ptr(), wrapped=wrapped)
        return self._live_state()

    def _check_terrain(self):
        if self._mj_model.n_boxes > 0 and not self._randomized:
            raise RuntimeError("task 'stairs': the terrain only enters through domain_randomize(model, rng, terrain_matrix) "
                               "(go2/randomize.py:97-108); apply it first, e.g. via wrap_for_brax_training(..., randomization_fn=...)")

    # -- domain randomisation hook (used by randomize.domain_randomize / the training wrapper) ----------
    def apply_randomization(self, keys, terrain_matrix=None, dynamics: bool = True):
        keys = prng.as_keys(keys, self._num_envs, self._rng_partitionable)
        abi = self._ensure_handle(keys.shape[0])
        if self._mj_model.n_boxes > 0:
            if terrain_matrix is None:
                raise ValueError("task 'stairs' needs a terrain_matrix [T,100,10]")
            abi.set_terrain(np.asarray(terrain_matrix, dtype=np.float32))
        abi.randomize(keys, dynamics=dynamics)
        self._randomized = True
        return {k: abi.buf[k] for k in ("body_mass", "body_ipos_base", "dof_armature", "dof_damping", "actuator_gain", "actuator_bias1", "qpos0",
                                        "box_friction", "floor_friction", "terrain_index")}

    # -- go2/joystick_pgtt.py:603-611 (host restatement for inspection; the kernel samples in-place) --------
    def sample_command(self, rng, x_k):
        """Host restatement for a single env (numpy, bit-compatible with jax.random); inside `step` the fused kernel
        resamples in place (csrc/pgtt_env.cuh:env_step). rng: uint32[2], x_k: float[3] -> float32[3]."""
        part = self._rng_partitionable
        _, y_rng, w_rng, z_rng = prng.split(np.asarray(rng, dtype=np.uint32), 4, part)
        y_k = prng.uniform(y_rng, (3,), self._cmd_u_min, self._cmd_u_max, part)
        z_k = prng.bernoulli(z_rng, self._cmd_b, (3,), part).astype(np.float32)
        w_k = prng.bernoulli(w_rng, 0.5, (3,), part).astype(np.float32)
        x_k = np.asarray(x_k, dtype=np.float32)
        return (x_k - w_k * (x_k - y_k * z_k)).astype(np.float32)
