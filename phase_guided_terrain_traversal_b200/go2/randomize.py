"""`domain_randomize(model, rng, terrain_matrix)` (go2/randomize.py:23-171).

Reference: returns `(batched mjx.Model, in_axes)` for the Playground vmap wrapper. Here the per-env
model fields live in device buffers of the env handle and are drawn by `pgtt_randomize`
(csrc/pgtt_env.cuh:env_randomize) from the same per-env key streams, in the same split order. The
return value keeps the reference's shape - a pair `(batched_model, in_axes)` - where `batched_model`
is a `RandomizedModel` the training wrapper applies to the env; `in_axes` names the randomised fields
(axis 0) like go2/randomize.py:140-154.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import numpy as np

IN_AXES_FIELDS = ["geom_friction", "body_ipos", "body_mass", "qpos0", "dof_frictionloss", "dof_armature", "dof_damping",
                  "actuator_gainprm", "actuator_biasprm", "geom_size", "body_pos", "body_quat"]


@dataclass
class RandomizedModel:
    model: Any
    rng: np.ndarray                       # uint32 [N,2]
    terrain_matrix: Optional[np.ndarray]  # float32 [T,100,10] or None (flat)
    dynamics: bool = True                 # False = terrain assignment only ("no DR", BASELINE config 2)
    fields: Optional[dict] = None         # per-env device views, filled when applied to an env

    def apply(self, env):
        self.fields = env.apply_randomization(self.rng, self.terrain_matrix, self.dynamics)
        return self


def domain_randomize(model, rng, terrain_matrix, dynamics: bool = True):
    rng = rng.detach().cpu().numpy() if hasattr(rng, "detach") else np.asarray(rng)
    rng = np.ascontiguousarray(rng, dtype=np.uint32).reshape(-1, 2)
    t = np.asarray(terrain_matrix, dtype=np.float32)
    if t.ndim != 3 or t.shape[1:] != (100, 10):
        raise ValueError(f"terrain_matrix must be [T,100,10], got {t.shape}")
    in_axes = {k: 0 for k in IN_AXES_FIELDS}
    return RandomizedModel(model, rng, t, dynamics), in_axes
