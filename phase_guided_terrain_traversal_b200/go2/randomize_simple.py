"""`domain_randomize(model, rng)` for the flat scene (go2/randomize_simple.py:24-138): same dynamics
draws as go2/randomize.py minus box frictions / terrain, plus a live floor-friction draw."""
from __future__ import annotations

import numpy as np

from .randomize import RandomizedModel

IN_AXES_FIELDS = ["geom_friction", "body_ipos", "body_mass", "qpos0", "dof_frictionloss", "dof_armature", "dof_damping",
                  "actuator_gainprm", "actuator_biasprm"]


def domain_randomize(model, rng, dynamics: bool = True):
    rng = rng.detach().cpu().numpy() if hasattr(rng, "detach") else np.asarray(rng)
    rng = np.ascontiguousarray(rng, dtype=np.uint32).reshape(-1, 2)
    return RandomizedModel(model, rng, None, dynamics), {k: 0 for k in IN_AXES_FIELDS}
