"""`Joystick` - the non-phase BASELINE joystick task (go2/joystick.py:35-611), batched.

The comparison method of the reference (`training/train.py --method baseline`, train.py:111-114,119-122). Differences
from the phase-guided task (`diff go2/joystick.py go2/joystick_pgtt.py`): observations carry no gait phase and no
gait frequency (162 / 206 instead of 171 / 215, joystick.py:333-341), `H_max` is the quadrant maximum (:186), the
clearance cost uses the world-frame foot height (:569-572), the air-time reward threshold is 0.5 s (:591), and the
default config is `baseline_config()` (go2/configs.py:82-152). Everything else - physics, ray grid, contacts,
commands, wrappers - is the same fused kernel with `variant = 1`.
"""
from __future__ import annotations

from typing import Dict, Optional, Union

from . import go2_constants as consts
from . import joystick_pgtt
from .configs import baseline_config, default_config  # noqa: F401  (re-exported like the reference module)


class Joystick(joystick_pgtt.Joystick):
    """Track a joystick command (baseline task)."""

    def __init__(self, task: str = "flat_terrain", config=None, config_overrides: Optional[Dict[str, Union[str, int, list]]] = None,
                 *, num_envs: Optional[int] = None, device: int = 0, rng_partitionable: bool = True):
        joystick_pgtt.Go2Env.__init__(self, xml_path=consts.task_to_xml(task).as_posix(), config=config if config is not None else baseline_config(),
                                      config_overrides=config_overrides, task=task, num_envs=num_envs, device=device,
                                      rng_partitionable=rng_partitionable, variant=1)
