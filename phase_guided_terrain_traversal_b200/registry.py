"""Environment registry with the call surface training/train.py:116-130,165-170,231 uses:
`register_environment`, `get_default_config`, `load`, `get_domain_randomizer` and the `_randomizer` dict
(mujoco_playground.registry / mujoco_playground.locomotion in the reference)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

_envs: Dict[str, Callable] = {}
_cfgs: Dict[str, Callable] = {}
_randomizer: Dict[str, Optional[Callable]] = {}


def register_environment(env_name: str, env_class: Callable, cfg_class: Callable) -> None:
    _envs[env_name] = env_class
    _cfgs[env_name] = cfg_class


def get_default_config(env_name: str):
    if env_name not in _cfgs:
        raise ValueError(f"Env '{env_name}' not found in default configs.")
    return _cfgs[env_name]()


def load(env_name: str, config=None, config_overrides: Optional[Dict[str, Any]] = None, **kw):
    if env_name not in _envs:
        raise ValueError(f"Env '{env_name}' not found. Available envs: {sorted(_envs)}")
    config = config if config is not None else get_default_config(env_name)
    return _envs[env_name](config=config, config_overrides=config_overrides, **kw)


def get_domain_randomizer(env_name: str) -> Optional[Callable]:
    return _randomizer.get(env_name)


ALL_ENVS = _envs


def _register_defaults():
    import functools
    from .go2 import joystick_pgtt
    from .go2.configs import default_config
    register_environment("Go2JoystickFlatTerrain", functools.partial(joystick_pgtt.Joystick, task="flat_terrain"), default_config)
    register_environment("Go2JoystickStairs", functools.partial(joystick_pgtt.Joystick, task="stairs"), default_config)


_register_defaults()
