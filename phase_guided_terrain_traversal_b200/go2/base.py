"""`Go2Env` + `State`: host-side mirror of go2/base.py:45-231 and of `mjx_env.State`.

The reference class compiles the MJCF with MuJoCo, overrides time step and PD gains
(go2/base.py:53-68) and offers sensor getters on a single-env `mjx.Data`. Here the model comes from
`model.compile_model` (same overrides), physics lives in the CUDA library behind the C ABI
(include/pgtt_b200.h) and every array has a leading `N` (num_envs) axis: the batching that the
reference adds with `jax.vmap` in the Playground wrapper is native to the kernels.

Value semantics: JAX returns a fresh `State` per call; here `State` fields are zero-copy views of
device buffers owned by the env handle and are updated IN PLACE by `reset` / `step`. A `State` from
an earlier step therefore aliases the current one; use `State.clone()` to keep a snapshot and
`env.set_state(snapshot)` (or pass the snapshot to `step`) to go back to it.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np

from .. import _native as nat
from .. import model as gm
from .. import prng
from ..abi_env import AbiEnv
from . import go2_constants as consts

# `info` dict keys (go2/joystick_pgtt.py:101-120) -> ABI buffer names
INFO_FIELDS = {
    "rng": "rng", "command": "command", "step": "step", "steps_until_next_cmd": "steps_until_next_cmd",
    "phase": "phase", "phase_dt": "phase_dt", "gait_freq": "gait_freq", "last_act": "last_act",
    "last_last_act": "last_last_act", "feet_air_time": "feet_air_time", "last_contact": "last_contact",
    "swing_peak": "swing_peak", "H_max": "H_max", "heightscan": "heightscan", "H_min": "H_min",
    "motor_targets": "motor_targets", "qpos_error_history": "qpos_error_history", "qvel_history": "qvel_history",
    # keys the brax / playground wrappers add
    "steps": "steps", "truncation": "truncation", "episode_done": "episode_done", "episode_metrics": "episode_metrics",
    # extras this implementation exposes (this step's contact flags, base.py:153-171)
    "contact": "contact", "first_contact": "first_contact",
}
SCALAR_INFO = {"step", "steps_until_next_cmd", "phase_dt", "gait_freq", "steps", "truncation", "episode_done"}
DATA_FIELDS = ["qpos", "qvel", "qacc", "qacc_warmstart", "ctrl", "time", "sensordata", "actuator_force", "site_xpos", "site_xmat",
               "contact_dist", "contact_geom", "solver_niter"]
METRIC_KEYS = [f"reward/{k}" for k in nat.REWARD_KEYS] + ["swing_peak"]


class Data:
    """The `mjx.Data` fields the task reads, as `[N, ...]` views."""

    def __init__(self, fields: Dict[str, Any]):
        self.__dict__.update(fields)

    def _fields(self):
        return dict(self.__dict__)

    def replace(self, **kw):
        f = self._fields()
        f.update(kw)
        return Data(f)


@dataclass
class State:
    """mjx_env.State(data, obs, reward, done, metrics, info) with a leading N on every leaf."""
    data: Data
    obs: Dict[str, Any]
    reward: Any
    done: Any
    metrics: Dict[str, Any]
    info: Dict[str, Any]
    _owner: Any = field(default=None, repr=False, compare=False)
    _live: bool = field(default=False, repr=False, compare=False)

    def replace(self, **kw) -> "State":
        s = copy.copy(self)
        for k, v in kw.items():
            setattr(s, k, v)
        return s

    def clone(self) -> "State":
        """Deep snapshot (device copies); no longer aliases the env's buffers."""
        c = lambda t: t.clone()
        return State(Data({k: c(v) for k, v in self.data._fields().items()}), {k: c(v) for k, v in self.obs.items()}, c(self.reward), c(self.done),
                     {k: c(v) for k, v in self.metrics.items()}, {k: c(v) for k, v in self.info.items()}, self._owner, False)


class Go2Env:
    """Base class for the GO2 (go2/base.py:45-113)."""

    def __init__(self, xml_path: str, config, config_overrides: Optional[Dict[str, Any]] = None, task: Optional[str] = None,
                 num_envs: Optional[int] = None, device: int = 0, rng_partitionable: bool = True, variant: int = 0):
        self._config = copy.deepcopy(config)
        self._variant = int(variant)   # 0 = go2/joystick_pgtt.py, 1 = go2/joystick.py (baseline task)
        if config_overrides:
            self._config.update_from_flattened_dict(config_overrides)
        self._xml_path = xml_path
        self._task = task
        # base.py:57-62: timestep <- sim_dt, damping <- Kd, gain <- Kp, bias[1] <- -Kp
        self._mj_model = gm.compile_model(task, sim_dt=self._config.sim_dt, Kp=self._config.Kp, Kd=self._config.Kd)
        self._mjx_model = self._mj_model
        self._default_pose = np.array(self._mj_model.home_qpos[7:], dtype=np.float32)
        lim = np.array(self._mj_model.jnt_range, dtype=np.float64) * self._config.soft_joint_pos_limit_factor  # base.py:76-79 (Q13)
        self._lowers, self._uppers = lim[:, 0], lim[:, 1]
        self._cmd_u_max = np.array(self._config.command_config.u_max)
        self._cmd_u_min = np.array(self._config.command_config.u_min)
        self._cmd_b = np.array(self._config.command_config.b)
        # ids the reference caches (base.py:81-113); values follow the compiled MJCF
        self._imu_site_id = 0
        self._feet_site_id = np.array([2, 1, 4, 3])          # FR FL RR RL
        self._feet_geom_id = np.array([32, 20, 56, 44])      # FR FL RR RL
        self._floor_geom_id = np.concatenate([[self._mj_model.floor_geom_id], self._mj_model.box_geom_id0 + np.arange(self._mj_model.n_boxes)])
        self._device = int(device)
        self._rng_partitionable = bool(rng_partitionable)
        self._num_envs = num_envs
        self._episode_length = None     # set by the training wrapper before the handle exists
        self._abi: Optional[AbiEnv] = None
        self._state: Optional[State] = None
        self._randomized = False

    # -- handle management ---------------------------------------------------------------------
    def _ensure_handle(self, n: int) -> AbiEnv:
        if self._abi is not None and self._abi.N == n:
            return self._abi
        if self._abi is not None:
            self._abi.close()
        cfg = copy.deepcopy(self._config)
        if self._episode_length is not None:
            cfg.episode_length = int(self._episode_length)
        self._abi = AbiEnv(self._mj_model, cfg, n, device=self._device, backend="torch", rng_partitionable=self._rng_partitionable, variant=self._variant)
        self._num_envs = n
        self._randomized = False
        self._state = None
        return self._abi

    @property
    def num_envs(self) -> Optional[int]:
        return self._num_envs

    def close(self):
        if self._abi is not None:
            self._abi.close()
            self._abi = None

    def _live_state(self) -> State:
        if self._state is None:
            b = self._abi.buf
            data = Data({k: b[k] for k in DATA_FIELDS})
            data.site_xpos = b["site_xpos"].view(-1, 5, 3)
            data.site_xmat = b["site_xmat"].view(-1, 3, 3)    # imu site (all robot sites share body frames only for imu; feet: see sensors)
            obs = {"state": b["obs_state"], "privileged_state": b["obs_privileged"]}
            metrics = {k: b["metrics"][:, i] for i, k in enumerate(METRIC_KEYS)}
            info = {}
            for k, name in INFO_FIELDS.items():
                t = b[name]
                info[k] = t[:, 0] if k in SCALAR_INFO else t
            info["heightscan"] = b["heightscan"].view(-1, consts.num_heightscans, consts.num_widthscans, 3)
            self._state = State(data, obs, b["reward"][:, 0], b["done"][:, 0], metrics, info, self, True)
        return self._state

    def set_state(self, state: State) -> None:
        """Copy a (cloned) State back into the handle's buffers."""
        live = self._live_state()
        if state is live or (state._live and state._owner is self):
            return
        for k, v in state.data._fields().items():
            getattr(live.data, k).copy_(v.reshape(getattr(live.data, k).shape))
        for k, v in state.obs.items():
            live.obs[k].copy_(v)
        live.reward.copy_(state.reward)
        live.done.copy_(state.done)
        for k, v in state.metrics.items():
            live.metrics[k].copy_(v)
        for k, v in state.info.items():
            if k in live.info:
                live.info[k].copy_(v.reshape(live.info[k].shape))

    # -- sensor readings (base.py:115-151); all [N, ...] -------------------------------------------
    def _sens(self, data: Data, name: str):
        a, n = self._mj_model.sensor_adr[name]
        return data.sensordata[:, a:a + n]

    def get_upvector(self, data):
        return self._sens(data, consts.UPVECTOR_SENSOR)

    def get_gravity(self, data):
        return -data.site_xmat[:, 2, :]      # imu_xmat^T (0,0,-1)

    def get_global_linvel(self, data):
        return self._sens(data, consts.GLOBAL_LINVEL_SENSOR)

    def get_global_angvel(self, data):
        return self._sens(data, consts.GLOBAL_ANGVEL_SENSOR)

    def get_local_linvel(self, data):
        return self._sens(data, consts.LOCAL_LINVEL_SENSOR)

    def get_accelerometer(self, data):
        return self._sens(data, consts.ACCELEROMETER_SENSOR)

    def get_gyro(self, data):
        return self._sens(data, consts.GYRO_SENSOR)

    def get_feet_pos(self, data):
        a, _ = self._mj_model.sensor_adr[consts.FEET_POS_SENSOR[0]]
        return data.sensordata[:, a:a + 12].reshape(-1, 4, 3)

    def get_yaw(self, data):
        from .utility import quat_to_yaw
        return quat_to_yaw(data.qpos[:, 3:7])

    def compute_contact(self, data):
        """4 feet x (floor + boxes) any-collision flags, order FR FL RR RL (base.py:153-171):
        a pair present in the contact list with dist < 0. Evaluated on the `data` contact list."""
        import torch
        g, d = data.contact_geom.view(-1, 8, 2), data.contact_dist
        feet = torch.as_tensor(self._feet_geom_id, device=g.device, dtype=g.dtype)
        is_foot = (g[:, :, None, 0] == feet) | (g[:, :, None, 1] == feet)        # [N,8,4]
        return (is_foot & (d < 0)[:, :, None] & (g[:, :, None, 0] >= 0)).any(1)

    # -- accessors (base.py:217-231 + MjxEnv) --------------------------------------------------------
    @property
    def xml_path(self) -> str:
        return self._xml_path

    @property
    def action_size(self) -> int:
        return gm.NU

    @property
    def mj_model(self):
        return self._mj_model

    @property
    def mjx_model(self):
        return self._mjx_model

    @property
    def dt(self) -> float:
        return self._config.ctrl_dt

    @property
    def sim_dt(self) -> float:
        return self._config.sim_dt

    @property
    def n_substeps(self) -> int:
        return int(round(self.dt / self.sim_dt))

    @property
    def observation_size(self):
        nobs = nat.BUFFER_FIELDS[13][2] - (9 if self._variant else 0)
        return {"state": (nobs,), "privileged_state": (nobs + 44,)}

    @property
    def unwrapped(self):
        return self
