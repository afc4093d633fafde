"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY - never imported by the product).

PARITY UNPINNED: see oracle/pgtt_oracle.h. Allowed importers: tests/, __graft_entry__.smoke(),
bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIBS: dict = {}

REWARD_KEYS = [
    "tracking_lin_vel", "tracking_ang_vel", "lin_vel_z", "ang_vel_xy", "orientation", "dof_pos_limits", "pose",
    "termination", "stand_still", "torques", "action_rate", "energy", "feet_clearance", "feet_height", "feet_slip",
    "feet_air_time", "feet_phase", "feet_swing", "body_height", "contact", "center",
]


def build(force: bool = False) -> None:
    """Compile both precisions with gcc (seconds). Safe to call repeatedly."""
    outs = [HERE / "_build" / "liborc_f64.so", HERE / "_build" / "liborc_f32.so"]
    src_m = max((HERE / "pgtt_oracle.c").stat().st_mtime, (HERE / "pgtt_oracle.h").stat().st_mtime)
    if force or not all(o.exists() and o.stat().st_mtime >= src_m for o in outs):
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s", "all"], check=True, capture_output=True)


def build_native() -> Path:
    """fp32 oracle compiled for THIS host (-O3 -march=native): the CPU-baseline build of bench.py. Built in-tree
    (oracle/_build/liborc_f32_native.so, git-ignored) so that whoever audits the process sees which library the CPU arm
    ran. Always rebuilt (a library built on another machine may use instructions this CPU lacks); written under a
    temporary name and renamed, so concurrent builders never load a half-written file."""
    import os
    bdir = HERE / "_build"
    bdir.mkdir(exist_ok=True)
    out = bdir / "liborc_f32_native.so"
    tmp = bdir / f".liborc_f32_native_{os.getpid()}.so"
    subprocess.run(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-fopenmp", "-std=gnu11", "-DORC_F32", "-o", str(tmp),
                    str(HERE / "pgtt_oracle.c"), "-lm"], check=True, capture_output=True)
    os.replace(tmp, out)
    return out


def _lib(precision: str):
    if precision not in _LIBS:
        if precision == "f32native":
            lib = ctypes.CDLL(str(build_native()))
        else:
            build()
            lib = ctypes.CDLL(str(HERE / "_build" / f"liborc_{precision}.so"))
        lib.orc_create.restype = ctypes.c_void_p
        lib.orc_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        lib.orc_destroy.argtypes = [ctypes.c_void_p]
        lib.orc_randomize.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.orc_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.orc_step_envs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.orc_physics.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.orc_scan.argtypes = [ctypes.c_void_p] * 4
        lib.orc_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        lib.orc_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        lib.orc_field_count.argtypes = [ctypes.c_char_p]
        lib.orc_ray.restype = ctypes.c_double
        lib.orc_ray.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.orc_get_contacts.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.orc_rng_split.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.orc_rng_uniform.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
        lib.orc_rng_bits.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.orc_rng_randint.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.orc_threefry2x32.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        lib.orc_gait_get_z.restype = ctypes.c_double
        lib.orc_gait_get_z.argtypes = [ctypes.c_double] * 3
        lib.orc_quat_to_yaw.restype = ctypes.c_double
        lib.orc_quat_to_yaw.argtypes = [ctypes.c_void_p]
        _LIBS[precision] = lib
    return _LIBS[precision]


def pack_model(m, n_model_bodies: int | None = None) -> np.ndarray:
    """Flatten a `model.Go2Model` in the order `orc_create` consumes."""
    nb = m.n_boxes
    park = np.zeros((100, 10))
    park[:, 3] = 1.0
    park[:, 7:10] = 1.0
    park[:, 0:3] = 1000.0  # unused slots (flat scene): far away, never selected
    if nb:
        park[:nb] = m.box_park
    parts = [
        [m.timestep], m.gravity, [m.impratio, m.tolerance, m.ls_tolerance, m.meaninertia],
        [m.iterations, m.ls_iterations, m.max_geom_pairs, m.max_contact_points, nb],
        m.body_parent, m.body_pos, m.body_quat, m.body_ipos, m.body_iquat, m.body_mass, m.body_inertia, m.body_invweight0,
        m.jnt_body, m.jnt_axis, m.jnt_range, m.jnt_solref, m.jnt_solimp,
        m.qpos0, m.dof_armature, m.dof_damping, m.dof_invweight0,
        m.act_dof, m.act_gainprm[:, 0], m.act_biasprm, m.act_ctrlrange, m.act_forcerange,
        m.foot_body, m.foot_geom_id, m.foot_pos, [m.foot_radius], m.foot_friction, m.foot_solref, m.foot_solimp, [m.foot_margin],
        [m.floor_geom_id, m.box_geom_id0], m.floor_friction, m.floor_solref, m.floor_solimp,
        [m.box_rbound], getattr(m, "box_solref", [0.02, 1.0]), getattr(m, "box_solimp", [0.9, 0.95, 0.001, 0.5, 2.0]), m.box_friction,
        park[:, 0:3], park[:, 3:7], park[:, 7:10], m.imu_pos,
        [n_model_bodies if n_model_bodies is not None else 14 + nb],
    ]
    return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])


def pack_task(cfg, m, rng_partitionable: bool = True, variant: int = 0) -> np.ndarray:
    """Flatten the task config (`go2.configs.default_config()` layout) for `orc_create`."""
    n = cfg.noise_config
    r = cfg.reward_config
    parts = [
        [cfg.ctrl_dt, cfg.action_scale, n.level],
        [n.scales.joint_pos, n.scales.joint_vel, n.scales.gyro, n.scales.gravity, n.scales.linvel, n.scales.heightscan],
        [r.scales[k] for k in REWARD_KEYS], [r.tracking_sigma, r.swing_height, r.base_feet_distance, r.phase_sigma],
        cfg.command_config.u_max, cfg.command_config.u_min, cfg.command_config.b, cfg.gait_freq,
        [cfg.soft_joint_pos_limit_factor], m.home_qpos[7:], m.home_qpos,
        [cfg.history_update_steps, cfg.episode_length, int(round(cfg.ctrl_dt / cfg.sim_dt)), int(rng_partitionable), int(variant)],
    ]
    return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])


class Oracle:
    """N independent CPU envs. All I/O as float64 / int numpy arrays with leading N."""

    def __init__(self, model, cfg, n_envs: int, precision: str = "f32", rng_partitionable: bool = True, variant: int = 0):
        self.lib = _lib(precision)
        self.n = n_envs
        self.model, self.cfg = model, cfg
        mc = pack_model(model)
        tc = pack_task(cfg, model, rng_partitionable, variant)
        self.variant = int(variant)
        self.h = self.lib.orc_create(n_envs, mc.ctypes.data, mc.size, tc.ctypes.data, tc.size)
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.part = int(rng_partitionable)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def get(self, name: str) -> np.ndarray:
        c = self.lib.orc_field_count(name.encode())
        if c < 0:
            raise KeyError(name)
        out = np.zeros((self.n, c))
        self.lib.orc_get(self.h, name.encode(), out.ctypes.data)
        return out

    def set(self, name: str, value) -> None:
        c = self.lib.orc_field_count(name.encode())
        if c < 0:
            raise KeyError(name)
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=np.float64).reshape(-1, c), (self.n, c)))
        self.lib.orc_set(self.h, name.encode(), v.ctypes.data)

    def randomize(self, keys, terrain=None, dynamics: bool = True) -> None:
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(self.n, 2)
        if terrain is not None:
            terrain = np.ascontiguousarray(terrain, dtype=np.float32)
            assert terrain.shape[1:] == (100, 10)
            self.lib.orc_randomize(self.h, keys.ctypes.data, terrain.ctypes.data, terrain.shape[0], int(dynamics))
        else:
            self.lib.orc_randomize(self.h, keys.ctypes.data, None, 0, int(dynamics))

    def reset(self, keys) -> None:
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(self.n, 2)
        self.lib.orc_reset(self.h, keys.ctypes.data)

    def step(self, action, wrapped: bool = True) -> None:
        a = np.ascontiguousarray(action, dtype=np.float64).reshape(self.n, 12)
        self.lib.orc_step_envs(self.h, a.ctypes.data, int(wrapped))

    def forward(self) -> None:
        self.lib.orc_physics(self.h, 0)

    def physics_step(self) -> None:
        self.lib.orc_physics(self.h, 1)

    def scan(self, center, yaw) -> np.ndarray:
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(center, dtype=np.float64).reshape(-1, 3), (self.n, 3)))
        y = np.ascontiguousarray(np.broadcast_to(np.asarray(yaw, dtype=np.float64).reshape(-1), (self.n,)))
        out = np.zeros((self.n, 117, 3))
        self.lib.orc_scan(self.h, c.ctypes.data, y.ctypes.data, out.ctypes.data)
        return out.reshape(self.n, 13, 9, 3)

    def ray(self, i: int, pnt, vec) -> float:
        p = np.ascontiguousarray(pnt, dtype=np.float64); v = np.ascontiguousarray(vec, dtype=np.float64)
        return float(self.lib.orc_ray(self.h, i, p.ctypes.data, v.ctypes.data))

    def contacts(self, i: int):
        f = np.zeros((8, 17))
        k = np.zeros((8, 5), dtype=np.int32)
        self.lib.orc_get_contacts(self.h, i, f.ctypes.data, k.ctypes.data)
        return f, k


# ---- jax.random probes ----------------------------------------------------------------------
def threefry2x32(key, x0, x1, precision="f64"):
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(2, dtype=np.uint32)
    _lib(precision).orc_threefry2x32(k.ctypes.data, int(x0), int(x1), out.ctypes.data)
    return out


def rng_split(key, num=2, partitionable=True):
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros((num, 2), dtype=np.uint32)
    _lib("f64").orc_rng_split(int(partitionable), k.ctypes.data, num, out.ctypes.data)
    return out


def rng_uniform(key, n, lo=0.0, hi=1.0, partitionable=True, precision="f32"):
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(n)
    _lib(precision).orc_rng_uniform(int(partitionable), k.ctypes.data, n, lo, hi, out.ctypes.data)
    return out


def rng_bits(key, n, partitionable=True):
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(n, dtype=np.uint32)
    _lib("f64").orc_rng_bits(int(partitionable), k.ctypes.data, n, out.ctypes.data)
    return out


def rng_randint(key, lo, hi, partitionable=True):
    k = np.asarray(key, dtype=np.uint32)
    return _lib("f64").orc_rng_randint(int(partitionable), k.ctypes.data, lo, hi)
