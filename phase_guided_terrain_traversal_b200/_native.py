"""ctypes binding of the C ABI in include/pgtt_b200.h (libpgtt_b200.so, CUDA, sm_100a).

There is deliberately no fallback: if the shared library is missing or fails to load the import of
any compute entry point raises `NativeLibraryError` (build it with `python __graft_entry__.py` or
`build_library()`; nvcc cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = CSRC / "libpgtt_b200.so"
INCLUDE = PKG_DIR.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


class NativeLibraryError(RuntimeError):
    pass


class PgttError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pgtt error {code}: {msg}")
        self.code = code


# --------------------------------------------------------------------------------------
# struct mirrors (keep in sync with include/pgtt_b200.h)
# --------------------------------------------------------------------------------------
d, i32 = C.c_double, C.c_int


class ModelDesc(C.Structure):
    _fields_ = [
        ("timestep", d), ("gravity", d * 3), ("impratio", d), ("tolerance", d), ("ls_tolerance", d), ("meaninertia", d),
        ("iterations", i32), ("ls_iterations", i32), ("max_geom_pairs", i32), ("max_contact_points", i32), ("n_boxes", i32),
        ("body_pos", d * 3 * 14), ("body_ipos", d * 3 * 14), ("body_iquat", d * 4 * 14), ("body_mass", d * 14), ("body_inertia", d * 3 * 14),
        ("body_invweight0", d * 2 * 14),
        ("jnt_range", d * 2 * 12), ("jnt_solref", d * 2), ("jnt_solimp", d * 5),
        ("qpos0", d * 19), ("dof_armature", d * 18), ("dof_damping", d * 18), ("dof_invweight0", d * 18),
        ("act_dof", i32 * 12),
        ("act_gain", d * 12), ("act_bias", d * 3 * 12), ("act_ctrlrange", d * 2 * 12), ("act_forcerange", d * 2 * 12),
        ("foot_geom_id", i32 * 4),
        ("foot_pos", d * 3), ("foot_radius", d), ("foot_friction", d * 3), ("foot_solref", d * 2), ("foot_solimp", d * 5), ("foot_margin", d),
        ("floor_geom_id", i32), ("box_geom_id0", i32),
        ("floor_friction", d * 3), ("floor_solref", d * 2), ("floor_solimp", d * 5),
        ("box_rbound", d), ("box_friction", d * 3), ("box_solref", d * 2), ("box_solimp", d * 5),
        ("imu_pos", d * 3),
        ("n_model_bodies", i32),
    ]


class TaskDesc(C.Structure):
    _fields_ = [
        ("ctrl_dt", d), ("action_scale", d), ("noise_level", d),
        ("noise_joint_pos", d), ("noise_joint_vel", d), ("noise_gyro", d), ("noise_gravity", d), ("noise_linvel", d), ("noise_heightscan", d),
        ("reward_scale", d * 21),
        ("tracking_sigma", d), ("swing_height", d), ("base_feet_distance", d), ("phase_sigma", d),
        ("cmd_u_max", d * 3), ("cmd_u_min", d * 3), ("cmd_b", d * 3), ("gait_freq", d * 2),
        ("soft_limit_factor", d),
        ("default_pose", d * 12), ("home_qpos", d * 19),
        ("history_update_steps", i32), ("episode_length", i32), ("n_substeps", i32), ("rng_partitionable", i32), ("variant", i32),
    ]


# name -> (ctype of element, trailing dim); order == struct pgtt_buffers after num_envs
BUFFER_FIELDS = [
    ("qpos", "f", 19), ("qvel", "f", 18), ("qacc", "f", 18), ("qacc_warmstart", "f", 18), ("ctrl", "f", 12), ("time", "f", 1),
    ("sensordata", "f", 49), ("actuator_force", "f", 12), ("site_xpos", "f", 15), ("site_xmat", "f", 9),
    ("contact_dist", "f", 8), ("contact_geom", "i", 16), ("solver_niter", "i", 4),
    ("obs_state", "f", 171), ("obs_privileged", "f", 215), ("reward", "f", 1), ("done", "f", 1), ("metrics", "f", 22),
    ("rng", "u", 2), ("command", "f", 3), ("step", "i", 1), ("steps_until_next_cmd", "i", 1),
    ("phase", "f", 4), ("phase_dt", "f", 1), ("gait_freq", "f", 1), ("last_act", "f", 12), ("last_last_act", "f", 12), ("feet_air_time", "f", 4),
    ("last_contact", "i", 4),
    ("swing_peak", "f", 4), ("H_max", "f", 4), ("H_min", "f", 4), ("heightscan", "f", 351), ("motor_targets", "f", 12),
    ("qpos_error_history", "f", 24), ("qvel_history", "f", 24),
    ("contact", "i", 4), ("first_contact", "i", 4),
    ("steps", "f", 1), ("truncation", "f", 1), ("episode_done", "f", 1), ("episode_metrics", "f", 24),
    ("first_qpos", "f", 19), ("first_qvel", "f", 18), ("first_qacc_warmstart", "f", 18), ("first_obs_state", "f", 171), ("first_obs_privileged", "f", 215),
    ("body_mass", "f", 13), ("body_ipos_base", "f", 3), ("dof_armature", "f", 12), ("dof_damping", "f", 12),
    ("actuator_gain", "f", 12), ("actuator_bias1", "f", 12), ("qpos0", "f", 12), ("box_friction", "f", 100), ("floor_friction", "f", 1),
    ("terrain_index", "i", 1),
]
_CT = {"f": C.c_float, "i": C.c_int32, "u": C.c_uint32}
NP_DTYPE = {"f": np.float32, "i": np.int32, "u": np.uint32}


class Buffers(C.Structure):
    _fields_ = [("num_envs", i32)] + [(n, C.POINTER(_CT[k])) for n, k, _ in BUFFER_FIELDS]


class RolloutBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs_state", "obs_privileged", "action", "raw_action", "log_prob", "reward", "discount", "truncation")]


REWARD_KEYS = [
    "tracking_lin_vel", "tracking_ang_vel", "lin_vel_z", "ang_vel_xy", "orientation", "dof_pos_limits", "pose",
    "termination", "stand_still", "torques", "action_rate", "energy", "feet_clearance", "feet_height", "feet_slip",
    "feet_air_time", "feet_phase", "feet_swing", "body_height", "contact", "center",
]
DEBUG_FLOATS = 2048


def _fill(arr, values):
    v = np.asarray(values, dtype=np.float64)
    flat = v.ravel()
    dst = np.ctypeslib.as_array(arr).reshape(-1)
    assert dst.size == flat.size, (dst.size, flat.size)
    dst[:] = flat


def model_desc(m) -> ModelDesc:
    """Pack a `model.Go2Model` (checks that it is inside the supported GO2 family)."""
    hip_axes = m.jnt_axis[0::3]
    assert np.allclose(hip_axes, [1, 0, 0]) and np.allclose(m.jnt_axis[1::3], [0, 1, 0]) and np.allclose(m.jnt_axis[2::3], [0, 1, 0]), \
        "kernels assume x-axis abduction and y-axis hip / knee joints"
    assert np.allclose(m.body_quat, [1, 0, 0, 0]), "kernels assume identity body frames"
    assert list(m.jnt_body) == list(range(2, 14)) and list(m.foot_body) == [4, 7, 10, 13]
    md = ModelDesc()
    md.timestep, md.impratio, md.tolerance, md.ls_tolerance, md.meaninertia = m.timestep, m.impratio, m.tolerance, m.ls_tolerance, m.meaninertia
    _fill(md.gravity, m.gravity)
    md.iterations, md.ls_iterations, md.max_geom_pairs, md.max_contact_points, md.n_boxes = \
        m.iterations, m.ls_iterations, m.max_geom_pairs, m.max_contact_points, m.n_boxes
    _fill(md.body_pos, m.body_pos); _fill(md.body_ipos, m.body_ipos); _fill(md.body_iquat, m.body_iquat)
    _fill(md.body_mass, m.body_mass); _fill(md.body_inertia, m.body_inertia); _fill(md.body_invweight0, m.body_invweight0)
    _fill(md.jnt_range, m.jnt_range); _fill(md.jnt_solref, m.jnt_solref); _fill(md.jnt_solimp, m.jnt_solimp)
    _fill(md.qpos0, m.qpos0); _fill(md.dof_armature, m.dof_armature); _fill(md.dof_damping, m.dof_damping); _fill(md.dof_invweight0, m.dof_invweight0)
    for a in range(12):
        md.act_dof[a] = int(m.act_dof[a])
    _fill(md.act_gain, m.act_gainprm[:, 0]); _fill(md.act_bias, m.act_biasprm); _fill(md.act_ctrlrange, m.act_ctrlrange); _fill(md.act_forcerange, m.act_forcerange)
    for g in range(4):
        md.foot_geom_id[g] = int(m.foot_geom_id[g])
    _fill(md.foot_pos, m.foot_pos); md.foot_radius = m.foot_radius; _fill(md.foot_friction, m.foot_friction)
    _fill(md.foot_solref, m.foot_solref); _fill(md.foot_solimp, m.foot_solimp); md.foot_margin = m.foot_margin
    md.floor_geom_id, md.box_geom_id0 = m.floor_geom_id, m.box_geom_id0
    _fill(md.floor_friction, m.floor_friction); _fill(md.floor_solref, m.floor_solref); _fill(md.floor_solimp, m.floor_solimp)
    md.box_rbound = m.box_rbound; _fill(md.box_friction, m.box_friction)
    _fill(md.box_solref, [0.02, 1.0]); _fill(md.box_solimp, [0.9, 0.95, 0.001, 0.5, 2.0])
    _fill(md.imu_pos, m.imu_pos)
    md.n_model_bodies = 14 + m.n_boxes
    return md


def task_desc(cfg, m, rng_partitionable: bool = True, variant: int = 0) -> TaskDesc:
    td = TaskDesc()
    n, r = cfg.noise_config, cfg.reward_config
    td.ctrl_dt, td.action_scale, td.noise_level = cfg.ctrl_dt, cfg.action_scale, n.level
    td.noise_joint_pos, td.noise_joint_vel, td.noise_gyro = n.scales.joint_pos, n.scales.joint_vel, n.scales.gyro
    td.noise_gravity, td.noise_linvel, td.noise_heightscan = n.scales.gravity, n.scales.linvel, n.scales.heightscan
    _fill(td.reward_scale, [r.scales[k] for k in REWARD_KEYS])
    td.tracking_sigma, td.swing_height, td.base_feet_distance, td.phase_sigma = r.tracking_sigma, r.swing_height, r.base_feet_distance, r.phase_sigma
    _fill(td.cmd_u_max, cfg.command_config.u_max); _fill(td.cmd_u_min, cfg.command_config.u_min); _fill(td.cmd_b, cfg.command_config.b)
    _fill(td.gait_freq, cfg.gait_freq)
    td.soft_limit_factor = cfg.soft_joint_pos_limit_factor
    _fill(td.default_pose, m.home_qpos[7:]); _fill(td.home_qpos, m.home_qpos)
    td.history_update_steps, td.episode_length = cfg.history_update_steps, cfg.episode_length
    td.n_substeps = int(round(cfg.ctrl_dt / cfg.sim_dt))
    td.rng_partitionable = int(rng_partitionable)
    td.variant = int(variant)
    return td


# --------------------------------------------------------------------------------------
# library loading
# --------------------------------------------------------------------------------------
def declare(lib):
    vp = C.c_void_p
    lib.pgtt_last_error.restype = C.c_char_p
    lib.pgtt_version.restype = C.c_int
    lib.pgtt_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(TaskDesc), C.c_int, C.c_int, C.POINTER(vp)]
    lib.pgtt_destroy.argtypes = [vp]
    lib.pgtt_sync.argtypes = [vp, vp]
    lib.pgtt_set_terrain_table.argtypes = [vp, vp, C.c_int]
    lib.pgtt_randomize.argtypes = [vp, vp, C.c_int, vp]
    lib.pgtt_reset.argtypes = [vp, vp, vp]
    lib.pgtt_step.argtypes = [vp, vp, C.c_int, vp]
    lib.pgtt_step_record.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.pgtt_forward.argtypes = [vp, vp]
    lib.pgtt_heightscan.argtypes = [vp, vp, vp, vp, vp]
    lib.pgtt_get_buffers.argtypes = [vp, C.POINTER(Buffers)]
    lib.pgtt_obs_dims.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.pgtt_debug_forward.argtypes = [vp, vp, vp]
    lib.pgtt_launch_count.argtypes = [vp]
    lib.pgtt_launch_count.restype = C.c_int64
    lib.pgtt_record.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.pgtt_step_kernel_generation.argtypes = [vp]
    if hasattr(lib, "pgtt_policy_create"):     # absent from the host-emulated test library (env kernels only)
        lib.pgtt_policy_last_error.restype = C.c_char_p
        lib.pgtt_policy_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
        lib.pgtt_policy_destroy.argtypes = [vp]
        lib.pgtt_policy_set_params.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, vp]
        lib.pgtt_policy_set_params_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp]
        lib.pgtt_policy_act.argtypes = [vp, vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp, vp, vp, vp]
        lib.pgtt_policy_launch_count.argtypes = [vp]
        lib.pgtt_policy_launch_count.restype = C.c_int64
        lib.pgtt_rollout.argtypes = [vp, vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(RolloutBuffers), vp]
        lib.pgtt_gae.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp, vp, vp]
        lib.pgtt_col_moments_scratch_doubles.argtypes = [C.c_int]
        lib.pgtt_col_moments_scratch_doubles.restype = C.c_longlong
        lib.pgtt_col_moments.argtypes = [vp, C.c_longlong, C.c_int, C.c_int, vp, vp, vp]
        lib.pgtt_gae_sums.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]
        lib.pgtt_moments_finalize.argtypes = [vp, vp, vp]
        lib.pgtt_minibatch_gather.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.pgtt_gae_moments.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]
        lib.pgtt_ppo_head.argtypes = [vp] * 8 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]
        lib.pgtt_adam_clip.argtypes = [vp] * 6 + [C.c_longlong] + [C.c_float] * 6 + [vp]
        lib.pgtt_adam_scratch_floats.restype = C.c_int
        lib.pgtt_learner_last_error.restype = C.c_char_p
        lib.pgtt_linear_forward.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        lib.pgtt_linear_backward_input.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp]
        lib.pgtt_linear_backward_params_splits.argtypes = [C.c_int]
        lib.pgtt_linear_backward_params_scratch.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.pgtt_linear_backward_params_scratch.restype = C.c_longlong
        lib.pgtt_linear_backward_params.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        lib.pgtt_silu_backward.argtypes = [vp, vp, vp, C.c_longlong, vp]
        lib.pgtt_mlp_last_error.restype = C.c_char_p
        lib.pgtt_mlp_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(vp)]
        lib.pgtt_mlp_destroy.argtypes = [vp]
        lib.pgtt_mlp_destroy.restype = None
        lib.pgtt_mlp_rows.argtypes = [vp]
        lib.pgtt_mlp_forward.argtypes = [vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), vp, vp]
        lib.pgtt_mlp_backward.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), vp]
        lib.pgtt_mlp_forward_gather.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, vp]
        lib.pgtt_mlp_debug_trace.argtypes = [vp, vp, C.c_int]
    return lib


ABI_SYMBOLS = [
    "pgtt_last_error", "pgtt_version", "pgtt_create", "pgtt_destroy", "pgtt_sync", "pgtt_set_terrain_table", "pgtt_randomize",
    "pgtt_reset", "pgtt_step", "pgtt_step_record", "pgtt_forward", "pgtt_heightscan", "pgtt_get_buffers", "pgtt_obs_dims", "pgtt_debug_forward", "pgtt_launch_count", "pgtt_record", "pgtt_step_kernel_generation", "pgtt_step_launches",
    "pgtt_policy_last_error", "pgtt_policy_create", "pgtt_policy_destroy", "pgtt_policy_set_params", "pgtt_policy_set_params_device", "pgtt_policy_act",
    "pgtt_policy_launch_count", "pgtt_rollout", "pgtt_gae", "pgtt_gae_moments", "pgtt_gae_sums", "pgtt_moments_finalize", "pgtt_minibatch_gather", "pgtt_col_moments_scratch_doubles", "pgtt_col_moments", "pgtt_ppo_head", "pgtt_adam_clip", "pgtt_adam_scratch_floats",
    "pgtt_learner_last_error", "pgtt_linear_forward", "pgtt_linear_backward_input", "pgtt_linear_backward_params_splits", "pgtt_linear_backward_params_scratch", "pgtt_linear_backward_params", "pgtt_silu_backward",
    "pgtt_mlp_last_error", "pgtt_mlp_create", "pgtt_mlp_destroy", "pgtt_mlp_rows", "pgtt_mlp_forward", "pgtt_mlp_forward_gather", "pgtt_mlp_backward", "pgtt_mlp_debug_trace",
]

_LIB = None


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -> csrc/libpgtt_b200.so (in-tree, so it travels to the GPU box with the snapshot)."""
    srcs = sorted(CSRC.glob("*.cu*")) + sorted(CSRC.glob("*.h")) + [INCLUDE / "pgtt_b200.h"]      # every kernel source and header
    newest = max(s.stat().st_mtime for s in srcs)
    if not force and LIB_PATH.exists() and LIB_PATH.stat().st_mtime >= newest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB_PATH), str(CSRC / "pgtt_api.cu"), str(CSRC / "pgtt_policy.cu"), str(CSRC / "pgtt_learner.cu"), str(CSRC / "pgtt_mlp.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise NativeLibraryError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


def load_library(path: Path | None = None):
    global _LIB
    if path is None and _LIB is not None:
        return _LIB
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise NativeLibraryError(f"{p} not found - build it first (python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    try:
        lib = declare(C.CDLL(str(p)))
    except OSError as e:  # e.g. libcudart missing
        raise NativeLibraryError(f"cannot load {p}: {e}") from e
    if path is None:
        _LIB = lib
    return lib


def check(lib, rc: int):
    if rc != 0:
        raise PgttError(rc, lib.pgtt_last_error().decode())
