#!/usr/bin/env python
"""Generates tests/golden/task_layer_*.npz by EXECUTING the reference's own task code.

The reference (`/root/reference/go2/*.py`) is pure Python over third-party packages that are not
installable here (jax, mujoco/mjx, mujoco_playground, ml_collections, etils - SURVEY.md 8c). This
script imports the reference modules UNMODIFIED from /root/reference with small stand-ins for
those packages:

  jax.numpy / jax.vmap / jax.jit   numpy-backed (float32-only arrays with `.at[].set()`)
  jax.random                       phase_guided_terrain_traversal_b200.prng (threefry2x32, pinned by the
                                   Random123 / JAX known-answer vectors in tests/test_golden.py)
  mjx_env.init/step, mjx.forward,  the CPU oracle (oracle/pgtt_oracle.c, fp32) - MJX itself is absent, so
  mjx.ray, contact list            the PHYSICS in these fixtures is the oracle's, not MuJoCo's
  collision.geoms_colliding        restated from mujoco_playground._src.collision (8 lines)
  mjx math helpers                 axis_angle_to_quat, quat_mul

and then runs `Joystick.reset` / `Joystick.step` (go2/joystick_pgtt.py:50-231) exactly as the
training loop would. What the fixtures pin is therefore the reference's TASK LAYER as written by its
authors: observation layout and noise wiring, the 21 reward terms and their scaling / clipping, gait
reference, heightscan grid geometry and quadrant statistics, contact-flag bookkeeping, command
resampling, history rolls, info updates. They do not pin MJX physics (parity unpinned there).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
"""
from __future__ import annotations

import functools
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from oracle.oracle import Oracle  # noqa: E402
from phase_guided_terrain_traversal_b200 import config_dict as our_config_dict  # noqa: E402
from phase_guided_terrain_traversal_b200 import model as gm  # noqa: E402
from phase_guided_terrain_traversal_b200 import prng, terrain  # noqa: E402

PART = True  # jax_threefry_partitionable


# ------------------------------------------------------------------------------------------------
# numpy stand-in for jax.numpy: float32 / int32 only, functional .at[].set()
# ------------------------------------------------------------------------------------------------
def _narrow(a):
    a = np.asarray(a)
    if a.dtype == np.float64:
        a = a.astype(np.float32)
    elif a.dtype == np.int64:
        a = a.astype(np.int32)
    return a


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Op:
            def set(self, v):
                out = np.array(arr, copy=True)
                out[idx] = v
                return J(out)

            def add(self, v):
                out = np.array(arr, copy=True)
                out[idx] += v
                return J(out)
        return _Op()


class JArr(np.ndarray):
    __array_priority__ = 100

    @property
    def at(self):
        return _At(self)

    def __array_wrap__(self, obj, context=None, return_scalar=False):
        out = _narrow(np.asarray(obj)).view(JArr)
        return out[()] if return_scalar and out.ndim == 0 else out

    def astype(self, dtype, *a, **k):
        return J(np.asarray(self).astype(dtype, *a, **k))


def J(x):
    return _narrow(x).view(JArr)


def _wrap(fn):
    @functools.wraps(fn)
    def f(*a, **k):
        out = fn(*a, **k)
        if isinstance(out, (list, tuple)):
            return type(out)(J(o) for o in out)
        return J(out)
    return f


jp = types.ModuleType("jax.numpy")
for _name in ["zeros", "ones", "zeros_like", "ones_like", "full", "arange", "meshgrid", "stack", "concatenate", "hstack", "vstack",
              "where", "sum", "square", "exp", "abs", "sqrt", "clip", "min", "max", "maximum", "minimum", "mean", "std", "cos", "sin",
              "fmod", "round", "roll", "any", "all", "dot", "cross", "argmin", "linspace", "squeeze", "ravel", "reshape", "sign", "log"]:
    setattr(jp, _name, _wrap(getattr(np, _name)))
jp.array = lambda x, dtype=None: J(np.array(x, dtype=dtype))
jp.asarray = jp.array
jp.pi = np.pi
jp.float32, jp.int32, jp.bool_ = np.float32, np.int32, np.bool_
jp.ndarray = np.ndarray
jp.linalg = types.SimpleNamespace(norm=_wrap(np.linalg.norm))


def _vmap(fn, in_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.asarray(a).shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            outs.append(fn(*[a if ax is None else (a[i] if ax == 0 else np.take(a, i, axis=ax)) for a, ax in zip(args, axes)]))
        return J(np.stack([np.asarray(o) for o in outs]))
    return mapped


def _jit(fn=None, **kw):
    return fn if fn is not None else (lambda f: f)


class _Random(types.ModuleType):
    @staticmethod
    def PRNGKey(seed):
        return prng.PRNGKey(seed)

    @staticmethod
    def split(key, num=2):
        return prng.split(np.asarray(key), num, PART)

    @staticmethod
    def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
        return J(prng.uniform(np.asarray(key), shape, np.asarray(minval), np.asarray(maxval), PART))

    @staticmethod
    def exponential(key, shape=()):
        return J(prng.exponential(np.asarray(key), PART))

    @staticmethod
    def bernoulli(key, p=0.5, shape=()):
        return J(prng.bernoulli(np.asarray(key), np.asarray(p), shape, PART))


def install_shims(physics):
    """Registers the stand-in packages. `physics` is the PhysicsShim the mjx calls are routed to."""
    jax = types.ModuleType("jax")
    jax.numpy = jp
    jax.Array = np.ndarray
    jax.vmap = _vmap
    jax.jit = _jit
    jax.random = _Random("jax.random")
    jax.debug = types.SimpleNamespace(print=lambda *a, **k: None)
    jsst = types.ModuleType("jax.scipy.spatial.transform")
    from scipy.spatial.transform import Rotation as _R

    class Rotation:
        def __init__(self, r):
            self.r = r

        @staticmethod
        def from_quat(q):
            return Rotation(_R.from_quat(np.asarray(q, dtype=np.float64)))

        def as_euler(self, seq):
            return J(self.r.as_euler(seq).astype(np.float32))
    jsst.Rotation = Rotation
    mods = {"jax": jax, "jax.numpy": jp, "jax.random": jax.random, "jax.scipy": types.ModuleType("jax.scipy"),
            "jax.scipy.spatial": types.ModuleType("jax.scipy.spatial"), "jax.scipy.spatial.transform": jsst}

    ml = types.ModuleType("ml_collections")
    ml.config_dict = our_config_dict
    mods["ml_collections"] = ml
    mods["ml_collections.config_dict"] = our_config_dict

    etils = types.ModuleType("etils")
    epath = types.ModuleType("etils.epath")
    epath.Path = Path
    etils.epath = epath
    mods["etils"] = etils
    mods["etils.epath"] = epath

    mujoco = types.ModuleType("mujoco")
    mjx = types.ModuleType("mujoco.mjx")
    mjx.Data = object
    mjx.Model = object
    mjx.forward = physics.forward
    mjx.ray = physics.ray
    mujoco.mjx = mjx
    mujoco.MjModel = object
    src = types.ModuleType("mujoco.mjx._src")
    mmath = types.ModuleType("mujoco.mjx._src.math")

    def axis_angle_to_quat(axis, angle):
        s, c = np.sin(np.asarray(angle, np.float32) * np.float32(0.5)), np.cos(np.asarray(angle, np.float32) * np.float32(0.5))
        return J(np.concatenate([np.atleast_1d(c), np.asarray(axis, np.float32) * s]).astype(np.float32))

    def quat_mul(u, v):
        u, v = np.asarray(u, np.float32), np.asarray(v, np.float32)
        return J(np.array([u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3], u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2],
                           u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1], u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0]], np.float32))
    mmath.axis_angle_to_quat, mmath.quat_mul = axis_angle_to_quat, quat_mul
    src.math = mmath
    mods.update({"mujoco": mujoco, "mujoco.mjx": mjx, "mujoco.mjx._src": src, "mujoco.mjx._src.math": mmath})

    mp = types.ModuleType("mujoco_playground")
    mps = types.ModuleType("mujoco_playground._src")
    mjx_env = types.ModuleType("mujoco_playground._src.mjx_env")

    class State:
        def __init__(self, data, obs, reward, done, metrics, info):
            self.data, self.obs, self.reward, self.done, self.metrics, self.info = data, obs, reward, done, metrics, info

        def replace(self, **kw):
            s = State(self.data, self.obs, self.reward, self.done, self.metrics, self.info)
            for k, v in kw.items():
                setattr(s, k, v)
            return s

    class MjxEnv:
        def __init__(self, config, config_overrides=None):
            self._config = config

        @property
        def dt(self):
            return self._config.ctrl_dt

        @property
        def sim_dt(self):
            return self._config.sim_dt

        @property
        def n_substeps(self):
            return int(round(self.dt / self.sim_dt))
    mjx_env.State, mjx_env.MjxEnv = State, MjxEnv
    mjx_env.init, mjx_env.step = physics.init, physics.step
    mjx_env.get_sensor_data = physics.get_sensor_data
    mjx_env.update_assets = lambda *a, **k: None
    coll = types.ModuleType("mujoco_playground._src.collision")

    def geoms_colliding(data, geom1, geom2):
        """mujoco_playground._src.collision.geoms_colliding / get_collision_info restated."""
        geom = np.asarray(data.contact.geom)
        mask = (np.array([geom1, geom2]) == geom).all(axis=1) | (np.array([geom2, geom1]) == geom).all(axis=1)
        idx = np.where(mask, np.asarray(data.contact.dist), 1e4).argmin()
        dist = np.asarray(data.contact.dist)[idx] * mask[idx]
        return J(np.asarray(dist < 0))
    coll.geoms_colliding = geoms_colliding
    mps.mjx_env, mps.collision = mjx_env, coll
    mp._src = mps
    mods.update({"mujoco_playground": mp, "mujoco_playground._src": mps, "mujoco_playground._src.mjx_env": mjx_env,
                 "mujoco_playground._src.collision": coll})
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    mods.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    sys.modules.update(mods)
    return State


# ------------------------------------------------------------------------------------------------
# physics stand-in: one oracle env behind the mjx / mjx_env call surface
# ------------------------------------------------------------------------------------------------
class Contact:
    pass


class Data:
    """The mjx.Data fields the task reads, refreshed from the oracle after each call."""

    def replace(self, **kw):
        d = Data()
        d.__dict__.update(self.__dict__)
        d.__dict__.update(kw)
        return d


class PhysicsShim:
    """Installed ONCE (the reference modules bind these functions at import); `bind` swaps the oracle env."""

    def __init__(self):
        self.orc, self.m = None, None

    def bind(self, orc: Oracle, model: gm.Go2Model):
        self.orc, self.m = orc, model

    def _pull(self) -> Data:
        o = self.orc
        d = Data()
        f32 = lambda name: J(o.get(name)[0].astype(np.float32))
        d.qpos, d.qvel, d.ctrl = f32("qpos"), f32("qvel"), f32("ctrl")
        d.sensordata, d.actuator_force = f32("sensordata"), f32("actuator_force")
        d.site_xpos = J(o.get("site_xpos")[0].astype(np.float32).reshape(5, 3))
        xm = o.get("site_xmat")[0].astype(np.float32).reshape(3, 3)
        d.site_xmat = J(np.stack([xm] * 5))
        d.xfrc_applied = J(np.zeros((14 + self.m.n_boxes, 6), np.float32))
        f, k = o.contacts(0)
        c = Contact()
        c.dist = J(f[:, 0].astype(np.float32))
        c.geom = J(k[:, :2].astype(np.int32))
        d.contact = c
        return d

    def _push(self, data: Data):
        self.orc.set("qpos", np.asarray(data.qpos, np.float64))
        self.orc.set("qvel", np.asarray(data.qvel, np.float64))

    def init(self, model, qpos, qvel, ctrl):
        self.orc.set("qpos", np.asarray(qpos, np.float64))
        self.orc.set("qvel", np.asarray(qvel, np.float64))
        self.orc.set("ctrl", np.asarray(ctrl, np.float64))
        self.orc.set("qacc_warmstart", np.zeros(18))
        self.orc.forward()
        return self._pull()

    def forward(self, model, data):
        self._push(data)
        self.orc.forward()
        return self._pull()

    def step(self, model, data, action, n_substeps):
        self._push(data)
        for _ in range(n_substeps):
            self.orc.set("ctrl", np.asarray(action, np.float64))
            self.orc.physics_step()
        return self._pull()

    def ray(self, model, data, pnt, vec, geomgroup=None):
        assert tuple(geomgroup) == (1, 0, 0, 0, 1, 1)
        dist = self.orc.ray(0, np.asarray(pnt, np.float64), np.asarray(vec, np.float64))
        return J(np.float32(dist)), -1

    def get_sensor_data(self, model, data, name):
        a, n = self.m.sensor_adr[name]
        return data.sensordata[a:a + n]


# ------------------------------------------------------------------------------------------------
def make_reference_env(task, cfg, model, State, method="pgtt"):
    """Joystick instance of the REFERENCE class, constructed without MuJoCo: the attributes
    Go2Env.__init__ (go2/base.py:45-113) derives from the compiled MJCF are filled from model.py."""
    sys.path.insert(0, str(REF))
    if method == "baseline":
        import go2.joystick as ref_joy          # the non-phase comparison task (train.py:111-114)
    else:
        import go2.joystick_pgtt as ref_joy
    env = object.__new__(ref_joy.Joystick)
    env._config = cfg
    env._mj_model = model
    env._mjx_model = types.SimpleNamespace(nv=18, nu=12, nbody=14 + model.n_boxes)
    env._imu_site_id = 0
    env._init_q = J(model.home_qpos)
    env._default_pose = J(model.home_qpos[7:])
    env.init_feet_pos = J(np.zeros((4, 3)))
    env._lowers, env._uppers = J(model.jnt_range[:, 0]), J(model.jnt_range[:, 1])
    env._soft_lowers = env._lowers * cfg.soft_joint_pos_limit_factor
    env._soft_uppers = env._uppers * cfg.soft_joint_pos_limit_factor
    env._torso_body_id = 1
    env._feet_site_id = np.array([2, 1, 4, 3])      # FR FL RR RL (site ids: imu 0, FL 1, FR 2, RL 3, RR 4)
    env._floor_geom_id = np.concatenate([[model.floor_geom_id], model.box_geom_id0 + np.arange(model.n_boxes)]).astype(np.int32)
    env._feet_geom_id = np.array([32, 20, 56, 44], np.int32)
    adr = [list(range(*(lambda a_n: (a_n[0], a_n[0] + a_n[1]))(model.sensor_adr[f"{s}_global_linvel"]))) for s in ["FR_foot", "FL_foot", "RR_foot", "RL_foot"]]
    env._foot_linvel_sensor_adr = np.array(adr)
    env._cmd_u_max, env._cmd_u_min = J(np.array(cfg.command_config.u_max, np.float32)), J(np.array(cfg.command_config.u_min, np.float32))
    env._cmd_b = J(np.array(cfg.command_config.b, np.float32))
    env._xml_path = "n/a"
    return env


FIELDS_INFO = ["command", "step", "steps_until_next_cmd", "phase", "phase_dt", "gait_freq", "last_act", "last_last_act", "feet_air_time",
               "last_contact", "swing_peak", "H_max", "H_min", "heightscan", "motor_targets", "qpos_error_history", "qvel_history", "rng"]


def snapshot(state, out, tag, reward_keys):
    out[f"{tag}/obs_state"] = np.asarray(state.obs["state"], np.float32)
    out[f"{tag}/obs_privileged"] = np.asarray(state.obs["privileged_state"], np.float32)
    out[f"{tag}/reward"] = np.float32(state.reward)
    out[f"{tag}/done"] = np.float32(state.done)
    out[f"{tag}/metrics"] = np.array([np.float32(state.metrics[f"reward/{k}"]) for k in reward_keys] + [np.float32(state.metrics["swing_peak"])], np.float32)
    for k in FIELDS_INFO:
        v = np.asarray(state.info[k])
        out[f"{tag}/info/{k}"] = v.astype(np.uint32) if k == "rng" else (v.astype(np.int32) if v.dtype.kind in "bi" else v.astype(np.float32))
    out[f"{tag}/qpos"] = np.asarray(state.data.qpos, np.float32)
    out[f"{tag}/qvel"] = np.asarray(state.data.qvel, np.float32)


def run_case(name, task, level, dr, seeds, n_steps, out_dir, method="pgtt"):
    from phase_guided_terrain_traversal_b200.go2.configs import baseline_config, default_config, training_overrides
    from phase_guided_terrain_traversal_b200._native import REWARD_KEYS
    cfg = training_overrides(baseline_config() if method == "baseline" else default_config())   # train.py:119-129 for both methods
    variant = int(method == "baseline")
    model = gm.compile_model(task)
    table = terrain.load_terrain(level) if task == "stairs" else None
    out = {"meta/task": task, "meta/level": level or "", "meta/dr": int(dr), "meta/seeds": np.array(seeds), "meta/n_steps": n_steps,
           "meta/partitionable": int(PART), "meta/variant": variant}
    for seed in seeds:
        orc = Oracle(model, cfg, 1, "f32", rng_partitionable=PART, variant=variant)
        dr_key = prng.env_keys(11, 1, offset=seed)
        orc.randomize(dr_key, table, bool(dr))
        PHYSICS.bind(orc, model)
        env = make_reference_env(task, cfg, model, STATE, method)
        reset_key = prng.env_keys(12, 1, offset=seed)[0]
        state = env.reset(reset_key)
        tag = f"seed{seed}"
        out[f"{tag}/dr_key"], out[f"{tag}/reset_key"] = dr_key[0], reset_key
        out[f"{tag}/terrain_index"] = np.int32(orc.get("terrain_index")[0, 0])
        snapshot(state, out, f"{tag}/reset", REWARD_KEYS)
        g = np.random.default_rng(1000 + seed)
        acts = g.uniform(-1, 1, (n_steps, 12)).astype(np.float32)
        out[f"{tag}/actions"] = acts
        for s in range(n_steps):
            state = env.step(state, J(acts[s]))
            snapshot(state, out, f"{tag}/step{s}", REWARD_KEYS)
    path = Path(out_dir) / f"task_layer_{name}.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, f"{path.stat().st_size / 1024:.0f} KB")


PHYSICS = PhysicsShim()
STATE = install_shims(PHYSICS)


def extra_vectors(out_dir):
    """Small closed-form pieces of the reference evaluated directly: gait.get_z, quat_to_yaw, grid geometry."""
    sys.path.insert(0, str(REF))
    import go2.gait as ref_gait
    from go2.utility import quat_to_yaw as ref_yaw
    phi = np.linspace(0, 2 * np.pi, 257, dtype=np.float32)[:-1]
    hs = np.array([-0.2, -0.1, 0.05], np.float32)
    z = np.stack([np.asarray(ref_gait.get_z(J(phi), swing_height=np.float32(h), swing_min=np.float32(-0.3))) for h in hs])
    g = np.random.default_rng(5)
    q = g.normal(size=(64, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    yaw = np.array([np.float32(ref_yaw(J(qq))) for qq in q], np.float32)
    # Joystick.sample_command (go2/joystick_pgtt.py:603-611) of the reference class on 32 keys
    from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
    cfg = training_overrides(default_config())
    ref_env = make_reference_env("flat_terrain", cfg, gm.compile_model("flat_terrain"), STATE)
    ckeys = prng.env_keys(77, 32)
    cx = g.uniform(-1, 1, (32, 3)).astype(np.float32)
    cmd = np.stack([np.asarray(ref_env.sample_command(ckeys[i], J(cx[i])), np.float32) for i in range(32)])
    np.savez_compressed(Path(out_dir) / "closed_form.npz", phi=phi, swing_heights=hs, get_z=z.astype(np.float32), quat=q, yaw=yaw,
                        phases=np.asarray(ref_gait.PHASES, np.float32), p_stance=np.float32(ref_gait.p_stance),
                        cmd_keys=ckeys, cmd_x=cx, cmd_out=cmd)
    print("wrote closed_form.npz")
    # the reference's config factories, evaluated (go2/configs.py:6-152) and its constants (go2/go2_constants.py)
    import json
    import go2.configs as ref_cfg
    import go2.go2_constants as ref_consts
    blob = {"default_config": ref_cfg.default_config().to_dict(), "baseline_config": ref_cfg.baseline_config().to_dict(),
            "constants": {k: getattr(ref_consts, k) for k in ["FEET_SITES", "FEET_GEOMS", "FEET_POS_SENSOR", "ROOT_BODY", "UPVECTOR_SENSOR",
                                                             "GLOBAL_LINVEL_SENSOR", "GLOBAL_ANGVEL_SENSOR", "LOCAL_LINVEL_SENSOR",
                                                             "ACCELEROMETER_SENSOR", "GYRO_SENSOR", "num_heightscans", "num_widthscans", "dist_x", "dist_y"]},
            "task_to_xml": {t: ref_consts.task_to_xml(t).as_posix() for t in ["flat_terrain", "stairs"]}}
    (Path(out_dir) / "reference_config.json").write_text(json.dumps(blob, indent=1, sort_keys=True, default=list))
    print("wrote reference_config.json")


def policy_vectors(out_dir):
    """policy177 (one of the two policies policy_folder/README.md:7 calls best): weights + running statistics as
    loaded by policy_io.load_policy, and the deployment forward `tanh(loc)` computed by the REFERENCE network
    class deploy/policy_net.py:35-64 (torch) on obs drawn around the normaliser statistics."""
    import importlib.util
    import torch
    from phase_guided_terrain_traversal_b200 import policy_io
    d = policy_io.load_policy(REF / "policy_folder" / "policy177")
    spec = importlib.util.spec_from_file_location("ref_policy_net", REF / "deploy" / "policy_net.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ks, bs = d["policy"]
    net = ref.MLP(ks, bs, torch.nn.SiLU(), d["mean"], d["std"])
    g = np.random.default_rng(7)
    obs = (d["mean"] + d["std"] * g.normal(size=(64, 171)).clip(-3, 3)).astype(np.float32)
    with torch.no_grad():
        act = net(torch.from_numpy(obs)).numpy()
    np.savez_compressed(Path(out_dir) / "policy177.npz", obs=obs, action_deterministic=act.astype(np.float32), mean=d["mean"], std=d["std"],
                        count=np.float64(d["count"]), **{f"kernel{i}": k for i, k in enumerate(ks)}, **{f"bias{i}": b for i, b in enumerate(bs)},
                        priv_mean=d["value_mean"], priv_std=d["value_std"])
    print("wrote policy177.npz")


def terrain_vectors(out_dir):
    """terrain/generator.py + terrain/getIndexes.py of the reference, imported unmodified (cv2 / noise / alive_progress / jax are
    only imported at module level there and get empty stand-ins): the 14-tile adjacency tables `generate_14` hands to the WFC
    solver, and the boxes `addElement` creates for every tile kind at three (width, step height, steps) settings."""
    import json
    import os
    for name in ("cv2", "noise", "alive_progress"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["alive_progress"].alive_bar = lambda *a, **k: None
    sys.modules["alive_progress"].alive_it = lambda x: x
    sys.path.insert(0, str(REF / "terrain"))
    sys.path.insert(0, str(REF))
    cwd = os.getcwd()
    os.chdir(REF)                 # the generator parses ./go2/xmls/scene_mjx_feetonly.xml in its constructor
    try:
        import generator as ref_gen
        import getIndexes as ref_idx
        left, right, up, down = (-1, 0), (1, 0), (0, 1), (0, -1)
        directions = [left, down, right, up]
        conn = {0: {left: (0, 4, 10, 11), down: (0, 5, 11, 12), right: (0, 2, 12, 13), up: (0, 3, 13, 10)},      # generator.py:295-298
                1: {left: (1, 2, 6, 7), down: (1, 3, 7, 8), right: (1, 4, 8, 9), up: (1, 5, 9, 6)}}
        conn.update(ref_idx.Stairs(directions)); conn.update(ref_idx.StairsTurningUp(directions)); conn.update(ref_idx.StairsTurningDown(directions))
        table = {str(t): {f"{d[0]},{d[1]}": sorted(int(x) for x in v) for d, v in per.items()} for t, per in conn.items()}
        (Path(out_dir) / "terrain_adjacency.json").write_text(json.dumps(table, indent=1, sort_keys=True))
        out = {}
        settings = [(0.3, 0.07, 2), (0.41, 0.13, 3), (0.45, 0.04, 4)]
        out["settings"] = np.array(settings)
        for si, (w, h, n) in enumerate(settings):
            for tile in range(14):
                tgen = ref_gen.TerrainGenerator(width=w, step_height=h, num_stairs=n, render=False)
                ref_gen.addElement(tgen, tile, [0.7, -1.1])
                rows = [np.concatenate([b["pos"], b["quat"], b["size"]]) for b in tgen.box_data]
                out[f"s{si}_tile{tile}"] = np.array(rows, np.float64).reshape(-1, 10)
        np.savez_compressed(Path(out_dir) / "terrain_tiles.npz", **out)
    finally:
        os.chdir(cwd)
    print("wrote terrain_adjacency.json, terrain_tiles.npz")


if __name__ == "__main__":
    out_dir = Path(__file__).resolve().parent
    terrain_vectors(out_dir)
    policy_vectors(out_dir)
    extra_vectors(out_dir)
    run_case("flat", "flat_terrain", None, dr=1, seeds=[0, 1], n_steps=12, out_dir=out_dir)
    run_case("stairs_level07", "stairs", "level07", dr=1, seeds=[0, 1, 2], n_steps=12, out_dir=out_dir)
    run_case("stairs_level1_nodr", "stairs", "level1", dr=0, seeds=[3], n_steps=8, out_dir=out_dir)
    run_case("baseline_stairs_level07", "stairs", "level07", dr=1, seeds=[0, 1], n_steps=12, out_dir=out_dir, method="baseline")
