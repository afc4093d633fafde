"""Golden vectors produced by the reference's own task code (tests/golden/make_golden.py).

`task_layer_*.npz` hold reset + step outputs of the UNMODIFIED `go2/joystick_pgtt.py:Joystick` run over
stand-ins for its missing third-party packages (physics = the oracle, jax.random = prng.py). They pin
the task layer - obs layout / noise wiring, the 21 rewards, gait reference, ray-grid geometry and
quadrant statistics, contact flags, command resampling, info bookkeeping - of
  (a) the CPU oracle itself (so the oracle is checked against the reference's code, not only restated),
  (b) the kernel source (host-emulated here, real CUDA under -m gpu) through the C ABI.
jax.random is pinned by published known answers: the Random123 / JAX threefry2x32 vectors and the
documented value of `jax.random.split(PRNGKey(0))`.
"""
from pathlib import Path

import numpy as np
import pytest

from backends import BACKENDS, make_env
from oracle import oracle as orc_mod
from oracle.oracle import Oracle
from phase_guided_terrain_traversal_b200 import model as gm
from phase_guided_terrain_traversal_b200 import prng, terrain
from phase_guided_terrain_traversal_b200.go2 import gait, utility

GOLD = Path(__file__).resolve().parent / "golden"
CASES = ["flat", "stairs_level07", "stairs_level1_nodr", "baseline_stairs_level07"]

# golden info key -> (oracle field, abi field)
INFO_MAP = {"command": ("command", "command"), "step": ("step", "step"), "steps_until_next_cmd": ("steps_until_next_cmd", "steps_until_next_cmd"),
            "phase": ("phase", "phase"), "phase_dt": ("phase_dt", "phase_dt"), "gait_freq": ("gait_freq", "gait_freq"),
            "last_act": ("last_act", "last_act"), "last_last_act": ("last_last_act", "last_last_act"),
            "feet_air_time": ("feet_air_time", "feet_air_time"), "last_contact": ("last_contact", "last_contact"),
            "swing_peak": ("swing_peak", "swing_peak"), "H_max": ("H_max", "H_max"), "H_min": ("H_min", "H_min"),
            "heightscan": ("heightscan", "heightscan"), "motor_targets": ("motor_targets", "motor_targets"),
            "qpos_error_history": ("qpos_error_history", "qpos_error_history"), "qvel_history": ("qvel_history", "qvel_history"), "rng": ("rng", "rng")}
EXACT = {"step", "steps_until_next_cmd", "last_contact", "rng", "last_act", "last_last_act"}


# ---- jax.random known answers -----------------------------------------------------------------------
THREEFRY_KAT = [((0x0, 0x0), (0x0, 0x0), (0x6B200159, 0x99BA4EFE)),
                ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
                ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]


def test_threefry_known_answers():
    """Random123 kat_vectors for threefry2x32_20 (the same three vectors JAX's own random_test checks)."""
    for key, ctr, want in THREEFRY_KAT:
        a, b = prng.threefry2x32(np.array(key, np.uint32), np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
        assert (int(a[0]), int(b[0])) == want
        assert tuple(int(x) for x in orc_mod.threefry2x32(key, ctr[0], ctr[1])) == want


def test_split_known_answer():
    """jax.random.split(jax.random.PRNGKey(0)) as printed in the JAX documentation (original, non-partitionable layout)."""
    want = np.array([[4146024105, 967050713], [2718843009, 1272950319]], dtype=np.uint32)
    assert np.array_equal(prng.split(prng.PRNGKey(0), 2, partitionable=False), want)
    assert np.array_equal(orc_mod.rng_split(prng.PRNGKey(0), 2, partitionable=False), want)


@pytest.mark.parametrize("part", [True, False])
def test_host_prng_matches_oracle(part):
    k = np.array([123, 456], np.uint32)
    assert np.array_equal(prng.split(k, 3, part), orc_mod.rng_split(k, 3, part))
    for n in (1, 3, 12, 117):
        assert np.array_equal(prng.random_bits(k, n, part), orc_mod.rng_bits(k, n, part))
        assert np.array_equal(prng.uniform(k, (n,), -0.5, 0.7, part), orc_mod.rng_uniform(k, n, -0.5, 0.7, part).astype(np.float32))
    assert prng.randint(k, 0, 100, part) == orc_mod.rng_randint(k, 0, 100, part)


# ---- closed-form pieces ---------------------------------------------------------------------------------
def test_gait_and_yaw_closed_form():
    g = np.load(GOLD / "closed_form.npz")
    assert np.allclose(gait.PHASES, g["phases"]) and gait.p_stance == float(g["p_stance"])
    for h, z in zip(g["swing_heights"], g["get_z"]):
        assert np.abs(gait.get_z(g["phi"], swing_height=h, swing_min=-0.3) - z).max() < 2e-6
        zo = np.array([orc_mod._lib("f32").orc_gait_get_z(float(p), float(h), -0.3) for p in g["phi"]])
        assert np.abs(zo - z).max() < 2e-6
    assert np.abs(utility.quat_to_yaw(g["quat"].astype(np.float64)) - g["yaw"]).max() < 2e-6


# ---- task layer --------------------------------------------------------------------------------------------
def _cfg(variant=0):
    from phase_guided_terrain_traversal_b200.go2.configs import baseline_config, default_config, training_overrides
    return training_overrides(baseline_config() if variant else default_config())


def _variant(g):
    return int(g["meta/variant"]) if "meta/variant" in g.files else 0


def _tol(name):
    if name in ("obs_state", "obs_privileged", "reward", "metrics"):
        return 2e-5
    return 1e-5


def _compare(get, g, tag, seed_tag, who):
    """get(name) -> array of env 0. Float fields: abs err <= tol * max(1, |field|_inf)."""
    def chk(name, got, want, exact=False):
        got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
        assert got.shape == want.shape, (who, seed_tag, tag, name, got.shape, want.shape)
        if exact:
            assert np.array_equal(got, want), (who, seed_tag, tag, name)
        else:
            err = np.abs(got - want).max()
            assert err <= _tol(name) * max(1.0, np.abs(want).max()), (who, seed_tag, tag, name, err)
    p = f"{seed_tag}/{tag}/"
    chk("obs_state", get("obs_state"), g[p + "obs_state"])
    chk("obs_privileged", get("obs_priv"), g[p + "obs_privileged"])
    chk("reward", get("reward"), g[p + "reward"])
    chk("done", get("done"), g[p + "done"], exact=True)
    chk("metrics", get("metrics"), g[p + "metrics"])
    chk("qpos", get("qpos"), g[p + "qpos"])
    for k, (of, af) in INFO_MAP.items():
        chk(k, get(of, af), g[p + "info/" + k], exact=k in EXACT)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_task_code(case):
    """(a): the oracle's own reset/step (restated task layer) against the reference-code fixtures."""
    g = np.load(GOLD / f"task_layer_{case}.npz")
    task, level, dr = str(g["meta/task"]), str(g["meta/level"]), bool(g["meta/dr"])
    var = _variant(g)
    cfg, m = _cfg(var), gm.compile_model(task)
    table = terrain.load_terrain(level) if task == "stairs" else None
    for seed in g["meta/seeds"]:
        st = f"seed{seed}"
        o = Oracle(m, cfg, 1, "f32", rng_partitionable=bool(g["meta/partitionable"]), variant=var)
        o.randomize(g[st + "/dr_key"][None], table, dr)
        assert int(o.get("terrain_index")[0, 0]) == int(g[st + "/terrain_index"])
        o.reset(g[st + "/reset_key"][None])
        nobs = 162 if var else 171
        cut = {"obs_state": nobs, "obs_priv": nobs + 44}          # the oracle keeps max-size arrays; the variant fills a prefix
        get = lambda of, af=None: o.get(of)[0][:cut.get(of)]
        _compare(get, g, "reset", st, "oracle")
        for s in range(int(g["meta/n_steps"])):
            o.step(g[st + "/actions"][s][None].astype(np.float64), wrapped=False)
            _compare(get, g, f"step{s}", st, "oracle")


@pytest.mark.parametrize("kind", BACKENDS)
@pytest.mark.parametrize("case", CASES)
def test_kernel_matches_reference_task_code(kind, case):
    """(b): the kernel through the C ABI against the same fixtures. Physics differs from the oracle's
    only by fp32 operation order; a few steps stay inside the tolerances before trajectories drift."""
    g = np.load(GOLD / f"task_layer_{case}.npz")
    task, level, dr = str(g["meta/task"]), str(g["meta/level"]), bool(g["meta/dr"])
    var = _variant(g)
    cfg, m = _cfg(var), gm.compile_model(task)
    table = terrain.load_terrain(level) if task == "stairs" else None
    seeds = list(g["meta/seeds"])
    n = len(seeds)
    env = make_env(kind, m, cfg, n, rng_partitionable=bool(g["meta/partitionable"]), variant=var)
    if table is not None:
        env.set_terrain(table)
    env.randomize(np.stack([g[f"seed{s}/dr_key"] for s in seeds]), dr)
    env.reset(np.stack([g[f"seed{s}/reset_key"] for s in seeds]))
    amap = {"obs_priv": "obs_privileged"}

    def getter(i):
        return lambda of, af=None: env.get(af or amap.get(of, of))[i]
    for i, s in enumerate(seeds):
        _compare_loose(getter(i), g, "reset", f"seed{s}", kind)
    for step in range(min(3, int(g["meta/n_steps"]))):
        env.step(np.stack([g[f"seed{s}/actions"][step] for s in seeds]).astype(np.float32), wrapped=False)
        for i, s in enumerate(seeds):
            _compare_loose(getter(i), g, f"step{step}", f"seed{s}", kind)


def _compare_loose(get, g, tag, seed_tag, who):
    """Kernel vs fixture: the kernel's physics differs from the oracle's by fp32 operation order (arrow
    matrices vs dense), so solver-dependent values carry the tolerances of test_kernel_parity.py: 1e-4 of
    the field's max-norm for obs / reward / info floats, 2e-3 for the accelerometer slice (reads qacc),
    exact for integer / boolean bookkeeping and the rng keys."""
    p = f"{seed_tag}/{tag}/"
    checks = [("obs_state", None, g[p + "obs_state"], 1e-4), ("reward", None, g[p + "reward"], 1e-4), ("qpos", None, g[p + "qpos"], 1e-5),
              ("metrics", None, g[p + "metrics"], 2e-4)]
    for k in ("command", "phase", "phase_dt", "gait_freq", "feet_air_time", "swing_peak", "H_max", "H_min", "heightscan", "motor_targets"):
        checks.append((INFO_MAP[k][0], INFO_MAP[k][1], g[p + "info/" + k], 1e-5))
    for name, abi, want, tol in checks:
        got = np.asarray(get(name, abi), np.float64).reshape(-1)
        want = np.asarray(want, np.float64).reshape(-1)
        err = np.abs(got - want).max()
        assert err <= tol * max(1.0, np.abs(want).max()), (who, seed_tag, tag, name, err)
    got, want = np.asarray(get("obs_priv"), np.float64), np.asarray(g[p + "obs_privileged"], np.float64)
    acc = slice(want.size - 44 + 3, want.size - 44 + 6)          # privileged extras start at nobs: linvel 3, accelerometer 3, ...
    assert np.abs(got[acc] - want[acc]).max() <= 2e-3 * max(1.0, np.abs(want[acc]).max()), (who, seed_tag, tag, "accelerometer")
    got[acc] = want[acc]
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), (who, seed_tag, tag, "obs_privileged")
    for k in ("step", "steps_until_next_cmd", "last_contact", "rng", "last_act", "last_last_act"):
        assert np.array_equal(np.asarray(get(INFO_MAP[k][0], INFO_MAP[k][1]), np.float64).reshape(-1),
                              np.asarray(g[p + "info/" + k], np.float64).reshape(-1)), (who, seed_tag, tag, k)


def test_sample_command_matches_reference_method():
    """Host `Joystick.sample_command` against the reference method evaluated by make_golden.py (32 keys)."""
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    g = np.load(GOLD / "closed_form.npz")
    env = Joystick(task="flat_terrain", config=_cfg())
    out = np.stack([env.sample_command(k, x) for k, x in zip(g["cmd_keys"], g["cmd_x"])])
    assert np.array_equal(out, g["cmd_out"])
    assert (out != g["cmd_x"]).any() and (out == g["cmd_x"]).any()      # both branches of the w_k coin occur
