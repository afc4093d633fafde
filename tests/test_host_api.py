"""Host-side mirror of the reference interface: configs, constants, registry, ABI surface, error behaviour.
No compute calls here (no GPU): the C-ABI library is only loaded and its exports / error paths checked."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest

import phase_guided_terrain_traversal_b200 as pkg
from phase_guided_terrain_traversal_b200 import _native as nat
from phase_guided_terrain_traversal_b200 import model as gm
from phase_guided_terrain_traversal_b200 import prng, registry, terrain
from phase_guided_terrain_traversal_b200.go2 import configs, go2_constants as consts

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def _norm(x):
    if isinstance(x, dict):
        return {k: _norm(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_norm(v) for v in x]
    return x


def test_configs_equal_the_reference_factories():
    """default_config() / baseline_config() key for key against go2/configs.py evaluated (tests/golden/make_golden.py)."""
    ref = json.loads((GOLD / "reference_config.json").read_text())
    assert _norm(configs.default_config().to_dict()) == ref["default_config"]
    assert _norm(configs.baseline_config().to_dict()) == ref["baseline_config"]
    # metrics / reward order is the dict order of the scales (go2/configs.py:31-59)
    assert list(configs.default_config().reward_config.scales.keys()) == nat.REWARD_KEYS


def test_constants_equal_the_reference():
    ref = json.loads((GOLD / "reference_config.json").read_text())
    for k, v in ref["constants"].items():
        assert getattr(consts, k) == v, k
    for t, p in ref["task_to_xml"].items():
        assert consts.task_to_xml(t).as_posix() == p
    with pytest.raises(KeyError):          # go2_constants.py:45-52 raises KeyError for an unknown task
        consts.task_to_xml("rough_terrain_nonexistent")


def test_training_overrides_match_train_py():
    cfg = configs.training_overrides(configs.default_config())   # training/train.py:127-129
    assert cfg.command_config.u_max == [0.6, 0.6, 1.0] and cfg.command_config.u_min == [-0.6, -0.6, -1.0] and cfg.gait_freq == [1, 3]


def test_eval_overrides_match_evaluate_py():
    import re
    from phase_guided_terrain_traversal_b200.go2.configs import default_config, eval_overrides
    cfg = eval_overrides(default_config())
    assert list(cfg.command_config.u_max) == [0.4, 0.4, 0.7] and list(cfg.command_config.u_min) == [-0.4, -0.4, -0.7] and list(cfg.gait_freq) == [1, 3]
    ref = Path("/root/reference/training/evaluate.py")
    if ref.exists():     # the values above are what the reference script sets (training/evaluate.py:127-129)
        src = ref.read_text()
        assert re.search(r"^\s*env_cfg\.command_config\.u_max=\[0\.4,0\.4,0\.7\]", src, re.M) and re.search(r"^\s*env_cfg\.command_config\.u_min=\[-0\.4,-0\.4,-0\.7\]", src, re.M)


def test_policy_loader_resolves_no_code_and_keeps_value_statistics(tmp_path):
    """A policy pickle is third-party input: the loader maps every global but numpy array reconstruction and plain
    containers to an inert bag (a crafted `builtins.eval` reduce does not run); both checkpoint layouts - the 2-tuple
    (normaliser, PPONetworkParams) and the 3-tuple (normaliser, policy, value) of e.g. policy3 - return the
    privileged_state statistics the value network was trained on."""
    import io
    import pickle
    from phase_guided_terrain_traversal_b200 import policy_io

    class Evil:
        def __reduce__(self):
            return (eval, ("__import__('pathlib').Path(%r).write_text('x')" % str(tmp_path / "pwned"),))
    out = policy_io._StubUnpickler(io.BytesIO(pickle.dumps((Evil(),)))).load()
    assert not (tmp_path / "pwned").exists() and isinstance(out[0], dict)
    rng = np.random.default_rng(0)
    ks = [rng.normal(size=s).astype(np.float32) for s in ((171, 8), (8, 24))]
    bs = [rng.normal(size=s).astype(np.float32) for s in ((8,), (24,))]
    vk = [rng.normal(size=s).astype(np.float32) for s in ((215, 8), (8, 1))]
    vb = [rng.normal(size=s).astype(np.float32) for s in ((8,), (1,))]
    tree = lambda k, b: {"params": {f"hidden_{i}": {"kernel": k[i], "bias": b[i]} for i in range(2)}}
    norm = {"mean": {"state": np.zeros(171, np.float32), "privileged_state": np.full(215, 2, np.float32)},
            "std": {"state": np.ones(171, np.float32), "privileged_state": np.full(215, 3, np.float32)}, "count": np.float64(7)}
    (tmp_path / "three").write_bytes(pickle.dumps((norm, tree(ks, bs), tree(vk, vb))))
    d = policy_io.load_policy(tmp_path / "three")
    assert d["value"] is not None and np.all(d["value_mean"] == 2) and np.all(d["value_std"] == 3) and d["count"] == 7


def test_unsupported_terrain_tables_fail_loudly():
    from phase_guided_terrain_traversal_b200 import terrain
    t = terrain.load_terrain("level1").copy()
    terrain.validate_terrain(t)
    for mut, msg in ((lambda a: a.__setitem__((0, 0, 4), 0.3), "yaw"), (lambda a: a.__setitem__((0, 0, 8), 0.0), "half-sizes"),
                     (lambda a: a.__setitem__((0, 0, 0), np.nan), "finite"), (lambda a: a.__setitem__((0, 0, slice(3, 7)), 0.0), "quaternion")):
        b = t.copy(); mut(b)
        with pytest.raises(ValueError, match=msg):
            terrain.validate_terrain(b)
    with pytest.raises(ValueError, match="shape"):
        terrain.validate_terrain(t[:, :50])


def test_registry_call_sequence_of_train_py():
    """register_environment -> get_default_config -> load -> _randomizer / get_domain_randomizer (train.py:116-130,165-170,231)."""
    import functools
    from phase_guided_terrain_traversal_b200.go2 import joystick_pgtt, randomize
    registry.register_environment("Go2", functools.partial(joystick_pgtt.Joystick, task="stairs"), configs.default_config)
    cfg = registry.get_default_config("Go2")
    env = registry.load("Go2", config=configs.training_overrides(cfg))
    assert env.action_size == 12 and env.dt == 0.02 and env.sim_dt == 0.005 and env.n_substeps == 4
    assert env.observation_size == {"state": (171,), "privileged_state": (215,)}
    assert env.xml_path.endswith("terrain_scene_mjx.xml") and env.mjx_model.n_boxes == 100
    table = terrain.load_terrain("level1")
    registry._randomizer["Go2"] = functools.partial(randomize.domain_randomize, terrain_matrix=table)
    rm, in_axes = registry.get_domain_randomizer("Go2")(env.mjx_model, rng=prng.env_keys(0, 8))
    assert rm.rng.shape == (8, 2) and rm.terrain_matrix.shape == (100, 100, 10) and in_axes["body_mass"] == 0
    with pytest.raises(ValueError):
        registry.load("NoSuchEnv")


def test_abi_header_and_library_agree():
    """Every function include/pgtt_b200.h declares is exported by the built library and listed in the binding."""
    hdr = (ROOT / "include" / "pgtt_b200.h").read_text()
    declared = set(re.findall(r"\b(pgtt_[a-z_0-9]+)\s*\(", hdr)) - {"pgtt_env"}
    assert declared == set(nat.ABI_SYMBOLS), declared ^ set(nat.ABI_SYMBOLS)
    nat.build_library()
    try:
        lib = ctypes.CDLL(str(nat.LIB_PATH))
    except OSError as e:
        pytest.skip(f"CUDA runtime not loadable on this host: {e}")
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert ctypes.sizeof(nat.Buffers) == 4 + 4 + 8 * len(nat.BUFFER_FIELDS)      # int + pad + pointers
    lib = nat.declare(lib)
    assert lib.pgtt_version() >= 100


def test_product_path_fails_loudly_without_a_gpu():
    """No CPU fallback: creating a handle on a machine without a CUDA device raises (PGTT_ERR_CUDA)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    env = Joystick(task="flat_terrain", config=configs.default_config())
    with pytest.raises((nat.PgttError, nat.NativeLibraryError, RuntimeError, AssertionError)):
        env.reset(prng.env_keys(0, 4))


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(nat.NativeLibraryError):
        nat.load_library(tmp_path / "libpgtt_b200.so")


def test_terrain_loader_and_fixture_invariants():
    """terrains/level*.npy format: [T,100,10] f32, boxes stand on z = 0, yaw-only rotations."""
    for name in ["level1", "level07", "level13"]:
        t = terrain.load_terrain(name)
        assert t.dtype == np.float32 and t.shape[1:] == (100, 10)
        act = t[..., 0] < 50
        assert np.allclose(t[act][:, 2], t[act][:, 9], atol=1e-6)
        assert np.abs(t[..., 4:6]).max() == 0
        terrain.validate_terrain(t)
    with pytest.raises(FileNotFoundError):
        terrain.load_terrain("level99")
    with pytest.raises(ValueError):
        from phase_guided_terrain_traversal_b200.go2.randomize import domain_randomize
        domain_randomize(None, prng.env_keys(0, 2), np.zeros((3, 50, 10), np.float32))


def test_key_helpers():
    k = prng.as_keys(7, 5)
    assert k.shape == (5, 2) and np.array_equal(k, prng.split(prng.PRNGKey(7), 5))
    assert np.array_equal(prng.as_keys(k), k)
    with pytest.raises(ValueError):
        prng.as_keys(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        prng.as_keys(k, 6)


def test_abi_rejects_bad_arguments_without_touching_a_device():
    """Error behaviour of the C ABI: null pointers / empty shapes return PGTT_ERR_ARG (-1) with a message from
    pgtt_last_error / pgtt_policy_last_error before any CUDA call is made, so this runs on a host without a GPU."""
    nat.build_library()
    try:
        lib = nat.declare(ctypes.CDLL(str(nat.LIB_PATH)))
    except OSError as e:
        pytest.skip(f"CUDA runtime not loadable on this host: {e}")
    lib.pgtt_last_error.restype = ctypes.c_char_p
    null = ctypes.c_void_p(0)
    out = ctypes.c_void_p(0)
    assert lib.pgtt_create(None, None, 0, 16, ctypes.byref(out)) == -1 and b"pgtt_create" in lib.pgtt_last_error()
    m = gm.compile_model("flat_terrain")
    md, td = nat.model_desc(m), nat.task_desc(configs.default_config(), m)
    assert lib.pgtt_create(ctypes.byref(md), ctypes.byref(td), 0, 0, ctypes.byref(out)) == -1          # num_envs <= 0
    assert out.value is None
    for call in (lambda: lib.pgtt_sync(null, null), lambda: lib.pgtt_set_terrain_table(null, null, 0),
                 lambda: lib.pgtt_randomize(null, null, 0, null), lambda: lib.pgtt_get_buffers(null, None),
                 lambda: lib.pgtt_obs_dims(null, None, None), lambda: lib.pgtt_record(null, null, null, null, null, null, null)):
        assert call() == -1 and lib.pgtt_last_error()
    assert lib.pgtt_step_kernel_generation(null) == -1
    assert lib.pgtt_destroy(null) == 0                                                                 # destroying nothing is fine
    assert lib.pgtt_gae(null, null, null, null, 0, 0, 0.95, 0.97, 1.0, null, null, null) == -1 and b"pgtt_gae" in lib.pgtt_policy_last_error()
    assert lib.pgtt_ppo_head(*([null] * 8), 0, 12, 0.3, 0.01, 0.001, null, null, null, null) == -1
    assert lib.pgtt_adam_clip(*([null] * 6), 0, 3e-4, 0.9, 0.999, 1e-8, 1.0, 1.0, null) == -1 and b"pgtt_adam_clip" in lib.pgtt_policy_last_error()
    assert lib.pgtt_adam_scratch_floats() > 0
    # round-2 learner entry points: the same contract (argument errors before any CUDA call)
    assert lib.pgtt_step_launches(null) == -1
    assert lib.pgtt_gae_moments(null, null, null, null, 0, 0, 0.95, 0.97, 1.0, null, null, null, null) == -1 and b"pgtt_gae_moments" in lib.pgtt_policy_last_error()
    assert lib.pgtt_gae_sums(null, null, null, null, 20, 2048, 0.95, 0.97, 1.0, null, null, null, null) == -1
    assert lib.pgtt_moments_finalize(null, null, null) == -1
    assert lib.pgtt_minibatch_gather(null, null, 0, 0, 0, 0, 0, null, null, null, null, null, null, null, null) == -1
    assert lib.pgtt_col_moments(null, 0, 0, 0, null, null, null) == -1 and lib.pgtt_col_moments_scratch_doubles(171) == 592 * 2 * 171
    assert lib.pgtt_policy_set_params_device(null, None, None, null, null, null) == -1
    lib.pgtt_mlp_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p(0)
    assert lib.pgtt_mlp_create(0, (ctypes.c_int * 1)(4), 8, 0, ctypes.byref(h)) == -1 and b"pgtt_mlp_create" in lib.pgtt_mlp_last_error()
    assert lib.pgtt_mlp_create(2, (ctypes.c_int * 3)(4, 5000, 2), 8, 0, ctypes.byref(h)) == -1            # width out of range
    assert h.value is None
    assert lib.pgtt_mlp_forward(null, null, 0, None, None, null, null) == -1 and lib.pgtt_mlp_backward(null, null, None, None, null) == -1
    assert lib.pgtt_mlp_forward_gather(null, null, 0, 0, null, 0, null, null, None, None, null, null) == -1
    assert lib.pgtt_mlp_rows(null) == 0
    lib.pgtt_mlp_destroy(null)                                                                          # destroying nothing is fine
