"""Policy kernel (tcgen05, bf16 operands / fp32 accumulate) and the rollout collector.

Numerics: against a plain-torch fp32 statement of the same acting step (policy.reference_forward).
Tolerances: vs the reference with bf16-rounded operands 1e-2 abs on logits with median < 1e-5 and 99 % < 1e-3 -
the typical error is 1e-7 (accumulation order only), but a hidden activation that lands on a bf16 rounding
boundary may round the other way (1 bf16 ulp = 2^-8 relative, times downstream weights: up to 5.7e-3 observed
on B200 in 1 % of rows); vs pure fp32 2 % of the logit scale (bf16 operand rounding through four layers)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"


def test_policy_pickle_roundtrip(tmp_path):
    """save_policy / load_policy keep the (RunningStatisticsState, PPONetworkParams) layout deploy/policy_net.py reads."""
    from phase_guided_terrain_traversal_b200 import policy_io
    g = np.random.default_rng(0)
    sizes = [171, 512, 256, 128, 24]
    ks = [g.normal(size=(i, o)).astype(np.float32) for i, o in zip(sizes[:-1], sizes[1:])]
    bs = [g.normal(size=o).astype(np.float32) for o in sizes[1:]]
    mean, std = g.normal(size=171).astype(np.float32), g.uniform(0.5, 2, 171).astype(np.float32)
    policy_io.save_policy(tmp_path / "p", mean, std, (ks, bs), count=5.0)
    d = policy_io.load_policy(tmp_path / "p")
    assert all(np.array_equal(a, b) for a, b in zip(d["policy"][0], ks)) and all(np.array_equal(a, b) for a, b in zip(d["policy"][1], bs))
    assert np.array_equal(d["mean"], mean) and np.array_equal(d["std"], std) and d["count"] == 5.0


def test_reference_forward_matches_reference_network_golden():
    """The torch fp32 statement used as kernel reference reproduces deploy/policy_net.py on policy177 (fixture)."""
    from phase_guided_terrain_traversal_b200.policy import reference_forward
    g = np.load(GOLD / "policy177.npz")
    ks, bs = [g[f"kernel{i}"] for i in range(4)], [g[f"bias{i}"] for i in range(4)]
    out = reference_forward(ks, bs, g["obs"], g["mean"], g["std"])
    assert np.abs(out["action"].numpy() - g["action_deterministic"]).max() < 2e-5
    # the shipped normaliser statistics carry the distributional pins of SURVEY 8c-3
    assert abs(g["mean"][5] + 0.991) < 5e-3 and np.allclose(g["std"][30:38], 0.7075, atol=2e-3) and abs(g["mean"][155] - 2.01) < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 100, 128, 4096])
def test_policy_kernel_vs_torch_reference(n):
    import torch
    from phase_guided_terrain_traversal_b200.policy import PolicyNet, reference_forward
    g = np.load(GOLD / "policy177.npz")
    ks, bs = [g[f"kernel{i}"] for i in range(4)], [g[f"bias{i}"] for i in range(4)]
    net = PolicyNet()
    net.set_params(ks, bs, g["mean"], g["std"])
    rng = np.random.default_rng(n)
    obs = (g["mean"] + g["std"] * rng.normal(size=(n, 171)).clip(-3, 3)).astype(np.float32)
    eps = rng.normal(size=(n, 12)).astype(np.float32)
    out = net.act(torch.from_numpy(obs).cuda(), eps=torch.from_numpy(eps).cuda(), want_logits=True)
    torch.cuda.synchronize()
    ref_bf = reference_forward(ks, bs, obs, g["mean"], g["std"], eps, bf16_operands=True)
    ref_32 = reference_forward(ks, bs, obs, g["mean"], g["std"], eps)
    lg = out["logits"].cpu()
    # measured on B200 (tools/policy_error_probe.py, n = 4096): median 5e-8, 99 % < 3.4e-4, max 5.7e-3 (45 rows with a
    # flipped bf16 rounding of a hidden activation); bf16-operand reference vs pure fp32: max 5.7e-2 at |logit| <= 7.2
    err = (lg - ref_bf["logits"]).abs()
    assert err.max() < 1e-2, err.max()
    assert err.median() < 1e-5 and torch.quantile(err.flatten(), 0.99) < 1e-3
    assert (lg - ref_32["logits"]).abs().max() < 2e-2 * max(1.0, float(ref_32["logits"].abs().max()))
    # sampling is 1-Lipschitz in the logits (tanh, softplus): |d raw| <= |d loc| + |d scale-logit| * |eps|, |d action| <= |d raw|
    dl = (lg - ref_bf["logits"]).abs()
    bound = dl[:, :12] + dl[:, 12:] * torch.from_numpy(eps).abs() + 1e-5
    assert ((out["raw_action"].cpu() - ref_bf["raw_action"]).abs() <= bound).all()
    assert ((out["action"].cpu() - ref_bf["action"]).abs() <= bound).all()
    assert (out["action"].cpu() - ref_bf["action"]).abs().median() < 1e-5
    assert ((out["log_prob"].cpu() - ref_bf["log_prob"]).abs() / (1 + ref_bf["log_prob"].abs())).max() < 2e-2
    det = net.act(torch.from_numpy(obs).cuda(), deterministic=True)
    assert (det["action"].cpu() - torch.tanh(ref_bf["logits"][:, :12])).abs().max() < 1e-2
    if n == 4096:   # deployment forward of the reference network on its own fixture obs
        o2 = net.act(torch.from_numpy(g["obs"]).cuda(), deterministic=True)
        assert np.abs(o2["action"].cpu().numpy() - g["action_deterministic"]).max() < 3e-2


@pytest.mark.gpu
def test_policy_internal_noise_is_standard_normal():
    import torch
    from phase_guided_terrain_traversal_b200.policy import PolicyNet
    net = PolicyNet().init_random(1)
    obs = torch.zeros((8192, 171), device="cuda")
    o = net.act(obs, seed=3, want_logits=True)
    loc, sr = o["logits"][:, :12], o["logits"][:, 12:]
    e = (o["raw_action"] - loc) / (torch.nn.functional.softplus(sr) + 0.001)
    assert abs(float(e.mean())) < 0.02 and abs(float(e.std()) - 1) < 0.02
    o2 = net.act(obs, seed=3)
    assert not torch.equal(o2["raw_action"], o["raw_action"])      # the step counter advances the stream


@pytest.mark.gpu
def test_rollout_collector_shapes_and_consistency(train_cfg):
    """20-step unroll on stairs/level07 driven by policy177: buffers are time-major, next_obs[t] == obs[t + 1],
    stored actions are what the env consumed (info.last_act), discount = 1 - done, and a trained policy keeps most
    robots upright (closed loop through physics, obs and policy)."""
    import functools
    import torch
    from phase_guided_terrain_traversal_b200 import prng, terrain
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    from phase_guided_terrain_traversal_b200.go2.randomize import domain_randomize
    from phase_guided_terrain_traversal_b200.policy import PolicyNet
    from phase_guided_terrain_traversal_b200.rollout import RolloutCollector
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
    g = np.load(GOLD / "policy177.npz")
    n, T = 512, 20
    keys = prng.env_keys(5, n)
    env = Joystick(task="stairs", config=train_cfg)
    wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=functools.partial(domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain("level07")))
    state = wenv.reset(keys + np.uint32(9))
    net = PolicyNet()
    net.set_params([g[f"kernel{i}"] for i in range(4)], [g[f"bias{i}"] for i in range(4)], g["mean"], g["std"])
    col = RolloutCollector(wenv, net, unroll_length=T, seed=1)
    fallen = 0.0
    for it in range(5):
        state, ro = col.collect(state)
        torch.cuda.synchronize()
        assert ro.obs_state.shape == (T + 1, n, 171) and ro.action.shape == (T, n, 12) and ro.reward.shape == (T, n)
        assert torch.equal(ro.next_observation["state"][:-1], ro.observation["state"][1:])
        assert torch.equal(ro.obs_state[-1], state.obs["state"]) and torch.equal(ro.action[-1], state.info["last_act"])
        assert torch.all((ro.discount == 0) | (ro.discount == 1)) and torch.isfinite(ro.log_prob).all() and (ro.action.abs() <= 1).all()
        fallen += float((1 - ro.discount).sum()) / n
    assert fallen < 0.25, fallen                                   # random actions lose ~all robots within 100 steps
    assert float(ro.reward.mean()) > 0.005


@pytest.mark.gpu
def test_rollouts_of_two_envs_with_different_configs_alternate(train_cfg):
    """Two handles on one device (a phase-guided stairs env, 171 observations, and a baseline flat env, 162 - as train + eval
    envs of training/train.py:242-263 would be) collected alternately through their CUDA-graphed rollouts: each reproduces its
    solo run bit for bit, i.e. a graph replay never runs with the other handle's constant table; a handle re-created with
    the same buffers does not inherit the old graph."""
    import functools
    import torch
    from phase_guided_terrain_traversal_b200 import prng, terrain
    from phase_guided_terrain_traversal_b200.go2 import joystick, joystick_pgtt, randomize, randomize_simple
    from phase_guided_terrain_traversal_b200.go2.configs import baseline_config, training_overrides
    from phase_guided_terrain_traversal_b200.policy import PolicyNet
    from phase_guided_terrain_traversal_b200.rollout import RolloutCollector
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
    n, T = 96, 5
    keys = prng.env_keys(8, n)

    def make(which):
        if which == "a":
            env = joystick_pgtt.Joystick(task="stairs", config=train_cfg)
            rfn = functools.partial(randomize.domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain("level07"))
            net = PolicyNet().init_random(3)
        else:
            env = joystick.Joystick(task="flat_terrain", config=training_overrides(baseline_config()))
            rfn = functools.partial(randomize_simple.domain_randomize, rng=keys)
            net = PolicyNet((162, 512, 256, 128, 24)).init_random(4)
        wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=rfn)
        wenv.reset(keys + np.uint32(1))
        return wenv, RolloutCollector(wenv, net, unroll_length=T, seed=2)

    def run(col, k):
        out = []
        for _ in range(k):
            _, ro = col.collect()
            out.append((ro.obs_state.clone(), ro.reward.clone(), ro.action.clone()))
        torch.cuda.synchronize()
        return out

    solo = {}
    for which in ("a", "b"):
        wenv, col = make(which)
        solo[which] = run(col, 3)
        del col, wenv
    (wa, ca), (wb, cb) = make("a"), make("b")
    got = {"a": [], "b": []}
    for _ in range(3):
        got["a"] += run(ca, 1)
        got["b"] += run(cb, 1)
    for which in ("a", "b"):
        for x, y in zip(got[which], solo[which]):
            for u, v in zip(x, y):
                assert torch.equal(u, v), which
    assert got["a"][0][0].shape[-1] == 171 and got["b"][0][0].shape[-1] == 162


@pytest.mark.gpu
def test_evaluator_closed_loop_with_a_reference_policy(train_cfg):
    """training/evaluate.py semantics (success = episode ends without termination) and a closed-loop plausibility pin
    (SURVEY 8c-3): policy177, trained by the reference in real MJX, walks in this env - measured on B200: level07, 1000
    envs, full 1000-step episodes: episode reward 12.4 +- 11 (a random-init policy that stands still: 0.05), 72 %
    successes, and the observation statistics it sees match its own normaliser (gravity_z -0.97 vs -0.99, phase cos std
    0.708 vs 0.708, gait_freq 2.02 vs 2.01, joint-position std 0.16 vs 0.14)."""
    import functools
    from phase_guided_terrain_traversal_b200 import prng, terrain, wrapper
    from phase_guided_terrain_traversal_b200.evaluate import evaluate
    from phase_guided_terrain_traversal_b200.go2 import joystick_pgtt, randomize
    from phase_guided_terrain_traversal_b200.policy import PolicyNet
    g = np.load(GOLD / "policy177.npz")
    res = {}
    for which in ("policy177", "random"):
        env = joystick_pgtt.Joystick(task="stairs", config=train_cfg)
        keys = prng.env_keys(4, 500)
        wenv = wrapper.wrap_for_brax_training(env, episode_length=400, randomization_fn=functools.partial(
            randomize.domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain("level07")))
        wenv.reset(keys + np.uint32(1))
        net = PolicyNet()
        if which == "policy177":
            net.set_params([g[f"kernel{i}"] for i in range(4)], [g[f"bias{i}"] for i in range(4)], g["mean"], g["std"])
        else:
            net.init_random(0)
        res[which] = evaluate(wenv, net, episode_length=400, collect_obs_stats=True)
        env.close()
    r, z = res["policy177"], res["random"]
    assert r["num_eval_envs"] == 500 and 0 <= r["success_count"] <= 500
    assert r["episode_reward"] > 2.0 and z["episode_reward"] < 0.5          # the trained policy tracks commands, a standing robot earns ~0
    assert r["success_rate"] > 0.6 and r["avg_episode_length"] > 300
    m, s = r["obs_mean"], r["obs_std"]
    assert abs(m[5] - g["mean"][5]) < 0.05                                    # gravity z
    assert abs(s[30] - g["std"][30]) < 0.02 and abs(m[155] - g["mean"][155]) < 0.1
    assert 0.5 < s[6:18].mean() / g["std"][6:18].mean() < 2.0


@pytest.mark.gpu
def test_device_side_parameter_hand_over_equals_the_host_one():
    """`pgtt_policy_set_params_device` (packing kernel reading CUDA fp32 tensors, stream-ordered) against `pgtt_policy_set_params` (host repack):
    the two handles produce bit-identical actions, log-probs and logits for the same observations, seed and step."""
    import torch
    from phase_guided_terrain_traversal_b200.policy import PolicyNet
    g = np.random.default_rng(5)
    sizes = (171, 512, 256, 128, 24)
    ks = [(g.uniform(-1, 1, (i, o)) * np.sqrt(3.0 / i)).astype(np.float32) for i, o in zip(sizes[:-1], sizes[1:])]
    bs = [g.normal(0, 0.1, o).astype(np.float32) for o in sizes[1:]]
    mean, std = g.normal(0, 1, 171).astype(np.float32), g.uniform(0.5, 2.0, 171).astype(np.float32)
    obs = torch.from_numpy(g.normal(0, 1, (300, 171)).astype(np.float32)).cuda()
    a, b = PolicyNet(sizes), PolicyNet(sizes)
    a.set_params(ks, bs, mean, std)
    tc = lambda x: torch.from_numpy(x).cuda()
    b.set_params_device([tc(k) for k in ks], [tc(x) for x in bs], tc(mean), tc(std))
    ra, rb = a.act(obs, seed=3, want_logits=True), b.act(obs, seed=3, want_logits=True)
    torch.cuda.synchronize()
    for k in ("action", "raw_action", "log_prob", "logits"):
        assert torch.equal(ra[k], rb[k]), k
    # a second hand-over overwrites every packed weight (nothing stale survives), identity normaliser included
    half = [(k * 0.5).astype(np.float32) for k in ks]
    b.set_params_device([tc(k) for k in half], [tc(x) for x in bs])
    a.set_params(half, bs)
    ra, rb = a.act(obs, seed=4, want_logits=True), b.act(obs, seed=4, want_logits=True)
    torch.cuda.synchronize()
    assert torch.equal(ra["logits"], rb["logits"]) and torch.equal(ra["action"], rb["action"])
