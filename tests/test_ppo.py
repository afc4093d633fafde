"""PPO learner: pure functions against independent restatements on CPU; the trainer end to end on the GPU."""
import math

import numpy as np
import pytest


def gae_numpy(trunc, term, rew, val, boot, lam, gamma):
    """Loop form of generalized advantage estimation with brax's truncation handling."""
    T, B = rew.shape
    vs = np.zeros((T, B)); adv = np.zeros((T, B))
    v_next = np.concatenate([val[1:], boot[None]], 0)
    acc = np.zeros(B)
    for t in reversed(range(T)):
        mask = 1.0 - trunc[t]
        delta = (rew[t] + gamma * (1 - term[t]) * v_next[t] - val[t]) * mask
        acc = delta + gamma * (1 - term[t]) * mask * lam * acc
        vs[t] = acc + val[t]
    vs_next = np.concatenate([vs[1:], boot[None]], 0)
    adv = (rew + gamma * (1 - term) * vs_next - val) * (1 - trunc)
    return vs, adv


def test_compute_gae_matches_loop_form():
    import torch
    from phase_guided_terrain_traversal_b200.ppo import compute_gae
    g = np.random.default_rng(0)
    T, B = 20, 37
    trunc = (g.random((T, B)) < 0.05).astype(np.float64)
    term = ((g.random((T, B)) < 0.05) & (trunc == 0)).astype(np.float64)
    rew, val, boot = g.normal(size=(T, B)), g.normal(size=(T, B)), g.normal(size=B)
    vs, adv = compute_gae(*(torch.from_numpy(x) for x in (trunc, term, rew, val, boot)), 0.95, 0.97)
    vs_n, adv_n = gae_numpy(trunc, term, rew, val, boot, 0.95, 0.97)
    assert np.allclose(vs.numpy(), vs_n, atol=1e-12) and np.allclose(adv.numpy(), adv_n, atol=1e-12)
    # no termination / truncation, lambda = 1: advantages are discounted returns minus values
    z = np.zeros((T, B))
    vs1, adv1 = compute_gae(*(torch.from_numpy(x) for x in (z, z, rew, val, boot)), 1.0, 0.9)
    ret = np.zeros((T, B)); acc = boot.copy()
    for t in reversed(range(T)):
        acc = rew[t] + 0.9 * acc; ret[t] = acc
    assert np.allclose(adv1.numpy(), ret - val, atol=1e-10) and np.allclose(vs1.numpy(), ret, atol=1e-10)


def test_tanh_normal_log_prob_and_entropy_against_torch_distributions():
    import torch
    from torch.distributions import Normal, TransformedDistribution
    from torch.distributions.transforms import TanhTransform
    from phase_guided_terrain_traversal_b200.ppo import tanh_normal_entropy, tanh_normal_log_prob
    torch.manual_seed(0)
    logits = torch.randn(50, 24, dtype=torch.float64)
    raw = torch.randn(50, 12, dtype=torch.float64)
    loc, sr = logits.chunk(2, -1)
    scale = torch.nn.functional.softplus(sr) + 0.001
    d = TransformedDistribution(Normal(loc, scale), [TanhTransform(cache_size=1)])
    ref = d.log_prob(torch.tanh(raw)).sum(-1)
    assert torch.allclose(tanh_normal_log_prob(logits, raw), ref, atol=1e-6)
    eps = torch.randn(50, 12, dtype=torch.float64)
    sample = loc + scale * eps
    ref_ent = (Normal(loc, scale).entropy() + torch.log(1 - torch.tanh(sample) ** 2)).sum(-1)
    assert torch.allclose(tanh_normal_entropy(logits, eps), ref_ent, atol=1e-6)


def test_ppo_loss_on_policy_properties():
    """With behaviour == target policy the ratio is 1: the policy loss is minus the mean normalised advantage (~0) and its
    gradient equals the REINFORCE gradient; the value loss is 0.25 * MSE against the GAE targets."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo
    torch.manual_seed(1)
    cfg = ppo.PPOConfig()
    gen = torch.Generator().manual_seed(0)
    pp = ppo.lecun_uniform_params((171, 64, 24), gen, "cpu")
    vp = ppo.lecun_uniform_params((215, 64, 1), gen, "cpu")
    T, B = 5, 16
    obs, obs_priv = torch.randn(T + 1, B, 171), torch.randn(T + 1, B, 215)
    logits = ppo.mlp(obs[:T], *pp)
    loc, sr = logits.chunk(2, -1)
    raw = (loc + (torch.nn.functional.softplus(sr) + 0.001) * torch.randn(T, B, 12)).detach()
    batch = {"obs": obs, "obs_priv": obs_priv, "raw_action": raw, "log_prob": ppo.tanh_normal_log_prob(logits, raw).detach(),
             "reward": torch.randn(T, B), "discount": torch.ones(T, B), "truncation": torch.zeros(T, B), "eps": torch.randn(T, B, 12)}
    total, m = ppo.ppo_loss(pp, vp, batch, cfg)
    assert abs(float(m["policy_loss"])) < 1e-5                       # mean of normalised advantages
    base = ppo.mlp(obs_priv, *vp).squeeze(-1)
    vs, _ = ppo.compute_gae(batch["truncation"], torch.zeros(T, B), batch["reward"], base[:T].detach(), base[T].detach(), cfg.gae_lambda, cfg.discounting)
    assert abs(float(m["v_loss"]) - 0.25 * float(((vs - base[:T].detach()) ** 2).mean())) < 1e-6
    assert abs(float(total) - float(m["policy_loss"] + m["v_loss"] - cfg.entropy_cost * m["entropy"])) < 1e-6
    total.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in pp[0] + vp[0])


def test_running_stats_merge_equals_batch_statistics():
    import torch
    from phase_guided_terrain_traversal_b200.ppo import RunningStats
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(100, 7, generator=g) * 3 + 1, torch.randn(50, 7, generator=g) - 2, torch.randn(300, 7, generator=g) * 0.1]
    rs = RunningStats(7, "cpu")
    for x in xs:
        rs.update(x)
    full = torch.cat(xs).double()
    assert float(rs.count) == 450 and torch.allclose(rs.mean, full.mean(0), atol=1e-10)
    assert torch.allclose(rs.std, full.std(0, unbiased=False), atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_trainer_runs_and_is_consistent_with_the_policy_kernel(train_cfg, tmp_path, graph):
    """Two training steps on 256 flat-terrain envs: the learner's log-prob of the stored raw actions under the
    COLLECTION parameters matches what the tcgen05 policy kernel reported (bf16 operands vs fp32: median < 2e-2), losses are finite,
    parameters move, and the exported pickle has the reference's layout."""
    import torch
    from phase_guided_terrain_traversal_b200 import policy_io, ppo, prng
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    from phase_guided_terrain_traversal_b200.go2.randomize_simple import domain_randomize
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
    n = 256
    cfg = ppo.PPOConfig(num_envs=n, batch_size=64, num_minibatches=8, num_updates_per_batch=2, use_cuda_graph=graph, seed=3)
    env = Joystick(task="flat_terrain", config=train_cfg)
    keys = prng.env_keys(1, n)
    wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=lambda m: domain_randomize(m, rng=keys))
    state = wenv.reset(keys)
    tr = ppo.PPOTrainer(wenv, state, cfg)
    assert tr.unrolls_per_step == 2 and tr.mb == 64
    # consistency of the two policy implementations at the collection parameters
    _, ro = tr.collector.collect()
    with torch.no_grad():   # (an autograd graph built on the legacy stream would pin the parameters' AccumulateGrad nodes to it,
        #                     and the captured backward may not touch the legacy stream)
        logits = ppo.mlp(ro.obs_state[:-1], *tr.policy_params)
        lp = ppo.tanh_normal_log_prob(logits, ro.raw_action)
    torch.cuda.synchronize()
    # the kernel rounds weights and activations to bf16, the learner is fp32: ~1e-2 on a log-prob summed over 12 dims
    # (an importance-ratio noise of 1 %, far inside the 0.3 clip)
    d = (lp.detach() - ro.log_prob).abs()
    assert float(d.max()) < 0.3 and float(d.median()) < 2e-2, (float(d.max()), float(d.median()))
    before = [p.detach().clone() for p in tr.params]
    for _ in range(2):
        m = tr.training_step()
        assert all(math.isfinite(v) for v in m.values()), m
    assert any(float((a - b.detach()).abs().max()) > 0 for a, b in zip(before, tr.params))
    assert m["env_steps"] == 2 * 2 * 20 * n                        # the extra collect above is not counted
    tr.save(tmp_path / "policy_test")
    d = policy_io.load_policy(tmp_path / "policy_test")
    assert [k.shape for k in d["policy"][0]] == [(171, 512), (512, 256), (256, 128), (128, 24)]
    assert [k.shape for k in d["value"][0]] == [(215, 512), (512, 256), (256, 128), (128, 1)]
    assert d["count"] == 2 * 2 * 20 * n


@pytest.mark.gpu
def test_baseline_task_trains(train_cfg):
    """`--method baseline` (go2/joystick.py, obs 162 / 206): env, rollout (policy kernel with a 162-wide input) and learner."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo, prng
    from phase_guided_terrain_traversal_b200.go2.configs import baseline_config, training_overrides
    from phase_guided_terrain_traversal_b200.go2.joystick import Joystick
    from phase_guided_terrain_traversal_b200.go2.randomize_simple import domain_randomize
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
    n = 256
    env = Joystick(task="flat_terrain", config=training_overrides(baseline_config()))
    assert env.observation_size == {"state": (162,), "privileged_state": (206,)}
    keys = prng.env_keys(2, n)
    wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=lambda m: domain_randomize(m, rng=keys))
    state = wenv.reset(keys)
    assert tuple(state.obs["state"].shape) == (n, 162) and tuple(state.obs["privileged_state"].shape) == (n, 206)
    tr = ppo.PPOTrainer(wenv, state, ppo.PPOConfig(num_envs=n, batch_size=64, num_minibatches=8, num_updates_per_batch=1, use_cuda_graph=False))
    m = tr.training_step()
    assert all(math.isfinite(v) for v in m.values()) and tr.collector.buf.obs_state.shape == (21, n, 162)


@pytest.mark.gpu
def test_gae_kernel_matches_torch_statement():
    import torch
    from phase_guided_terrain_traversal_b200 import ppo
    g = torch.Generator(device="cuda").manual_seed(0)
    T, B = 20, 1000
    trunc = (torch.rand(T, B, generator=g, device="cuda") < 0.05).float()
    done = torch.maximum(trunc, (torch.rand(T, B, generator=g, device="cuda") < 0.05).float())
    disc = 1.0 - done
    rew, vals = torch.randn(T, B, generator=g, device="cuda"), torch.randn(T + 1, B, generator=g, device="cuda")
    vs, adv = ppo.compute_gae_native(trunc, disc, rew, vals, 0.95, 0.97, 2.0)
    term = (1 - disc) * (1 - trunc)
    vs_r, adv_r = ppo.compute_gae(trunc.double(), term.double(), rew.double() * 2.0, vals[:T].double(), vals[T].double(), 0.95, 0.97)
    assert float((vs.double() - vs_r).abs().max()) < 1e-4 and float((adv.double() - adv_r).abs().max()) < 1e-4


@pytest.mark.gpu
def test_fused_ppo_head_matches_autograd():
    """`pgtt_ppo_head` (loss terms + gradients in one launch) against the torch statement of the same loss: values 1e-5 rel,
    parameter gradients 1e-4 of their max-norm; includes clipped ratios (behaviour log-probs perturbed)."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo
    torch.manual_seed(2)
    dev = "cuda"
    cfg = ppo.PPOConfig()
    gen = torch.Generator(device=dev).manual_seed(0)
    pp = ppo.lecun_uniform_params((171, 128, 24), gen, dev)
    vp = ppo.lecun_uniform_params((215, 128, 1), gen, dev)
    T, B = 20, 64
    obs, obs_priv = torch.randn(T + 1, B, 171, device=dev), torch.randn(T + 1, B, 215, device=dev)
    with torch.no_grad():
        logits = ppo.mlp(obs[:T], *pp)
        loc, sr = logits.chunk(2, -1)
        raw = loc + (torch.nn.functional.softplus(sr) + 0.001) * torch.randn(T, B, 12, device=dev)
        old_lp = ppo.tanh_normal_log_prob(logits, raw) + 0.4 * torch.randn(T, B, device=dev)      # ratios on both sides of the clip range
    done = (torch.rand(T, B, device=dev) < 0.05).float()
    batch = {"obs": obs, "obs_priv": obs_priv, "raw_action": raw, "log_prob": old_lp, "reward": torch.randn(T, B, device=dev),
             "discount": 1 - done, "truncation": done * (torch.rand(T, B, device=dev) < 0.5).float(), "eps": torch.randn(T, B, 12, device=dev)}
    params = pp[0] + pp[1] + vp[0] + vp[1]
    out = {}
    for fused in (False, True):
        for p in params:
            p.grad = None
        total, m = ppo.ppo_loss(pp, vp, batch, cfg, fused=fused)
        total.backward()
        out[fused] = ({k: float(v) for k, v in m.items()}, [p.grad.clone() for p in params])
    for k in out[False][0]:
        assert abs(out[True][0][k] - out[False][0][k]) <= 1e-5 * max(1.0, abs(out[False][0][k])), (k, out[True][0][k], out[False][0][k])
    for a, b in zip(out[True][1], out[False][1]):
        assert float((a - b).abs().max()) <= 1e-4 * max(float(b.abs().max()), 1e-6)


@pytest.mark.gpu
def test_native_tensor_core_layers_match_torch_fp32_through_the_whole_loss():
    """`PPOConfig.native_mlp`: both MLPs (forward, SiLU, input / weight / bias gradients) on the hand-written tcgen05 layers of
    csrc/pgtt_learner.cu against torch fp32 GEMMs ('highest', the reference's precision) on the full PPO loss at the
    reference network sizes and a 5120-transition minibatch: loss terms to 1e-5, every parameter gradient to 1e-4 of its
    max-norm; with the value network on a side stream and the parameter gradients on auxiliary streams as in training."""
    import dataclasses
    import torch
    from phase_guided_terrain_traversal_b200 import ppo
    torch.set_float32_matmul_precision("highest")
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(3)
    T, B = 20, 256
    r = lambda *s: torch.randn(s, generator=g, device=dev)
    obs, obs_priv = torch.zeros(T + 1, B, 172, device=dev), torch.zeros(T + 1, B, 216, device=dev)      # zero-padded rows, as the trainer gathers them
    obs[..., :171] = r(T + 1, B, 171); obs_priv[..., :215] = r(T + 1, B, 215)
    pol = ppo.lecun_uniform_params([171, 512, 256, 128, 24], g, dev)
    val = ppo.lecun_uniform_params([215, 512, 256, 128, 1], g, dev)
    with torch.no_grad():
        logits = ppo.mlp(obs[:T], *pol)
        loc, sr = logits.chunk(2, -1)
        raw = loc + (torch.nn.functional.softplus(sr) + 0.001) * r(T, B, 12)
        old_lp = ppo.tanh_normal_log_prob(logits, raw) + 0.3 * r(T, B)
    done = (torch.rand((T, B), generator=g, device=dev) < 0.05).float()
    batch = {"obs": obs, "obs_priv": obs_priv, "raw_action": raw, "log_prob": old_lp, "reward": r(T, B).abs(), "discount": 1 - done,
             "truncation": torch.zeros((T, B), device=dev), "eps": r(T, B, 12)}
    params = pol[0] + pol[1] + val[0] + val[1]
    out = {}
    for native in (False, True):
        cfg = dataclasses.replace(ppo.PPOConfig(), native_mlp=native)
        streams = (torch.cuda.Stream(dev), (torch.cuda.Stream(dev), torch.cuda.Stream(dev))) if native else (None, (None, None))
        for p in params:
            p.grad = None
        loss, m = ppo.ppo_loss(pol, val, batch, cfg, fused=True, side_stream=streams[0], aux_streams=streams[1])
        loss.backward()
        torch.cuda.synchronize()
        out[native] = ({k: float(v) for k, v in m.items()}, [p.grad.clone() for p in params])
    for k in out[False][0]:
        assert abs(out[True][0][k] - out[False][0][k]) <= 1e-5 * max(1.0, abs(out[False][0][k])), (k, out[True][0][k], out[False][0][k])
    for a, b in zip(out[True][1], out[False][1]):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-4 * max(float(b.abs().max()), 1e-6), (a.shape, float((a - b).abs().max()), float(b.abs().max()))


@pytest.mark.gpu
def test_checkpoint_resume_roundtrip(train_cfg, tmp_path):
    """`--checkpoint_folder` semantics: a trainer restored from a saved pickle has the same parameters and observation
    statistics, so its deterministic policy outputs are identical; shipped reference policies restore as well."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo, prng
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    from phase_guided_terrain_traversal_b200.go2.randomize_simple import domain_randomize
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training

    def make(seed):
        n = 128
        env = Joystick(task="flat_terrain", config=train_cfg)
        keys = prng.env_keys(seed, n)
        wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=lambda m: domain_randomize(m, rng=keys))
        st = wenv.reset(keys)
        return ppo.PPOTrainer(wenv, st, ppo.PPOConfig(num_envs=n, batch_size=32, num_minibatches=8, num_updates_per_batch=1, use_cuda_graph=False, seed=seed))

    a = make(1)
    a.training_step()
    a.save(tmp_path / "123")
    b = make(2)
    b.restore(tmp_path / "123")
    for x, y in zip(a.params, b.params):
        assert torch.equal(x, y)
    assert torch.allclose(a.norm_state.mean, b.norm_state.mean, atol=1e-6) and torch.allclose(a.norm_priv.std, b.norm_priv.std, rtol=1e-6)
    obs = torch.randn(128, 171, device="cuda")
    assert torch.equal(a.net.act(obs, deterministic=True)["action"], b.net.act(obs, deterministic=True)["action"])
    m = b.training_step()
    assert all(math.isfinite(v) for v in m.values())


@pytest.mark.gpu
def test_value_net_on_a_side_stream_gives_the_same_gradients():
    """`PPOConfig.parallel_nets`: the value network's forward/backward run on a second stream and the parameter-gradient GEMMs on two more; loss and gradients are the
    ones of the single-stream evaluation."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(0)
    T, B = 20, 64
    cfg = ppo.PPOConfig()
    r = lambda *s: torch.randn(s, generator=g, device=dev)
    batch = {"obs": r(T + 1, B, 171), "obs_priv": r(T + 1, B, 215), "raw_action": r(T, B, 12), "log_prob": -r(T, B).abs(), "reward": r(T, B).abs(),
             "discount": (torch.rand((T, B), generator=g, device=dev) > 0.05).float(), "truncation": torch.zeros((T, B), device=dev), "eps": r(T, B, 12)}
    pol = ppo.lecun_uniform_params([171, 512, 256, 128, 24], g, dev)
    val = ppo.lecun_uniform_params([215, 512, 256, 128, 1], g, dev)
    params = pol[0] + pol[1] + val[0] + val[1]
    out = []
    for side, aux in ((None, (None, None)), (torch.cuda.Stream(dev), (torch.cuda.Stream(dev), torch.cuda.Stream(dev)))):
        for p in params:
            p.grad = None
        loss, _ = ppo.ppo_loss(pol, val, batch, cfg, fused=True, side_stream=side, aux_streams=aux)
        loss.backward()
        torch.cuda.synchronize()
        out.append((float(loss.detach()), [p.grad.clone() for p in params]))
        del loss
    # (the head kernel accumulates its sums with float atomics: equal up to summation order)
    assert abs(out[0][0] - out[1][0]) <= 1e-5 * abs(out[0][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max()) + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm", [1.0, None])
def test_native_clip_adam_matches_torch(max_norm):
    """`pgtt_adam_clip` (flat vector, two launches) against `clip_grad_norm_` + `torch.optim.Adam` over six steps with gradients
    large enough to be clipped: parameters agree to float rounding."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(1)
    shapes = [(171, 512), (512,), (512, 256), (256,), (128, 24), (24,), (215, 512), (1,)]
    ref = [torch.randn(s, generator=g, device=dev).requires_grad_() for s in shapes]
    mine = [p.detach().clone().requires_grad_() for p in ref]
    opt = torch.optim.Adam(ref, lr=3e-4, eps=1e-8)
    flat = ppo.FlatAdam(mine, 3e-4, eps=1e-8)
    for step in range(6):
        grads = [torch.randn(s, generator=g, device=dev) * (0.05 if step % 2 else 0.001) for s in shapes]
        for p, q, gr in zip(ref, mine, grads):
            p.grad = gr.clone(); q.grad = gr.clone() * 4.0       # the native step is given 4 x the gradient and grad_scale = 1/4
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt.step()
        flat.step(flat.flat_grad(), max_norm, 0.25)
        flat.zero_grad()
    torch.cuda.synchronize()
    assert float(flat.t) == 6.0
    for p, q in zip(ref, mine):
        assert q.data_ptr() >= flat.flat.data_ptr() and float((p - q).detach().abs().max()) <= 2e-6 * max(float(p.detach().abs().max()), 1.0)


def test_train_follows_the_brax_evaluation_schedule_and_honours_the_early_stop(monkeypatch):
    """`ppo.train` with an eval env: `num_evals` evaluations, the first before any training, ceil(num_timesteps / ((num_evals - 1)
    * env steps per training step)) training steps between two of them (brax ppo.train as called by training/train.py:242-263), and a
    truthy `progress_fn` return (the convergence verdict of training/train.py:224-229) ends the run. Trainer and evaluator are stubs:
    this is the host-side loop only."""
    from phase_guided_terrain_traversal_b200 import ppo
    log = []

    class FakeTrainer:
        world, rank, group, dev = 1, 0, None, "cpu"

        def __init__(self, wenv, state, cfg):
            self.env_steps, self.cfg = 0, cfg

        def training_step(self):
            self.env_steps += self.cfg.unroll_length * self.cfg.batch_size * self.cfg.num_minibatches
            log.append(("train", self.env_steps))
            return {"reward_per_step": 0.0}

    class FakeEvaluator:
        def __init__(self, eval_env, wrap_env_fn, randomization_fn, cfg, trainer, num_eval_envs, deterministic_eval, seed):
            assert eval_env == "eval-env" and num_eval_envs == 128 and deterministic_eval is False
            self.trainer = trainer

        def run_evaluation(self, training_metrics=None):
            log.append(("eval", self.trainer.env_steps))
            return {"eval/episode_reward": float(self.trainer.env_steps)}

    class FakeWrapped:
        def reset(self, keys):
            return "state"

    monkeypatch.setattr(ppo, "PPOTrainer", FakeTrainer)
    monkeypatch.setattr(ppo, "Evaluator", FakeEvaluator)
    cfg = ppo.PPOConfig(num_envs=64, batch_size=8, num_minibatches=8, unroll_length=20)
    per = 20 * 8 * 8
    cfg.num_timesteps = 5 * per            # 3 epochs (num_evals = 4) of ceil(5 / 3) = 2 training steps
    seen = []
    ppo.train("env", lambda e, **kw: FakeWrapped(), None, None, cfg, progress_fn=lambda n, m: seen.append((n, m["eval/episode_reward"])) and False,
              eval_env="eval-env", num_evals=4)
    assert log == [("eval", 0), ("train", per), ("train", 2 * per), ("eval", 2 * per), ("train", 3 * per), ("train", 4 * per), ("eval", 4 * per),
                   ("train", 5 * per), ("train", 6 * per), ("eval", 6 * per)]
    assert seen == [(0, 0.0), (2 * per, 2.0 * per), (4 * per, 4.0 * per), (6 * per, 6.0 * per)]
    # early stop: the verdict of the second evaluation ends the run; policy_params_fn still saw that evaluation's parameters
    log.clear()
    saved = []
    tr = ppo.train("env", lambda e, **kw: FakeWrapped(), None, None, cfg, progress_fn=lambda n, m: n >= 2 * per, policy_params_fn=lambda n, t: saved.append(n),
                   eval_env="eval-env", num_evals=4)
    assert log == [("eval", 0), ("train", per), ("train", 2 * per), ("eval", 2 * per)] and saved == [2 * per] and tr.stopped_early
    # num_evals = 1: no evaluation before training, one epoch holding every training step, one evaluation at the end
    log.clear()
    ppo.train("env", lambda e, **kw: FakeWrapped(), None, None, cfg, eval_env="eval-env", num_evals=1)
    assert log == [("train", k * per) for k in range(1, 6)] + [("eval", 5 * per)]
    # without an eval env: progress_fn after every training step, no evaluator
    log.clear()
    ppo.train("env", lambda e, **kw: FakeWrapped(), None, None, cfg, progress_fn=lambda n, m: None)
    assert log == [("train", k * per) for k in range(1, 6)]


@pytest.mark.gpu
def test_in_training_evaluation_reports_the_keys_progress_reads(train_cfg):
    """The real evaluator on a second env handle next to the training env (training/train.py:242-263 passes `eval_env`): 64 eval envs run one
    40-step episode with sampled actions; the metrics carry the first-episode sums `progress` divides by scale * episode length
    (train.py:214-216); the sums are consistent (reward terms add up to the episode reward, nobody exceeds the episode length); and the
    training env is not disturbed by the interleaved handle (same training metrics as a run without evaluation)."""
    from phase_guided_terrain_traversal_b200 import ppo, prng
    from phase_guided_terrain_traversal_b200.go2.base import METRIC_KEYS
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    from phase_guided_terrain_traversal_b200.go2.randomize_simple import domain_randomize
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
    n = 256
    cfg = ppo.PPOConfig(num_envs=n, batch_size=64, num_minibatches=8, num_updates_per_batch=1, episode_length=40, seed=5, use_cuda_graph=False)
    cfg.num_timesteps = 2 * 20 * 64 * 8
    out = {}
    for with_eval in (True, False):
        got = []
        tr = ppo.train(Joystick(task="flat_terrain", config=train_cfg), wrap_for_brax_training, domain_randomize, prng.env_keys(1, n), cfg,
                       progress_fn=lambda s, m: got.append((s, dict(m))) and False,
                       eval_env=Joystick(task="flat_terrain", config=train_cfg) if with_eval else None, num_evals=3, num_eval_envs=64)
        out[with_eval] = (got, dict(tr.metrics))
    got, _ = out[True]
    assert [s for s, _ in got] == [0, 20 * 64 * 8, 2 * 20 * 64 * 8]
    for _, m in got:
        assert {"eval/episode_reward", "eval/episode_reward_std", "eval/avg_episode_length", "eval/episode_reward/tracking_lin_vel",
                "eval/episode_reward/tracking_ang_vel"} <= set(m)
        assert all(math.isfinite(v) for v in m.values())
        assert 1.0 <= m["eval/avg_episode_length"] <= 40.0
        assert 0.0 <= m["eval/episode_reward/tracking_lin_vel"] <= 1.0 * 40 and 0.0 <= m["eval/episode_reward/tracking_ang_vel"] <= 0.5 * 40
    # Joystick.step: reward = clip(sum of the terms * dt, 0, 10000) (go2/joystick_pgtt.py:202): the summed terms bound the reward from above
    m = got[-1][1]
    terms = sum(m[f"eval/episode_{k}"] for k in METRIC_KEYS if k.startswith("reward/"))
    assert m["eval/episode_reward"] >= terms * 0.02 - 1e-3
    assert "training/reward_per_step" in got[-1][1] and "training/reward_per_step" not in got[0][1]
    a, b = out[True][1], out[False][1]
    assert a["env_steps"] == b["env_steps"] and abs(a["reward_per_step"] - b["reward_per_step"]) < 1e-6 and abs(a["total_loss"] - b["total_loss"]) < 1e-4, (a, b)


@pytest.mark.gpu
def test_native_sgd_step_equals_the_autograd_step(train_cfg):
    """The product SGD step (`PPOTrainer._sgd_body_native`: gathers -> pgtt_mlp_forward_gather x 2 -> pgtt_gae_moments -> pgtt_ppo_head ->
    pgtt_mlp_backward x 2 -> pgtt_adam_clip, no autograd) against the autograd statement of the same step (`_sgd_body` over `ppo_loss`) on the
    same minibatch from the same parameter / optimiser state: loss terms to 1e-5, the flat gradient to 1e-4 of its largest entry, the
    parameters after the Adam step to 2 lr (Adam's first steps are sign-like: an entry whose gradient is ~0 may step the other way)."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo, prng
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    from phase_guided_terrain_traversal_b200.go2.randomize_simple import domain_randomize
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
    n = 256
    cfg = ppo.PPOConfig(num_envs=n, batch_size=64, num_minibatches=8, num_updates_per_batch=1, use_cuda_graph=False, seed=9)
    env = Joystick(task="flat_terrain", config=train_cfg)
    keys = prng.env_keys(4, n)
    wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=lambda m: domain_randomize(m, rng=keys))
    tr = ppo.PPOTrainer(wenv, wenv.reset(keys), cfg)
    assert tr._native_step_ok()
    tr.training_step()                                   # fills the transition store, the normaliser statistics, the permutation, the noise
    opt = tr.flat_opt
    saved = [t.clone() for t in (opt.flat, opt.m, opt.v, opt.t)]
    tr._mbi.fill_(3)
    # autograd statement
    m_auto = {k: float(v) for k, v in tr._sgd_body(tr._minibatch()).items()}
    g_auto = torch.cat([p.grad.reshape(-1) for p in tr.params]).clone()
    p_auto = opt.flat.clone()
    for dst, src in zip((opt.flat, opt.m, opt.v, opt.t), saved):
        dst.copy_(src)
    # product path
    m_nat = {k: float(v) for k, v in tr._sgd_body_native().items()}
    torch.cuda.synchronize()
    g_nat, p_nat = tr._nb["flat_g"], opt.flat
    for k in m_auto:
        assert abs(m_nat[k] - m_auto[k]) <= 1e-5 * max(1.0, abs(m_auto[k])), (k, m_nat[k], m_auto[k])
    assert float((g_nat - g_auto).abs().max()) <= 1e-4 * float(g_auto.abs().max()), (float((g_nat - g_auto).abs().max()), float(g_auto.abs().max()))
    assert float(g_auto.abs().max()) > 0
    assert float((p_nat - p_auto).abs().max()) <= 2.0 * cfg.learning_rate + 1e-7
    assert float((p_nat - p_auto).abs().mean()) <= 0.02 * cfg.learning_rate


@pytest.mark.gpu
def test_column_moment_kernel_matches_float64_statistics():
    """`sharding.allreduce_moments` on a large CUDA fp32 matrix runs the hand-written float64 column-moment kernel (`pgtt_col_moments`); a
    column slice of a padded store (row stride 172 for 171 columns) must give the float64 mean / variance of torch to 1e-12, and the
    running statistics built from it must equal the ones built from the generic path."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo, sharding
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    store = torch.randn(21, 512, 172, device="cuda", generator=g) * 3.0 + 0.7
    x = store[:20, :, :171].reshape(-1, 171)
    assert x.data_ptr() == store.data_ptr() and x.stride(0) == 172                       # a view: nothing was copied
    n, mean, var = sharding.allreduce_moments(x)
    xd = x.double()
    assert float(n) == x.shape[0]
    assert float((mean - xd.mean(0)).abs().max()) < 1e-12 and float((var - xd.var(0, unbiased=False)).abs().max()) < 1e-10
    a, b = ppo.RunningStats(171, x.device), ppo.RunningStats(171, x.device)
    a.update(store[:20, :, :171])
    b.update(store[:20, :, :171].cpu().to("cuda").contiguous()[:, :64])                  # small batches take the generic path ...
    b.update(store[:20, :, :171].contiguous()[:, 64:])                                  # ... large ones the kernel: merged statistics agree
    assert float((a.mean - b.mean).abs().max()) < 1e-12 and float((a.std - b.std).abs().max()) < 1e-10


@pytest.mark.gpu
def test_gae_moment_variants_agree():
    """`pgtt_gae_moments` (single rank: mean / std inside the GAE launch) against `pgtt_gae_sums` + `pgtt_moments_finalize` (multi-rank form: float64 sums,
    all-reduce in between - here one rank) and against torch on the advantages the plain `pgtt_gae` writes: same vs / adv bit for bit, moments to 1e-6."""
    import ctypes as C
    import torch
    from phase_guided_terrain_traversal_b200 import _native as nat
    lib = nat.load_library()
    g = torch.Generator(device="cuda"); g.manual_seed(8)
    T, B = 20, 256
    r = lambda *s: torch.randn(s, device="cuda", generator=g)
    trunc = (torch.rand((T, B), device="cuda", generator=g) < 0.02).float()
    disc = 1.0 - (torch.rand((T, B), device="cuda", generator=g) < 0.05).float()
    rew, val = r(T, B).abs(), r(T + 1, B)
    out = {}
    for kind in ("plain", "moments", "sums"):
        vs, adv, mom = torch.empty(T, B, device="cuda"), torch.empty(T, B, device="cuda"), torch.zeros(2, device="cuda")
        sums = torch.zeros(3, dtype=torch.float64, device="cuda")
        args = (trunc.data_ptr(), disc.data_ptr(), rew.data_ptr(), val.data_ptr(), T, B, 0.95, 0.97, 1.0, vs.data_ptr(), adv.data_ptr())
        if kind == "plain":
            assert lib.pgtt_gae(*args, None) == 0
            mom = torch.stack([adv.mean(), adv.std(unbiased=False)])
        elif kind == "moments":
            assert lib.pgtt_gae_moments(*args, mom.data_ptr(), None) == 0
        else:
            assert lib.pgtt_gae_sums(*args, sums.data_ptr(), None) == 0
            assert lib.pgtt_moments_finalize(sums.data_ptr(), mom.data_ptr(), None) == 0
        torch.cuda.synchronize()
        out[kind] = (vs, adv, mom)
    for kind in ("moments", "sums"):
        assert torch.equal(out[kind][0], out["plain"][0]) and torch.equal(out[kind][1], out["plain"][1])
        assert float((out[kind][2] - out["plain"][2]).abs().max()) < 1e-6, (kind, out[kind][2], out["plain"][2])
