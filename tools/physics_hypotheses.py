#!/usr/bin/env python
"""Which reading of the un-pinned MJX details do the reference's own policies prefer?  

The reference ships 88 policies trained in real MJX; each pickle carries the brax observation normaliser (running mean /
std of the 171 policy + 215 privileged observations over the whole training run). Closed loop, a policy only reproduces
its own statistics in a simulator that behaves like the one it was trained in - so for every hypothesis about a detail
SURVEY App. A could not verify, this tool runs the CPU oracle closed loop with shipped policies (sampled actions, DR on, the
terrain level the policy was trained on, full 1000-step episodes with auto-reset) and scores the mismatch between the
observation statistics it produces and the statistics stored in the policy:

    score = mean over dims of |mean_sim - mean_ref| / std_ref   (z units; physics-sensitive groups listed separately)

Caveat stated in the output: the stored statistics are cumulative over training (early, bad policies included), so the
absolute scores are not zero for the true simulator; the comparison BETWEEN hypotheses is what counts.

    python tools/physics_hypotheses.py [--envs 512] [--steps 1000] [--policies policy177:level13 policy3:level07 policy180:level1]
"""
import argparse, copy, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from oracle.oracle import Oracle
from phase_guided_terrain_traversal_b200 import model as gm, policy_io, prng, terrain
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides

REF = Path("/root/reference/policy_folder")

GROUPS = {   # slices of the 215-dim privileged observation (state = first 171)
    "gyro": slice(0, 3), "gravity": slice(3, 6), "joint_pos": slice(6, 18), "joint_vel": slice(18, 30), "scan": slice(38, 155),
    "local_linvel": slice(171, 174), "accelerometer": slice(174, 177), "global_angvel": slice(177, 180), "actuator_force": slice(180, 192),
    "last_contact": slice(192, 196), "feet_linvel": slice(196, 208), "feet_air_time": slice(208, 212),
}


def hypotheses():
    def base(m): return m
    def no_act_damping(m): m.act_biasprm = m.act_biasprm.copy(); m.act_biasprm[:, 2] = 0.0; return m            # Q12: <position> drops the class default biasprm[2]
    def no_ctrl_clamp(m): m.act_ctrlrange = m.act_ctrlrange.copy(); m.act_ctrlrange[:, 0] = -1e3; m.act_ctrlrange[:, 1] = 1e3; return m   # Q11
    def no_culling(m): m.max_geom_pairs = -1; return m                                                       # Q3: every penetrating pair competes for the 4 slots
    def cull_10(m): m.max_geom_pairs = 10; return m
    def solimp_foot(m):                                                                                        # contact solimp = the foot's, no mixing
        m.floor_solimp = m.foot_solimp.copy(); m.box_solimp = list(m.foot_solimp); return m
    def solimp_other(m):                                                                                       # contact solimp = the floor / box default
        m.foot_solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0]); return m
    def impratio_1(m): m.impratio = 1.0; return m
    def timestep_004(m): m.timestep = 0.004; return m                                                        # base.py:57 override not applied (XML value)
    def damping_2(m): m.dof_damping = m.dof_damping.copy(); m.dof_damping[6:] = 2.0; return m                # base.py:60 override not applied (XML class value)
    def kp_50(m):                                                                                              # base.py:61-62 override not applied
        m.act_gainprm = m.act_gainprm.copy(); m.act_biasprm = m.act_biasprm.copy(); m.act_gainprm[:, 0] = 50.0; m.act_biasprm[:, 1] = -50.0; return m
    return [("current (SURVEY App. A as built)", base), ("Q12 actuator biasprm[2] dropped (joint damping 0.5, not 1.0)", no_act_damping),
            ("Q11 ctrl not clamped to ctrlrange", no_ctrl_clamp), ("Q3 no broad-phase culling", no_culling), ("Q3 max_geom_pairs = 10", cull_10),
            ("contact solimp = foot's (no mix)", solimp_foot), ("contact solimp = floor/box default (no mix)", solimp_other),
            ("impratio 1", impratio_1), ("timestep 0.004 (XML, override lost)", timestep_004), ("joint damping 2.0 (XML class value)", damping_2),
            ("Kp 50 (XML class value)", kp_50)]


def mlp(kernels, biases, x):
    for i, (k, b) in enumerate(zip(kernels, biases)):
        x = x @ k + b
        if i + 1 < len(kernels):
            x = x / (1.0 + np.exp(-x))          # swish, deploy/policy_net.py:56-64
    return x


def run(m, cfg, pol, level, n, steps, seed):
    orc = Oracle(m, cfg, n, "f32native")
    keys = prng.env_keys(seed, n)
    orc.randomize(keys, terrain.load_terrain(level), True)
    orc.reset(keys + np.uint32(1))
    rng = np.random.default_rng(seed)
    ks, bs = pol["policy"]
    s1 = np.zeros(215); s2 = np.zeros(215); cnt = 0
    done_sum = 0.0; rew = 0.0
    for t in range(steps):
        ob = orc.get("obs_state")
        pv = orc.get("obs_priv")
        s1 += pv.sum(0); s2 += (pv * pv).sum(0); cnt += n
        logits = mlp(ks, bs, ((ob - pol["mean"]) / pol["std"]).astype(np.float32))
        loc, sc = logits[:, :12], np.logaddexp(0.0, logits[:, 12:]) + 0.001
        act = np.tanh(loc + sc * rng.standard_normal(loc.shape))
        orc.step(act.astype(np.float64), wrapped=True)
        d = orc.get("done")[:, 0]; tr = orc.get("truncation")[:, 0]
        done_sum += float((d * (1 - tr)).sum()); rew += float(orc.get("reward").sum())
    mean = s1 / cnt
    std = np.sqrt(np.maximum(s2 / cnt - mean * mean, 0))
    return mean, std, done_sum / (n * steps) * 1000.0, rew / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=384)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--policies", nargs="*", default=["policy177:level13", "policy3:level07", "policy180:level1"])
    ap.add_argument("--only", nargs="*", type=int, default=None, help="indices of hypotheses to run")
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r02_physics_hypotheses.json"))
    a = ap.parse_args()
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    cfg = training_overrides(default_config())
    results = {}
    for spec in a.policies:
        name, level = spec.split(":")
        pol = policy_io.load_policy(REF / name)
        ref_mean = np.concatenate([pol["mean"], pol["value_mean"][171:]]) if "value_mean" in pol else None
        ref_std = np.concatenate([pol["std"], pol["value_std"][171:]])
        # the policy statistics of the first 171 dims and the value statistics of the same dims agree (same stream of observations)
        results[spec] = {}
        for hi, (hname, fn) in enumerate(hypotheses()):
            if a.only is not None and hi not in a.only:
                continue
            m = fn(copy.deepcopy(gm.compile_model("stairs", sim_dt=cfg.sim_dt, Kp=cfg.Kp, Kd=cfg.Kd)))
            t0 = time.time()
            mean, std, falls, ep_rew = run(m, cfg, pol, level, a.envs, a.steps, seed=11)
            z = np.abs(mean - ref_mean) / np.maximum(ref_std, 1e-6)
            ls = np.abs(np.log(np.maximum(std, 1e-9) / np.maximum(ref_std, 1e-9)))
            rec = {"z_mean_all": float(z.mean()), "logstd_all": float(ls[ref_std > 1e-4].mean()), "terminations_per_1000_env_steps": falls, "reward_per_env": ep_rew,
                   "groups": {g: {"z": float(z[sl].mean()), "logstd": float(ls[sl].mean()), "sim_mean": float(mean[sl].mean()), "ref_mean": float(ref_mean[sl].mean()),
                                  "sim_std": float(std[sl].mean()), "ref_std": float(ref_std[sl].mean())} for g, sl in GROUPS.items()}}
            results[spec][hname] = rec
            g = rec["groups"]
            print(f"{spec:20s} {hname:62s} z {rec['z_mean_all']:.3f} logstd {rec['logstd_all']:.3f} falls/1k {falls:5.2f} | qvel std {g['joint_vel']['sim_std']:.2f}/{g['joint_vel']['ref_std']:.2f} "
                  f"contact {g['last_contact']['sim_mean']:.2f}/{g['last_contact']['ref_mean']:.2f} acc_z {mean[176]:.2f}/{ref_mean[176]:.2f} air {g['feet_air_time']['sim_mean']:.3f}/{g['feet_air_time']['ref_mean']:.3f} "
                  f"foot vel std {g['feet_linvel']['sim_std']:.2f}/{g['feet_linvel']['ref_std']:.2f} force std {g['actuator_force']['sim_std']:.2f}/{g['actuator_force']['ref_std']:.2f}  ({time.time() - t0:.0f} s)", flush=True)
    Path(a.out).write_text(json.dumps({"envs": a.envs, "steps": a.steps, "results": results}, indent=1) + "\n")


if __name__ == "__main__":
    main()
