import sys, functools; sys.path.insert(0,'.')
import numpy as np, torch
from phase_guided_terrain_traversal_b200 import prng, terrain, wrapper
from phase_guided_terrain_traversal_b200.evaluate import evaluate
from phase_guided_terrain_traversal_b200.go2 import joystick_pgtt, randomize
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
from phase_guided_terrain_traversal_b200.policy import PolicyNet
g = np.load("tests/golden/policy177.npz")
cfg = training_overrides(default_config())
for level in ("level07", "level13"):
    for which in ("policy177",):
        env = joystick_pgtt.Joystick(task="stairs", config=cfg)
        keys = prng.env_keys(4, 1000)
        wenv = wrapper.wrap_for_brax_training(env, episode_length=1000, randomization_fn=functools.partial(randomize.domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain(level)))
        wenv.reset(keys + np.uint32(1))
        net = PolicyNet()
        if which == "policy177": net.set_params([g[f"kernel{i}"] for i in range(4)], [g[f"bias{i}"] for i in range(4)], g["mean"], g["std"])
        else: net.init_random(0)
        r = evaluate(wenv, net, collect_obs_stats=True)
        print(level, which, {k: (round(v,3) if isinstance(v,float) else v) for k,v in r.items() if not (k.startswith("obs_") or k.startswith("priv_"))})
        if which == "policy177":
            m, s = r["obs_mean"], r["obs_std"]
            print("   obs stats ours vs policy normaliser: gravity_z", round(m[5],3), round(float(g["mean"][5]),3), "| cos phase std", round(s[30],3), round(float(g["std"][30]),3),
                  "| scan mean", round(m[38:155].mean(),3), round(float(g["mean"][38:155].mean()),3), "| gait_freq", round(m[155],3), round(float(g["mean"][155]),3),
                  "| cmd std", s[168:171].round(2), g["std"][168:171].round(2), "| qvel std mean", round(s[18:30].mean(),2), round(float(g["std"][18:30].mean()),2),
                  "| joint pos std", round(s[6:18].mean(),3), round(float(g["std"][6:18].mean()),3))
            pm, ps, gm_, gs_ = r["priv_mean"], r["priv_std"], g["priv_mean"], g["priv_std"]
            def row(name, sl):
                print(f"      {name:28s} ours mean {np.round(pm[sl],3)} std {np.round(ps[sl],3)} | policy normaliser mean {np.round(gm_[sl],3)} std {np.round(gs_[sl],3)}")
            row("local linvel", slice(171,174)); row("accelerometer", slice(174,177)); row("global angvel", slice(177,180))
            row("actuator force FR (hip,thigh,knee)", slice(180,183)); row("last_contact", slice(192,196)); row("feet air time", slice(208,212))
            row("foot linvel FR", slice(196,199)); row("gyro", slice(0,3)); row("joint vel FL", slice(18,21))
        env.close()
