"""Kernel-time table of the graphed SGD step (torch profiler / CUPTI): python tools/learner_profile.py [highest|high]"""
import functools, os, sys
sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from phase_guided_terrain_traversal_b200 import ppo, prng, terrain
from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
from phase_guided_terrain_traversal_b200.go2.randomize import domain_randomize
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training

prec = sys.argv[1] if len(sys.argv) > 1 else "high"
n = 4096
cfg = ppo.PPOConfig(num_envs=n, matmul_precision=prec, parallel_nets=os.environ.get("PGTT_PAR", "1") == "1")
env = Joystick(task="stairs", config=training_overrides(default_config()))
keys = prng.env_keys(1, n)
wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=functools.partial(domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain("level1"), dynamics=True))
tr = ppo.PPOTrainer(wenv, wenv.reset(keys), cfg)
tr.training_step(); tr.training_step()
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    tr._sgd_step(i)
e1.record(); torch.cuda.synchronize()
print(f"{prec}: graphed SGD step {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(reps):
        tr._sgd_step(i)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print(f"sum of kernel time per step {tot / reps:.1f} us over {sum(r[1] for r in rows) / reps:.0f} kernels")
for k, c, t in sorted(rows, key=lambda r: -r[2])[:28]:
    print(f"{100 * t / tot:5.1f}%  n/step={c / reps:5.1f}  mean={t / c:7.2f} us  {k[:110]}")
