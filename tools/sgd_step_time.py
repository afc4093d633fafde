"""Device time of ONE graphed SGD step of the PPO learner (reference sizes: 4096 envs, minibatch 256 x 20)."""
import sys, os, functools; sys.path.insert(0, '.')
import torch
from phase_guided_terrain_traversal_b200 import ppo, prng, terrain
from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
from phase_guided_terrain_traversal_b200.go2.randomize import domain_randomize
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
n = 4096
nm = os.environ.get("PGTT_NATIVE_MLP", "1")
cfg = ppo.PPOConfig(num_envs=n, parallel_nets=os.environ.get("PGTT_PAR", "1") == "1", native_mlp={"1": True, "0": False, "layers": "layers"}[nm])
env = Joystick(task="stairs", config=training_overrides(default_config()))
keys = prng.env_keys(1, n)
wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=functools.partial(domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain("level1"), dynamics=True))
tr = ppo.PPOTrainer(wenv, wenv.reset(keys), cfg)
tr.training_step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e0.record()
    for _ in range(100): tr._graph.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"native_mlp={nm} parallel_nets={cfg.parallel_nets}: SGD step {e0.elapsed_time(e1) * 10:.1f} us")
