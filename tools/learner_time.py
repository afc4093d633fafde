import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from phase_guided_terrain_traversal_b200 import ppo, prng, terrain
from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
from phase_guided_terrain_traversal_b200.go2.randomize import domain_randomize
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
import functools
for prec in (sys.argv[1:] or ["highest", "high"]):
    n=4096
    import os
    cfg=ppo.PPOConfig(num_envs=n, matmul_precision=prec, parallel_nets=os.environ.get("PGTT_PAR","1")=="1", native_optimizer=os.environ.get("PGTT_NOPT","1")=="1")
    env=Joystick(task="stairs", config=training_overrides(default_config()))
    keys=prng.env_keys(1,n)
    wenv=wrap_for_brax_training(env, episode_length=1000, randomization_fn=functools.partial(domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain("level1"), dynamics=True))
    st=wenv.reset(keys)
    tr=ppo.PPOTrainer(wenv, st, cfg)
    tr.training_step(); torch.cuda.synchronize()
    t0=time.time()
    for _ in range(3): m=tr.training_step()
    torch.cuda.synchronize(); dt=(time.time()-t0)/3
    # rollout-only time
    t1=time.time()
    for _ in range(2*3): tr.collector.collect()
    torch.cuda.synchronize(); dr=(time.time()-t1)/3
    print(prec, "training step", round(dt*1e3,1), "ms; rollout part", round(dr*1e3,1), "ms;", round(163840/dt/1e6,3), "M env-steps/s", m["total_loss"])
    env.close()
