"""Where one PPO training step spends its time (reference sizes), by wrapping the phases with host-synchronised timers."""
import sys, os, functools, time; sys.path.insert(0, '.')
import torch
from phase_guided_terrain_traversal_b200 import ppo, prng, terrain
from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
from phase_guided_terrain_traversal_b200.go2.randomize import domain_randomize
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("PGTT_N", 4096))                  # envs per rank
bs = n * world // 16                                      # one unroll per training step at 32 minibatches... (4096 envs, 1 GPU: 256 -> 2 unrolls)
cfg = ppo.PPOConfig(num_envs=n * world, batch_size=int(os.environ.get("PGTT_BS", 256 * world)))
env = Joystick(task="stairs", config=training_overrides(default_config()), device=local)
keys = prng.env_keys(1, n, offset=rank * n)
wenv = wrap_for_brax_training(env, episode_length=1000, randomization_fn=functools.partial(domain_randomize, rng=keys, terrain_matrix=terrain.load_terrain(os.environ.get("PGTT_LEVEL", "level1")), dynamics=True))
tr = ppo.PPOTrainer(wenv, wenv.reset(keys), cfg)
tr.training_step(); torch.cuda.synchronize()
acc = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r
    return w
tr.collector.collect = timed("collect", tr.collector.collect)
tr.norm_state.update = timed("norm_update", tr.norm_state.update); tr.norm_priv.update = timed("norm_update", tr.norm_priv.update)
tr.norm_state.normalize = timed("normalize", tr.norm_state.normalize); tr.norm_priv.normalize = timed("normalize", tr.norm_priv.normalize)
tr._sgd_step = timed("sgd", tr._sgd_step)
tr._sync_policy = timed("sync_policy", tr._sync_policy)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): tr.training_step()
torch.cuda.synchronize(); tot = time.perf_counter() - t0
if rank == 0: print(f"world {world}, {n} envs per rank, batch {cfg.batch_size}, unrolls per step {tr.unrolls_per_step}, minibatch segments per rank {tr.mb}:", {k: round(v / 3 * 1e3, 2) for k, v in acc.items()}, "total ms", round(tot / 3 * 1e3, 2), "unaccounted", round((tot - sum(acc.values())) / 3 * 1e3, 2))

if world > 1:
    tr._graph = None
    import gc; gc.collect(); torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
