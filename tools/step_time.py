"""Device time of whole wrapped control steps through pgtt_step (the product call), back to back and with the L2 flushed between steps.
python tools/step_time.py [task] [N] [level] [steps] [dr]   (PGTT_FUSE_TASK=0|1, PGTT_KERNEL=warp|quad select the variant)"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from phase_guided_terrain_traversal_b200 import model as gm, terrain
from phase_guided_terrain_traversal_b200.abi_env import AbiEnv
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
task = sys.argv[1] if len(sys.argv) > 1 else "stairs"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
level = sys.argv[3] if len(sys.argv) > 3 else "level1"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
dr = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
m = gm.compile_model(task); cfg = training_overrides(default_config())
env = AbiEnv(m, cfg, N)
keys = np.stack([np.zeros(N, dtype=np.uint32), np.arange(N, dtype=np.uint32)], 1)
if task == "stairs":
    env.set_terrain(terrain.load_terrain(level)); env.randomize(keys, dynamics=dr)
env.reset(keys)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
acts = [torch.rand((N, 12), generator=g, device="cuda") * 2 - 1 for _ in range(16)]
for i in range(30): env.step(acts[i % 16])
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
l0 = env.launch_count()
for cold in (False, True):
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    for i in range(steps):
        if cold: flush.fill_(i & 0xFF)
        ev[i][0].record(); env.step_ptr(acts[i % 16].data_ptr(), wrapped=True); ev[i][1].record()
    torch.cuda.synchronize()
    ms = np.mean([e[0].elapsed_time(e[1]) for e in ev])
    print(f"{'L2-flushed' if cold else 'back-to-back'} N={N} {level} dr={dr}: step {ms:.4f} ms  {N / ms * 1e3:.3e} env-steps/s  launches/step {(env.launch_count() - l0) / (2 * steps if cold else steps):.1f} "
          f"niter {env.get('solver_niter').mean():.2f} done-rate {env.get('done').mean():.3f}")
