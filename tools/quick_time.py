"""Quick device timing of the fused step kernel (development aid; bench.py is the contract)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from phase_guided_terrain_traversal_b200 import model as gm, terrain
from phase_guided_terrain_traversal_b200.abi_env import AbiEnv
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides

task = sys.argv[1] if len(sys.argv) > 1 else "stairs"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
level = sys.argv[3] if len(sys.argv) > 3 else "level1"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
m = gm.compile_model(task); cfg = training_overrides(default_config())
env = AbiEnv(m, cfg, N)
keys = np.stack([np.zeros(N, dtype=np.uint32), np.arange(N, dtype=np.uint32)], 1)
if task == "stairs":
    env.set_terrain(terrain.load_terrain(level))
    env.randomize(keys, dynamics=False)
env.reset(keys)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
acts = [torch.rand((N, 12), generator=g, device="cuda") * 2 - 1 for _ in range(16)]
for i in range(20):
    env.step(acts[i % 16])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    env.step_ptr(acts[i % 16].data_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"task={task} N={N} level={level}: {ms:.4f} ms/step  {N / ms * 1e3:.3e} env-steps/s  niter mean {env.get('solver_niter').mean():.2f} "
      f"done-rate {env.get('done').mean():.3f} nan {np.isnan(env.get('qpos')).any()}")
