"""Device time of the two kernels of a control step (physics, task), each bracketed by its own CUDA event pair
(development aid; bench.py is the contract).  python tools/kernel_times.py [task] [N] [level] [steps] [dr]"""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
from phase_guided_terrain_traversal_b200 import model as gm, terrain
from phase_guided_terrain_traversal_b200.abi_env import AbiEnv
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides

task = sys.argv[1] if len(sys.argv) > 1 else "stairs"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
level = sys.argv[3] if len(sys.argv) > 3 else "level1"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
dr = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
m = gm.compile_model(task); cfg = training_overrides(default_config())
env = AbiEnv(m, cfg, N)
keys = np.stack([np.zeros(N, dtype=np.uint32), np.arange(N, dtype=np.uint32)], 1)
if task == "stairs":
    env.set_terrain(terrain.load_terrain(level))
    env.randomize(keys, dynamics=dr)
env.reset(keys)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
acts = [torch.rand((N, 12), generator=g, device="cuda") * 2 - 1 for _ in range(16)]
for i in range(30):
    env.step(acts[i % 16])
torch.cuda.synchronize()
part = env.lib.pgtt_internal_step_part
part.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cold in (False, True):
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for i in range(steps):
        if cold:
            flush.fill_(i & 0xFF)
        ev[i][0].record()
        part(env.h, acts[i % 16].data_ptr(), 1, 0, None)
        ev[i][1].record()
        part(env.h, acts[i % 16].data_ptr(), 1, 1, None)
        ev[i][2].record()
    torch.cuda.synchronize()
    ph = np.mean([e[0].elapsed_time(e[1]) for e in ev]); tk = np.mean([e[1].elapsed_time(e[2]) for e in ev])
    print(f"{'L2-flushed' if cold else 'back-to-back'} task={task} N={N} level={level} dr={dr} gen={env.lib.pgtt_step_kernel_generation(env.h)}: "
          f"physics {ph:.4f} ms  task {tk:.4f} ms  sum {ph + tk:.4f} ms  {N / (ph + tk) * 1e3:.3e} env-steps/s  "
          f"niter {env.get('solver_niter').mean():.2f} done-rate {env.get('done').mean():.3f} nan {np.isnan(env.get('qpos')).any()}")
