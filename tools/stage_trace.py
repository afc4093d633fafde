"""Where does the warp-per-env physics kernel spend its cycles? One launch with per-warp SM-clock stamps after each stage of
each substep (pgtt_internal_physics_trace); prints mean stage durations and barrier waits per substep.
    python tools/stage_trace.py [N] [level]"""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
from phase_guided_terrain_traversal_b200 import model as gm, terrain
from phase_guided_terrain_traversal_b200.abi_env import AbiEnv
from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
level = sys.argv[2] if len(sys.argv) > 2 else "level1"
m = gm.compile_model("stairs"); cfg = training_overrides(default_config())
env = AbiEnv(m, cfg, N)
keys = np.stack([np.zeros(N, dtype=np.uint32), np.arange(N, dtype=np.uint32)], 1)
env.set_terrain(terrain.load_terrain(level)); env.randomize(keys, dynamics=False); env.reset(keys)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
acts = [torch.rand((N, 12), generator=g, device="cuda") * 2 - 1 for _ in range(16)]
for i in range(30):
    env.step(acts[i % 16])
fn = env.lib.pgtt_internal_physics_trace
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
tr = torch.zeros((N, 40), dtype=torch.int64, device="cuda")
rc = fn(env.h, acts[0].data_ptr(), tr.data_ptr(), None); assert rc == 0, rc
torch.cuda.synchronize()
t = tr.cpu().numpy().astype(np.float64)
niter = env.get("solver_niter")
names = ["position", "velocity", "collision work", "collision barrier", "rows+factor", "solver work", "solver barrier"]
wpb = 28
print(f"N={N} level={level}: kernel span of a warp, cycles: mean {np.mean(t[:, 32] - t[:, 0]):.0f} max {np.max(t[:, 32] - t[:, 0]):.0f}  (1965 MHz: {np.mean(t[:, 32] - t[:, 0]) / 1.965e6:.4f} ms)")
tot = np.zeros(7)
for s in range(4):
    b = 8 * s
    d = np.stack([t[:, b + k + 1] - t[:, b + k] for k in range(7)], 1)
    tot += d.mean(0)
    print(f"substep {s}: " + "  ".join(f"{names[k]} {d[:, k].mean():7.0f}" for k in range(7)) + f"   niter mean {niter[:, s].mean():.2f}")
    if s < 3:
        print(f"           euler -> next forward {np.mean(t[:, b + 8] - t[:, b + 7]):7.0f}")
print("sum over substeps: " + "  ".join(f"{names[k]} {tot[k]:7.0f} ({100 * tot[k] / np.mean(t[:, 32] - t[:, 0]):4.1f} %)" for k in range(7)))
# solver duration by iteration count, first substep
for it in range(1, 6):
    sel = niter[:, 0] == it
    if sel.any():
        print(f"  substep 0, {it} Newton iterations: {sel.mean() * 100:4.1f} % of envs, solver own work {np.mean(t[sel, 6] - t[sel, 5]):7.0f} cycles")
