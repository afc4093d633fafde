"""Device time of the hand-written learner GEMMs (csrc/pgtt_learner.cu) per layer shape, against torch fp32 / tf32 matmuls."""
import ctypes as C, sys
sys.path.insert(0, ".")
import torch
from phase_guided_terrain_traversal_b200 import _native as nat
lib = nat.load_library()
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M = 5120


def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tot = {"fwd": 0, "bx": 0, "bp": 0, "tfwd": 0, "tbx": 0, "tbp": 0}
for K, N, ldx in [(171, 512, 172), (512, 256, 512), (256, 128, 256), (128, 24, 128), (215, 512, 216), (512, 256, 512), (256, 128, 256), (128, 1, 128)]:
    x = torch.randn(M, ldx, device="cuda"); w = torch.randn(K, N, device="cuda"); b = torch.randn(N, device="cuda"); dy = torch.randn(M, N, device="cuda")
    y = torch.empty(M, N, device="cuda"); z = torch.empty(M, N, device="cuda"); dx = torch.empty(M, ldx, device="cuda"); dw = torch.empty(K, N, device="cuda"); db = torch.empty(N, device="cuda")
    sc = torch.empty(int(lib.pgtt_linear_backward_params_scratch(M, K, N)), device="cuda")
    xk = x[:, :K].contiguous()
    t_f = timeit(lambda: lib.pgtt_linear_forward(p(x), ldx, p(w), p(b), M, K, N, 1, p(y), p(z), st))
    t_x = timeit(lambda: lib.pgtt_linear_backward_input(p(dy), p(w), M, K, N, p(dx), ldx, None, st))
    t_p = timeit(lambda: lib.pgtt_linear_backward_params(p(x), ldx, p(dy), M, K, N, p(dw), p(db), p(sc), st))
    torch.set_float32_matmul_precision("highest")
    r_f = timeit(lambda: torch.addmm(b, xk, w)); r_x = timeit(lambda: dy @ w.t()); r_p = timeit(lambda: xk.t() @ dy)
    print(f"K={K:4d} N={N:4d}: native fwd {t_f:6.1f} us  dX {t_x:6.1f}  dW+db {t_p:6.1f} | torch fp32 fwd {r_f:6.1f}  dX {r_x:6.1f}  dW {r_p:6.1f}")
    for k, v in zip(tot, (t_f, t_x, t_p, r_f, r_x, r_p)): tot[k] += v
print("sum over the 8 layers:", {k: round(v, 1) for k, v in tot.items()})
