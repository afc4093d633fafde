"""Hand-written tcgen05 dense layers of the learner (csrc/pgtt_learner.cu) against torch fp32 statements of the same
products (float64 for the reference values). Tolerance: the split-bf16 products carry ~16 mantissa bits -> relative error of
a length-K dot product ~2^-17 sqrt(K) of |x||w|; asserted at 1e-4 of the result's max-norm (the gradient-parity bar)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from phase_guided_terrain_traversal_b200 import _native as nat
    return nat.load_library()


def _ptr(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("M,K,N,ldx", [(5120, 171, 512, 172), (5120, 512, 256, 512), (5120, 256, 128, 256), (5120, 128, 24, 128), (5120, 215, 512, 216),
                                       (5120, 128, 1, 128), (300, 37, 50, 40), (128, 64, 128, 64)])
def test_linear_forward_backward_match_fp64(M, K, N, ldx):
    import torch
    lib = _lib()
    g = torch.Generator(device="cuda"); g.manual_seed(M + K + N)
    x = torch.randn((M, ldx), generator=g, device="cuda"); x[:, K:] = 7.0          # padding columns must be ignored
    w = torch.randn((K, N), generator=g, device="cuda") / np.sqrt(K)
    b = torch.randn((N,), generator=g, device="cuda")
    dy = torch.randn((M, N), generator=g, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    xd, wd, bd, dyd = x[:, :K].double(), w.double(), b.double(), dy.double()
    # forward, plain and with the fused SiLU
    y = torch.empty((M, N), device="cuda"); z = torch.empty((M, N), device="cuda")
    assert lib.pgtt_linear_forward(_ptr(x), ldx, _ptr(w), _ptr(b), M, K, N, 0, _ptr(y), None, st) == 0, lib.pgtt_learner_last_error()
    ref = xd @ wd + bd
    assert float((y.double() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert lib.pgtt_linear_forward(_ptr(x), ldx, _ptr(w), _ptr(b), M, K, N, 1, _ptr(y), _ptr(z), st) == 0
    assert float((z.double() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert float((y.double() - torch.nn.functional.silu(ref)).abs().max()) <= 1e-4 * float(ref.abs().max())
    # input gradient
    dx = torch.full((M, ldx), 3.0, device="cuda")
    assert lib.pgtt_linear_backward_input(_ptr(dy), _ptr(w), M, K, N, _ptr(dx), ldx, None, st) == 0
    ref = dyd @ wd.t()
    assert float((dx[:, :K].double() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert bool((dx[:, K:] == 3.0).all())                                          # nothing written beyond the K columns
    # ... fused with the derivative of the SiLU that produced x: gradient wrt that SiLU's pre-activation
    zin = torch.randn((M, ldx), generator=g, device="cuda")
    dzin = torch.empty((M, ldx), device="cuda")
    assert lib.pgtt_linear_backward_input(_ptr(dy), _ptr(w), M, K, N, _ptr(dzin), ldx, _ptr(zin), st) == 0
    zz = zin[:, :K].double().requires_grad_(True)
    torch.nn.functional.silu(zz).backward(ref)
    assert float((dzin[:, :K].double() - zz.grad).abs().max()) <= 1e-4 * float(zz.grad.abs().max())
    # parameter gradients, twice: the split sum is ordered, so the result is reproducible bit for bit
    scratch = torch.empty(int(lib.pgtt_linear_backward_params_scratch(M, K, N)), device="cuda")
    outs = []
    for _ in range(2):
        dw = torch.empty((K, N), device="cuda"); db = torch.empty((N,), device="cuda")
        assert lib.pgtt_linear_backward_params(_ptr(x), ldx, _ptr(dy), M, K, N, _ptr(dw), _ptr(db), _ptr(scratch), st) == 0
        outs.append((dw, db))
    ref = xd.t() @ dyd
    assert float((outs[0][0].double() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert float((outs[0][1].double() - dyd.sum(0)).abs().max()) <= 1e-4 * float(dyd.sum(0).abs().max() + 1)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # silu backward
    dz = torch.empty_like(dy)
    assert lib.pgtt_silu_backward(_ptr(dy), _ptr(z), _ptr(dz), M * N, st) == 0
    zz = z.double().requires_grad_(True)
    torch.nn.functional.silu(zz).backward(dyd)
    assert float((dz.double() - zz.grad).abs().max()) <= 1e-5 * float(zz.grad.abs().max())


def test_split_precision_is_far_tighter_than_plain_bf16():
    """The three-MMA split is what makes the products fp32-grade: the same layer in torch bf16 misses by ~1e-2."""
    import torch
    lib = _lib()
    M, K, N = 1024, 512, 256
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn((M, K), generator=g, device="cuda"); w = torch.randn((K, N), generator=g, device="cuda") / np.sqrt(K)
    y = torch.empty((M, N), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.pgtt_linear_forward(_ptr(x), K, _ptr(w), None, M, K, N, 0, _ptr(y), None, st) == 0
    ref = x.double() @ w.double()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    err_bf16 = float(((x.bfloat16() @ w.bfloat16()).double() - ref).abs().max() / ref.abs().max())
    err_fp32 = float(((x @ w).double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5 and err < err_bf16 / 100, (err, err_bf16, err_fp32)
