"""Whole-MLP forward / backward on blocked split-bf16 tcgen05 GEMMs (csrc/pgtt_mlp.cu, `pgtt_mlp_*` of include/pgtt_b200.h) against
a float64 torch statement of the same network (SiLU between layers, as brax's make_ppo_networks builds it for
training/train.py:135-161). Tolerance: the split keeps ~16 mantissa bits per product - errors stay at a few 1e-6 of the
output scale, the same as an fp32 GEMM; asserted at 2e-5 (outputs) / 5e-5 (gradients) of each tensor's largest entry."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(dims, rows, ldx, seed):
    import torch
    from phase_guided_terrain_traversal_b200 import _native as nat
    lib = nat.load_library()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(seed)
    L = len(dims) - 1
    x = torch.randn(rows, ldx, device=dev, generator=g)
    ws = [torch.randn(dims[l], dims[l + 1], device=dev, generator=g) / np.sqrt(dims[l]) for l in range(L)]
    bs = [torch.randn(dims[l + 1], device=dev, generator=g) * 0.3 for l in range(L)]
    dy = torch.randn(rows, dims[-1], device=dev, generator=g)
    h = C.c_void_p()
    arr = (C.c_int * (L + 1))(*dims)
    assert lib.pgtt_mlp_create(L, arr, rows, 0, C.byref(h)) == 0, lib.pgtt_mlp_last_error()
    try:
        y = torch.empty(rows, dims[-1], device=dev)
        dws = [torch.full_like(w, float("nan")) for w in ws]
        dbs = [torch.full_like(b, float("nan")) for b in bs]
        vp = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for rep in range(2):      # the second pass re-uses every workspace buffer: nothing stale may leak
            assert lib.pgtt_mlp_forward(h, x.data_ptr(), ldx, vp(ws), vp(bs), y.data_ptr(), st) == 0, lib.pgtt_mlp_last_error()
            assert lib.pgtt_mlp_backward(h, dy.data_ptr(), vp(dws), vp(dbs), st) == 0, lib.pgtt_mlp_last_error()
        torch.cuda.synchronize()
    finally:
        lib.pgtt_mlp_destroy(h)
    # float64 reference
    xr = x[:, :dims[0]].double()
    wr = [w.double().requires_grad_() for w in ws]
    br = [b.double().requires_grad_() for b in bs]
    a = xr
    for l in range(L):
        a = a @ wr[l] + br[l]
        if l + 1 < L:
            a = torch.nn.functional.silu(a)
    (a * dy.double()).sum().backward()
    rel = lambda got, ref: float((got.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    return rel(y, a.detach()), [rel(dws[l], wr[l].grad) for l in range(L)], [rel(dbs[l], br[l].grad) for l in range(L)]


@pytest.mark.parametrize("dims,rows,ldx", [
    ((171, 512, 256, 128, 24), 5120, 172),      # policy network, reference minibatch (256 x 20 transitions), padded observation rows
    ((215, 512, 256, 128, 1), 5376, 216),       # value network incl. the bootstrap row
    ((162, 512, 256, 128, 24), 1280, 162),      # baseline-task observation width, small minibatch
    ((37, 96, 40, 8, 5), 333, 37),              # ragged everything: widths off the 8 / 32 / 128 grids, rows off the 32 / 128 grids
    ((64, 3), 77, 80),                          # a single layer
])
def test_forward_and_gradients_match_float64(dims, rows, ldx):
    ey, edw, edb = _run(dims, rows, ldx, seed=len(dims) * 1000 + rows)
    assert ey < 2e-5, ("y", ey)
    assert max(edw) < 5e-5, ("dW", edw)
    assert max(edb) < 5e-5, ("db", edb)


def test_argument_checks():
    from phase_guided_terrain_traversal_b200 import _native as nat
    lib = nat.load_library()
    h = C.c_void_p()
    assert lib.pgtt_mlp_create(0, (C.c_int * 1)(4), 8, 0, C.byref(h)) != 0
    assert lib.pgtt_mlp_create(2, (C.c_int * 3)(4, 0, 2), 8, 0, C.byref(h)) != 0
    assert b"pgtt_mlp_create" in lib.pgtt_mlp_last_error()
    assert lib.pgtt_mlp_create(1, (C.c_int * 2)(4, 2), 8, 0, C.byref(h)) == 0
    try:
        assert lib.pgtt_mlp_backward(h, None, None, None, None) != 0
    finally:
        lib.pgtt_mlp_destroy(h)


def test_fused_gather_and_normalisation_equal_the_materialised_minibatch():
    """`pgtt_mlp_forward_gather` (minibatch rows picked out of the time-major transition store and normalised inside the input
    kernel) against `pgtt_mlp_forward` on the minibatch gathered and normalised by torch: same network output and gradients
    (the two inputs differ by one fp32 rounding of the normalisation: asserted at 1e-5 of the largest entry)."""
    import torch
    from phase_guided_terrain_traversal_b200 import ppo
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(11)
    T, S, mb, width, ld = 21, 96, 24, 37, 40
    data = torch.randn(T, S, ld, device=dev, generator=g) * 3 + 1
    idx = torch.randperm(S, device=dev, generator=g)[:mb]
    mean = torch.randn(width, device=dev, generator=g)
    inv_std = 1.0 / (torch.rand(width, device=dev, generator=g) + 0.5)
    ks = [(torch.randn(a, b, device=dev, generator=g) / np.sqrt(a)).requires_grad_() for a, b in ((37, 64), (64, 48), (48, 5))]
    bs = [torch.zeros(b, device=dev, requires_grad=True) for b in (64, 48, 5)]
    out = {}
    for fused in (True, False):
        spec = ppo.GatherInput(data, idx, T - 1, width, mean, inv_std)
        x = spec if fused else spec.materialise()
        for p in ks + bs:
            p.grad = None
        y = ppo.mlp(x, ks, bs, None, True)
        assert tuple(y.shape) == (T - 1, mb, 5)
        (y * torch.linspace(-1, 1, y.numel(), device=dev).reshape(y.shape)).sum().backward()
        out[fused] = (y.detach().clone(), [p.grad.clone() for p in ks + bs])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(out[True][0], out[False][0]) < 1e-5
    assert max(rel(a, b) for a, b in zip(out[True][1], out[False][1])) < 1e-5
    # and against float64 end to end
    xr = ((data[:T - 1].index_select(1, idx)[..., :width].double() - mean.double()) * inv_std.double())
    a = xr
    for l in range(3):
        a = a @ ks[l].detach().double() + bs[l].detach().double()
        if l < 2:
            a = torch.nn.functional.silu(a)
    assert rel(out[True][0].double(), a) < 2e-5
