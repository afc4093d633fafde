"""Device time of pgtt_mlp_forward / pgtt_mlp_backward (csrc/pgtt_mlp.cu) at the reference network sizes, back to back."""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
from phase_guided_terrain_traversal_b200 import _native as nat
lib = nat.load_library()
dev = torch.device("cuda", 0)
for dims, rows, ldx in (((171, 512, 256, 128, 24), 5120, 172), ((215, 512, 256, 128, 1), 5376, 216)):
    L = len(dims) - 1
    x = torch.randn(rows, ldx, device=dev)
    ws = [torch.randn(dims[l], dims[l + 1], device=dev) / np.sqrt(dims[l]) for l in range(L)]
    bs = [torch.randn(dims[l + 1], device=dev) for l in range(L)]
    dws = [torch.empty_like(w) for w in ws]; dbs = [torch.empty_like(b) for b in bs]
    dy = torch.randn(rows, dims[-1], device=dev); y = torch.empty(rows, dims[-1], device=dev)
    h = C.c_void_p()
    assert lib.pgtt_mlp_create(L, (C.c_int * (L + 1))(*dims), rows, 0, C.byref(h)) == 0
    vp = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cs = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fwd = lambda: lib.pgtt_mlp_forward(h, x.data_ptr(), ldx, vp(ws), vp(bs), y.data_ptr(), cs())
    bwd = lambda: lib.pgtt_mlp_backward(h, dy.data_ptr(), vp(dws), vp(dbs), cs())
    for name, fn in (("forward", fwd), ("backward", bwd), ("forward+backward", lambda: (fwd(), bwd()))):
        g = torch.cuda.CUDAGraph()
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(10): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"{dims} rows {rows}: {name:18s} {e0.elapsed_time(e1) * 10:.1f} us")
    import os
    if os.environ.get("PGTT_MLP_TRACE") == "1":
        fwd(); bwd(); torch.cuda.synchronize()
        buf = (C.c_ulonglong * (8 * 32))()
        n = lib.pgtt_mlp_debug_trace(h, buf, 32)
        t0 = buf[0]
        print("GEMM launches of one forward + backward (CTA 0; us since the first kernel entered): entered, first chunk, MMAs issued, accumulators done, epilogue warp done, CTA done")
        for i in range(n):
            print(f"  launch {i:2d}: " + "  ".join(f"{(buf[8 * i + k] - t0) / 1e3:8.2f}" for k in range(6)))
    lib.pgtt_mlp_destroy(h)
