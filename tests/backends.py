"""Back-ends the kernel-parity tests run on: the host-emulated kernel source (CPU, always) and the real CUDA
library (marked gpu), each with both kernel generations (warp-per-env, quad-per-env; PGTT_KERNEL is read by
pgtt_create)."""
import os

import pytest


def make_env(kind, model, cfg, n, **kw):
    base, _, gen = kind.partition("-")
    if gen != "auto":   # "cuda-auto": the generation pgtt_create picks from the env count
        os.environ["PGTT_KERNEL"] = "quad" if gen.startswith("quad") else "warp"
    os.environ["PGTT_QUAD_FULLSCAN"] = "1" if gen in ("quadfull", "warpfull") else "0"   # full box scans instead of the near lists
    try:
        if base == "emu":
            import emu_backend
            return emu_backend.make_env(model, cfg, n, **kw)
        from phase_guided_terrain_traversal_b200.abi_env import AbiEnv
        return AbiEnv(model, cfg, n, backend="torch", **kw)
    finally:
        os.environ.pop("PGTT_KERNEL", None)
        os.environ.pop("PGTT_QUAD_FULLSCAN", None)


BACKENDS = [pytest.param("emu", id="emu"), pytest.param("emu-warpfull", id="emu-warpfull"), pytest.param("emu-quad", id="emu-quad"), pytest.param("emu-quadfull", id="emu-quadfull"),
            pytest.param("cuda", id="cuda", marks=pytest.mark.gpu), pytest.param("cuda-quad", id="cuda-quad", marks=pytest.mark.gpu)]
