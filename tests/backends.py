"""Back-ends the kernel-parity tests run on: the host-emulated kernel source (CPU, always) and
the real CUDA library (marked gpu)."""
import pytest


def make_env(kind, model, cfg, n, **kw):
    if kind == "emu":
        import emu_backend
        return emu_backend.make_env(model, cfg, n, **kw)
    from phase_guided_terrain_traversal_b200.abi_env import AbiEnv
    return AbiEnv(model, cfg, n, backend="torch", **kw)


BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
