"""Builds and loads the host-emulated kernel library (TEST INFRASTRUCTURE ONLY).

`libpgtt_emu.so` is csrc/pgtt_api.cu compiled by g++ with -DPGTT_HOST_EMU: identical kernel
source, each warp run as 32 fibers (tests/simt_emu/simt_emu.h). It lets the `-m "not gpu"` suite
check the warp-cooperative kernel logic against the CPU oracle. The package never loads it.
"""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tests" / "simt_emu"
CSRC = ROOT / "phase_guided_terrain_traversal_b200" / "csrc"
EMU_LIB = EMU_DIR / "libpgtt_emu.so"
_lib = None


def build(force=False):
    srcs = list(CSRC.glob("*.cu*")) + list(CSRC.glob("*.h")) + [EMU_DIR / "simt_emu.h", ROOT / "include" / "pgtt_b200.h"]
    newest = max(s.stat().st_mtime for s in srcs)
    if not force and EMU_LIB.exists() and EMU_LIB.stat().st_mtime >= newest:
        return EMU_LIB
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-DPGTT_HOST_EMU", "-DEMU_IMPL", "-ffp-contract=off",
           "-Wno-unused-function", "-I", str(EMU_DIR), "-I", str(CSRC), "-x", "c++", str(CSRC / "pgtt_api.cu"), "-o", str(EMU_LIB)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emu build failed:\n" + res.stderr[-4000:])
    return EMU_LIB


def load():
    global _lib
    if _lib is None:
        import ctypes
        from phase_guided_terrain_traversal_b200 import _native as nat
        _lib = nat.declare(ctypes.CDLL(str(build())))
    return _lib


def make_env(model, cfg, n, **kw):
    from phase_guided_terrain_traversal_b200.abi_env import AbiEnv
    return AbiEnv(model, cfg, n, backend="numpy", lib=load(), **kw)
