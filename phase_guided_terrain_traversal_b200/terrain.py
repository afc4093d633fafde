"""Terrain box tables: `[n_terrains, 100, 10]` float32 = pos xyz | quat wxyz | half-size xyz, the
format `terrain/generator.py:288-391` writes and `training/train.py:165-170` loads with `jnp.load`.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

TERRAIN_DIR = Path(__file__).resolve().parent / "assets" / "terrains"


def load_terrain(name_or_path) -> np.ndarray:
    """Accepts a reference-style `.npy` path, a packed fixture `.npz` path, or a bare level name
    such as "level07" / "terrains/level07.npy" (resolved against the bundled fixtures)."""
    p = Path(str(name_or_path))
    if p.exists():
        a = np.load(p)
        a = a["boxes"] if hasattr(a, "files") else a
    else:
        q = TERRAIN_DIR / (p.stem + ".npz")
        if not q.exists():
            raise FileNotFoundError(f"terrain file {name_or_path!r} not found (bundled levels: {sorted(x.stem for x in TERRAIN_DIR.glob('*.npz'))})")
        a = np.load(q)["boxes"]
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 3 or a.shape[1:] != (100, 10):
        raise ValueError(f"terrain table must have shape [T,100,10], got {a.shape}")
    return a


def validate_terrain(a: np.ndarray) -> None:
    """The kernels support yaw-only box rotations standing on z = 0 (true of every shipped level)."""
    if np.abs(a[..., 4:6]).max() != 0:
        raise ValueError("only yaw-rotated boxes (quat x = y = 0) are supported")
