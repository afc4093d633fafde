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
    validate_terrain(a)
    return a


def validate_terrain(a: np.ndarray) -> None:
    """What the kernels support (true of every table terrain/generator.py writes): shape [T,100,10], finite values,
    yaw-only box rotations (quat x = y = 0, non-zero quaternion) and positive half-sizes. The ray grid and the sphere/box
    collision use only the yaw of a box, so anything else would give silently wrong hits - fail loudly instead."""
    a = np.asarray(a)
    if a.ndim != 3 or a.shape[1:] != (100, 10):
        raise ValueError(f"terrain table must have shape [T,100,10], got {a.shape}")
    if not np.isfinite(a).all():
        raise ValueError("terrain table holds non-finite values")
    if np.abs(a[..., 4:6]).max() != 0:
        raise ValueError("only yaw-rotated boxes (quat x = y = 0) are supported")
    if (np.abs(a[..., 3]) + np.abs(a[..., 6])).min() <= 1e-9:
        raise ValueError("terrain table holds a zero quaternion")
    if a[..., 7:10].min() <= 0:
        raise ValueError("box half-sizes must be positive")
