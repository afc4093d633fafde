#!/usr/bin/env python3
"""Repack the reference's terrain box tables (terrains/level*.npy, float32 [T,100,10] =
pos xyz | quat wxyz | half-size xyz, written by terrain/generator.py:288-391) into one compressed
fixture per level under phase_guided_terrain_traversal_b200/assets/terrains/.

Run HERE (the container that mounts /root/reference); the GPU box only sees the fixtures.
Values are stored bit-exact (float32); `terrain.load_terrain` accepts both this fixture format and
a reference-style .npy path, so a user's own `--terrain_file` keeps working.
"""
import sys
from pathlib import Path

import numpy as np

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference") / "terrains"
out = Path(__file__).resolve().parent.parent / "phase_guided_terrain_traversal_b200" / "assets" / "terrains"
out.mkdir(parents=True, exist_ok=True)
for f in sorted(ref.glob("level*.npy")):
    a = np.load(f)
    assert a.dtype == np.float32 and a.shape[1:] == (100, 10), (f, a.shape, a.dtype)
    np.savez_compressed(out / (f.stem + ".npz"), boxes=a)
    print(f.name, a.shape, (out / (f.stem + ".npz")).stat().st_size)
