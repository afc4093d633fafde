"""Terrain generator against fixtures produced by the reference's terrain/generator.py + getIndexes.py
(tests/golden/make_golden.py:terrain_vectors) and against the invariants of the shipped terrains/level*.npy files."""
import json
from pathlib import Path

import numpy as np
import pytest

from phase_guided_terrain_traversal_b200 import terrain, terrain_gen as tg

GOLD = Path(__file__).resolve().parent / "golden"


def test_adjacency_equals_reference_tables():
    want = json.loads((GOLD / "terrain_adjacency.json").read_text())
    got = tg.adjacency()
    assert sorted(int(t) for t in want) == list(range(14))
    for t, per in want.items():
        for d, tiles in per.items():
            di, dj = (int(x) for x in d.split(","))
            assert sorted(got[int(t)][(di, dj)]) == tiles, (t, d)


def test_tile_boxes_equal_reference_geometry():
    g = np.load(GOLD / "terrain_tiles.npz")
    for si, (w, h, n) in enumerate(g["settings"]):
        for tile in range(14):
            want = g[f"s{si}_tile{tile}"]
            got = np.array(tg.tile_boxes(tile, 0.7, -1.1, w, h, int(n))).reshape(-1, 10)
            assert got.shape == want.shape, (si, tile, got.shape, want.shape)
            if len(want):
                # quaternions: same rotation (q and -q both describe it)
                sign = np.sign((got[:, 3:7] * want[:, 3:7]).sum(1, keepdims=True))
                assert np.allclose(got[:, :3], want[:, :3], atol=1e-9) and np.allclose(got[:, 7:], want[:, 7:], atol=1e-9), (si, tile)
                assert np.allclose(got[:, 3:7] * sign, want[:, 3:7], atol=1e-9), (si, tile)


def test_collapsed_grids_respect_the_rules():
    rng = np.random.default_rng(0)
    table = tg.adjacency()
    for _ in range(50):
        wave = tg.collapse_grid(5, rng)
        assert (wave[0] == 1).all() and (wave[-1] == 1).all() and (wave[:, 0] == 1).all() and (wave[:, -1] == 1).all() and wave[2, 2] == 0
        # every adjacent pair is allowed by the table of at least one of the two tiles (whichever collapsed first restricted the other)
        for i in range(5):
            for j in range(5):
                for (di, dj) in tg.DIRS:
                    a, b = i + di, j + dj
                    if 0 <= a < 5 and 0 <= b < 5:
                        assert wave[a, b] in table[int(wave[i, j])][(di, dj)] or wave[i, j] in table[int(wave[a, b])][(-di, -dj)]


def test_generated_tables_have_the_invariants_of_the_shipped_levels():
    m = tg.create_random_matrix(30, 100, 5, 0.07, 0.07, seed=1)
    terrain.validate_terrain(m)
    assert m.shape == (30, 100, 10) and m.dtype == np.float32
    ship = terrain.load_terrain("level07")
    for tab in (m, ship):
        act = tab[..., 0] < 50
        assert np.allclose(tab[..., 2][act], tab[..., 9][act], atol=1e-6)                    # every box stands on z = 0
        q = tab[..., 3:7][act]
        assert np.allclose(q[:, 1:3], 0) and np.allclose(np.linalg.norm(q, axis=1), 1, atol=1e-6)      # yaw-only rotations
        yaw = 2 * np.arctan2(q[:, 3], q[:, 0])
        assert np.allclose(np.round(yaw / (np.pi / 2)) * (np.pi / 2), yaw, atol=1e-5)
        tops = 2 * tab[..., 9][act]
        assert np.allclose(np.round(tops / 0.07) * 0.07, tops, atol=1e-5) and tops.max() <= 4 * 0.07 + 1e-6
        assert np.abs(tab[..., :2][act]).max() < 4.6
        n_act = act.sum(1)
        assert n_act.min() >= 16 and n_act.max() <= 100
    # unused rows: parked far away with unit size and identity rotation, numbered like the reference does
    inact = ~(m[..., 0] < 50)
    k = np.arange(100, 100 + 30 * 100, dtype=np.float32).reshape(30, 100)
    assert np.array_equal(m[..., 0][inact], k[inact]) and np.array_equal(m[..., 7:][inact], np.ones((inact.sum(), 3), np.float32))
    # similar box budgets to the shipped level (31 .. 100 active boxes there)
    assert abs(np.median((m[..., 0] < 50).sum(1)) - np.median((ship[..., 0] < 50).sum(1))) < 30


@pytest.mark.gpu
def test_env_steps_on_a_generated_terrain(train_cfg):
    import functools
    import torch
    from phase_guided_terrain_traversal_b200 import prng, wrapper
    from phase_guided_terrain_traversal_b200.go2 import joystick_pgtt, randomize
    m = tg.create_random_matrix(16, seed=3, height_min=0.05, height_max=0.12)
    n = 256
    env = joystick_pgtt.Joystick(task="stairs", config=train_cfg)
    keys = prng.env_keys(6, n)
    wenv = wrapper.wrap_for_brax_training(env, episode_length=1000, randomization_fn=functools.partial(randomize.domain_randomize, rng=keys, terrain_matrix=m))
    state = wenv.reset(keys)
    for _ in range(30):
        state = wenv.step(state, torch.rand((n, 12), device="cuda") * 2 - 1)
    torch.cuda.synchronize()
    assert torch.isfinite(state.obs["state"]).all() and float(state.info["heightscan"][..., 2].max()) > 0.04
