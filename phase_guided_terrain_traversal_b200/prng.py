"""Host-side jax.random key plumbing (numpy, uint32): `PRNGKey`, `split`, `fold_in`.

The reference hands `jax.random.PRNGKey`s to `env.reset` and to `domain_randomize`
(`training/train.py:231-263` through brax `ppo.train`, `go2/joystick_pgtt.py:50`,
`go2/randomize.py:23`). The per-env streams themselves are consumed on the device
(csrc/pgtt_env.cuh:threefry); this module only makes and splits the *top-level* keys on the host so
a training script can build the `[N, 2]` key arrays the ABI takes. Same threefry2x32 block function
and the same two stream layouts as JAX (`jax_threefry_partitionable` on = JAX >= 0.5 default).
"""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    x = x.astype(np.uint64)
    return (((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & _M32).astype(np.uint32)


def threefry2x32(key, x0, x1):
    """Vectorised Threefry-2x32 (20 rounds). key: uint32[2]; x0, x1: uint32 arrays of equal shape."""
    key = np.asarray(key, dtype=np.uint32)
    x0 = np.array(x0, dtype=np.uint32, copy=True)
    x1 = np.array(x1, dtype=np.uint32, copy=True)
    ks = (key[0], key[1], np.uint32(key[0] ^ key[1] ^ np.uint32(0x1BD11BDA)))
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r) ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed) for the default (threefry2x32, 32-bit seed handling) implementation."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2, partitionable: bool = True) -> np.ndarray:
    """jax.random.split(key, num) -> uint32[num, 2]."""
    key = np.asarray(key, dtype=np.uint32).reshape(2)
    if partitionable:
        a, b = threefry2x32(key, np.zeros(num, dtype=np.uint32), np.arange(num, dtype=np.uint32))
        return np.stack([a, b], 1)
    cnt = np.arange(2 * num, dtype=np.uint32)
    a, b = threefry2x32(key, cnt[:num], cnt[num:])
    return np.concatenate([a, b]).reshape(num, 2)


def fold_in(key, data: int, partitionable: bool = True) -> np.ndarray:
    """jax.random.fold_in(key, data)."""
    key = np.asarray(key, dtype=np.uint32).reshape(2)
    d = PRNGKey(data)
    if partitionable:
        a, b = threefry2x32(key, d[:1], d[1:])
    else:
        a, b = threefry2x32(key, d[:1], d[1:])
    return np.array([a[0], b[0]], dtype=np.uint32)


def env_keys(seed: int, num_envs: int, offset: int = 0) -> np.ndarray:
    """Benchmark / test convention of SURVEY.md 8d: key[i] = (0, seed * 2**20 + offset + i)."""
    lo = (np.arange(num_envs, dtype=np.uint64) + np.uint64(offset) + (np.uint64(seed) << np.uint64(20))) & np.uint64(0xFFFFFFFF)
    return np.stack([np.zeros(num_envs, dtype=np.uint32), lo.astype(np.uint32)], 1)


def as_keys(rng, num_envs: int | None = None, partitionable: bool = True) -> np.ndarray:
    """Normalise what `reset` / `domain_randomize` accept: an int seed, one key (split into N), or [N,2] keys."""
    if isinstance(rng, (int, np.integer)):
        if num_envs is None:
            raise ValueError("an integer seed needs num_envs")
        return split(PRNGKey(int(rng)), num_envs, partitionable)
    a = rng.detach().cpu().numpy() if hasattr(rng, "detach") else np.asarray(rng)
    a = a.astype(np.uint32, copy=False)
    if a.shape == (2,):
        return a.reshape(1, 2) if num_envs in (None, 1) else split(a, num_envs, partitionable)
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError(f"rng must be an int seed, a key uint32[2] or keys uint32[N,2]; got shape {a.shape}")
    if num_envs is not None and a.shape[0] != num_envs:
        raise ValueError(f"got {a.shape[0]} keys for {num_envs} envs")
    return np.ascontiguousarray(a)


# --------------------------------------------------------------------------------------------------
# draws (host restatement of jax.random for tests / fixtures; the env draws on the device)
# --------------------------------------------------------------------------------------------------
def random_bits(key, n: int, partitionable: bool = True) -> np.ndarray:
    """jax.random.bits(key, (n,), uint32)."""
    key = np.asarray(key, dtype=np.uint32).reshape(2)
    if partitionable:
        a, b = threefry2x32(key, np.zeros(n, dtype=np.uint32), np.arange(n, dtype=np.uint32))
        return a ^ b
    half = (n + 1) // 2
    cnt = np.arange(2 * half, dtype=np.uint32)
    if n & 1:
        cnt[-1] = 0  # odd sizes are zero-padded
    a, b = threefry2x32(key, cnt[:half], cnt[half:])
    return np.concatenate([a, b])[:n]


def uniform(key, shape=(), minval=0.0, maxval=1.0, partitionable: bool = True) -> np.ndarray:
    """jax.random.uniform(key, shape, float32, minval, maxval): mantissa trick, affine map, clamp at minval."""
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    bits = random_bits(key, n, partitionable)
    u = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.broadcast_to(np.asarray(minval, dtype=np.float32), shape).reshape(-1) if shape else np.asarray(minval, dtype=np.float32).reshape(-1)
    hi = np.broadcast_to(np.asarray(maxval, dtype=np.float32), shape).reshape(-1) if shape else np.asarray(maxval, dtype=np.float32).reshape(-1)
    v = np.maximum(lo, (u * (hi - lo) + lo).astype(np.float32))
    return v.reshape(shape).astype(np.float32)


def exponential(key, partitionable: bool = True) -> np.float32:
    u = uniform(key, (), partitionable=partitionable)
    return np.float32(-np.log1p(-u))


def bernoulli(key, p, shape=(), partitionable: bool = True) -> np.ndarray:
    return uniform(key, shape, partitionable=partitionable) < np.asarray(p, dtype=np.float32)


def randint(key, minval: int, maxval: int, partitionable: bool = True) -> int:
    """jax.random.randint(key, (), minval, maxval) for int32 (two 32-bit draws + modular combination)."""
    k1, k2 = split(key, 2, partitionable)
    hb, lb = int(random_bits(k1, 1, partitionable)[0]), int(random_bits(k2, 1, partitionable)[0])
    span = max(maxval - minval, 1)
    mult = ((65536 % span) * (65536 % span)) % span
    return minval + ((hb % span) * mult + (lb % span)) % span
