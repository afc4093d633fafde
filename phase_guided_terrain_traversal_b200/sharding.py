"""Env-index sharding across ranks and the one collective of the path.

Envs are independent, so `env.step` needs no communication: rank r owns the contiguous index range
`shard(n_total, r, world)` with per-env keys `prng.env_keys(seed, n_local, offset=start)` - the union over
ranks is exactly the key set of a single-process run over `n_total` envs, so a sharded run steps the very
same environments. The only exchange is the learner-side moment reduction (advantage normalisation of the
north star, brax's running observation statistics): `(count, sum, sum of squares)` packed into one small
all-reduce - NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import prng


def shard(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the envs rank `rank` owns; sizes differ by at most one, earlier ranks get the remainder."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_keys(seed: int, n_total: int, rank: int, world: int) -> np.ndarray:
    """Per-env jax-style keys of this rank's shard (== rows [start, stop) of `env_keys(seed, n_total)`)."""
    start, stop = shard(n_total, rank, world)
    return prng.env_keys(seed, stop - start, offset=start)


def allreduce_moments(x, group=None):
    """Global (count, mean, variance) of a sharded tensor `x[n_local, ...]` over its first axis.

    One all-reduce of `1 + 2 * prod(x.shape[1:])` float64 values; without an initialised process group
    (single process) it reduces locally. Works on CUDA tensors (NCCL) and CPU tensors (gloo).
    """
    import torch
    import torch.distributed as dist
    feat = x.shape[1:]
    if x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.shape[0] * x.shape[1] >= 1 << 16:
        # large fp32 matrices (the observation batches of the normaliser, possibly a column slice of a padded store): one pass of the hand-written
        # float64 column-moment kernel instead of a float64 copy and two reductions
        import ctypes as C
        from . import _native as nat
        lib = nat.load_library()
        rows, cols = x.shape
        sums = torch.empty(2 * cols, dtype=torch.float64, device=x.device)
        scratch = torch.empty(int(lib.pgtt_col_moments_scratch_doubles(cols)), dtype=torch.float64, device=x.device)
        rc = lib.pgtt_col_moments(x.data_ptr(), rows, cols, x.stride(0), sums.data_ptr(), scratch.data_ptr(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        if rc:
            raise nat.PgttError(rc, lib.pgtt_policy_last_error().decode())
        packed = torch.cat([torch.full((1,), float(rows), dtype=torch.float64, device=x.device), sums])
    else:
        xf = x.reshape(x.shape[0], -1).to(torch.float64)
        # (torch.full, not torch.tensor: no host-to-device copy, so the reduction can sit inside a CUDA-graph capture)
        packed = torch.cat([torch.full((1,), float(x.shape[0]), dtype=torch.float64, device=x.device), xf.sum(0), (xf * xf).sum(0)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    d = (packed.numel() - 1) // 2
    count = packed[0]
    mean = packed[1:1 + d] / count
    var = (packed[1 + d:] / count - mean * mean).clamp_min(0.0)
    return count, mean.reshape(feat), var.reshape(feat)


def normalize_advantages(adv, eps: float = 1e-8, group=None):
    """(adv - global mean) / (global std + eps) over every env of every rank (north-star GAE all-reduce)."""
    _, mean, var = allreduce_moments(adv.reshape(-1, 1), group)
    return ((adv.to(mean.dtype) - mean) / (var.sqrt() + eps)).to(adv.dtype)
