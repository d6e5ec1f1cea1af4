"""Data-parallel plumbing: one process per GPU (torch.distributed, NCCL over
NVLink on B200; gloo for the CPU tests).

The reference is single-device: its only parallelism is jax.vmap over the
`n_env_train` environment keys (dgppo/trainer/trainer.py:132-134,
dgppo/algo/informarl.py:183-184).  Environments are independent, so the keys
shard across ranks with NO collective inside the rollout; the only collective
of the training step is the mean all-reduce of the PPO gradients (policy + Vl +
Vh, ~627 KiB fp32) before clipping / Adam (informarl.py:440-447,
dgppo.py:316-319), issued as ONE flat buffer because the message is
latency-bound.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; sizes differ by at most 1."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_keys(keys, rank: int | None = None, world_size: int | None = None):
    """This rank's slice of the per-environment keys (trainer.py:132-134)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(len(keys), rank, world_size)
    return keys[lo:hi]


def allreduce_mean_flat(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Mean all-reduce of a list of gradient tensors as one flat buffer.
    With equal shard sizes the mean of per-rank means equals the single-device
    mean, so the update matches the reference's."""
    r, w = world()
    if w == 1:
        return list(tensors)
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= w
    out, off = [], 0
    for t in tensors:
        out.append(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return out


def same_shuffle(n: int, seed: int) -> np.ndarray:
    """The host-side env shuffle of DGPPO.update (dgppo.py:155-156) must be
    identical on every rank: derive it from a shared seed."""
    idx = np.arange(n)
    np.random.default_rng(seed).shuffle(idx)
    return idx


def max_over_ranks(value: float, device=None) -> float:
    r, w = world()
    if w == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
