"""Multi-GPU sharding of the evaluation path: one process per GPU (torch.distributed).

Every point (and every spline of a batch) is independent, so the path shards by splitting the
point or spline index range into contiguous, balanced slices; the spline itself (2 KB - 3 MB) is
replicated.  There is no collective on the data path.  ``gather_last_dim`` is the optional final
gather of struct-of-arrays outputs (NCCL all-gather over NVLink on GPUs, gloo in CPU tests); big
outputs (config 5: 240 GB) simply stay sharded.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

__all__ = ["shard_range", "rank_world", "init_from_env", "shard_points", "gather_last_dim"]


def shard_range(n: int, rank: int, world: int):
    """Half-open slice ``[lo, hi)`` of ``range(n)`` owned by ``rank``: sizes differ by at most one,
    the first ``n % world`` ranks take the extra item, slices are contiguous and ordered by rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Join the job described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT (torchrun).
    Returns (rank, world, local_rank).  A single process (no env) is rank 0 of 1 without a group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local if local < torch.cuda.device_count() else 0)
    return rank, world, local


def shard_points(uvw, rank=None, world=None, dim=0):
    """This rank's contiguous slice of a point array / tensor along ``dim`` (a view)."""
    if rank is None or world is None:
        rank, world = rank_world()
    lo, hi = shard_range(uvw.shape[dim], rank, world)
    index = [slice(None)] * uvw.ndim
    index[dim] = slice(lo, hi)
    return uvw[tuple(index)]


def gather_last_dim(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather struct-of-arrays shards ``(..., n_local)`` (split by ``shard_range``) back into
    ``(..., n_total)`` on every rank.  Uneven shards are padded to the largest one for the
    collective and trimmed afterwards."""
    rank, world = rank_world()
    if world == 1:
        return local
    most = (n_total + world - 1) // world
    lead = tuple(local.shape[:-1])
    padded = local.new_zeros((*lead, most))
    padded[..., : local.shape[-1]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    out = local.new_empty((*lead, n_total))
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out[..., lo:hi] = parts[r][..., : hi - lo]
    return out
