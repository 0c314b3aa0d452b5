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

__all__ = ["shard_range", "rank_world", "init_from_env", "shard_points", "gather_last_dim", "gather_records",
           "bind_to_gpu_numa_node"]


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


def gather_records(local: torch.Tensor, n_total: int, out: torch.Tensor = None, group=None) -> torch.Tensor:
    """All-gather array-of-structs shards ``(n_local, stride)`` (rows split by ``shard_range``) into ``(n_total, stride)``
    on every rank.  Records are point-major, so each rank's shard is one contiguous block of the result: equal shards
    go through a single ``all_gather_into_tensor`` straight into ``out`` (no staging copy); ragged shards are padded to
    the largest one and trimmed."""
    rank, world = rank_world()
    stride = local.shape[1]
    if out is None:
        out = local.new_empty((n_total, stride))
    if world == 1:
        out.copy_(local)
        return out
    local = local.contiguous()
    if n_total % world == 0 and out.is_contiguous():
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    most = (n_total + world - 1) // world
    padded = local.new_zeros((most, stride))
    padded[: local.shape[0]] = local
    parts = local.new_empty((world * most, stride))
    dist.all_gather_into_tensor(parts, padded, group=group)
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        out[lo:hi] = parts[r * most: r * most + hi - lo]
    return out


def bind_to_gpu_numa_node(device_index=None):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers (first touched by
    this process) and the copy-engine traffic stay on that node.  Linux sysfs only; returns a short description, or the
    reason nothing was changed.  Purely a host-side placement hint: results do not depend on it."""
    try:
        if not torch.cuda.is_available():
            return "no CUDA device"
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        props = torch.cuda.get_device_properties(idx)
        bus = f"{getattr(props, 'pci_domain_id', 0):04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        with open(f"{base}/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return f"{bus}: local cpus {text} not in this process's affinity mask"
        try:
            with open(f"{base}/numa_node") as f:
                node = f.read().strip()
        except OSError:
            node = "?"
        if use != allowed:
            os.sched_setaffinity(0, use)
        return f"gpu {idx} ({bus}) numa node {node}: {len(use)} of {len(allowed)} cpus"
    except Exception as exc:  # pragma: no cover - depends on the box
        return f"not bound: {type(exc).__name__}: {exc}"
