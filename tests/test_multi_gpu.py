"""Multi-GPU sharding on real devices (NCCL): skipped unless the box has at least two GPUs.

One process per GPU (torchrun): every rank evaluates its contiguous shard of one point batch through the public API,
there is no collective on the data path, and the optional final gather (`gather_last_dim` for struct-of-arrays outputs,
`gather_records` for array-of-structs records) re-assembles the unsharded result bit for bit on every rank."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from bspy_b200.sharding import init_from_env, shard_points, shard_range, gather_last_dim, gather_records
import bspy_b200 as bspy
from bspy_b200 import _cuda
rank, world, local = init_from_env("nccl")
dev = torch.device("cuda", local)
rng = np.random.default_rng(3)
def K(o, n):
    w = rng.uniform(0.25, 1.75, n - o + 1); inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
    return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))
s = bspy.Spline(3, 3, (4, 4, 4), (18, 18, 18), [K(4, 18) for _ in range(3)], rng.standard_normal((3, 18, 18, 18)))
for N in (300_001, 300_000):                      # ragged and even shards; big enough for the cell-sorted path
    g = torch.Generator(device=dev).manual_seed(11)   # the same batch on every rank
    pts = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=g)
    mine = shard_points(pts, rank, world)
    lo, hi = shard_range(N, rank, world)
    assert mine.shape[0] == hi - lo
    # default dispatch: whole batch and shards may take different kernels (cell polynomials need a few points per cell):
    # the gathered shards agree with the unsharded result to the strict parity bar
    whole = s.evaluate_points(pts, jacobian=True)
    part = s.evaluate_points(mine, jacobian=True)
    jac = gather_last_dim(part.jacobian.contiguous(), N)
    assert bool(torch.isclose(jac, whole.jacobian, rtol=1e-12, atol=1e-13).all()), "gathered shards outside the bar"
    # the recurrence kernels do the same arithmetic per point whatever the batch: bit for bit
    _cuda.set_option("CELL_POLY", 0)
    whole = s.evaluate_points(pts, jacobian=True)
    part = s.evaluate_points(mine, jacobian=True)
    vals = gather_last_dim(part.values.contiguous(), N)
    jac = gather_last_dim(part.jacobian.contiguous(), N)
    assert torch.equal(vals, whole.values) and torch.equal(jac, whole.jacobian), "gathered struct-of-arrays shards differ"
    rec = s.evaluate_points(mine, jacobian=True, out_layout="aos").records
    full = gather_records(rec, N)
    want = s.evaluate_points(pts, jacobian=True, out_layout="aos").records
    assert full.shape == want.shape and torch.equal(full, want), "gathered records differ"
    _cuda.set_option("CELL_POLY", None)
    count = torch.tensor([float(mine.shape[0])], device=dev); dist.all_reduce(count); assert int(count.item()) == N
batch = bspy.SplineBatch(1, 3, (4,), (32,), [torch.from_numpy(np.stack([K(4, 32) for _ in range(7)])).to(dev)],
                         torch.from_numpy(rng.standard_normal((7, 3, 32))).to(dev))
u = torch.rand((7, 64), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
lo, hi = shard_range(7, rank, world)
mine = batch.shard(rank, world).evaluate(u[lo:hi]).values          # splines sharded by index range
assert torch.equal(mine, batch.evaluate(u).values[lo:hi])
empty = batch.shard(world + 5 - 1, world + 5) if False else None    # (empty shards are covered by the CPU tests)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
def test_nccl_sharding_and_gathers(tmp_path):
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29741", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
