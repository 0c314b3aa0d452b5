"""CPU-only checks of the boundary and of the N>1 host logic:
* libbspy_cuda.so builds/loads and exports exactly the symbols include/bspy_cuda.h declares,
  the ctypes struct matches the C struct, argument errors come back as status codes (no GPU needed:
  argument validation happens before any CUDA call);
* shard_range / shard_points / gather_last_dim with a world_size-2 gloo group."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "bspy_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bspy_cuda_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bspy_b200 import _cuda
    lib = _cuda.library()
    declared = _header_symbols()
    assert declared == sorted(_cuda.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _cuda.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\b(bspy_cuda_\w+)\b", out)))
    assert exported == declared
    assert lib.bspy_cuda_abi_version() == 2


def test_library_is_sm100a_with_dmma_and_lineinfo():
    from bspy_b200 import _cuda
    out = subprocess.run(["cuobjdump", "-lelf", _cuda.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4bspy17grid2_dmma_kernelILi3ELi4ELb0EEEvNS_11Grid2ParamsE", _cuda.LIB_PATH],
                          capture_output=True, text=True).stdout
    assert "DMMA" in sass, "the grid kernel must run on the FP64 tensor pipe"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4bspy21eval_curve_tab_kernelILi4ELi3ELb0ELi2ELb1EEEvNS_11CurveParamsENS_11TableLayoutE",
                           _cuda.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass, "the cached curve tables must arrive by bulk asynchronous copy (TMA) on an mbarrier"


def test_struct_layout_matches_header():
    from bspy_b200 import _cuda
    # int32 x2, int32[8] x2, ptr[8], ptr, int32 x2, ptr, int64  ->  8 + 64 + 64 + 8 + 8 + 8 + 8 = 168 bytes on LP64
    assert C.sizeof(_cuda.CSpline) == 168
    assert _cuda.CSpline.knots.offset == 72 and _cuda.CSpline.coefs.offset == 136 and _cuda.CSpline.normalSign.offset == 144
    assert _cuda.CSpline.curveTable.offset == 152 and _cuda.CSpline.curveTableBytes.offset == 160
    assert _cuda.library().bspy_cuda_abi_version() == 2


def test_argument_errors_are_status_codes():
    from bspy_b200 import _cuda
    lib = _cuda.library()
    assert lib.bspy_cuda_spans(None, 8, 4, None, 10, None, None) == _cuda.E_ARG
    assert b"bad argument" in lib.bspy_cuda_last_error_string()
    s = _cuda.CSpline()
    s.nInd, s.nDep = 9, 1
    s.coefs = 8
    assert lib.bspy_cuda_eval_points(C.byref(s), None, 1, 1, 0, None, 0, 0, None, None, None, None, None, None, None) == _cuda.E_UNSUPPORTED
    s.nInd, s.nDep = 1, 3
    s.order[0], s.nCoef[0], s.knots[0] = 4, 8, 8
    dummy = C.c_void_p(8)
    assert lib.bspy_cuda_eval_points(C.byref(s), dummy, 1, 1, 0, None, 0, 0, None, None, None, dummy, None, None, None) == _cuda.E_NORMAL_DIMS
    assert b"one different" in lib.bspy_cuda_last_error_string()
    assert lib.bspy_cuda_eval_many(9, 16, 3, 1, dummy, 25, dummy, 48, dummy, 4, dummy, None, None, None) == _cuda.E_UNSUPPORTED
    with pytest.raises(ValueError):
        _cuda._check(_cuda.E_NORMAL_DIMS, "x")
    with pytest.raises(NotImplementedError):
        _cuda._check(_cuda.E_UNSUPPORTED, "x")
    with pytest.raises(_cuda.CudaPathError):
        _cuda._check(700, "x")


def test_shard_range_partitions():
    from bspy_b200.sharding import shard_range, shard_points
    for n in (0, 1, 7, 32, 1000003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    x = np.arange(22).reshape(11, 2)
    assert np.array_equal(np.concatenate([shard_points(x, r, 3) for r in range(3)]), x)
    with pytest.raises(ValueError):
        shard_range(5, 2, 2)


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from bspy_b200.sharding import init_from_env, shard_points, shard_range, gather_last_dim, gather_records
import bspy_b200 as bspy, fake_cuda
from bspy_b200 import _cuda
for name in ("device", "new_flag", "launch_count", "eval_points", "eval_points_host", "eval_points_aos", "record_layout", "eval_grid", "spans", "basis"):
    setattr(_cuda, name, getattr(fake_cuda, name))
rank, world, local = init_from_env("gloo")
assert world == 2 and dist.get_backend() == "gloo"
rng = np.random.default_rng(3)
kn = np.array([0, 0, 0, 0, .2, .5, .7, 1, 1, 1, 1.])
s = bspy.Spline(1, 3, (4,), (7,), [kn], rng.standard_normal((3, 7)))
u = rng.uniform(0, 1, (1001, 1))                       # same on both ranks (same seed)
mine = shard_points(u, rank, world)
lo, hi = shard_range(1001, rank, world)
assert mine.shape[0] == hi - lo
local_vals = torch.from_numpy(s.evaluate_points(mine).values)          # (3, n_local), no collective on the data path
full = gather_last_dim(local_vals, 1001)                               # optional final gather
ref = s.evaluate_points(u).values
assert full.shape == (3, 1001) and np.array_equal(full.numpy(), ref), "gathered shards differ from the unsharded result"
b = torch.tensor([float(local_vals.shape[1])]); dist.all_reduce(b); assert int(b) == 1001
# array-of-structs records: ragged shards (1001 points) and even shards (1000 points) through gather_records
for total in (1001, 1000):
    pts = u[:total]
    rec = torch.from_numpy(s.evaluate_points(shard_points(pts, rank, world), jacobian=True, out_layout="aos").records)
    whole = gather_records(rec, total)
    want = s.evaluate_points(pts, jacobian=True, out_layout="aos")
    assert whole.shape == (total, 8) and np.array_equal(whole.numpy(), want.records), "gathered records differ"
    assert np.array_equal(whole.numpy()[:, :3].T, want.values) and np.array_equal(whole.numpy()[:, 3:6].T, want.jacobian[:, 0])
# a rank with an empty shard (world > items) still takes part
tiny = shard_points(u[:1], rank, world)
assert tiny.shape[0] == (1 if rank == 0 else 0)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(script), ROOT]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
