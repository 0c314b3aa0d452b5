"""GPU experiment: upper bound for cell-binned scattered evaluation -- time the scattered kernel on points
that are already sorted by knot-span cell (warp-coherent coefficient windows) vs random order."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import bspy_b200 as bspy

def run(cfg, N):
    wl = bench.CONFIGS[cfg]()
    rng = np.random.default_rng(wl.seed)
    s = wl.make_spline(rng, bspy)
    s.freeze()
    pts = torch.rand((N, s.nInd), dtype=torch.float64, device="cuda")
    spans = s.evaluate_points(pts, values=False, spans=True, check_domain=False).spans.long()   # (nInd, N)
    key = torch.zeros(N, dtype=torch.long, device="cuda")
    for i in range(s.nInd):
        key = key * (s.nCoef[i] + 1) + spans[i]
    order = torch.argsort(key)
    sorted_pts = pts[order].contiguous()
    block = pts.reshape(-1, 32, s.nInd)   # control: random
    from bspy_b200 import _cuda
    from bspy_b200._spline_evaluation import device_spline
    ds = device_spline(s)
    for name, p, binned in (("random direct", pts, False), ("sorted direct", sorted_pts, False), ("random binned", pts, True), ("sorted binned", sorted_pts, True)):
        run1 = lambda: _cuda.eval_points(ds, p, s.nInd, 1, N, values=True, jacobian=True, binned=binned)
        for _ in range(2):
            run1()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run1()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{cfg} N={N} {name:16s} {ms:8.3f} ms  {N/ms/1e6:7.2f} Gpts/s  fp64 {wl.flops_per_point*N/ms/1e9:6.2f} TF/s", flush=True)

run("cfg4", 20_000_000)
run("cfg5", 10_000_000)
