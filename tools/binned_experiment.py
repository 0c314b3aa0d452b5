"""GPU experiment: cost of the pieces of the binned scattered path (run under ncu launch list)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import bspy_b200 as bspy
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4 << 19
wl = bench.CONFIGS[cfg]()
s = wl.make_spline(np.random.default_rng(wl.seed), bspy)
s.freeze()
pts = torch.rand((N, s.nInd), dtype=torch.float64, device="cuda")
for _ in range(3):
    r = s.evaluate_points(pts, jacobian=True, check_domain=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    r = s.evaluate_points(pts, jacobian=True, check_domain=False)
b.record(); torch.cuda.synchronize()
print(cfg, N, f"{a.elapsed_time(b)/3:.3f} ms  {N*3/a.elapsed_time(b)/1e6:.2f} Gpts/s")
