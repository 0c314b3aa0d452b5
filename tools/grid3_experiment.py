"""GPU experiment: volume grid (nInd == 3) on the tensor pipe: value + jacobian on a 256^3 grid."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, bspy_b200 as bspy
wl = bench.CONFIGS["cfg4"]()
s = wl.make_spline(np.random.default_rng(1004), bspy); s.freeze()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ax = [torch.linspace(0, 1, n, dtype=torch.float64, device="cuda") for _ in range(3)]
for kw, bpp in ((dict(jacobian=True), 96), (dict(), 24)):
    for _ in range(2):
        r = s.evaluate_grid(*ax, check_domain=False, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        r = s.evaluate_grid(*ax, check_domain=False, **kw)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"volume grid {n}^3 {kw}: {ms:.3f} ms  {n**3/ms/1e6:.2f} Gpts/s  {n**3*bpp/ms/1e6:.0f} GB/s")
