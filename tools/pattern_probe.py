"""GPU experiment: HBM rate of the grid kernel's store pattern alone (no arithmetic), for several CTA
tile shapes, at the size of the teapot workload (32 patches stacked along rows, 12 output planes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bspy_b200 import _cuda
n = 2048
for planes, nU in ((12, n * 32), (3, n * 32), (1, n * 32 * 12)):
    dst = torch.empty(planes * nU * n, dtype=torch.float64, device="cuda")
    for rows, cols in ((64, 256), (64, 2048), (32, 512), (16, 1024), (8, 2048), (8, 256), (8, 512), (16, 256), (64, 64), (64, 128), (32, 128)):
        for _ in range(2):
            _cuda.probe_tiles(dst, planes, nU, n, rows, cols)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            nb = _cuda.probe_tiles(dst, planes, nU, n, rows, cols)
        b.record(); torch.cuda.synchronize()
        print(f"planes={planes:4d} nU={nU} tile {rows:3d}x{cols:4d}  {nb*3/(a.elapsed_time(b)*1e-3)/1e9:8.1f} GB/s", flush=True)
    del dst
src = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); dst = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
for kind in (0, 1):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _cuda.probe_hbm(kind, src, dst); a.record(); nb = _cuda.probe_hbm(kind, src, dst); b.record(); torch.cuda.synchronize()
    print("linear", "copy" if kind == 0 else "fill", f"{nb/(a.elapsed_time(b)*1e-3)/1e9:8.1f} GB/s")
