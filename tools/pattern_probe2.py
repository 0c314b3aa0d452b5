"""GPU experiment: is the power-of-two row pitch (2048 doubles = 16 KB) camping on HBM channels?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bspy_b200 import _cuda
for nV in (2048, 2000, 2064, 2176, 1920):
    planes, nU = 12, 2048 * 32
    dst = torch.empty(planes * nU * nV, dtype=torch.float64, device="cuda")
    for rows, cols in ((64, 256), (8, 256), (8, 2048), (16, 256), (64, 64)):
        for _ in range(2):
            _cuda.probe_tiles(dst, planes, nU, nV, rows, cols)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            nb = _cuda.probe_tiles(dst, planes, nU, nV, rows, cols)
        b.record(); torch.cuda.synchronize()
        print(f"nV={nV} planes={planes} tile {rows:3d}x{cols:4d}  {nb*3/(a.elapsed_time(b)*1e-3)/1e9:8.1f} GB/s", flush=True)
    del dst
# interleaved warps: the 8 warps of a CTA write side by side (1 KB contiguous per row per step)
planes, nU, nV = 12, 2048 * 32, 2048
dst = torch.empty(planes * nU * nV, dtype=torch.float64, device="cuda")
for cols in (128, 256, 512, 1024, 2048):
    for _ in range(2):
        _cuda.probe_tiles(dst, planes, nU, nV, 8, cols | 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        nb = _cuda.probe_tiles(dst, planes, nU, nV, 8, cols | 1)
    b.record(); torch.cuda.synchronize()
    print(f"interleaved 8x{cols:4d} planes={planes}  {nb*3/(a.elapsed_time(b)*1e-3)/1e9:8.1f} GB/s", flush=True)
