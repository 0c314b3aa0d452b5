"""Throughput of the surface grid kernel with float64 and float32 outputs (32 teapot-like bicubic patches, 2048^2 grid,
value + jacobian + unit normal): device tensors in, preallocated device outputs, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bspy_b200 as bspy

rng = np.random.default_rng(0)
kn = np.array([0, 0, 0, 0, 1, 1, 1, 1.0])
batch = bspy.SplineBatch(2, 3, (4, 4), (4, 4), [kn, kn], rng.standard_normal((32, 3, 4, 4)))
n = 2048
u = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
for name, dt, tdt in (("float64", None, torch.float64), ("float32", np.float32, torch.float32)):
    out = {"values": torch.empty((32, 3, n, n), dtype=tdt, device="cuda"), "jacobian": torch.empty((32, 3, 2, n, n), dtype=tdt, device="cuda"),
           "normal": torch.empty((32, 3, n, n), dtype=tdt, device="cuda")}
    for _ in range(3):
        batch.evaluate_grid(u, u, jacobian=True, normal=True, check_domain=False, out=out, dtype=dt)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        batch.evaluate_grid(u, u, jacobian=True, normal=True, check_domain=False, out=out, dtype=dt)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    pts = 32 * n * n
    bytes_pt = 12 * (8 if dt is None else 4)
    print(f"{name}: {ms:.3f} ms, {pts / ms / 1e6:.1f} Gpts/s, {pts * bytes_pt / ms / 1e6:.0f} GB/s written")

