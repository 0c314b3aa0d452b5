"""FP64 vector (DFMA) and tensor (DMMA.8x8x4) throughput of the device: the roofline denominators for configs 4/5."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bspy_b200 import _cuda

dev = torch.device("cuda:0")
for kind, name in ((0, "dfma"), (1, "dmma"), (2, "dmma+dfma interleaved per warp"), (3, "dmma / dfma on alternate warps"),
                   (4, "dfma outer-product tile (3 varying operands), 16 warps per scheduler"), (5, "the same at 4 warps per scheduler")):
    _cuda.probe_fp64(kind, 2000, dev)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        flops = _cuda.probe_fp64(kind, 20000, dev)
        b.record()
        torch.cuda.synchronize()
        best = max(best, flops / (a.elapsed_time(b) * 1e-3))
    print(f"{name}: {best / 1e12:.2f} TFLOP/s")
