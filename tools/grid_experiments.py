"""GPU experiment: where does the grid kernel's time go?  Times bspy_cuda_eval_grid_batch on the teapot
batch (32 patches, n x n grid) for different output subsets; prints GB/s of algorithmic output bytes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bspy_b200 as bspy
from bspy_b200 import _cuda

def main(n=2048, S=32):
    t = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "teapot.npz"))
    kn = t["knots"]
    splines = [bspy.Spline(2, 3, (4, 4), (4, 4), (kn, kn), t["coefs"][p % 32]) for p in range(S)]
    batch = bspy.SplineBatch.from_splines(splines)
    ax = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
    out = {"values": torch.empty((S, 3, n, n), dtype=torch.float64, device="cuda"),
           "jacobian": torch.empty((S, 3, 2, n, n), dtype=torch.float64, device="cuda"),
           "normal": torch.empty((S, 3, n, n), dtype=torch.float64, device="cuda")}
    combos = [("values", dict(values=True), 24), ("jacobian", dict(values=False, jacobian=True), 48),
              ("normal", dict(values=False, normal=True), 24), ("val+jac", dict(values=True, jacobian=True), 72),
              ("all", dict(values=True, jacobian=True, normal=True), 96),
              ("all-raw-normal", dict(values=True, jacobian=True, normal=True, normalize=False), 96)]
    for name, kw, bpp in combos:
        o = {k: (v if kw.get(k, k == "values" and kw.get("values", True)) else None) for k, v in out.items()}
        o = {"values": out["values"] if kw.get("values", True) else None,
             "jacobian": out["jacobian"] if kw.get("jacobian") else None,
             "normal": out["normal"] if kw.get("normal") else None}
        for _ in range(3):
            batch.evaluate_grid(ax, ax, check_domain=False, out=o, **kw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            batch.evaluate_grid(ax, ax, check_domain=False, out=o, **kw)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{os.environ.get('BSPY_GRID_STORE','cs'):8s} n={n} S={S} {name:16s} {ms:8.3f} ms  {S*n*n*bpp/ms/1e6:8.1f} GB/s  {S*n*n/ms/1e6:7.2f} Gpts/s", flush=True)

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2048, int(sys.argv[2]) if len(sys.argv) > 2 else 32)
