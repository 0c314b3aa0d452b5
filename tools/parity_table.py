"""How far the CUDA path is from the STRICT parity bar on every golden case (run on a GPU box):

    python tools/parity_table.py > gpurun_out/parity_table.md

For each case of tests/golden/ref_cases.npz (outputs of the unmodified reference) the worst ratio
|x - ref| / (1e-13 + 1e-12 |ref|) over values, jacobian entries, the case's mixed partials and raw / unit normals; <= 1 is
inside the bar.  The same ratio for the C restatement of the reference (oracle/bspy_oracle.c, a different summation order
than numpy's) is printed next to it: on the ill-conditioned fixtures both exceed 1 by similar factors, which is why the
tests hold those files to the condition-aware bar."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bspy_b200 as bspy  # noqa: E402
from golden_io import load_cases  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402


def ratio(x, ref):
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    fin = np.isfinite(ref) & np.isfinite(x)
    if not fin.any():
        return 0.0
    return float((np.abs(x - ref)[fin] / (1e-13 + 1e-12 * np.abs(ref[fin]))).max())


print("| case | nInd | nDep | order | values | jacobian | mixed partials | raw normal | unit normal | C oracle: values | C oracle: jacobian |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for c in load_cases():
    s = bspy.Spline(c.nInd, c.nDep, c.order, c.nCoef, c.knots, c.coefs, c.metadata)
    with np.errstate(all="ignore"):
        r = s.evaluate_points(c.uvw, values=True, jacobian=True)
        rv, rj = ratio(r.values.T, c["values"]), ratio(np.transpose(r.jacobian, (2, 0, 1)), c["jacobian"])
        rw = max([ratio(s.evaluate_points(c.uvw, values=False, with_respect_to=w).derivative.T, c["deriv_" + "_".join(map(str, w))])
                  for w in c.meta["wrt"]] or [0.0])
        rn = ru = float("nan")
        if c.meta["normal"]:
            rn = ratio(s.evaluate_points(c.uvw, values=False, normal=True, normalize=False).normal.T, c["normal_raw"])
            ru = ratio(s.evaluate_points(c.uvw, values=False, normal=True).normal.T, c["normal_unit"])
        o = CO.evaluate(s, c.uvw, values=True, jacobian=True)
        ov, oj = ratio(o["values"], c["values"]), ratio(o["jacobian"], c["jacobian"])
    print(f"| {c.tag} | {c.nInd} | {c.nDep} | {'x'.join(map(str, c.order))} | {rv:.3g} | {rj:.3g} | {rw:.3g} | {rn:.3g} | {ru:.3g} | {ov:.3g} | {oj:.3g} |")
