"""Known-answer fixture for curvature, generated from the UNMODIFIED reference.

    python tests/golden/make_golden_curvature.py [/root/reference]

Reproduces the inputs of the reference's own test (``tests/bspy_test.py:693-700``): the circular-arc-like curve built by
``Spline.section([[1, 0, 90, 1], [0, 1, 180, 2]])`` (curvature 1.0 at u = 0 and 2.0 at u = 1, pinned there to 2e-15), its
planar projection ``testCurve @ [0, 1]`` at 101 parameters, and ``mySurface`` (``tests/bspy_test.py:119-122``, Gaussian
curvature 1.024 at (0.25, 0.5), pinned to 1e-14).  ``Spline.section`` is fitting code (out of scope here), so the fixture
stores the splines it returns together with the reference's curvature values.  Output: ``ref_curvature_kat.npz``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    bspy = import_reference(root)
    out = {}

    def put(tag, s):
        out[f"{tag}/nInd"], out[f"{tag}/nDep"] = np.int64(s.nInd), np.int64(s.nDep)
        out[f"{tag}/order"], out[f"{tag}/nCoef"] = np.array(s.order), np.array(s.nCoef)
        for i, k in enumerate(s.knots):
            out[f"{tag}/knots{i}"] = np.asarray(k, dtype=np.float64)
        out[f"{tag}/coefs"] = np.ascontiguousarray(s.coefs, dtype=np.float64)

    section = bspy.Spline.section([[1.0, 0.0, 90.0, 1.0], [0.0, 1.0, 180.0, 2.0]])
    put("section", section)
    out["section/u"] = np.array([0.0, 1.0])
    out["section/curvature"] = np.array([section.curvature(0.0), section.curvature(1.0)])
    planar = section @ [0, 1]
    put("planar", planar)
    u = np.linspace(0.0, 1.0, 101)
    out["planar/u"] = u
    out["planar/curvature"] = np.array([planar.curvature(x) for x in u])
    surf = bspy.Spline(2, 3, [3, 4], [4, 5], [[0, 0, 0, .5, 1, 1, 1], [0, 0, 0, 0, .5, 1, 1, 1, 1]],
                       [[0, 0, 0, 0, 0, .3, .3, .3, .3, .3, .7, .7, .7, .7, .7, 1, 1, 1, 1, 1],
                        [0, .25, .5, .75, 1, 0, .25, .5, .75, 1, 0, .25, .5, .75, 1, 0, .25, .5, .75, 1],
                        [0, 0, 0, 0, 0, 0, 1, 2, 1, 0, 0, 2, 1, 2, 0, 0, 0, 0, 0, 0]])
    put("surface", surf)
    g = np.linspace(0.05, 0.95, 7)
    uv = np.array([[0.25, 0.5]] + [[a, b] for a in g for b in g])
    out["surface/uv"] = uv
    with np.errstate(all="ignore"):
        out["surface/curvature"] = np.array([surf.curvature(p) for p in uv])
    assert abs(out["section/curvature"][0] - 1.0) < 2.0e-15 and abs(out["section/curvature"][1] - 2.0) < 2.0e-15
    assert abs(out["surface/curvature"][0] - 1.024) < 1.0e-14
    np.savez_compressed(os.path.join(HERE, "ref_curvature_kat.npz"), **out)
    print("wrote ref_curvature_kat.npz:", {k: v.shape for k, v in out.items() if k.endswith("curvature")})


if __name__ == "__main__":
    main()
