"""bspy_b200 -- B200-native batched evaluation behind the BSpy ``Spline`` API.

Drop-in for ONE path of ericbrec/BSpy: ``Spline.evaluate / derivative / jacobian / normal /
tangent_space / bspline_values / domain`` and the JSON load/save, plus the added vectorised entry
points ``Spline.evaluate_points`` / ``Spline.evaluate_grid``, the ``SplineBatch`` container, and -- the first of
the path's callers -- ``Spline.contract`` and ``SplineBlock`` evaluation.
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/bspy_cuda.h``; there is no CPU fallback.

    import bspy_b200 as bspy
    s = bspy.Spline.load("surface.json")[0]
    xyz = s(0.25, 0.5)                                     # reference API, one kernel launch
    r = s.evaluate_points(uv, jacobian=True, normal=True)  # (N, 2) numpy array or CUDA tensor in
"""
from bspy_b200.manifold import Manifold
from bspy_b200.spline import Spline
from bspy_b200._spline_evaluation import EvalResult
from bspy_b200.batch import SplineBatch
from bspy_b200.spline_block import SplineBlock

__version__ = "0.1.0"
__all__ = ["Manifold", "Spline", "SplineBatch", "SplineBlock", "EvalResult"]
