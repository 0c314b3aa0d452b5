"""``bspy`` import shim: makes the north star's module path literally true.

Put this directory on the path (``PYTHONPATH=/root/repo/shim:/root/repo``) and ``import bspy`` resolves to the B200-native
evaluation path: ``bspy.Spline`` (same constructor, ``evaluate / derivative / jacobian / normal / tangent_space /
bspline_values / domain``, JSON ``load / save``), ``bspy.SplineBlock``, ``bspy.Manifold`` and ``bspy._cuda`` -- the ctypes
binding of libbspy_cuda.so (``bspy_b200._cuda``).  Everything outside the evaluation path (fitting, intersection, solids,
the viewer) is NOT here: this shim is for callers that only evaluate.  It deliberately lives outside the repository root so
that it can never shadow the unmodified reference package that bench.py's CPU arm imports from baseline/_ref."""
import sys as _sys

import bspy_b200 as _impl
from bspy_b200 import *  # noqa: F401,F403
from bspy_b200 import _cuda, _spline_evaluation  # noqa: F401

__all__ = list(getattr(_impl, "__all__", [n for n in dir(_impl) if not n.startswith("_")]))
_sys.modules[__name__ + "._cuda"] = _cuda
_sys.modules[__name__ + "._spline_evaluation"] = _spline_evaluation
