"""``Spline``: the part of the reference's ``bspy.Spline`` (``bspy/spline.py``) that the evaluation
path and JSON I/O touch, with identical constructor semantics, attribute names, argument-form
dispatch, return conventions, exceptions and file format -- evaluated on the GPU.

Kept from the reference API: ``__init__``, ``__call__``, ``__repr__``, ``bspline_values`` (static),
``copy``, ``derivative``, ``domain``, ``domain_dimension``, ``evaluate``, ``from_dict``, ``jacobian``,
``load``, ``negate_normal``, ``normal``, ``range_dimension``, ``save``, ``tangent_space``, ``to_dict``.
Added: ``evaluate_points``, ``evaluate_grid``, ``bspline_values_batch``, ``freeze`` / ``unfreeze``.
Everything else on the reference class (fitting, arithmetic, intersection, ...) is outside the
path this package replaces.
"""
from __future__ import annotations

import json
from os import path

import numpy as np

from bspy_b200 import _spline_evaluation as _ev
from bspy_b200.manifold import Manifold


@Manifold.register
class Spline(Manifold):
    """Tensor-product B-spline: ``nInd`` independent and ``nDep`` dependent variables.

    Parameters (as in the reference, ``bspy/spline.py:46-76``)
    ----------
    nInd, nDep : int
    order, nCoef : length-``nInd`` sequences of int
    knots : ``nInd`` knot sequences, ``knots[i]`` of length ``order[i] + nCoef[i]``
    coefs : coefficients as ``(nDep, *nCoef)``, or as a list of ``prod(nCoef)`` points of length
        ``nDep`` (first independent variable varying fastest), or as ``nDep`` arrays
    metadata : dict, optional
    """

    def __init__(self, nInd, nDep, order, nCoef, knots, coefs, metadata={}):
        if not (nInd >= 0):
            raise ValueError("nInd < 0")
        if not (nDep >= 0):
            raise ValueError("nDep < 0")
        self.nInd, self.nDep = int(nInd), int(nDep)
        if len(order) != self.nInd:
            raise ValueError("len(order) != nInd")
        if len(nCoef) != self.nInd:
            raise ValueError("len(nCoef) != nInd")
        if len(knots) != nInd:
            raise ValueError("len(knots) != nInd")
        self.order = tuple(int(o) for o in order)
        self.nCoef = tuple(int(n) for n in nCoef)
        for i, kk in enumerate(knots):
            expected = self.order[i] + self.nCoef[i]
            if len(kk) != expected:
                raise ValueError(f"Knots array for variable {i} should have length {expected}")
        self.knots = tuple(np.array(kk) for kk in knots)
        for kk, o, n in zip(self.knots, self.order, self.nCoef):
            # non-decreasing, and no knot of multiplicity above the order (every B-spline has support)
            if not (np.all(kk[:n] <= kk[1:n + 1]) and np.all(kk[o:o + n] - kk[:n] > 0)):
                raise ValueError("Improper knot order or multiplicity")
        total = int(np.prod(self.nCoef, dtype=np.int64)) if self.nInd else 1
        if not (len(coefs) == total or len(coefs) == self.nDep):
            raise ValueError(f"Length of coefs should be {total} or {self.nDep}")
        c = np.array(coefs)
        wanted = (self.nDep, *self.nCoef)
        if c.shape != wanted:
            if len(c) == total:
                # list of points, first variable fastest -> (nDep, n_0, ..., n_last) view
                c = c.reshape((*self.nCoef[::-1], self.nDep)).T
            else:
                c = np.array([block.T for block in c]).reshape(wanted)
        self.coefs = c
        self.metadata = dict(metadata)

    # ---- evaluation (reference dispatch: bspy/spline.py:904-949, 720-770) ------------------
    def __call__(self, *uvw, **kwargs):
        return self.evaluate(*uvw, **kwargs)

    def __repr__(self):
        return f"Spline({self.nInd}, {self.nDep}, {self.order}, {self.nCoef}, {self.knots}, {self.coefs}, {self.metadata})"

    def _ufunc_style(self, uvw, kwargs, with_respect_to):
        """numpy-ufunc calling convention of the reference: ``nInd`` broadcastable arrays in, a tuple
        of ``nDep`` arrays (or one array when nDep == 1) out, cast to ``coefs.dtype``."""
        kwargs = dict(kwargs)
        where, outs = kwargs.pop("where", True), kwargs.pop("out", None)
        if kwargs:
            raise NotImplementedError(f"ufunc keyword arguments other than where= / out= are not supported by the CUDA path: {sorted(kwargs)}")
        if len(uvw) != self.nInd:
            raise ValueError("invalid number of arguments")
        if where is not True or outs is not None:
            return self._ufunc_where_out(uvw, where, outs, with_respect_to)
        arrays = np.broadcast_arrays(*[np.asarray(a, dtype=np.float64) for a in uvw])
        shape = arrays[0].shape
        pts = np.stack([a.reshape(-1) for a in arrays], axis=1)
        res = _ev.evaluate_points(self, pts, values=with_respect_to is None, with_respect_to=with_respect_to)
        soa = res.values if with_respect_to is None else res.derivative
        dt = self.coefs.dtype
        if self.nDep > 1:
            return tuple(soa[d].reshape(shape).astype(dt, copy=False) for d in range(self.nDep))
        out = soa[0].reshape(shape)
        if out.ndim >= 2:
            # the reference iterates the rows of the result array here (bspy/spline.py:947) and keeps
            # entry 0 of each; reproduced so that callers see the same shape
            out = out[:, 0][..., None]
        return np.array(out, dt)

    def _ufunc_where_out(self, uvw, where, outs, with_respect_to):
        """``where=`` / ``out=`` of the reference's ``np.frompyfunc`` ufunc (bspy/spline.py:940-947): only the selected points
        are evaluated (and domain-checked); results land in the ``out`` arrays -- object arrays as ``frompyfunc`` produces
        them when none are given, so that entries left out by ``where`` stay ``None`` and the final cast to ``coefs.dtype``
        raises exactly where the reference's does."""
        arrays = np.broadcast_arrays(*[np.asarray(a, dtype=np.float64) for a in uvw], np.asarray(where, dtype=bool))
        mask = arrays[-1]
        shape = mask.shape
        if outs is None:
            outs = tuple(np.empty(shape, dtype=object) for _ in range(self.nDep))
        elif not isinstance(outs, tuple):
            outs = (outs,)
        if len(outs) != self.nDep:
            raise ValueError("The 'out' tuple must have exactly one entry per ufunc output")
        for o in outs:
            if o.shape != shape:
                raise ValueError(f"non-broadcastable output operand with shape {o.shape} doesn't match the broadcast shape {shape}")
        sel = np.nonzero(mask.reshape(-1))[0]
        if sel.size:
            pts = np.stack([a.reshape(-1)[sel] for a in arrays[:-1]], axis=1)
            res = _ev.evaluate_points(self, pts, values=with_respect_to is None, with_respect_to=with_respect_to)
            soa = res.values if with_respect_to is None else res.derivative
            for d, o in enumerate(outs):
                o[mask] = soa[d]
        dt = self.coefs.dtype
        if self.nDep > 1:
            return tuple(o.astype(dt, copy=False) for o in outs)
        # nDep == 1: the reference iterates the first axis of the result and keeps entry 0 (a 1-tuple) of each row (spline.py:947);
        # an entry left out by where= is None there and cannot be subscripted
        if outs[0].ndim >= 2:
            return np.array([[x[0]] for x in outs[0]], dt)
        if any(x is None for x in outs[0]):
            raise TypeError("'NoneType' object is not subscriptable")
        return np.array(outs[0], dt)

    def evaluate(self, *uvw, **kwargs):
        """Value of the spline.  ``s(0.2, 0.3)``, ``s([0.2, 0.3])`` -> ``ndarray (nDep,)``;
        ``s(uArray, vArray)`` (numpy ufunc style) -> tuple of ``nDep`` arrays."""
        if len(uvw) == 0 and self.nInd == 0:
            return self.coefs
        if np.isscalar(uvw[0]):
            return _ev.evaluate(self, uvw)
        if len(uvw) > 1 or len(uvw[0]) > self.nInd:
            return self._ufunc_style(uvw, kwargs, None)
        return _ev.evaluate(self, *uvw)

    def derivative(self, with_respect_to, *uvw, **kwargs):
        """Mixed partial of order ``with_respect_to[i]`` in variable ``i``; same argument forms as
        ``evaluate``."""
        if len(uvw) == 0 and self.nInd == 0:
            return np.zeros(self.nDep, self.coefs.dtype)
        if np.isscalar(uvw[0]):
            return _ev.derivative(self, with_respect_to, uvw)
        if len(uvw) > 1 or len(uvw[0]) > self.nInd:
            return self._ufunc_style(uvw, kwargs, [int(w) for w in with_respect_to])
        return _ev.derivative(self, with_respect_to, *uvw)

    def jacobian(self, uvw):
        """``(nDep, nInd)`` matrix of first partials at one point."""
        return _ev.jacobian(self, uvw)

    def tangent_space(self, uvw):
        """Same as ``jacobian`` (tangents are the columns)."""
        return _ev.jacobian(self, uvw)

    def normal(self, uvw, normalize=True, indices=None):
        """Normal at one point; needs ``|nInd - nDep| == 1``."""
        return _ev.normal(self, uvw, normalize, indices)

    def contract(self, uvw):
        return _ev.contract(self, uvw)
    contract.__doc__ = _ev.contract.__doc__

    def curvature(self, uv):
        """Curvature of a curve (signed if planar) or Gaussian curvature of a surface at one point."""
        return _ev.curvature(self, uv)

    def curvature_points(self, uvw, **kwargs):
        return _ev.curvature_points(self, uvw, **kwargs)
    curvature_points.__doc__ = _ev.curvature_points.__doc__

    def domain(self):
        """``(nInd, 2)`` array of lower / upper parameter bounds."""
        return _ev.domain(self)

    def domain_dimension(self):
        return self.nInd

    def range_dimension(self):
        return self.nDep

    @staticmethod
    def bspline_values(knot, knots, splineOrder, u, derivativeOrder=0, taylorCoefs=False):
        """``(knot, values)``: the ``splineOrder`` non-zero B-spline values (or derivatives, or Taylor
        coefficients) at ``u``; ``knot=None`` finds the span by binary search."""
        return _ev.bspline_values(knot, knots, splineOrder, u, derivativeOrder, taylorCoefs)

    # ---- added vectorised entry points -----------------------------------------------------
    def evaluate_points(self, uvw, **kwargs):
        return _ev.evaluate_points(self, uvw, **kwargs)
    evaluate_points.__doc__ = _ev.evaluate_points.__doc__

    def evaluate_grid(self, *axes, **kwargs):
        return _ev.evaluate_grid(self, *axes, **kwargs)
    evaluate_grid.__doc__ = _ev.evaluate_grid.__doc__

    @staticmethod
    def bspline_values_batch(knot, knots, splineOrder, u, derivativeOrder=0, taylorCoefs=False):
        return _ev.bspline_values_batch(knot, knots, splineOrder, u, derivativeOrder, taylorCoefs)
    bspline_values_batch.__doc__ = _ev.bspline_values_batch.__doc__

    @staticmethod
    def collocation_matrix(knots, splineOrder, uValues, derivativeOrders=None):
        return _ev.collocation_matrix(knots, splineOrder, uValues, derivativeOrders)
    collocation_matrix.__doc__ = _ev.collocation_matrix.__doc__

    def freeze(self, device=None):
        return _ev.freeze(self, device)
    freeze.__doc__ = _ev.freeze.__doc__

    def unfreeze(self):
        return _ev.unfreeze(self)

    # ---- object plumbing -------------------------------------------------------------------
    def copy(self):
        return type(self)(self.nInd, self.nDep, self.order, self.nCoef, self.knots, self.coefs, self.metadata)

    def negate_normal(self):
        """Copy of the spline whose normal points the other way (same tangent space)."""
        s = self.copy()
        s.metadata["negateNormal"] = not self.metadata.get("negateNormal", False)
        return s

    # ---- JSON (reference bspy/spline.py:1099-1125, 1542-1583, 1998-2026, 2254-2267) ---------
    def to_dict(self):
        return {"type": "Spline", "nInd": self.nInd, "nDep": self.nDep, "order": self.order, "nCoef": self.nCoef,
                "knots": self.knots, "coefs": self.coefs, "metadata": self.metadata}

    @staticmethod
    def from_dict(dictionary):
        s = Spline(dictionary["nInd"], dictionary["nDep"], dictionary["order"], dictionary["nCoef"],
                   [np.array(k) for k in dictionary["knots"]], np.array(dictionary["coefs"]),
                   dictionary.get("metadata", {}))
        if s.metadata.get("flipNormal", False):     # files written by old versions
            s.metadata["negateNormal"] = True
            del s.metadata["flipNormal"]
        return s

    @staticmethod
    def load(fileName):
        """List of the splines stored in ``fileName`` (json; the legacy npz layout is tried first)."""
        try:
            with np.load(fileName) as kw:
                order = kw["order"]
                knots = [kw[f"knots{i}"] for i in range(len(order))]
                c = kw["coefficients"]
            return [Spline(len(order), c.shape[0], order, c.shape[1:], knots, c,
                           metadata=dict(Path=path, Name=path.splitext(path.split(fileName)[1])[0]))]
        except Exception:
            pass
        with open(fileName, "r", encoding="utf-8") as f:
            data = json.load(f)
        if isinstance(data, dict):
            data = [data]
        return [Spline.from_dict(d) for d in data]

    def save(self, fileName, *additional_splines):
        """Write this spline (and any further ones) as json, ``indent=4``: one dict, or a list."""
        class _Encoder(json.JSONEncoder):
            def default(self, obj):
                if isinstance(obj, np.ndarray):
                    return obj.tolist()
                if isinstance(obj, Spline):
                    return obj.to_dict()
                return super().default(obj)

        with open(fileName, "w", encoding="utf-8") as f:
            json.dump((self, *additional_splines) if additional_splines else self, f, indent=4, cls=_Encoder)
