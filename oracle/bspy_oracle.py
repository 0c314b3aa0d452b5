"""CPU oracle for the BSpy evaluation path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's evaluation hot path so
that the CUDA kernels in ``bspy_b200/_cuda`` can be checked against it.  It is **not**
part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import it.  The product path
(``bspy_b200``) never imports anything from ``oracle/`` and fails loudly without the
CUDA library.

Parity status: **pinned**.  ``tests/test_oracle_golden.py`` checks every function here
against (i) the reference's own golden tables ``truthCurve`` / ``truthSurface``
(reference ``tests/bspy_test.py:15-564``) and (ii) outputs of the unmodified reference
generated in the build container by ``tests/golden/make_golden.py`` (spans and basis
values bit-for-bit; values, derivatives, jacobians and normals to a few ulp).

Reference lines followed (all in ``bspy/_spline_evaluation.py`` unless noted):

* span search ............ ``:7-8``    (``np.searchsorted(..., 'right')`` + clamp)
* basis recurrence ....... ``:9-26``   (value stages, derivative stages, taylorCoefs)
* domain ................. ``:135-138``
* evaluate / derivative .. ``:140-164`` / ``:109-133`` (window slice, last variable first)
* jacobian ............... ``:205-213``
* normal ................. ``:215-246``  (cofactors via ``np.linalg.det``)
* ufunc-style dispatch ... ``bspy/spline.py:757-770, 936-949``
* contract ............... ``bspy/_spline_operations.py:184-223``
* SplineBlock sums ....... ``bspy/spline_block.py:37-44, 179-282``

Two tiers live here:

``*_pt``  functions  evaluate ONE point with Python loops and numpy float64 scalars, in
          the same operation order as the reference (this is also the cost model of
          the reference: one interpreter pass per point), and
``*_vec`` functions  evaluate N points with numpy array arithmetic.  Every elementwise
          operation is performed in the same order as in the ``_pt`` tier, so spans and
          basis values are bit-identical; contractions use matmul like the reference.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "OracleSpline", "span_pt", "basis_pt", "domain", "derivative_pt", "evaluate_pt",
    "jacobian_pt", "normal_pt", "span_vec", "basis_vec", "derivative_vec", "evaluate_vec",
    "jacobian_vec", "normal_vec", "check_domain_vec", "curvature_vec", "contract", "OracleBlock",
]


class OracleSpline:
    """Plain container: the attributes the evaluation path reads (reference
    ``bspy/spline.py:46-76``): ``nInd, nDep, order, nCoef, knots, coefs, metadata``.
    ``coefs`` has shape ``(nDep, *nCoef)``; ``knots[i]`` has ``order[i] + nCoef[i]`` entries."""

    def __init__(self, nInd, nDep, order, nCoef, knots, coefs, metadata=None):
        self.nInd = int(nInd)
        self.nDep = int(nDep)
        self.order = tuple(int(o) for o in order)
        self.nCoef = tuple(int(n) for n in nCoef)
        self.knots = tuple(np.asarray(k, dtype=np.float64) for k in knots)
        self.coefs = np.asarray(coefs, dtype=np.float64).reshape((self.nDep, *self.nCoef))
        self.metadata = dict(metadata or {})

    @classmethod
    def of(cls, s):
        """Build from any object with the Spline attributes (product Spline, reference Spline)."""
        return cls(s.nInd, s.nDep, s.order, s.nCoef, s.knots, np.asarray(s.coefs, dtype=np.float64),
                   getattr(s, "metadata", None))


# --------------------------------------------------------------------------- scalar tier

def span_pt(knots, order, u):
    """Rightmost knot index of the span holding ``u`` (``_spline_evaluation.py:7-8``):
    number of knots <= u, clamped to [order, len(knots) - order]."""
    ix = int(np.searchsorted(knots, u, side="right"))
    lo, hi = order, len(knots) - order
    return hi if ix > hi else (lo if ix < lo else ix)


def basis_pt(ix, knots, order, u, deriv=0, taylor=False):
    """The ``order`` non-zero B-spline values (or ``deriv``-th derivatives) on the span whose
    rightmost knot index is ``ix`` (``_spline_evaluation.py:4-27``).  ``ix=None`` searches.
    Returns ``(ix, basis)``.  Separate multiply and add, true divisions: bit-identical to
    the reference."""
    knots = np.asarray(knots)
    out = np.zeros(order, knots.dtype)
    if ix is None:
        ix = span_pt(knots, order, u)
    if deriv >= order:
        return ix, out
    out[order - 1] = 1.0
    nValueStages = order - deriv          # degrees 1 .. nValueStages-1 are value stages
    for deg in range(1, order):
        slot = order - deg
        if deg < nValueStages:
            for i in range(ix - deg, ix):
                a = (u - knots[i]) / (knots[i + deg] - knots[i])
                out[slot - 1] += (1.0 - a) * out[slot]
                out[slot] *= a
                slot += 1
        else:
            scale = deg / ((order - deg) if taylor else 1.0)
            for i in range(ix - deg, ix):
                a = scale / (knots[i + deg] - knots[i])
                out[slot - 1] += -a * out[slot]
                out[slot] *= a
                slot += 1
    return ix, out


def domain(s):
    """``[[k_i[o_i-1], k_i[n_i]]]`` (``_spline_evaluation.py:135-138``)."""
    return np.array([[s.knots[i][s.order[i] - 1], s.knots[i][s.nCoef[i]]] for i in range(s.nInd)])


def _check_point(s, uvw):
    uvw = np.atleast_1d(uvw)
    if len(uvw) != s.nInd:
        raise ValueError(f"Incorrect number of parameter values: {len(uvw)}")
    box = domain(s)
    for i in range(s.nInd):
        if uvw[i] < box[i][0] or uvw[i] > box[i][1]:
            raise ValueError(f"Spline evaluation outside domain: {uvw}")
    return uvw


def derivative_pt(s, wrt, uvw):
    """Mixed partial of order ``wrt[i]`` in variable ``i`` at one point
    (``_spline_evaluation.py:109-133``); contraction runs from the last variable to the first."""
    uvw = _check_point(s, uvw)
    window = [slice(0, s.nDep)]
    rows = []
    for i in range(s.nInd):
        ix, b = basis_pt(None, s.knots[i], s.order[i], uvw[i], wrt[i])
        rows.append(b)
        window.append(slice(ix - s.order[i], ix))
    acc = s.coefs[tuple(window)]
    for i in reversed(range(s.nInd)):
        acc = acc @ rows[i]
    return acc


def evaluate_pt(s, uvw):
    """Value at one point (``_spline_evaluation.py:140-164``)."""
    return derivative_pt(s, [0] * s.nInd, uvw)


def jacobian_pt(s, uvw):
    """``(nDep, nInd)`` matrix of first partials (``_spline_evaluation.py:205-213``)."""
    J = np.empty((s.nDep, s.nInd), s.coefs.dtype)
    for i in range(s.nInd):
        e = [0] * s.nInd
        e[i] = 1
        J[:, i] = derivative_pt(s, e, uvw)
    return J


def normal_pt(s, uvw, normalize=True, indices=None):
    """Cofactor normal (``_spline_evaluation.py:215-246``)."""
    uvw = np.atleast_1d(uvw)
    if abs(s.nInd - s.nDep) != 1:
        raise ValueError("The number of independent variables must be one different than the number of dependent variables.")
    T = jacobian_pt(s, uvw)
    if s.nInd > s.nDep:
        T = T.T
    D = T.shape[0]
    sign = -1 if s.metadata.get("negateNormal", False) else 1
    which = range(D) if indices is None else indices
    n = np.empty(len(which), s.coefs.dtype)
    for slot, i in enumerate(which):
        keep = [j for j in range(D) if j != i]
        n[slot] = sign * ((-1) ** i) * np.linalg.det(T[keep])
    if normalize:
        n /= np.linalg.norm(n)
    return n


# ----------------------------------------------------------------------- vectorised tier

def span_vec(knots, order, u):
    """``span_pt`` for an array of parameters (int32).  NaN sorts last, as in numpy."""
    knots = np.asarray(knots, dtype=np.float64)
    ix = np.searchsorted(knots, np.asarray(u, dtype=np.float64), side="right")
    return np.clip(ix, order, len(knots) - order).astype(np.int32)


def basis_vec(knots, order, u, deriv=0, taylor=False, ix=None):
    """``basis_pt`` for N parameters: returns ``(ix[N] int32, basis[N, order])``.
    Elementwise operations are issued in the reference's order, so the result is
    bit-identical to N scalar calls."""
    knots = np.asarray(knots, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64).reshape(-1)
    if ix is None:
        ix = span_vec(knots, order, u)
    ix = np.asarray(ix).astype(np.int64)
    out = np.zeros((u.shape[0], order))
    if deriv >= order:
        return ix.astype(np.int32), out
    out[:, order - 1] = 1.0
    nValueStages = order - deriv
    with np.errstate(all="ignore"):
        for deg in range(1, order):
            slot = order - deg
            for t in range(deg):
                lo = knots[ix - deg + t]
                gap = knots[ix + t] - lo
                if deg < nValueStages:
                    a = (u - lo) / gap
                    out[:, slot - 1] += (1.0 - a) * out[:, slot]
                else:
                    scale = deg / ((order - deg) if taylor else 1.0)
                    a = scale / gap
                    out[:, slot - 1] += -a * out[:, slot]
                out[:, slot] *= a
                slot += 1
    return ix.astype(np.int32), out


def check_domain_vec(s, uvw):
    """Index of the first point outside the closed domain, or -1.  NaN is inside
    (both comparisons are false), as in the reference."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    box = domain(s)
    bad = np.zeros(uvw.shape[0], bool)
    for i in range(s.nInd):
        bad |= (uvw[:, i] < box[i, 0]) | (uvw[:, i] > box[i, 1])
    hits = np.flatnonzero(bad)
    return int(hits[0]) if hits.size else -1


def derivative_vec(s, wrt, uvw, return_spans=False):
    """``derivative_pt`` for ``uvw[N, nInd]`` → ``(N, nDep)`` (and ``spans[N, nInd]``)."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    N = uvw.shape[0]
    spans = np.empty((N, s.nInd), np.int32)
    rows = []
    for i in range(s.nInd):
        ix, b = basis_vec(s.knots[i], s.order[i], uvw[:, i], wrt[i])
        spans[:, i] = ix
        rows.append(b)
    # gather the coefficient window of every point: (N, nDep, o_0, ..., o_last)
    index = [np.arange(s.nDep).reshape((1, s.nDep) + (1,) * s.nInd)]
    for i in range(s.nInd):
        shape = [1] * (2 + s.nInd)
        shape[2 + i] = s.order[i]
        offs = np.arange(s.order[i]).reshape(shape)
        base = (spans[:, i].astype(np.int64) - s.order[i]).reshape((N,) + (1,) * (1 + s.nInd))
        index.append(base + offs)
    acc = s.coefs[tuple(index)]
    with np.errstate(all="ignore"):
        for i in reversed(range(s.nInd)):
            b = rows[i].reshape((N,) + (1,) * (acc.ndim - 3) + (s.order[i], 1))
            acc = np.matmul(acc, b)[..., 0]
    return (acc, spans) if return_spans else acc


def evaluate_vec(s, uvw, return_spans=False):
    return derivative_vec(s, [0] * s.nInd, uvw, return_spans)


def jacobian_vec(s, uvw):
    """``(N, nDep, nInd)``."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    J = np.empty((uvw.shape[0], s.nDep, s.nInd))
    for i in range(s.nInd):
        e = [0] * s.nInd
        e[i] = 1
        J[:, :, i] = derivative_vec(s, e, uvw)
    return J


def normal_vec(s, uvw, normalize=True, indices=None):
    """``(N, D)`` with ``D = max(nInd, nDep)`` (or ``len(indices)``)."""
    if abs(s.nInd - s.nDep) != 1:
        raise ValueError("The number of independent variables must be one different than the number of dependent variables.")
    T = jacobian_vec(s, uvw)
    if s.nInd > s.nDep:
        T = np.swapaxes(T, 1, 2)
    D = T.shape[1]
    sign = -1 if s.metadata.get("negateNormal", False) else 1
    which = list(range(D)) if indices is None else list(indices)
    n = np.empty((T.shape[0], len(which)))
    with np.errstate(all="ignore"):
        for slot, i in enumerate(which):
            keep = [j for j in range(D) if j != i]
            n[:, slot] = sign * ((-1) ** i) * np.linalg.det(T[:, keep, :])
        if normalize:
            n /= np.sqrt(np.sum(n * n, axis=1))[:, None]
    return n


def curvature_vec(s, uvw):
    """Curvature at N points, reference ``:80-107``: curves -> cross-product formula (signed for planar curves),
    surfaces -> Gaussian curvature from the fundamental forms; nDep == 1 -> graph of the function (the reference
    builds ``self.graph()``, whose extra coordinates are the parameters themselves: derivative 1, second derivative 0)."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    N = uvw.shape[0]
    graph = s.nDep == 1
    with np.errstate(all="ignore"):
        if s.nInd == 1:
            fp, fpp = derivative_vec(s, [1], uvw), derivative_vec(s, [2], uvw)
            if graph:
                fp = np.concatenate([np.ones((N, 1)), fp], axis=1)
                fpp = np.concatenate([np.zeros((N, 1)), fpp], axis=1)
            pp, pq, qq = (fp * fp).sum(1), (fp * fpp).sum(1), (fpp * fpp).sum(1)
            if fp.shape[1] == 2:
                num = fp[:, 0] * fpp[:, 1] - fp[:, 1] * fpp[:, 0]
            else:
                num = np.sqrt(qq * pp - pq ** 2)
            return num / pp ** 1.5
        d = {w: derivative_vec(s, list(w), uvw) for w in ((1, 0), (0, 1), (2, 0), (1, 1), (0, 2))}
        if graph:
            z, o = np.zeros((N, 1)), np.ones((N, 1))
            su, sv = np.hstack([o, z, d[1, 0]]), np.hstack([z, o, d[0, 1]])
            suu, suv, svv = (np.hstack([z, z, d[w]]) for w in ((2, 0), (1, 1), (0, 2)))
        else:
            su, sv, suu, suv, svv = d[1, 0], d[0, 1], d[2, 0], d[1, 1], d[0, 2]
        n = np.cross(su, sv)
        n = n / np.sqrt((n * n).sum(1))[:, None]
        E, F, G = (su * su).sum(1), (su * sv).sum(1), (sv * sv).sum(1)
        L, M, Nn = (suu * n).sum(1), (suv * n).sum(1), (svv * n).sum(1)
        return (L * Nn - M ** 2) / (E * G - F ** 2)


def curvature_condition_vec(s, uvw):
    """Relative condition number of the curvature formula at N points: sum of |terms| over |result| of every difference
    in it (the cross-product numerator, L N - M^2, E G - F^2).  Where it is large the reference's own value carries a
    relative error of about eps * condition, whatever computes the derivatives; tests use it to state the curvature bar
    (strict 1e-12 relative where the condition is <= 100)."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    N = uvw.shape[0]
    graph = s.nDep == 1
    with np.errstate(all="ignore"):
        if s.nInd == 1:
            fp, fpp = derivative_vec(s, [1], uvw), derivative_vec(s, [2], uvw)
            if graph:
                fp = np.concatenate([np.ones((N, 1)), fp], axis=1)
                fpp = np.concatenate([np.zeros((N, 1)), fpp], axis=1)
            pp, pq, qq = (fp * fp).sum(1), (fp * fpp).sum(1), (fpp * fpp).sum(1)
            if fp.shape[1] == 2:
                a, b = fp[:, 0] * fpp[:, 1], fp[:, 1] * fpp[:, 0]
                return (np.abs(a) + np.abs(b)) / np.abs(a - b)
            return (qq * pp + pq ** 2) / np.abs(qq * pp - pq ** 2)
        d = {w: derivative_vec(s, list(w), uvw) for w in ((1, 0), (0, 1), (2, 0), (1, 1), (0, 2))}
        if graph:
            z, o = np.zeros((N, 1)), np.ones((N, 1))
            su, sv = np.hstack([o, z, d[1, 0]]), np.hstack([z, o, d[0, 1]])
            suu, suv, svv = (np.hstack([z, z, d[w]]) for w in ((2, 0), (1, 1), (0, 2)))
        else:
            su, sv, suu, suv, svv = d[1, 0], d[0, 1], d[2, 0], d[1, 1], d[0, 2]
        n = np.cross(su, sv)
        nlen = np.sqrt((n * n).sum(1))
        n = n / nlen[:, None]
        E, F, G = (su * su).sum(1), (su * sv).sum(1), (sv * sv).sum(1)
        L, M, Nn = (suu * n).sum(1), (suv * n).sum(1), (svv * n).sum(1)
        dots = sum((np.abs(x * n)).sum(1) / np.maximum(np.abs((x * n).sum(1)), 1e-300) for x in (suu, suv, svv)) / 3
        cross = np.sqrt(E * G) / nlen
        return (np.abs(L * Nn) + M ** 2) / np.abs(L * Nn - M ** 2) + (E * G + F ** 2) / np.abs(E * G - F ** 2) + dots + cross


def _curvature_formula(nInd, ders):
    """reference ``:80-107`` from the derivative arrays (each (N, nDep'))"""
    with np.errstate(all="ignore"):
        if nInd == 1:
            fp, fpp = ders
            pp, pq, qq = (fp * fp).sum(1), (fp * fpp).sum(1), (fpp * fpp).sum(1)
            num = fp[:, 0] * fpp[:, 1] - fp[:, 1] * fpp[:, 0] if fp.shape[1] == 2 else np.sqrt(qq * pp - pq ** 2)
            return num / pp ** 1.5
        su, sv, suu, suv, svv = ders
        n = np.cross(su, sv)
        n = n / np.sqrt((n * n).sum(1))[:, None]
        E, F, G = (su * su).sum(1), (su * sv).sum(1), (sv * sv).sum(1)
        L, M, Nn = (suu * n).sum(1), (suv * n).sum(1), (svv * n).sum(1)
        return (L * Nn - M ** 2) / (E * G - F ** 2)


def curvature_uncertainty_vec(s, uvw, k=16.0, trials=12, seed=7):
    """How much the reference's curvature at N points moves when every derivative entering the formula is perturbed by its
    own rounding uncertainty k * eps * sum|terms| (derivative_abs_vec): the largest |change| over a few random sign
    patterns.  Where derivatives are sums that cancel (flat regions of examples/TomsNasty.json: sum|terms| / |value| up to
    1e17) the formula amplifies rounding noise and the reference's value is only defined to that uncertainty -- for anything
    that adds the terms in another order than numpy.  Tests state the curvature bar as strict + this."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    N = uvw.shape[0]
    graph = s.nDep == 1
    wrts = ([1], [2]) if s.nInd == 1 else ([1, 0], [0, 1], [2, 0], [1, 1], [0, 2])
    with np.errstate(all="ignore"):
        ders = [derivative_vec(s, list(w), uvw) for w in wrts]
        unc = [k * np.finfo(float).eps * derivative_abs_vec(s, list(w), uvw) for w in wrts]

        def lift(arrs):
            if not graph:
                return arrs
            z, o = np.zeros((N, 1)), np.ones((N, 1))
            if s.nInd == 1:
                return [np.hstack([o, arrs[0]]), np.hstack([z, arrs[1]])]
            return [np.hstack([o, z, arrs[0]]), np.hstack([z, o, arrs[1]])] + [np.hstack([z, z, a]) for a in arrs[2:]]

        base = _curvature_formula(s.nInd, lift(ders))
        rng = np.random.default_rng(seed)
        worst = np.zeros(N)
        for _ in range(trials):
            moved = [d + u * rng.choice([-1.0, 1.0], size=d.shape) for d, u in zip(ders, unc)]
            worst = np.fmax(worst, np.abs(_curvature_formula(s.nInd, lift(moved)) - base))
        return np.where(np.isfinite(worst), worst, np.inf)


# ------------------------------------------------------------- conditioning of the sums
# The parity bar of the path is |x - ref| <= 1e-13 + 1e-12*|ref|.  A value/derivative is a
# sum of products coefficient x basis values; any implementation that adds those terms in a
# different order than numpy's matmul differs from the reference by up to a few
# eps * sum(|terms|).  For splines with nearly coincident knots (the reference's
# tests/trim-issue.json has a knot gap of 1.9e-6 at order 7) derivative basis values reach
# 5e5 and cancel, so eps*sum(|terms|) is far above 1e-12*|result| -- for the reference
# itself as much as for anything compared with it.  The functions below return
# sum(|terms|) so that tests can state the bar in its condition-aware form
#     |x - ref| <= 1e-13 + 1e-12*|ref| + k*eps*sum(|terms|)
# (the extra term is negligible for the well-conditioned north-star configurations).

def derivative_abs_vec(s, wrt, uvw):
    """sum |coef| * prod |basis|  for every point and dependent variable: ``(N, nDep)``."""
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    N = uvw.shape[0]
    index = [np.arange(s.nDep).reshape((1, s.nDep) + (1,) * s.nInd)]
    rows = []
    for i in range(s.nInd):
        ix, b = basis_vec(s.knots[i], s.order[i], uvw[:, i], wrt[i])
        rows.append(np.abs(b))
        shape = [1] * (2 + s.nInd)
        shape[2 + i] = s.order[i]
        base = (ix.astype(np.int64) - s.order[i]).reshape((N,) + (1,) * (1 + s.nInd))
        index.append(base + np.arange(s.order[i]).reshape(shape))
    acc = np.abs(s.coefs)[tuple(index)]
    with np.errstate(all="ignore"):
        for i in reversed(range(s.nInd)):
            acc = np.matmul(acc, rows[i].reshape((N,) + (1,) * (acc.ndim - 3) + (s.order[i], 1)))[..., 0]
    return acc


def jacobian_abs_vec(s, uvw):
    uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, s.nInd)
    S = np.empty((uvw.shape[0], s.nDep, s.nInd))
    for i in range(s.nInd):
        e = [0] * s.nInd
        e[i] = 1
        S[:, :, i] = derivative_abs_vec(s, e, uvw)
    return S


def normal_abs_vec(s, uvw):
    """Bound on sum(|terms|) of every cofactor: the permanent of the corresponding minor of
    the abs-scale jacobian (``(N, D)``)."""
    from itertools import permutations
    S = jacobian_abs_vec(s, uvw)
    if s.nInd > s.nDep:
        S = np.swapaxes(S, 1, 2)
    D = S.shape[1]
    out = np.zeros((S.shape[0], D))
    for i in range(D):
        keep = [j for j in range(D) if j != i]
        M = S[:, keep, :]
        for perm in permutations(range(D - 1)):
            term = np.ones(S.shape[0])
            for r, c in enumerate(perm):
                term = term * M[:, r, c]
            out[:, i] += term
    return out


# --------------------------------------------------------------------------- callers of the path

def contract(s, uvw):
    """``Spline.contract`` (``bspy/_spline_operations.py:184-223``): variables whose entry of ``uvw`` is not None are
    fixed; the coefficient window of each is contracted against its basis values, variable by variable, with the
    contracted axis moved last and ``coefs @ bValues`` like the reference."""
    box = domain(s)
    section = [slice(None)]
    bValues = []
    contracting = False
    for iv in range(s.nInd):
        if uvw[iv] is not None:
            if uvw[iv] < box[iv][0] or uvw[iv] > box[iv][1]:
                raise ValueError(f"Spline evaluation outside domain: {uvw}")
            ix, b = basis_pt(None, s.knots[iv], s.order[iv], uvw[iv])
            bValues.append(b)
            section.append(slice(ix - s.order[iv], ix))
            contracting = True
        else:
            bValues.append([])
            section.append(slice(None))
    if not contracting:
        return s
    order, nCoef, knots = list(s.order), list(s.nCoef), list(s.knots)
    coefs = s.coefs[tuple(section)]
    ix = 0
    for iv in range(s.nInd):
        if uvw[iv] is not None:
            del order[ix], nCoef[ix], knots[ix]
            coefs = np.moveaxis(coefs, ix + 1, -1) @ bValues[iv]
        else:
            ix += 1
    return OracleSpline(len(order), s.nDep, order, nCoef, knots, coefs, s.metadata)


class OracleBlock:
    """Rows of ``(map, OracleSpline)``: the sums of ``bspy/spline_block.py:37-44`` (values, derivatives), ``:231-245``
    (jacobian) and the cofactor normal of ``bspy/_spline_evaluation.py:215-246`` applied to the block jacobian."""

    def __init__(self, rows):
        self.rows = [[(list(m), sp) for m, sp in row] for row in rows]
        self.nDep = sum(row[0][1].nDep for row in self.rows)
        self.nInd = 1 + max(i for row in self.rows for m, _ in row for i in m)
        self.metadata = {}

    def derivative_vec(self, wrt, uvw):
        uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, self.nInd)
        out = np.zeros((uvw.shape[0], self.nDep))
        at = 0
        for row in self.rows:
            n = row[0][1].nDep
            for m, sp in row:
                out[:, at:at + n] += derivative_vec(sp, [wrt[i] for i in m], uvw[:, m])
            at += n
        return out

    def evaluate_vec(self, uvw):
        return self.derivative_vec([0] * self.nInd, uvw)

    def jacobian_vec(self, uvw):
        uvw = np.asarray(uvw, dtype=np.float64).reshape(-1, self.nInd)
        J = np.zeros((uvw.shape[0], self.nDep, self.nInd))
        at = 0
        for row in self.rows:
            n = row[0][1].nDep
            for m, sp in row:
                J[:, at:at + n, m] += jacobian_vec(sp, uvw[:, m])
            at += n
        return J

    def normal_vec(self, uvw, normalize=True, indices=None):
        if abs(self.nInd - self.nDep) != 1:
            raise ValueError("The number of independent variables must be one different than the number of dependent variables.")
        T = self.jacobian_vec(uvw)
        if self.nInd > self.nDep:
            T = np.swapaxes(T, 1, 2)
        D = T.shape[1]
        which = list(range(D)) if indices is None else list(indices)
        n = np.empty((T.shape[0], len(which)))
        with np.errstate(all="ignore"):
            for slot, i in enumerate(which):
                n[:, slot] = ((-1) ** i) * np.linalg.det(T[:, [j for j in range(D) if j != i], :])
            if normalize:
                n /= np.sqrt(np.sum(n * n, axis=1))[:, None]
        return n
