"""Host side of the evaluation path: same module-level functions as the reference's
``bspy/_spline_evaluation.py`` (``bspline_values, domain, evaluate, derivative, jacobian, normal``;
spline first, same argument meaning, same exceptions and messages), each of which launches the
sm_100a kernels of ``bspy_b200._cuda`` -- plus the vectorised entries ``evaluate_points`` and
``evaluate_grid`` that the reference lacks.

No arithmetic of the path happens in this file: it validates arguments, moves bytes (torch is the
plumbing for device memory, pinned host memory and streams) and reshapes results.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import numpy as np
import torch

from bspy_b200 import _cuda

__all__ = ["collocation_matrix", "contract", "curvature", "curvature_points", "EvalResult", "bspline_values", "bspline_values_batch", "domain", "evaluate", "derivative", "jacobian",
           "normal", "evaluate_points", "evaluate_grid", "device_spline", "freeze"]


@dataclass
class EvalResult:
    """Struct-of-arrays results of a batched evaluation (None = not requested).

    values (nDep, N) | derivative (nDep, N) | jacobian (nDep, nInd, N) |
    normal (len(indices) or max(nInd, nDep), N) | spans (nInd, N) int32.
    Arrays are numpy when the points were numpy / array-like, CPU tensors for CPU tensors and CUDA
    tensors for CUDA tensors.  For grids, N is replaced by the grid shape.

    ``out_layout="aos"``: ``records`` is the (N, stride) array of per-point records [values | jacobian (d, i) | normal]
    (stride = record length rounded up to 4 doubles) and values / jacobian / normal are strided VIEWS into it with the
    shapes above.  ``check_domain="defer"``: ``first_outside`` is the device int64[1] flag the kernels wrote (-1 = every
    point inside the domain); ``raise_if_outside()`` reads it and raises the reference's ValueError."""
    values: Any = None
    derivative: Any = None
    jacobian: Any = None
    normal: Any = None
    spans: Any = None
    records: Any = None
    first_outside: Any = None

    def raise_if_outside(self):
        if self.first_outside is not None and int(self.first_outside.item()) >= 0:
            raise ValueError(f"Spline evaluation outside domain: point index {int(self.first_outside.item())}")
        return self


# ----------------------------------------------------------------------------- device residency

def _result_dtype(self):
    dt = np.result_type(np.asarray(self.coefs).dtype, *[np.asarray(k).dtype for k in self.knots])
    return dt if np.issubdtype(dt, np.floating) else np.dtype(np.float64)


def _normal_sign(self):
    meta = getattr(self, "metadata", None)
    return -1 if (meta and meta.get("negateNormal", False)) else 1


class _Resident:
    __slots__ = ("knots_host", "coefs_host", "ds")


def device_spline(self, dev=None) -> _cuda.DeviceSpline:
    """The spline's knots and coefficients on ``dev`` (float64, contiguous).  Splines are mutable
    objects (users assign into ``spline.coefs``), so the cached device copy is revalidated against
    a host snapshot on every call and re-uploaded when anything changed; ``freeze`` skips that."""
    dev = _cuda.device(dev)
    frozen = getattr(self, "_bspy_frozen", None)
    if frozen is not None and frozen.device == dev:
        frozen.normal_sign = _normal_sign(self)
        frozen.c.normalSign = frozen.normal_sign
        return frozen
    cache = self.__dict__.setdefault("_bspy_device_cache", {})
    coefs = np.ascontiguousarray(self.coefs, dtype=np.float64)
    knots = [np.ascontiguousarray(k, dtype=np.float64) for k in self.knots]
    if coefs.shape != (self.nDep, *self.nCoef):
        raise ValueError(f"coefs shape {coefs.shape} does not match (nDep, *nCoef) = {(self.nDep, *self.nCoef)}")
    for i, k in enumerate(knots):
        if k.shape != (self.order[i] + self.nCoef[i],):
            raise ValueError(f"Knots array for variable {i} should have length {self.order[i] + self.nCoef[i]}")
    hit = cache.get(dev)
    if hit is not None and hit.coefs_host.shape == coefs.shape and np.array_equal(hit.coefs_host, coefs, equal_nan=True) \
            and len(hit.knots_host) == len(knots) and all(a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
                                                          for a, b in zip(hit.knots_host, knots)):
        hit.ds.normal_sign = _normal_sign(self)
        hit.ds.c.normalSign = hit.ds.normal_sign
        return hit.ds
    res = _Resident()
    res.knots_host = [k.copy() for k in knots]
    res.coefs_host = coefs.copy()
    res.ds = _cuda.DeviceSpline(self.nInd, self.nDep, self.order, self.nCoef,
                                [torch.from_numpy(k).to(dev) for k in res.knots_host],
                                torch.from_numpy(res.coefs_host).to(dev), _normal_sign(self))
    if self.nInd == 1:
        res.ds.build_curve_table()          # span tables of the curve, once per upload (fetched by TMA by big batches)
    cache[dev] = res
    return res.ds


def freeze(self, dev=None):
    """Upload once and stop revalidating: later calls reuse the device copy even if the host arrays
    are mutated (call ``freeze`` again, or ``unfreeze``, after editing the spline)."""
    self.__dict__.pop("_bspy_frozen", None)
    ds = device_spline(self, dev)
    self.__dict__["_bspy_frozen"] = ds
    return self


def unfreeze(self):
    self.__dict__.pop("_bspy_frozen", None)
    return self


# ----------------------------------------------------------------------------- reference API

def bspline_values(knot, knots, splineOrder, u, derivativeOrder=0, taylorCoefs=False):
    """``(knot, basis[splineOrder])`` for one parameter -- reference
    ``bspy/_spline_evaluation.py:4-27`` -- computed by ``bspy_cuda_basis`` (bit-identical)."""
    ix, b = bspline_values_batch(knot, knots, splineOrder, np.array([u], dtype=np.float64), derivativeOrder, taylorCoefs)
    return (int(ix[0]) if knot is None else knot), b[0]


def bspline_values_batch(knot, knots, splineOrder, u, derivativeOrder=0, taylorCoefs=False, dev=None):
    """Vectorised ``bspline_values``: ``u`` is a 1-D array / tensor of N parameters, ``knot`` is
    None (search), an int (same span for all) or an int array.  Returns ``(spans[N], basis[N, order])``
    as numpy arrays, or as CUDA tensors when ``u`` is a CUDA tensor."""
    on_device = isinstance(u, torch.Tensor) and u.is_cuda
    dev = u.device if on_device else _cuda.device(dev)
    kt = knots if (isinstance(knots, torch.Tensor) and knots.is_cuda) else \
        torch.from_numpy(np.ascontiguousarray(knots, dtype=np.float64)).to(dev)
    ut = u.contiguous().to(torch.float64) if on_device else \
        torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64).reshape(-1)).to(dev)
    spans_in = None
    if knot is not None:
        if np.isscalar(knot):
            spans_in = torch.full((ut.numel(),), int(knot), dtype=torch.int32, device=dev)
        elif isinstance(knot, torch.Tensor):
            spans_in = knot.to(device=dev, dtype=torch.int32).contiguous()
        else:
            spans_in = torch.from_numpy(np.ascontiguousarray(knot, dtype=np.int32)).to(dev)
        # a supplied span must keep its knot window knots[ix-order+1 .. ix+order-2] inside the array: the reference
        # would raise IndexError (or wrap a negative index) there; the kernel refuses such rows (NaN) and we raise
        if spans_in.numel():
            lo, hi = int(spans_in.min().item()), int(spans_in.max().item())
            if lo < int(splineOrder) - 1 or hi > kt.numel() - int(splineOrder) + 1:
                raise IndexError(f"knot index {lo if lo < int(splineOrder) - 1 else hi} is out of bounds for "
                                 f"{kt.numel()} knots of order {int(splineOrder)}")
    sp, b = _cuda.basis(kt, int(splineOrder), ut, int(derivativeOrder), bool(taylorCoefs), spans_in)
    if on_device:
        return sp, b
    kdt = np.asarray(knots).dtype if not isinstance(knots, torch.Tensor) else np.dtype(np.float64)
    b = b.cpu().numpy()
    if np.issubdtype(kdt, np.floating) and kdt != np.float64:
        b = b.astype(kdt)
    return sp.cpu().numpy(), b


def domain(self):
    """``(nInd, 2)`` array of parameter bounds (reference ``:135-138``).  Pure indexing."""
    return np.array([[self.knots[i][self.order[i] - 1], self.knots[i][self.nCoef[i]]] for i in range(self.nInd)])


def _single_point(self, uvw):
    uvw = np.atleast_1d(uvw)
    if len(uvw) != self.nInd:
        raise ValueError(f"Incorrect number of parameter values: {len(uvw)}")
    box = domain(self)
    for i in range(self.nInd):
        if uvw[i] < box[i][0] or uvw[i] > box[i][1]:
            raise ValueError(f"Spline evaluation outside domain: {uvw}")
    return uvw


def _launch_single(self, uvw, **request):
    ds = device_spline(self)
    pts = torch.from_numpy(np.ascontiguousarray(uvw, dtype=np.float64).reshape(1, self.nInd)).to(ds.device)
    return _cuda.eval_points(ds, pts, self.nInd, 1, 1, **request)


def evaluate(self, uvw):
    """Value at one point, ``ndarray (nDep,)`` (reference ``:140-164``)."""
    uvw = _single_point(self, uvw)
    if self.nInd == 0:
        return np.array(self.coefs)
    out = _launch_single(self, uvw, values=True)
    return out["values"][:, 0].cpu().numpy().astype(_result_dtype(self), copy=False)


def derivative(self, with_respect_to, uvw):
    """Mixed partial at one point, ``ndarray (nDep,)`` (reference ``:109-133``)."""
    uvw = _single_point(self, uvw)
    if self.nInd == 0:
        return np.array(self.coefs)
    wrt = [int(with_respect_to[i]) for i in range(self.nInd)]
    out = _launch_single(self, uvw, values=False, wrt=wrt)
    return out["derivative"][:, 0].cpu().numpy().astype(_result_dtype(self), copy=False)


def jacobian(self, uvw):
    """``(nDep, nInd)`` matrix of first partials at one point (reference ``:205-213``); one fused
    launch instead of nInd derivative calls."""
    uvw = _single_point(self, uvw)
    if self.nInd == 0:
        return np.empty((self.nDep, 0), np.asarray(self.coefs).dtype)
    out = _launch_single(self, uvw, values=False, jacobian=True)
    return out["jacobian"][:, :, 0].cpu().numpy().astype(np.asarray(self.coefs).dtype, copy=False)


def _normal_request(self, indices):
    if abs(self.nInd - self.nDep) != 1:
        raise ValueError("The number of independent variables must be one different than the number of dependent variables.")
    D = max(self.nInd, self.nDep)
    if indices is None:
        return None, 0
    idx = [int(i) for i in indices]
    if any(i < 0 or i >= D for i in idx):
        raise IndexError(f"normal index out of range for a normal of length {D}")
    mask = 0
    for i in idx:
        mask |= 1 << i
    if len(set(idx)) != len(idx):
        # repeated indices (the reference allows them, bspy/_spline_evaluation.py:234-240, and counts a repeated component
        # once per repetition in the norm): the kernels normalise over a component SET, so such requests take the raw
        # normal from the kernel and are rescaled by _normalize_selected below
        return _Repeated(idx), mask
    return idx, mask


class _Repeated(list):
    """marker: an ``indices`` list with repetitions"""


def _normalize_selected(raw, idx):
    """raw normals (D, N) -> components ``idx`` divided by the 2-norm of that (possibly repeating) selection"""
    sel = raw[list(idx)]
    if isinstance(sel, torch.Tensor):
        return sel / torch.sqrt((sel * sel).sum(dim=0))
    with np.errstate(all="ignore"):
        return sel / np.sqrt((sel * sel).sum(axis=0))


def normal(self, uvw, normalize=True, indices=None):
    """Cofactor normal at one point (reference ``:215-246``): ``ndarray (max(nInd,nDep),)`` or
    ``(len(indices),)``; the norm is taken over the selected components."""
    uvw = np.atleast_1d(uvw)
    idx, mask = _normal_request(self, indices)
    uvw = _single_point(self, uvw)
    repeated = isinstance(idx, _Repeated)
    out = _launch_single(self, uvw, values=False, normal=True, normalize=bool(normalize) and not repeated, normal_mask=mask)
    n = out["normal"][:, 0].cpu().numpy()
    if repeated and normalize:
        n = _normalize_selected(n[:, None], idx)[:, 0]
    elif idx is not None:
        n = n[idx]
    dt = np.asarray(self.coefs).dtype if hasattr(self, "coefs") else getattr(self, "coefsDtype", np.float64)
    return n.astype(dt if np.issubdtype(dt, np.floating) else np.float64, copy=False)


def collocation_matrix(knots, splineOrder, uValues, derivativeOrders=None, dev=None):
    """Dense collocation matrix ``A`` (N, nCoef) with ``A[r, ix-order:ix] = bspline_values(None, knots, order, u[r], d[r])``
    -- the rows ``Spline.least_squares`` assembles one Python call at a time (reference ``bspy/_spline_fitting.py:736-750``;
    also ``contour``, ``:190-219``).  ``derivativeOrders=None`` follows ``least_squares``: a parameter equal to its
    predecessor raises the derivative order by one (Hermite rows), otherwise values.  Returns ``(spans, A)`` as numpy
    arrays, or as CUDA tensors when ``uValues`` is a CUDA tensor.  Basis values are bit-identical to the reference."""
    on_device = isinstance(uValues, torch.Tensor) and uValues.is_cuda
    dev = uValues.device if on_device else _cuda.device(dev)
    kt = knots if (isinstance(knots, torch.Tensor) and knots.is_cuda) else \
        torch.from_numpy(np.ascontiguousarray(knots, dtype=np.float64)).to(dev)
    ut = uValues.contiguous().to(torch.float64) if on_device else \
        torch.from_numpy(np.ascontiguousarray(uValues, dtype=np.float64).reshape(-1)).to(dev)
    if derivativeOrders is None:
        uh = ut.cpu().numpy() if on_device else np.ascontiguousarray(uValues, dtype=np.float64).reshape(-1)
        d = np.zeros(uh.shape[0], np.int32)                      # run lengths of equal parameters (index logic only)
        for r in range(1, uh.shape[0]):
            if uh[r] == uh[r - 1]:
                d[r] = d[r - 1] + 1
    else:
        d = np.ascontiguousarray(derivativeOrders.cpu().numpy() if isinstance(derivativeOrders, torch.Tensor) else derivativeOrders,
                                 dtype=np.int32).reshape(-1)
    sp, A = _cuda.collocation(kt, int(splineOrder), ut, torch.from_numpy(d).to(dev) if d.any() else None)
    if on_device:
        return sp, A
    return sp.cpu().numpy(), A.cpu().numpy()


def contract(self, uvw):
    """``Spline.contract`` (reference ``bspy/_spline_operations.py:184-223``): fix the independent variables whose entry
    of ``uvw`` is not ``None``; returns a spline in the remaining variables (``self`` when nothing is fixed).  The basis
    values come from ``bspy_cuda_basis`` (bit-identical to the reference), every fixed variable is one
    ``bspy_cuda_contract_axis`` launch over the coefficient array."""
    box = domain(self)
    fixed = []
    for iv in range(self.nInd):
        if uvw[iv] is not None:
            if uvw[iv] < box[iv][0] or uvw[iv] > box[iv][1]:
                raise ValueError(f"Spline evaluation outside domain: {uvw}")
            fixed.append(iv)
    if not fixed:
        return self
    ds = device_spline(self)
    dev = ds.device
    coefs = ds.coefs                                            # (nDep, *nCoef) contiguous float64 on the device
    removed = 0
    for iv in fixed:
        u = torch.tensor([float(uvw[iv])], dtype=torch.float64, device=dev)
        _, b = _cuda.basis(ds.knots[iv], int(self.order[iv]), u, 0, False, None)
        # the span on the host, as the reference finds it (np.searchsorted 'right' + clamp, _spline_evaluation.py:7-8; the kernel's
        # own span is bit-exact with it): no device read-back between the launches of the fixed variables
        kn = np.asarray(self.knots[iv], dtype=np.float64)
        ix = int(min(max(np.searchsorted(kn, float(uvw[iv]), side="right"), self.order[iv]), len(kn) - self.order[iv]))
        coefs = _cuda.contract_axis(coefs, 1 + iv - removed, ix - self.order[iv], self.order[iv], b.reshape(-1))
        removed += 1
    keep = [iv for iv in range(self.nInd) if uvw[iv] is None]
    out = coefs.cpu().numpy().astype(np.asarray(self.coefs).dtype, copy=False)
    return type(self)(len(keep), self.nDep, [self.order[i] for i in keep], [self.nCoef[i] for i in keep],
                      [self.knots[i] for i in keep], out, self.metadata)


# ----------------------------------------------------------------------------- vectorised entries

def _classify(points):
    if isinstance(points, torch.Tensor):
        return "cuda" if points.is_cuda else "cpu_tensor"
    return "numpy"


def _raise_outside(self, flag_value, fetch_point):
    if flag_value >= 0:
        raise ValueError(f"Spline evaluation outside domain: {fetch_point(int(flag_value))}")


def _record_views(ds, records, jacobian, normal, idx):
    """values / jacobian / normal as strided views into the (N, stride) record array."""
    nDep, nInd = ds.nDep, ds.nInd
    rt = records.T if isinstance(records, torch.Tensor) else records.T            # (stride, N) view
    vals = rt[:nDep]
    jac = rt[nDep:nDep + nDep * nInd].reshape(nDep, nInd, -1) if (jacobian or normal) else None
    nrm = None
    if normal:
        at = nDep + nDep * nInd
        nrm = rt[at:at + ds.normal_dim]
        if idx is not None:
            nrm = nrm[idx]
    return vals, (jac if jacobian else None), nrm


def evaluate_points(self, uvw, *, with_respect_to=None, values=True, jacobian=False, normal=False, normalize=True,
                    indices=None, spans=False, layout="points", check_domain=True, device=None, out_layout="soa") -> EvalResult:
    """Evaluate the spline at N points in one go (the vectorised entry point the reference lacks).

    uvw : (N, nInd) array-like, numpy array, CPU tensor or CUDA tensor of float64
          (``layout="variables"``: (nInd, N); for curves a flat (N,) is accepted).
    with_respect_to : optional derivative multi-index -> ``derivative`` (nDep, N)
    values / jacobian / normal / spans : which outputs to produce (one fused pass over the window)
    normalize, indices : as in ``Spline.normal``
    check_domain : raise ``ValueError("Spline evaluation outside domain: ...")`` like the reference
          when any point is outside the closed domain (costs one device->host flag read); ``"defer"`` (CUDA inputs):
          the flag is written on the device and returned as ``first_outside`` without a synchronisation
    out_layout : ``"soa"`` (default) struct-of-arrays outputs; ``"aos"``: one record [values | jacobian | normal] per
          point (``records`` (N, stride), the per-point shape of the reference's returns) with values / jacobian /
          normal as views into it -- big scattered batches are then written straight to their original positions by
          the cell-sorted kernels, with no un-permute pass over the outputs

    Returns an ``EvalResult`` of outputs of the same kind as ``uvw``."""
    kind = _classify(uvw)
    idx, mask = (None, 0)
    if normal:
        idx, mask = _normal_request(self, indices)
    wrt = None if with_respect_to is None else [int(with_respect_to[i]) for i in range(self.nInd)]
    if layout not in ("points", "variables"):
        raise ValueError("layout must be 'points' or 'variables'")
    if out_layout not in ("soa", "aos"):
        raise ValueError("out_layout must be 'soa' or 'aos'")
    aos = out_layout == "aos"
    repeated = isinstance(idx, _Repeated)            # indices with repetitions: raw normals from the kernel, rescaled here
    unit = bool(normalize) and not repeated
    if aos and wrt is not None:
        raise ValueError("out_layout='aos' holds values, jacobian and normal; request with_respect_to with out_layout='soa'")
    defer = isinstance(check_domain, str) and check_domain == "defer"
    if defer and kind != "cuda":
        raise ValueError("check_domain='defer' needs CUDA inputs (host inputs are checked when the results are copied back)")
    request = dict(wrt=wrt, values=bool(values), jacobian=bool(jacobian), normal=bool(normal), normalize=unit,
                   normal_mask=mask, spans=bool(spans))

    def pick(nrm):
        if nrm is None or idx is None:
            return nrm
        return _normalize_selected(nrm, idx) if (repeated and normalize) else nrm[list(idx)]

    if kind == "cuda":
        pts = uvw if uvw.dtype == torch.float64 else uvw.to(torch.float64)
        if pts.dim() == 1 and self.nInd == 1:
            pts = pts.reshape(-1, 1) if layout == "points" else pts.reshape(1, -1)
        if pts.dim() != 2 or pts.shape[1 if layout == "points" else 0] != self.nInd:
            raise ValueError(f"Incorrect number of parameter values: {pts.shape}")
        N = pts.shape[0] if layout == "points" else pts.shape[1]
        ps, vs = (pts.stride(0), pts.stride(1)) if layout == "points" else (pts.stride(1), pts.stride(0))
        ds = device_spline(self, pts.device)
        flag = _cuda.new_flag(pts.device) if check_domain else None
        if aos:
            rec, sp = _cuda.eval_points_aos(ds, pts, ps, vs, N, jacobian=bool(jacobian), normal=bool(normal), normalize=unit,
                                            normal_mask=mask, spans=bool(spans), flag=flag)
            v, j, nrm = _record_views(ds, rec, jacobian, normal, None)
            res = EvalResult(v, None, j, pick(nrm), sp, records=rec)
        else:
            out = _cuda.eval_points(ds, pts, ps, vs, N, flag=flag, **request)
            res = EvalResult(out["values"], out["derivative"], out["jacobian"], pick(out["normal"]), out["spans"])
        if defer:
            res.first_outside = flag
        elif check_domain:
            _raise_outside(self, int(flag.item()),
                           lambda p: (pts[p] if layout == "points" else pts[:, p]).cpu().numpy())
        return res

    # ---- host input: chunked H2D -> kernel -> D2H pipeline (bspy_b200._cuda.eval_points_host) ----
    if kind == "numpy":
        host = torch.from_numpy(np.ascontiguousarray(np.asarray(uvw, dtype=np.float64)))
    else:
        host = uvw.to(torch.float64).contiguous()
    if host.dim() == 1 and self.nInd == 1:
        host = host.reshape(-1, 1) if layout == "points" else host.reshape(1, -1)
    if host.dim() != 2 or host.shape[1 if layout == "points" else 0] != self.nInd:
        raise ValueError(f"Incorrect number of parameter values: {tuple(host.shape)}")
    ds = device_spline(self, device)
    res, first = _cuda.eval_points_host(ds, host, layout, check=bool(check_domain), aos=aos, **request)
    if first >= 0:
        raise ValueError(f"Spline evaluation outside domain: {(host[first] if layout == 'points' else host[:, first]).numpy()}")
    if aos:
        rec = res["records"].numpy() if kind == "numpy" else res["records"]
        sp = res["spans"]
        if sp is not None and kind == "numpy":
            sp = sp.numpy()
        v, j, nrm = _record_views(ds, rec, jacobian, normal, None)
        return EvalResult(v, None, j, pick(nrm), sp, records=rec)
    if kind == "numpy":
        res = {k: (None if v is None else v.numpy()) for k, v in res.items()}
    return EvalResult(res["values"], res["derivative"], res["jacobian"], pick(res["normal"]), res["spans"])


def evaluate_grid(self, *axes, values=True, jacobian=False, normal=False, normalize=True, indices=None,
                  check_domain=True, device=None, dtype=None) -> EvalResult:
    """Evaluate on the tensor grid ``axes[0] x axes[1] x ...`` (one 1-D axis per independent
    variable).  Outputs have the grid shape in place of N, last variable fastest: values
    ``(nDep, n_0, .., n_last)``, jacobian ``(nDep, nInd, n_0, ..)``, normal ``(D, n_0, ..)`` -- the same
    numbers as ``spline(*np.meshgrid(*axes, indexing="ij"))`` in the reference.  Surfaces run on the
    FP64 tensor pipe.  ``dtype=np.float32`` (surfaces only): float32 outputs, computed in float64 and rounded on the
    store -- the tessellation buffers of the reference's viewer (``bspy/splineOpenGLFrame.py:1461-1513``)."""
    f32 = dtype is not None and np.dtype(dtype) == np.float32
    if dtype is not None and not f32 and np.dtype(dtype) != np.float64:
        raise ValueError("dtype must be float64 (default) or float32")
    if f32 and self.nInd != 2:
        raise NotImplementedError("float32 grid outputs are available for surfaces (nInd == 2)")
    if len(axes) == 1 and self.nInd != 1 and not np.isscalar(axes[0]) and len(axes[0]) == self.nInd \
            and not isinstance(axes[0], (np.ndarray, torch.Tensor)):
        axes = tuple(axes[0])
    if len(axes) != self.nInd:
        raise ValueError(f"Incorrect number of parameter values: {len(axes)}")
    kinds = {_classify(a) for a in axes}
    on_device = kinds == {"cuda"}
    idx, mask = (None, 0)
    if normal:
        idx, mask = _normal_request(self, indices)
        if isinstance(idx, _Repeated):
            raise NotImplementedError("repeated normal indices: use evaluate_points / normal (grids normalise over a component set)")
    dev = axes[0].device if on_device else _cuda.device(device)
    ds = device_spline(self, dev)
    d_axes = []
    for a in axes:
        if isinstance(a, torch.Tensor):
            d_axes.append(a.to(device=dev, dtype=torch.float64).contiguous().reshape(-1))
        else:
            d_axes.append(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(dev))
    flag = _cuda.new_flag(dev) if check_domain else None
    out = _cuda.eval_grid(ds, d_axes, values=bool(values), jacobian=bool(jacobian), normal=bool(normal),
                          normalize=bool(normalize), normal_mask=mask, flag=flag, **({"out_f32": True} if f32 else {}))
    defer = isinstance(check_domain, str) and check_domain == "defer"
    if check_domain and not defer:
        off = int(flag.item())
        if off >= 0:
            shape = [int(a.numel()) for a in d_axes]
            multi = np.unravel_index(off, shape)
            raise ValueError(f"Spline evaluation outside domain: {np.array([float(d_axes[i][multi[i]]) for i in range(self.nInd)])}")
    nrm = out["normal"]
    if nrm is not None and idx is not None:
        nrm = nrm[idx]
    res = EvalResult(out["values"], None, out["jacobian"], nrm, None, first_outside=flag if defer else None)
    if not on_device:
        # results go back through pinned staging buffers (recycled by torch's host allocator), all copies in flight together
        host = {}
        for name in ("values", "jacobian", "normal"):
            t = getattr(res, name)
            host[name] = None if t is None else \
                (torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t, non_blocking=True) if t.is_cuda else t)
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()
        conv = (lambda t: None if t is None else t.numpy()) if kinds <= {"numpy"} else (lambda t: t)
        res = EvalResult(conv(host["values"]), None, conv(host["jacobian"]), conv(host["normal"]), None)
    return res


# ----------------------------------------------------------------------------- curvature

def curvature_points(self, uvw, check_domain=True, device=None):
    """Curvature at N points (reference ``bspy/_spline_evaluation.py:80-107``, one point per call there):
    curves (nInd == 1): signed curvature for planar curves, unsigned otherwise; surfaces (nInd == 2, nDep == 3):
    Gaussian curvature; nDep == 1 is treated as the graph of the function, like the reference's ``self.graph()``.
    ``uvw``: (N, nInd) numpy / CUDA tensor (flat (N,) for curves); returns (N,) of the same kind."""
    if self.nInd not in (1, 2):
        raise ValueError("curvature is defined for curves and surfaces (nInd 1 or 2)")
    graph = self.nDep == 1
    if self.nInd == 2 and not graph and self.nDep != 3:
        raise ValueError("The number of independent variables must be one different than the number of dependent variables.")
    on_device = isinstance(uvw, torch.Tensor) and uvw.is_cuda
    if on_device:
        pts = uvw.to(torch.float64)
    else:
        dev = _cuda.device(device)
        pts = torch.from_numpy(np.ascontiguousarray(np.asarray(uvw, dtype=np.float64))).to(dev)
    pts = pts.reshape(-1, self.nInd)
    if max(self.order) <= _cuda.CURVATURE_MAX_ORDER and self.nDep <= 3:
        # one fused launch: second-derivative basis rows + one window walk + the curvature formula per point
        ds = device_spline(self, pts.device)
        flag = _cuda.new_flag(pts.device) if check_domain else None
        k = _cuda.curvature_points(ds, pts, pts.stride(0), pts.stride(1), pts.shape[0], flag=flag)
        if check_domain:
            _raise_outside(self, int(flag.item()), lambda p: pts[p].cpu().numpy())
        return k if on_device else k.cpu().numpy()
    ev = lambda **kw: evaluate_points(self, pts, values=False, check_domain=kw.pop("check", False), **kw)
    if self.nInd == 1:
        d1 = ev(with_respect_to=[1], check=check_domain).derivative
        d2 = ev(with_respect_to=[2]).derivative
        k = _cuda.curvature(1, self.nDep, graph, d1.contiguous(), d2.contiguous(), None)
    else:
        first = ev(jacobian=True, normal=not graph, check=check_domain)
        d2 = torch.stack([ev(with_respect_to=w).derivative for w in ([2, 0], [1, 1], [0, 2])])
        d1 = first.jacobian.reshape(self.nDep * 2, -1) if not graph else first.jacobian.reshape(2, -1)
        k = _cuda.curvature(2, self.nDep, graph, d1.contiguous(), d2.reshape(-1, d2.shape[-1]).contiguous(),
                            None if graph else first.normal.contiguous())
    return k if on_device else k.cpu().numpy()


def curvature(self, uv):
    """Curvature at one point (``float``), reference ``Spline.curvature``."""
    uv = np.atleast_1d(np.asarray(uv, dtype=np.float64))
    _single_point(self, uv)
    return float(curvature_points(self, uv.reshape(1, -1), check_domain=False)[0])
