"""``SplineBatch``: many splines of one shape packed for the device.

The reference evaluates a list of splines with a Python loop (``[s(u) for s in splines]``, or the
viewer's per-patch loop in ``examples/teapot.py:350-359``); here the list is one packed tensor per
attribute -- ``knots[i]``: ``(S, order[i]+nCoef[i])`` (or 1-D when all splines share the knots),
``coefs``: ``(S, nDep, *nCoef)`` -- and one launch evaluates all of them:

* curves (``nInd == 1``): ``evaluate(u)`` with ``u`` of shape ``(S, nPts)`` -> ``(S, nDep, nPts)``
  (``bspy_cuda_eval_many``: one warp per curve, knots/coefficients staged in shared memory);
* surfaces (``nInd == 2``): ``evaluate_grid(uAxis, vAxis)`` -> ``(S, nDep, nU, nV)`` (+ jacobian,
  normal) on the FP64 tensor pipe (``bspy_cuda_eval_grid_batch``).

``shard(rank, world)`` gives the contiguous slice of splines a rank owns: batches are sharded across
GPUs by splitting the spline index range, no collective on the data path.
"""
from __future__ import annotations

import numpy as np
import torch

from bspy_b200 import _cuda
from bspy_b200._spline_evaluation import EvalResult, _normal_request
from bspy_b200.sharding import shard_range


def _to_device(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)


class SplineBatch:
    def __init__(self, nInd, nDep, order, nCoef, knots, coefs, metadata=None, device=None):
        self.nInd, self.nDep = int(nInd), int(nDep)
        self.order = tuple(int(o) for o in order)
        self.nCoef = tuple(int(n) for n in nCoef)
        if len(self.order) != self.nInd or len(self.nCoef) != self.nInd or len(knots) != self.nInd:
            raise ValueError("order, nCoef and knots must have nInd entries")
        dev = coefs.device if isinstance(coefs, torch.Tensor) and coefs.is_cuda else _cuda.device(device)
        self.device = dev
        self.coefs = _to_device(coefs, dev)
        if self.coefs.dim() != 2 + self.nInd or tuple(self.coefs.shape[1:]) != (self.nDep, *self.nCoef):
            raise ValueError(f"coefs must have shape (S, nDep, *nCoef) = (S, {self.nDep}, {self.nCoef})")
        self.nSplines = int(self.coefs.shape[0])
        self.knots = []
        for i, k in enumerate(knots):
            k = _to_device(k, dev)
            want = self.order[i] + self.nCoef[i]
            if k.dim() == 1 and k.shape[0] == want:
                pass                                    # shared by all splines
            elif k.dim() == 2 and tuple(k.shape) == (self.nSplines, want):
                pass
            else:
                raise ValueError(f"Knots array for variable {i} should have length {want}")
            self.knots.append(k)
        self.metadata = dict(metadata or {})

    # ---- construction helpers --------------------------------------------------------------
    @classmethod
    def from_splines(cls, splines, device=None):
        """Pack a list of ``Spline`` objects of identical nInd / nDep / order / nCoef."""
        first = splines[0]
        for s in splines:
            if (s.nInd, s.nDep, s.order, s.nCoef) != (first.nInd, first.nDep, first.order, first.nCoef):
                raise ValueError("all splines of a batch must share nInd, nDep, order and nCoef")
        coefs = np.stack([np.ascontiguousarray(s.coefs, dtype=np.float64) for s in splines])
        knots = []
        for i in range(first.nInd):
            kk = np.stack([np.asarray(s.knots[i], dtype=np.float64) for s in splines])
            knots.append(kk[0] if np.all(kk == kk[0]) else kk)
        negate = {bool(s.metadata.get("negateNormal", False)) for s in splines}
        if len(negate) > 1:
            raise ValueError("all splines of a batch must agree on metadata['negateNormal']")
        return cls(first.nInd, first.nDep, first.order, first.nCoef, knots, coefs,
                   {"negateNormal": True} if negate == {True} else {}, device)

    @classmethod
    def load(cls, fileName, device=None):
        """Bulk loader: every spline stored in a BSpy json file -- a single ``Spline`` dict, a list of
        them (``Spline.save`` / ``Spline.load``, reference ``bspy/spline.py:1542-1583, 1998-2026``), or splines nested as
        the ``manifold`` of a ``Solid``'s boundaries (``tests/teapots.json``) -- packed straight into device batches, one
        per shape ``(nInd, nDep, order, nCoef, negateNormal)``.  Returns a list of batches in order of first appearance;
        ``batch.indices`` holds the positions (file order) of its splines."""
        import json

        from bspy_b200.spline import Spline
        with open(fileName, "r", encoding="utf-8") as f:
            data = json.load(f)
        found = []

        def walk(node):
            if isinstance(node, dict):
                if node.get("type") == "Spline" and "coefs" in node and "knots" in node:
                    found.append(Spline.from_dict(node))
                    return
                for v in node.values():
                    walk(v)
            elif isinstance(node, list):
                for v in node:
                    walk(v)

        walk(data)
        groups = {}
        for i, s in enumerate(found):
            key = (s.nInd, s.nDep, tuple(s.order), tuple(s.nCoef), bool(s.metadata.get("negateNormal", False)))
            groups.setdefault(key, []).append(i)
        batches = []
        for idx in groups.values():
            b = cls.from_splines([found[i] for i in idx], device)
            b.indices = list(idx)
            batches.append(b)
        return batches

    def __len__(self):
        return self.nSplines

    def spline(self, i):
        """Host-side ``Spline`` copy of element ``i``."""
        from bspy_b200.spline import Spline
        knots = [(k if k.dim() == 1 else k[i]).cpu().numpy() for k in self.knots]
        return Spline(self.nInd, self.nDep, self.order, self.nCoef, knots, self.coefs[i].cpu().numpy(), self.metadata)

    def shard(self, rank, world):
        """The contiguous slice of splines owned by ``rank`` of ``world`` (views, no copy)."""
        lo, hi = shard_range(self.nSplines, rank, world)
        knots = [k if k.dim() == 1 else k[lo:hi] for k in self.knots]
        return SplineBatch(self.nInd, self.nDep, self.order, self.nCoef, knots, self.coefs[lo:hi], self.metadata)

    def _descriptor(self):
        first_knots = [k if k.dim() == 1 else k[0] for k in self.knots]
        sign = -1 if self.metadata.get("negateNormal", False) else 1
        return _cuda.DeviceSpline(self.nInd, self.nDep, self.order, self.nCoef, first_knots, self.coefs[0], sign)

    # ---- curves ------------------------------------------------------------------------------
    def evaluate(self, u, derivative=False, check_domain=True, out=None) -> EvalResult:
        """Curves only: ``u`` is ``(S, nPts)`` (numpy or CUDA tensor); returns values
        ``(S, nDep, nPts)`` and, with ``derivative=True``, first derivatives of the same shape."""
        if self.nInd != 1:
            raise NotImplementedError("SplineBatch.evaluate handles curves (nInd == 1); use evaluate_grid for surfaces")
        on_device = isinstance(u, torch.Tensor) and u.is_cuda
        ut = _to_device(u, self.device)
        if ut.dim() != 2 or ut.shape[0] != self.nSplines:
            raise ValueError(f"u must have shape (nSplines, nPts) = ({self.nSplines}, nPts)")
        nPts = int(ut.shape[1])
        if self.nSplines == 0 or nPts == 0:                    # a rank whose shard is empty: empty outputs, no launch
            shape = (self.nSplines, self.nDep, nPts)
            vals = torch.empty(shape, dtype=torch.float64, device=self.device)
            der = torch.empty(shape, dtype=torch.float64, device=self.device) if derivative else None
            if not on_device:
                vals, der = vals.cpu().numpy(), None if der is None else der.cpu().numpy()
            return EvalResult(values=vals, derivative=der)
        k = self.knots[0]
        if k.dim() == 1:
            k = k.unsqueeze(0).expand(self.nSplines, -1)       # stride 0: every warp stages the same knots
        flag = _cuda.new_flag(self.device) if check_domain else None
        coefs = self.coefs.reshape(self.nSplines, self.nDep, self.nCoef[0])
        table = self._curve_images(k, coefs)
        if self.order[0] > _cuda.MANY_MAX_ORDER:
            res = self._evaluate_per_spline(k, coefs, ut, derivative, flag, out)
        elif table is not None:
            res = _cuda.eval_many_tab(self.order[0], self.nCoef[0], self.nDep, _ExpandedKnots(k), coefs, table, ut, deriv1=derivative,
                                      flag=flag, out=out)
        else:
            res = _cuda.eval_many(self.order[0], self.nCoef[0], self.nDep, _ExpandedKnots(k), coefs, ut, deriv1=derivative,
                                  flag=flag, out=out)
        defer = isinstance(check_domain, str) and check_domain == "defer"
        if check_domain and not defer:
            off = int(flag.item())
            if off >= 0:
                s, p = divmod(off, ut.shape[1])
                raise ValueError(f"Spline evaluation outside domain: spline {s}, u = {float(ut[s, p])}")
        vals, der = res["values"], res.get("derivative")
        if not on_device:
            vals = vals.cpu().numpy()
            der = None if der is None else der.cpu().numpy()
        return EvalResult(values=vals, derivative=der, first_outside=flag if defer else None)

    def _curve_images(self, k, coefs):
        """Cached per-curve images of a batch of curves (values and first derivatives): built by the first evaluation on the device the
        batch lives on and kept with the batch -- like the knots and coefficients they are derived from, a batch is immutable
        once uploaded.  They cost 3-4x the memory of the raw curves and one pass to build, so they are made by the SECOND
        evaluation of a batch; ``batch.cache_tables = False`` turns them off, ``= "now"`` builds them at once."""
        if not getattr(self, "cache_tables", True) or self.nSplines < 1024 or not coefs.is_cuda:
            return None
        hit = self.__dict__.get("_curve_images_cache")
        if hit is None:
            # a batch that is evaluated once never pays for the build: the images are made by the second call
            seen = self.__dict__.get("_curve_images_calls", 0)
            self.__dict__["_curve_images_calls"] = seen + 1
            if seen < 1 and getattr(self, "cache_tables", True) != "now":
                return None
            hit = (_cuda.many_table(self.order[0], self.nCoef[0], self.nDep, _ExpandedKnots(k), coefs),)
            self.__dict__["_curve_images_cache"] = hit
        return hit[0]

    def _evaluate_per_spline(self, k, coefs, ut, derivative, flag, out):
        """Curves whose order exceeds the warp-per-curve kernel's limit: one scattered-point launch per curve
        (any order up to 32), same outputs and the same flat first-outside index."""
        S, nPts = self.nSplines, int(ut.shape[1])
        if out is None:
            out = {"values": torch.empty((S, self.nDep, nPts), dtype=torch.float64, device=self.device),
                   "derivative": torch.empty((S, self.nDep, nPts), dtype=torch.float64, device=self.device) if derivative else None}
        for s in range(S):
            ds = _cuda.DeviceSpline(1, self.nDep, self.order, self.nCoef, [k[s].contiguous()], coefs[s].contiguous())
            one = _cuda.new_flag(self.device) if flag is not None else None
            r = _cuda.eval_points(ds, ut[s], 1, 1, nPts, wrt=[1] if derivative else None, values=True, flag=one)
            out["values"][s].copy_(r["values"])
            if derivative:
                out["derivative"][s].copy_(r["derivative"])
            if flag is not None and int(flag.item()) < 0 and int(one.item()) >= 0:
                flag.fill_(s * nPts + int(one.item()))
        return out

    # ---- surfaces ----------------------------------------------------------------------------
    def evaluate_grid(self, uAxis, vAxis, values=True, jacobian=False, normal=False, normalize=True, indices=None,
                      check_domain=True, out=None, dtype=None) -> EvalResult:
        """Surfaces only: every spline of the batch on the grid ``uAxis x vAxis``; outputs
        ``(S, nDep, nU, nV)``, ``(S, nDep, 2, nU, nV)``, ``(S, D, nU, nV)`` (v fastest).  ``dtype=np.float32``: float32
        outputs (device tensors, or ``out`` buffers of that type), computed in float64 and rounded on the store."""
        if self.nInd != 2:
            raise NotImplementedError("SplineBatch.evaluate_grid handles surfaces (nInd == 2)")
        f32 = dtype is not None and np.dtype(dtype) == np.float32
        if dtype is not None and not f32 and np.dtype(dtype) != np.float64:
            raise ValueError("dtype must be float64 (default) or float32")
        idx, mask = (None, 0)
        if normal:
            idx, mask = _normal_request(self, indices)
            if type(idx).__name__ == "_Repeated":
                raise NotImplementedError("repeated normal indices: use Spline.evaluate_points / Spline.normal")
        on_device = all(isinstance(a, torch.Tensor) and a.is_cuda for a in (uAxis, vAxis))
        axes = [_to_device(uAxis, self.device).reshape(-1), _to_device(vAxis, self.device).reshape(-1)]
        strides = [0 if k.dim() == 1 else int(k.stride(0)) for k in self.knots]
        request = dict(values=values, jacobian=jacobian, normal=normal, normalize=normalize, normal_mask=mask)
        defer = isinstance(check_domain, str) and check_domain == "defer"
        flag = None
        if self.nSplines == 0 or axes[0].numel() == 0 or axes[1].numel() == 0:
            # a rank whose shard is empty (world > nSplines): empty outputs of the right shape, no launch
            nU, nV = int(axes[0].numel()), int(axes[1].numel())
            odt, S, D = (torch.float32 if f32 else torch.float64), self.nSplines, max(self.nInd, self.nDep)
            mk = lambda *shape: torch.empty(shape, dtype=odt, device=self.device)
            r = EvalResult(values=mk(S, self.nDep, nU, nV) if values else None,
                           jacobian=mk(S, self.nDep, 2, nU, nV) if jacobian else None,
                           normal=mk(S, D if idx is None else len(idx), nU, nV) if normal else None)
            if not on_device and out is None and not f32:
                conv = lambda t: None if t is None else t.cpu().numpy()
                r = EvalResult(values=conv(r.values), jacobian=conv(r.jacobian), normal=conv(r.normal))
            return r
        if on_device or out is not None or f32:
            flag = _cuda.new_flag(self.device) if check_domain else None
            res = _cuda.eval_grid_batch(self._descriptor(), self.nSplines, strides, int(self.coefs.stride(0)), axes,
                                        flag=flag, out=out, **request, **({"out_f32": True} if f32 else {}))
            off = int(flag.item()) if (check_domain and not defer) else -1
        else:
            res, off = _cuda.eval_grid_batch_host(self._descriptor(), self.nSplines, strides, int(self.coefs.stride(0)),
                                                  axes, check=bool(check_domain), **request)
        if off >= 0:
            nU, nV = int(axes[0].numel()), int(axes[1].numel())
            s, rem = divmod(off, nU * nV)
            a, b = divmod(rem, nV)
            raise ValueError(f"Spline evaluation outside domain: spline {s}, uv = [{float(axes[0][a])} {float(axes[1][b])}]")
        nrm = res.get("normal")
        if nrm is not None and idx is not None:
            nrm = nrm[:, idx]
        r = EvalResult(values=res.get("values"), jacobian=res.get("jacobian"), normal=nrm, first_outside=flag if defer else None)
        if not on_device:
            conv = lambda t: None if t is None else (t.numpy() if not t.is_cuda else t.cpu().numpy())
            r = EvalResult(values=conv(r.values), jacobian=conv(r.jacobian), normal=conv(r.normal))
        return r


class _ExpandedKnots:
    """Adapter so that a stride-0 (shared) knot tensor passes the contiguity check of the binding:
    the kernel only needs a base pointer and the per-spline stride."""

    def __init__(self, t):
        self._t = t
        self.dtype, self.device = t.dtype, t.device

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return self._t.data_ptr()

    def stride(self, i):
        return self._t.stride(i)
