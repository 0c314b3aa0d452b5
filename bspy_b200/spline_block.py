"""``SplineBlock``: a block of splines that represents a system of equations -- the evaluation part of the
reference's ``bspy/spline_block.py`` (constructor ``:46-109``, ``_block_evaluation`` ``:37-44``, ``derivative /
evaluate / jacobian / normal`` ``:179-282``) on the CUDA path, plus the batched ``evaluate_points`` the reference
lacks.  Splines in the same row are added together; optional maps route the block's independent variables to
each spline's own.  (Row sums: ``bspy_cuda_block_accumulate``; block normals: ``bspy_cuda_normal_from_jacobian``.)

Out of scope here, as for ``Spline``: ``contours``, ``zeros``, ``split``, ``normal_spline`` and the other methods that
build new splines with adaptive host loops; they only *consume* this path.
"""
from __future__ import annotations

import numpy as np
import torch

from bspy_b200 import _cuda
from bspy_b200 import _spline_evaluation as _ev
from bspy_b200.spline import Spline

__all__ = ["SplineBlock"]


class SplineBlock:
    """A block (list of rows) of splines; each entry a ``Spline`` or a ``(map, Spline)`` tuple.  Same constructor
    rules and error messages as the reference (``bspy/spline_block.py:46-109``)."""

    def __init__(self, block):
        rows = self._as_rows(block)
        self.block, self.nInd, self.nDep, self.size = [], 0, 0, 0
        self.knotsDtype = self.coefsDtype = None
        bounds = {}                                   # block variable -> (lower, upper), first spline that uses it wins
        for entries in rows:
            members, used, depRow, nextDefault = [], set(), 0, 0
            for entry in entries:
                if isinstance(entry, Spline):         # no map given: the row's variables are numbered consecutively
                    spline, varMap = entry, list(range(nextDefault, nextDefault + entry.nInd))
                else:
                    varMap, spline = list(entry[0]), entry[1]
                nextDefault += spline.nInd
                if not members:
                    depRow = spline.nDep
                    if self.coefsDtype is None:
                        self.knotsDtype = np.asarray(spline.knots[0]).dtype
                        self.coefsDtype = np.asarray(spline.coefs).dtype
                elif spline.nDep != depRow:
                    raise ValueError("All splines in the same row must have the same nDep")
                box = spline.domain()
                for own, var in enumerate(varMap):
                    if var in used:
                        raise ValueError(f"Multiple splines in the same row map to independent variable {var}")
                    used.add(var)
                    known = bounds.get(var)
                    if known is None:
                        bounds[var] = box[own]
                    elif known[0] != box[own, 0] or known[1] != box[own, 1]:
                        raise ValueError("Domains of independent variables must match")
                members.append((varMap, spline))
            if depRow > 0:
                self.block.append(members)
                self.nDep += depRow
                self.size += len(entries)
        self.nInd = len(bounds)
        missing = [var for var in range(self.nInd) if var not in bounds]
        if missing:
            raise ValueError(f"Block is missing independent variable {missing[0]}")
        self._domain = np.array([bounds[var] for var in range(self.nInd)], self.knotsDtype)

    @staticmethod
    def _as_rows(block):
        """The accepted spellings (reference ``:56-59``): one spline, one row of splines, or a list of rows."""
        if isinstance(block, Spline):
            return [[block]]
        if isinstance(block[0], Spline) or (len(block) > 1 and isinstance(block[1], Spline)):
            return [block]
        return block

    def __call__(self, uvw):
        return self.evaluate(uvw)

    def __repr__(self):
        return f"SplineBlock({self.block})"

    def domain(self):
        """``(nInd, 2)`` bounds of the block's independent variables (reference ``:201-210``)."""
        return self._domain

    # ---- batched entry (added) ------------------------------------------------------------------------------
    def evaluate_points(self, uvw, *, with_respect_to=None, values=True, jacobian=False, normal=False, normalize=True,
                        indices=None, check_domain=True, device=None) -> _ev.EvalResult:
        """Evaluate the block at N points: ``uvw`` (N, nInd) numpy array or CUDA tensor.  Returns struct-of-arrays
        ``values`` (nDep, N), ``derivative`` (nDep, N) for ``with_respect_to``, ``jacobian`` (nDep, nInd, N),
        ``normal`` (max(nInd, nDep) or len(indices), N), of the same kind as ``uvw``."""
        on_device = isinstance(uvw, torch.Tensor) and uvw.is_cuda
        dev = uvw.device if on_device else _cuda.device(device)
        pts = uvw.to(torch.float64) if on_device else torch.from_numpy(np.ascontiguousarray(np.asarray(uvw, dtype=np.float64))).to(dev)
        if pts.dim() == 1 and self.nInd == 1:
            pts = pts.reshape(-1, 1)
        if pts.dim() != 2 or pts.shape[1] != self.nInd:
            raise ValueError(f"Incorrect number of parameter values: {tuple(pts.shape)}")
        N = pts.shape[0]
        idx, mask = (None, 0)
        if normal:
            idx, mask = _ev._normal_request(self, indices)
        wrt = None if with_respect_to is None else [int(with_respect_to[i]) for i in range(self.nInd)]
        need_jac = bool(jacobian or normal)
        out_v = torch.zeros((self.nDep, N), dtype=torch.float64, device=dev) if values else None
        out_d = torch.zeros((self.nDep, N), dtype=torch.float64, device=dev) if wrt is not None else None
        out_j = torch.zeros((self.nDep, self.nInd, N), dtype=torch.float64, device=dev) if need_jac else None
        flag = _cuda.new_flag(dev) if check_domain else None
        row0 = 0
        for row in self.block:
            nDepRow = row[0][1].nDep
            for map, spline in row:
                ds = _ev.device_spline(spline, dev)
                sub = pts[:, map].contiguous()                     # the block's variables routed to this spline's own
                r = _cuda.eval_points(ds, sub, sub.stride(0), sub.stride(1), N, flag=flag, values=bool(values), jacobian=need_jac,
                                      wrt=None if wrt is None else [wrt[i] for i in map])
                rows = list(range(row0, row0 + nDepRow))
                if values:
                    _cuda.block_accumulate(out_v, r["values"], rows)
                if wrt is not None:
                    _cuda.block_accumulate(out_d, r["derivative"], rows)
                if need_jac:
                    jrows = [(row0 + d) * self.nInd + map[iv] for d in range(nDepRow) for iv in range(spline.nInd)]
                    _cuda.block_accumulate(out_j.reshape(self.nDep * self.nInd, N), r["jacobian"].reshape(nDepRow * spline.nInd, N), jrows)
            row0 += nDepRow
        if check_domain:
            first = int(flag.item())
            if first >= 0:
                raise ValueError(f"Spline evaluation outside domain: {pts[first].cpu().numpy()}")
        out_n = None
        if normal:
            out_n = _cuda.normal_from_jacobian(out_j, self.nDep, self.nInd, 1, bool(normalize), mask)
            if idx is not None:
                out_n = out_n[idx]
        res = _ev.EvalResult(out_v, out_d, out_j if jacobian else None, out_n, None)
        if not on_device:
            res = _ev.EvalResult(*[None if t is None else t.cpu().numpy() for t in (res.values, res.derivative, res.jacobian, res.normal)], None)
        return res

    # ---- reference API: one point ---------------------------------------------------------------------------
    def _point(self, uvw):
        uvw = np.atleast_1d(np.asarray(uvw, dtype=np.float64))
        if len(uvw) != self.nInd:
            raise ValueError(f"Incorrect number of parameter values: {len(uvw)}")
        return uvw.reshape(1, self.nInd)

    def evaluate(self, uvw):
        """Value of the block at one point, ``ndarray (nDep,)`` (reference ``:212-226``)."""
        return self.evaluate_points(self._point(uvw)).values[:, 0].astype(self.coefsDtype, copy=False)

    def derivative(self, with_respect_to, uvw):
        """Mixed partial of the block at one point, ``ndarray (nDep,)`` (reference ``:179-199``)."""
        r = self.evaluate_points(self._point(uvw), values=False, with_respect_to=with_respect_to)
        return r.derivative[:, 0].astype(self.coefsDtype, copy=False)

    def jacobian(self, uvw):
        """``(nDep, nInd)`` jacobian of the block at one point (reference ``:228-245``)."""
        r = self.evaluate_points(self._point(uvw), values=False, jacobian=True)
        return r.jacobian[:, :, 0].astype(self.coefsDtype, copy=False)

    def normal(self, uvw, normalize=True, indices=None):
        """Cofactor normal of the block at one point (reference ``:247-282``); needs ``|nInd - nDep| == 1``."""
        if abs(self.nInd - self.nDep) != 1:
            raise ValueError("The number of independent variables must be one different than the number of dependent variables.")
        r = self.evaluate_points(self._point(uvw), values=False, normal=True, normalize=normalize, indices=indices)
        return r.normal[:, 0].astype(self.coefsDtype, copy=False)

    def contract(self, uvw):
        """Block of the member splines contracted at the given parameter values (``None`` keeps a variable), reference
        ``:143-177``; the remaining variables are renumbered consecutively."""
        kept = [var for var, value in enumerate(uvw) if value is None]
        renumber = {var: position for position, var in enumerate(kept)}
        rows = [[([renumber[var] for var in varMap if var in renumber], spline.contract([uvw[var] for var in varMap]))
                 for varMap, spline in members] for members in self.block]
        return SplineBlock(rows)
