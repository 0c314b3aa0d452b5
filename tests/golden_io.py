"""Helpers shared by the tests: load the committed golden fixtures (tests/golden/*.npz)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# the parity bar stated by BASELINE.json north_star
RTOL, ATOL = 1e-12, 1e-13


def close(x, ref, rtol=RTOL, atol=ATOL):
    """|x - ref| <= atol + rtol*|ref| with NaNs (and infs) required to match."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return bool(np.allclose(x, ref, rtol=rtol, atol=atol, equal_nan=True))


def worst(x, ref):
    x = np.asarray(x, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    with np.errstate(all="ignore"):
        e = np.abs(x - ref) - RTOL * np.abs(ref)
    e = e[np.isfinite(e)]
    return float(e.max()) if e.size else 0.0


class Case:
    def __init__(self, meta, arrays):
        self.meta = meta
        self.tag = meta["tag"]
        self.nInd, self.nDep = meta["nInd"], meta["nDep"]
        self.order, self.nCoef = tuple(meta["order"]), tuple(meta["nCoef"])
        self.metadata = meta["metadata"]
        self._a = arrays
        self.knots = [arrays[f"{self.tag}/knots{i}"] for i in range(self.nInd)]
        self.coefs = arrays[f"{self.tag}/coefs"]
        self.uvw = arrays[f"{self.tag}/uvw"]

    def __getitem__(self, key):
        return self._a[f"{self.tag}/{key}"]

    def has(self, key):
        return f"{self.tag}/{key}" in self._a

    def __repr__(self):
        return f"Case({self.tag})"


_cache = {}


def load_cases():
    if "cases" not in _cache:
        arrays = dict(np.load(os.path.join(GOLDEN, "ref_cases.npz")))
        meta = json.load(open(os.path.join(GOLDEN, "ref_cases.json")))
        _cache["cases"] = [Case(m, arrays) for m in meta]
    return _cache["cases"]


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def close_cond(x, ref, scale, k=16.0, rtol=RTOL, atol=ATOL):
    """Condition-aware form of the bar: |x-ref| <= atol + rtol*|ref| + k*eps*sum(|terms|)
    (see oracle/bspy_oracle.py, 'conditioning of the sums').  NaN/inf must match."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.broadcast_to(np.asarray(scale, dtype=np.float64), ref.shape)
    with np.errstate(all="ignore"):
        fin = np.isfinite(ref)
        if not np.array_equal(np.isnan(x), np.isnan(ref)):
            return False
        inf = ~fin & ~np.isnan(ref)
        if not np.array_equal(x[inf], ref[inf]):
            return False
        tol = atol + rtol * np.abs(ref) + k * np.finfo(float).eps * np.where(np.isfinite(scale), scale, np.inf)
        return bool(np.all(np.abs(x - ref)[fin] <= tol[fin]))


def well_conditioned_subset(normal_raw, idx, floor=1e-6):
    """Rows where normalising over the component subset ``idx`` is meaningful: the selected
    components are not rounding noise next to the full normal (or the whole normal is exactly
    zero, the NaN case that must match).  Reference fixtures such as examples/TomsNasty.json
    have flat regions where components 0 and 2 of the normal are ~1e-17 of component 1;
    dividing noise by its own norm is not a parity question."""
    normal_raw = np.asarray(normal_raw)
    full = np.sqrt((normal_raw ** 2).sum(axis=1))
    part = np.sqrt((normal_raw[:, list(idx)] ** 2).sum(axis=1))
    with np.errstate(all="ignore"):
        return (part >= floor * full) | (full == 0.0) | np.isnan(full)


def nondegenerate_normal(normal_raw, scale, floor=1e-9):
    """Rows where the raw normal is more than rounding noise next to the terms it is summed from
    (|n| > floor * max sum|terms|).  At singular points of a surface -- the collapsed control rows
    of the Utah teapot's lid and bottom patches -- the raw normal is an exact or inexact zero
    depending on how the terms happen to cancel, so the reference's unit normal there is either
    NaN (0/0) or an arbitrary unit vector; that is not a parity question.  Elsewhere NaNs must match."""
    normal_raw = np.asarray(normal_raw)
    with np.errstate(all="ignore"):
        n = np.sqrt((normal_raw ** 2).sum(axis=-1))
        return n > floor * np.asarray(scale).max(axis=-1)


def _spline_from(arrays, tag):
    """(nInd, nDep, order, nCoef, knots, coefs) of a spline stored by tests/golden/make_golden_block.py"""
    shape = [int(x) for x in arrays[f"{tag}/shape"]]
    nInd, nDep = shape[0], shape[1]
    order, nCoef = shape[2:2 + nInd], shape[2 + nInd:2 + 2 * nInd]
    return nInd, nDep, tuple(order), tuple(nCoef), [arrays[f"{tag}/knots{i}"] for i in range(nInd)], arrays[f"{tag}/coefs"]


def block_members(arrays, tag):
    """rows of (map, spline-tuple) of block ``tag`` in ref_block.npz"""
    rows, j = [], 0
    for n in arrays[f"block/{tag}/rows"]:
        row = []
        for _ in range(int(n)):
            row.append(([int(i) for i in arrays[f"block/{tag}/member{j}/map"]], _spline_from(arrays, f"block/{tag}/member{j}")))
            j += 1
        rows.append(row)
    return rows
