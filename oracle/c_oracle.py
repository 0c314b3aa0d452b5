"""ctypes binding of oracle/libbspy_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by bspy_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbspy_oracle.so")

VALUES, DERIV, JACOBIAN, NORMAL, NORMALIZE, SPANS = 1, 2, 4, 8, 16, 32
_lib = None


def build(force=False):
    src = os.path.join(HERE, "bspy_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.bspy_oracle_spans.restype = C.c_int
        _lib.bspy_oracle_basis.restype = C.c_int
        _lib.bspy_oracle_eval.restype = C.c_int
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def spans(knots, order, u):
    knots = np.ascontiguousarray(knots, np.float64)
    u = np.ascontiguousarray(u, np.float64).reshape(-1)
    out = np.empty(u.shape[0], np.int32)
    lib().bspy_oracle_spans(_p(knots), C.c_int(len(knots)), C.c_int(order), _p(u), C.c_int64(u.shape[0]), _p(out, C.c_int32))
    return out


def basis(knots, order, u, deriv=0, taylor=False, ix=None):
    knots = np.ascontiguousarray(knots, np.float64)
    u = np.ascontiguousarray(u, np.float64).reshape(-1)
    N = u.shape[0]
    ixo = np.empty(N, np.int32)
    out = np.empty((N, order))
    ixi = None if ix is None else np.ascontiguousarray(ix, np.int32)
    rc = lib().bspy_oracle_basis(_p(knots), C.c_int(len(knots)), C.c_int(order), _p(u), _p(ixi, C.c_int32),
                                 C.c_int64(N), C.c_int(deriv), C.c_int(bool(taylor)), _p(ixo, C.c_int32), _p(out))
    assert rc == 0, rc
    return ixo, out


def evaluate(s, uvw, *, wrt=None, values=True, jacobian=False, normal=False, normalize=True, indices=None, spans=False):
    """``s``: any object with nInd, nDep, order, nCoef, knots, coefs, metadata.  ``uvw[N, nInd]``.
    Returns a dict with AoS arrays: values[N,nDep], deriv[N,nDep], jacobian[N,nDep,nInd], normal[N,D]
    (all D components; the norm is taken over ``indices``), spans[N,nInd], first_oob."""
    uvw = np.ascontiguousarray(uvw, np.float64).reshape(-1, s.nInd)
    N, nInd, nDep = uvw.shape[0], s.nInd, s.nDep
    D = max(nInd, nDep)
    kn = [np.ascontiguousarray(k, np.float64) for k in s.knots]
    kptr = (C.POINTER(C.c_double) * max(nInd, 1))(*[_p(k) for k in kn])
    coefs = np.ascontiguousarray(s.coefs, np.float64)
    order = np.array(s.order, np.int32)
    nCoef = np.array(s.nCoef, np.int32)
    flags = (VALUES if values else 0) | (DERIV if wrt is not None else 0) | (JACOBIAN if jacobian else 0) | \
            (NORMAL if normal else 0) | (NORMALIZE if normalize else 0) | (SPANS if spans else 0)
    out = {}
    v = out["values"] = np.empty((N, nDep)) if values else None
    d = out["deriv"] = np.empty((N, nDep)) if wrt is not None else None
    j = out["jacobian"] = np.empty((N, nDep, nInd)) if jacobian else None
    n = out["normal"] = np.empty((N, D)) if normal else None
    sp = out["spans"] = np.empty((N, nInd), np.int32) if spans else None
    w = None if wrt is None else np.array(wrt, np.int32)
    mask = (1 << D) - 1 if indices is None else sum(1 << int(i) for i in set(indices))
    sign = -1 if getattr(s, "metadata", {}).get("negateNormal", False) else 1
    oob = C.c_int64(-1)
    rc = lib().bspy_oracle_eval(C.c_int(nInd), C.c_int(nDep), _p(order, C.c_int32), _p(nCoef, C.c_int32), kptr, _p(coefs),
                                _p(uvw), C.c_int64(N), _p(w, C.c_int32), C.c_int(flags), C.c_int(sign), C.c_uint32(mask),
                                _p(v), _p(d), _p(j), _p(n), _p(sp, C.c_int32), C.byref(oob))
    if rc != 0:
        raise ValueError(f"bspy_oracle_eval failed: {rc}")
    out["first_oob"] = oob.value
    return out
