"""bspy_b200._cuda -- ctypes binding of libbspy_cuda.so (include/bspy_cuda.h).

This is the thin layer the north star calls ``bspy/_cuda``: Python host code hands raw device
pointers (taken from torch tensors) and a CUDA stream to hand-written sm_100a kernels through a
C ABI.  torch is used for device memory, streams and copies only.

There is NO CPU fallback: if the shared library is missing it is built with nvcc, and if that is
impossible, or no CUDA device is present when a computation is requested, the call raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading

import torch

MAX_IND = 8
MAX_ORDER = 32
MANY_MAX_ORDER = 8     # bspy_cuda_eval_many (warp per curve)
E_ARG, E_UNSUPPORTED, E_NORMAL_DIMS = -1, -2, -3
NORMALIZE = 1
OUT_F32 = 2
WANT_JACOBIAN = 4
WANT_NORMAL = 8

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSPY_CUDA_LIB") or os.path.join(HERE, "libbspy_cuda.so")   # override: kernel A/B experiments only

# every symbol include/bspy_cuda.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "bspy_cuda_abi_version", "bspy_cuda_last_error_string", "bspy_cuda_launch_count", "bspy_cuda_set_option", "bspy_cuda_copy_2d",
    "bspy_cuda_spans", "bspy_cuda_basis", "bspy_cuda_eval_points", "bspy_cuda_eval_points_binned",
    "bspy_cuda_binned_workspace_bytes", "bspy_cuda_eval_points_aos", "bspy_cuda_aos_workspace_bytes", "bspy_cuda_curve_table_bytes", "bspy_cuda_curve_table_build", "bspy_cuda_eval_grid",
    "bspy_cuda_eval_grid_batch", "bspy_cuda_eval_many", "bspy_cuda_many_table_bytes", "bspy_cuda_many_table_build",
    "bspy_cuda_eval_many_tab", "bspy_cuda_probe_fp64", "bspy_cuda_probe_hbm",
    "bspy_cuda_probe_tiles", "bspy_cuda_curvature", "bspy_cuda_curvature_points", "bspy_cuda_contract_axis", "bspy_cuda_block_accumulate",
    "bspy_cuda_normal_from_jacobian", "bspy_cuda_collocation",
)


class CSpline(C.Structure):
    """``struct bspy_spline`` of include/bspy_cuda.h."""
    _fields_ = [
        ("nInd", C.c_int32), ("nDep", C.c_int32),
        ("order", C.c_int32 * MAX_IND), ("nCoef", C.c_int32 * MAX_IND),
        ("knots", C.c_void_p * MAX_IND), ("coefs", C.c_void_p),
        ("normalSign", C.c_int32), ("reserved", C.c_int32),
        ("curveTable", C.c_void_p), ("curveTableBytes", C.c_int64),
    ]


class CudaPathError(RuntimeError):
    """The CUDA path is unavailable or a kernel launch failed.  Never caught to fall back."""


_lock = threading.Lock()
_lib = None


def library():
    """Load (building first if needed) libbspy_cuda.so.  Raises CudaPathError if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if "BSPY_CUDA_LIB" not in os.environ:
            try:
                from . import build as _build
                _build.build()          # no-op unless the .so is missing or older than its sources; inter-process lock
            except Exception as exc:  # no nvcc, compile error
                if not os.path.exists(LIB_PATH):
                    raise CudaPathError(f"libbspy_cuda.so is missing and could not be built: {exc}") from exc
                raise CudaPathError(f"libbspy_cuda.so is older than its sources and could not be rebuilt: {exc}") from exc
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as exc:
            raise CudaPathError(f"cannot load {LIB_PATH}: {exc}") from exc
        lib.bspy_cuda_abi_version.restype = C.c_int
        lib.bspy_cuda_last_error_string.restype = C.c_char_p
        lib.bspy_cuda_launch_count.restype = C.c_int64
        vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
        sig = {
            "bspy_cuda_set_option": [C.c_char_p, i64, i32],
            "bspy_cuda_copy_2d": [vp, i64, vp, i64, i64, i64, vp],
            "bspy_cuda_spans": [vp, i32, i32, vp, i64, vp, vp],
            "bspy_cuda_basis": [vp, i32, i32, vp, vp, i64, i32, i32, vp, vp, vp],
            "bspy_cuda_eval_points": [C.POINTER(CSpline), vp, i64, i64, i64, C.POINTER(i32), u32, u32, vp, vp, vp, vp, vp, vp, vp],
            "bspy_cuda_eval_points_binned": [C.POINTER(CSpline), vp, i64, i64, i64, C.POINTER(i32), u32, u32, vp, vp, vp, vp, vp, vp, vp, i64, vp],
            "bspy_cuda_curve_table_build": [C.POINTER(CSpline), vp, i64, vp],
            "bspy_cuda_eval_points_aos": [C.POINTER(CSpline), vp, i64, i64, i64, u32, u32, vp, i64, vp, vp, vp, i64, vp],
            "bspy_cuda_eval_grid": [C.POINTER(CSpline), C.POINTER(vp), C.POINTER(i64), u32, u32, vp, vp, vp, vp, vp],
            "bspy_cuda_eval_grid_batch": [C.POINTER(CSpline), i64, C.POINTER(i64), i64, C.POINTER(vp), C.POINTER(i64), u32, u32, vp, vp, vp, vp, vp],
            "bspy_cuda_eval_many": [i32, i32, i32, i64, vp, i64, vp, i64, vp, i32, vp, vp, vp, vp],
            "bspy_cuda_many_table_build": [i32, i32, i32, i64, vp, i64, vp, i64, vp, i64, vp],
            "bspy_cuda_eval_many_tab": [i32, i32, i32, i64, vp, i64, vp, i64, vp, i64, vp, i32, vp, vp, vp, vp],
            "bspy_cuda_probe_fp64": [i32, i32, vp, C.POINTER(C.c_double), vp],
            "bspy_cuda_probe_hbm": [i32, vp, vp, i64, C.POINTER(C.c_double), vp],
            "bspy_cuda_probe_tiles": [vp, i32, i64, i64, i32, i32, C.POINTER(C.c_double), vp],
            "bspy_cuda_curvature": [i32, i32, i32, i64, vp, vp, vp, vp, vp],
            "bspy_cuda_curvature_points": [C.POINTER(CSpline), vp, i64, i64, i64, vp, vp, vp],
            "bspy_cuda_contract_axis": [vp, i64, i64, i64, i32, i32, vp, vp, vp],
            "bspy_cuda_block_accumulate": [vp, i64, vp, i64, i32, C.POINTER(i32), i64, vp],
            "bspy_cuda_normal_from_jacobian": [vp, i32, i32, i64, i32, u32, u32, vp, vp],
            "bspy_cuda_collocation": [vp, i32, i32, vp, vp, i64, vp, vp, i64, vp],
        }
        for name, args in sig.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        lib.bspy_cuda_binned_workspace_bytes.argtypes = [C.POINTER(CSpline), i64]
        lib.bspy_cuda_binned_workspace_bytes.restype = C.c_int64
        lib.bspy_cuda_curve_table_bytes.argtypes = [C.POINTER(CSpline)]
        lib.bspy_cuda_curve_table_bytes.restype = C.c_int64
        lib.bspy_cuda_many_table_bytes.argtypes = [i32, i32, i32, i64]
        lib.bspy_cuda_many_table_bytes.restype = C.c_int64
        lib.bspy_cuda_aos_workspace_bytes.argtypes = [C.POINTER(CSpline), i64]
        lib.bspy_cuda_aos_workspace_bytes.restype = C.c_int64
        if lib.bspy_cuda_abi_version() != 2:
            raise CudaPathError("libbspy_cuda.so ABI version mismatch; rebuild with python -m bspy_b200._cuda.build --force")
        _lib = lib
    return _lib


def launch_count() -> int:
    return int(library().bspy_cuda_launch_count())


def set_option(name: str, value=None):
    """Experiment switch of the library (``bspy_cuda_set_option``): ``value=None`` returns it to its default.
    Names as in the environment variables without the ``BSPY_`` prefix (``BIN_MODE``, ``CURVE_REPL``, ...)."""
    _check(library().bspy_cuda_set_option(name.encode(), 0 if value is None else int(value), 0 if value is None else 1),
           "bspy_cuda_set_option")


def device(index=None) -> torch.device:
    """The CUDA device computations run on.  Raises when there is none (no CPU fallback)."""
    if not torch.cuda.is_available():
        raise CudaPathError("bspy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if isinstance(index, torch.device):
        if index.type != "cuda":
            raise CudaPathError(f"bspy_b200 computes on CUDA devices only, not {index}")
        return index if index.index is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.device("cuda", torch.cuda.current_device() if index is None else int(index))


def _check(rc: int, what: str):
    if rc == 0:
        return
    msg = library().bspy_cuda_last_error_string().decode("utf-8", "replace")
    if rc == E_NORMAL_DIMS:
        raise ValueError(msg)
    if rc == E_ARG:
        raise ValueError(f"{what}: {msg}")
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise CudaPathError(f"{what}: CUDA error {rc}: {msg}")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f64(t, dev):
    assert t.dtype == torch.float64 and t.device == dev and t.is_contiguous(), "expected contiguous float64 device tensor"
    return t


class DeviceSpline:
    """Knots and coefficients of one spline resident on a device + the C struct describing them."""

    def __init__(self, nInd, nDep, order, nCoef, knots, coefs, normal_sign=1):
        self.nInd, self.nDep = int(nInd), int(nDep)
        self.order, self.nCoef = tuple(int(o) for o in order), tuple(int(n) for n in nCoef)
        if self.nInd > MAX_IND:
            raise NotImplementedError(f"nInd {self.nInd} > {MAX_IND} is not supported by the CUDA kernels")
        if any(o > MAX_ORDER for o in self.order):
            raise NotImplementedError(f"order > {MAX_ORDER} is not supported by the CUDA kernels")
        self.knots = list(knots)      # device float64 tensors
        self.coefs = coefs            # device float64 tensor, (nDep, *nCoef) contiguous
        self.device = coefs.device
        self.normal_sign = -1 if normal_sign < 0 else 1
        c = CSpline()
        c.nInd, c.nDep = self.nInd, self.nDep
        for i in range(self.nInd):
            c.order[i], c.nCoef[i] = self.order[i], self.nCoef[i]
            c.knots[i] = self.knots[i].data_ptr()
        c.coefs = self.coefs.data_ptr()
        c.normalSign = self.normal_sign
        self.c = c
        self.curve_table = None

    def build_curve_table(self):
        """Curves: build the span tables of this spline once (bspy_cuda_curve_table_build) and attach them to the C
        struct; big batches then fetch them by TMA instead of rebuilding them in every thread block.  Called by
        ``freeze()`` / the device cache of ``bspy_b200._spline_evaluation`` when the device copy is made."""
        if self.nInd != 1 or not self.coefs.is_cuda:
            return self
        lib = library()
        need = int(lib.bspy_cuda_curve_table_bytes(C.byref(self.c)))
        if need <= 0:
            return self
        table = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.bspy_cuda_curve_table_build(C.byref(self.c), _ptr(table), need, _stream(self.device))
        _check(rc, "bspy_cuda_curve_table_build")
        self.curve_table = table
        self.c.curveTable = table.data_ptr()
        self.c.curveTableBytes = need
        return self

    @property
    def normal_dim(self):
        return max(self.nInd, self.nDep)


# ------------------------------------------------------------------------------ entry points

def spans(knots, order, u):
    """int32 spans for a 1-D device tensor of parameters (bspy_cuda_spans)."""
    dev = u.device
    out = torch.empty(u.numel(), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_spans(_ptr(_f64(knots, dev)), knots.numel(), int(order), _ptr(_f64(u, dev)), u.numel(),
                                       _ptr(out), _stream(dev))
    _check(rc, "bspy_cuda_spans")
    return out


def basis(knots, order, u, deriv=0, taylor=False, spans_in=None):
    """(spans int32[N], basis float64[N, order]) -- bit-exact bspline_values (bspy_cuda_basis)."""
    dev = u.device
    N = u.numel()
    sp = torch.empty(N, dtype=torch.int32, device=dev)
    out = torch.empty((N, int(order)), dtype=torch.float64, device=dev)
    if spans_in is not None:
        assert spans_in.dtype == torch.int32 and spans_in.device == dev and spans_in.is_contiguous()
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_basis(_ptr(_f64(knots, dev)), knots.numel(), int(order), _ptr(_f64(u, dev)), _ptr(spans_in),
                                       N, int(deriv), int(bool(taylor)), _ptr(sp), _ptr(out), _stream(dev))
    _check(rc, "bspy_cuda_basis")
    return sp, out


def new_flag(dev):
    return torch.full((1,), -1, dtype=torch.int64, device=dev)


def eval_points(ds: DeviceSpline, uvw, point_stride, var_stride, N, *, wrt=None, values=True, jacobian=False,
                normal=False, normalize=True, normal_mask=0, spans=False, flag=None, binned=True):
    """Launch bspy_cuda_eval_points.  ``uvw`` is a float64 device tensor addressed through the two
    strides (elements).  Returns a dict of SoA device tensors; ``flag`` (int64[1], -1) receives the
    first out-of-domain index."""
    dev = ds.device
    D = ds.normal_dim
    out = {
        "values": torch.empty((ds.nDep, N), dtype=torch.float64, device=dev) if values else None,
        "derivative": torch.empty((ds.nDep, N), dtype=torch.float64, device=dev) if wrt is not None else None,
        "jacobian": torch.empty((ds.nDep, ds.nInd, N), dtype=torch.float64, device=dev) if jacobian else None,
        "normal": torch.empty((D, N), dtype=torch.float64, device=dev) if normal else None,
        "spans": torch.empty((ds.nInd, N), dtype=torch.int32, device=dev) if spans else None,
    }
    w = None
    if wrt is not None:
        w = (C.c_int32 * max(ds.nInd, 1))(*[int(x) for x in wrt])
    lib = library()
    ws_bytes = int(lib.bspy_cuda_binned_workspace_bytes(C.byref(ds.c), int(N))) if binned else 0
    with torch.cuda.device(dev):
        if ws_bytes > 0:
            # big scattered batch on a spline that lives in L2: cell-binned evaluation (workspace is ours)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            rc = lib.bspy_cuda_eval_points_binned(C.byref(ds.c), _ptr(uvw), int(point_stride), int(var_stride), int(N), w,
                                                  NORMALIZE if normalize else 0, int(normal_mask), _ptr(out["values"]),
                                                  _ptr(out["derivative"]), _ptr(out["jacobian"]), _ptr(out["normal"]),
                                                  _ptr(out["spans"]), _ptr(flag), _ptr(ws), ws_bytes, _stream(dev))
        else:
            rc = lib.bspy_cuda_eval_points(C.byref(ds.c), _ptr(uvw), int(point_stride), int(var_stride), int(N), w,
                                           NORMALIZE if normalize else 0, int(normal_mask), _ptr(out["values"]),
                                           _ptr(out["derivative"]), _ptr(out["jacobian"]), _ptr(out["normal"]),
                                           _ptr(out["spans"]), _ptr(flag), _stream(dev))
    _check(rc, "bspy_cuda_eval_points")
    return out


def record_layout(ds: DeviceSpline, jacobian, normal):
    """(length, stride) in doubles of an array-of-structs record [values | jacobian | normal]; the stride is the length
    rounded up to whole 32-byte sectors."""
    length = ds.nDep + (ds.nDep * ds.nInd if (jacobian or normal) else 0) + (ds.normal_dim if normal else 0)
    return length, (length + 3) // 4 * 4


def eval_points_aos(ds: DeviceSpline, uvw, point_stride, var_stride, N, *, jacobian=False, normal=False, normalize=True,
                    normal_mask=0, spans=False, flag=None, records=None):
    """Launch bspy_cuda_eval_points_aos: one record [values | jacobian (d, i) | normal] per point, ``records`` is
    (N, stride) float64 on the device (allocated here unless given).  Returns (records, spans or None)."""
    dev = ds.device
    length, stride = record_layout(ds, jacobian, normal)
    if records is None:
        records = torch.empty((N, stride), dtype=torch.float64, device=dev)
    assert records.dtype == torch.float64 and records.device == dev and records.stride(1) == 1 and records.shape[0] >= N
    sp = torch.empty((ds.nInd, N), dtype=torch.int32, device=dev) if spans else None
    flags = (NORMALIZE if normalize else 0) | (WANT_JACOBIAN if (jacobian or normal) else 0) | (WANT_NORMAL if normal else 0)
    lib = library()
    ws_bytes = int(lib.bspy_cuda_aos_workspace_bytes(C.byref(ds.c), int(N)))
    with torch.cuda.device(dev):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes > 0 else None
        rc = lib.bspy_cuda_eval_points_aos(C.byref(ds.c), _ptr(uvw), int(point_stride), int(var_stride), int(N), flags,
                                           int(normal_mask), _ptr(records), int(records.stride(0)), _ptr(sp), _ptr(flag),
                                           _ptr(ws), ws_bytes, _stream(dev))
    _check(rc, "bspy_cuda_eval_points_aos")
    return records, sp


def eval_grid(ds: DeviceSpline, axes, *, values=True, jacobian=False, normal=False, normalize=True, normal_mask=0,
              flag=None, out_f32=False):
    """Launch bspy_cuda_eval_grid; outputs (nDep, *nAxis), (nDep, nInd, *nAxis), (D, *nAxis); ``out_f32``: float32
    outputs (surfaces only: BSPY_OUT_F32)."""
    dev = ds.device
    shape = tuple(int(a.numel()) for a in axes)
    odt = torch.float32 if out_f32 else torch.float64
    out = {
        "values": torch.empty((ds.nDep, *shape), dtype=odt, device=dev) if values else None,
        "jacobian": torch.empty((ds.nDep, ds.nInd, *shape), dtype=odt, device=dev) if jacobian else None,
        "normal": torch.empty((ds.normal_dim, *shape), dtype=odt, device=dev) if normal else None,
    }
    n = max(ds.nInd, 1)
    ax = (C.c_void_p * n)(*[a.data_ptr() for a in axes])
    na = (C.c_int64 * n)(*shape)
    for a in axes:
        _f64(a, dev)
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_eval_grid(C.byref(ds.c), ax, na, (NORMALIZE if normalize else 0) | (OUT_F32 if out_f32 else 0), int(normal_mask),
                                           _ptr(out["values"]), _ptr(out["jacobian"]), _ptr(out["normal"]), _ptr(flag),
                                           _stream(dev))
    _check(rc, "bspy_cuda_eval_grid")
    return out


def eval_grid_batch(ds: DeviceSpline, n_splines, knot_strides, coef_stride, axes, *, values=True, jacobian=False,
                    normal=False, normalize=True, normal_mask=0, flag=None, out=None, out_f32=False):
    """Launch bspy_cuda_eval_grid_batch for ``n_splines`` surfaces; ``ds`` describes element 0; ``out_f32``: float32
    outputs (BSPY_OUT_F32)."""
    dev = ds.device
    shape = tuple(int(a.numel()) for a in axes)
    S = int(n_splines)
    odt = torch.float32 if out_f32 else torch.float64
    if out is None:
        out = {
            "values": torch.empty((S, ds.nDep, *shape), dtype=odt, device=dev) if values else None,
            "jacobian": torch.empty((S, ds.nDep, 2, *shape), dtype=odt, device=dev) if jacobian else None,
            "normal": torch.empty((S, ds.normal_dim, *shape), dtype=odt, device=dev) if normal else None,
        }
    for t in out.values():
        if t is not None and (t.dtype != odt or not t.is_contiguous()):
            raise ValueError(f"eval_grid_batch: output buffers must be contiguous {odt} tensors")
    ax = (C.c_void_p * 2)(*[a.data_ptr() for a in axes])
    na = (C.c_int64 * 2)(*shape)
    ks = (C.c_int64 * 2)(*[int(k) for k in knot_strides])
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_eval_grid_batch(C.byref(ds.c), S, ks, int(coef_stride), ax, na,
                                                 (NORMALIZE if normalize else 0) | (OUT_F32 if out_f32 else 0), int(normal_mask), _ptr(out.get("values")),
                                                 _ptr(out.get("jacobian")), _ptr(out.get("normal")), _ptr(flag), _stream(dev))
    _check(rc, "bspy_cuda_eval_grid_batch")
    return out


def eval_many(order, nCoef, nDep, knots, coefs, u, *, deriv1=False, flag=None, out=None):
    """Launch bspy_cuda_eval_many: knots (S, order+nCoef), coefs (S, nDep, nCoef), u (S, nPts)
    -> values (S, nDep, nPts) [, first derivatives of the same shape]."""
    dev = u.device
    S, nPts = int(u.shape[0]), int(u.shape[1])
    _f64(knots, dev), _f64(coefs, dev), _f64(u, dev)
    if out is None:
        out = {"values": torch.empty((S, nDep, nPts), dtype=torch.float64, device=dev),
               "derivative": torch.empty((S, nDep, nPts), dtype=torch.float64, device=dev) if deriv1 else None}
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_eval_many(int(order), int(nCoef), int(nDep), S, _ptr(knots), int(knots.stride(0)),
                                           _ptr(coefs), int(coefs.stride(0)), _ptr(u), nPts, _ptr(out["values"]),
                                           _ptr(out.get("derivative")), _ptr(flag), _stream(dev))
    _check(rc, "bspy_cuda_eval_many")
    return out


def many_table(order, nCoef, nDep, knots, coefs):
    """Build the cached per-curve images of a batch of curves (bspy_cuda_many_table_build): knots (S, order+nCoef) (stride 0
    allowed), coefs (S, nDep, nCoef).  Returns a uint8 device tensor, or None when the shape has no tables."""
    dev = coefs.device
    S = int(coefs.shape[0])
    lib = library()
    need = int(lib.bspy_cuda_many_table_bytes(int(order), int(nCoef), int(nDep), S))
    if need <= 0:
        return None
    table = torch.empty(need, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.bspy_cuda_many_table_build(int(order), int(nCoef), int(nDep), S, _ptr(knots), int(knots.stride(0)), _ptr(coefs),
                                            int(coefs.stride(0)), _ptr(table), need, _stream(dev))
    _check(rc, "bspy_cuda_many_table_build")
    return table


def eval_many_tab(order, nCoef, nDep, knots, coefs, table, u, *, deriv1=False, flag=None, out=None):
    """Launch bspy_cuda_eval_many_tab: values (S, nDep, nPts) [, first derivatives] of a batch of curves from its cached images."""
    dev = u.device
    S, nPts = int(u.shape[0]), int(u.shape[1])
    _f64(knots, dev), _f64(coefs, dev), _f64(u, dev)
    if out is None:
        out = {"values": torch.empty((S, nDep, nPts), dtype=torch.float64, device=dev),
               "derivative": torch.empty((S, nDep, nPts), dtype=torch.float64, device=dev) if deriv1 else None}
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_eval_many_tab(int(order), int(nCoef), int(nDep), S, _ptr(knots), int(knots.stride(0)), _ptr(coefs),
                                               int(coefs.stride(0)), _ptr(table), int(table.numel()), _ptr(u), nPts,
                                               _ptr(out["values"]), _ptr(out.get("derivative")), _ptr(flag), _stream(dev))
    _check(rc, "bspy_cuda_eval_many_tab")
    return out


def curvature(nInd, nDep, graph, d1, d2, normal):
    """Launch bspy_cuda_curvature on SoA derivative tensors (see include/bspy_cuda.h); returns (N,)."""
    dev = d1.device
    N = d1.shape[-1]
    out = torch.empty(N, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_curvature(int(nInd), int(nDep), int(bool(graph)), int(N), _ptr(_f64(d1, dev)), _ptr(_f64(d2, dev)),
                                           _ptr(normal), _ptr(out), _stream(dev))
    _check(rc, "bspy_cuda_curvature")
    return out


CURVATURE_MAX_ORDER = 8


def curvature_points(ds: DeviceSpline, uvw, point_stride, var_stride, N, flag=None):
    """Launch bspy_cuda_curvature_points (one fused pass per point); returns (N,) float64 on the device."""
    dev = ds.device
    out = torch.empty(N, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_curvature_points(C.byref(ds.c), _ptr(uvw), int(point_stride), int(var_stride), int(N), _ptr(out),
                                                  _ptr(flag), _stream(dev))
    _check(rc, "bspy_cuda_curvature_points")
    return out


def contract_axis(coefs, axis, first, order, basis):
    """Launch bspy_cuda_contract_axis: ``coefs`` (contiguous device tensor) with dimension ``axis`` contracted against
    ``basis`` (device tensor of ``order`` doubles) over indices first .. first+order-1; returns the tensor without
    that dimension."""
    dev = coefs.device
    shape = list(coefs.shape)
    outer = math.prod(shape[:axis])
    inner = math.prod(shape[axis + 1:])
    out = torch.empty(shape[:axis] + shape[axis + 1:], dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_contract_axis(_ptr(coefs), outer, int(shape[axis]), inner, int(first), int(order), _ptr(basis),
                                               _ptr(out), _stream(dev))
    _check(rc, "bspy_cuda_contract_axis")
    return out


def block_accumulate(dst, src, dst_rows):
    """dst[dst_rows[r], :] += src[r, :] for 2-D views (rows, N) of device tensors (bspy_cuda_block_accumulate)."""
    dev = dst.device
    rows = (C.c_int32 * max(len(dst_rows), 1))(*[int(r) for r in dst_rows])
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_block_accumulate(_ptr(dst), int(dst.stride(0)), _ptr(src), int(src.stride(0)), len(dst_rows), rows,
                                                  int(dst.shape[1]), _stream(dev))
    _check(rc, "bspy_cuda_block_accumulate")


def normal_from_jacobian(jac, nDep, nInd, sign, normalize, mask):
    """Cofactor normals (D, N) of the jacobians ``jac`` (nDep, nInd, N) (bspy_cuda_normal_from_jacobian)."""
    dev = jac.device
    N = jac.shape[-1]
    out = torch.empty((max(nDep, nInd), N), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_normal_from_jacobian(_ptr(jac), int(nDep), int(nInd), int(N), int(sign), NORMALIZE if normalize else 0,
                                                      int(mask), _ptr(out), _stream(dev))
    _check(rc, "bspy_cuda_normal_from_jacobian")
    return out


def collocation(knots, order, u, deriv_orders=None):
    """(spans int32[N], A float64[N, nCoef]) -- dense collocation rows (bspy_cuda_collocation)."""
    dev = u.device
    N = u.numel()
    nCoef = knots.numel() - int(order)
    sp = torch.empty(N, dtype=torch.int32, device=dev)
    A = torch.empty((N, nCoef), dtype=torch.float64, device=dev)
    if deriv_orders is not None:
        assert deriv_orders.dtype == torch.int32 and deriv_orders.device == dev and deriv_orders.is_contiguous()
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_collocation(_ptr(_f64(knots, dev)), knots.numel(), int(order), _ptr(_f64(u, dev)), _ptr(deriv_orders),
                                             N, _ptr(sp), _ptr(A), nCoef, _stream(dev))
    _check(rc, "bspy_cuda_collocation")
    return sp, A


def probe_fp64(kind, iters, dev):
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    flops = C.c_double(0.0)
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_probe_fp64(int(kind), int(iters), _ptr(sink), C.byref(flops), _stream(dev))
    _check(rc, "bspy_cuda_probe_fp64")
    return flops.value


def probe_tiles(dst, planes, nU, nV, tile_rows, tile_cols):
    nbytes = C.c_double(0.0)
    dev = dst.device
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_probe_tiles(_ptr(dst), int(planes), int(nU), int(nV), int(tile_rows), int(tile_cols),
                                             C.byref(nbytes), _stream(dev))
    _check(rc, "bspy_cuda_probe_tiles")
    return nbytes.value


def probe_hbm(kind, src, dst):
    nbytes = C.c_double(0.0)
    dev = dst.device
    with torch.cuda.device(dev):
        rc = library().bspy_cuda_probe_hbm(int(kind), _ptr(src), _ptr(dst), dst.numel(), C.byref(nbytes), _stream(dev))
    _check(rc, "bspy_cuda_probe_hbm")
    return nbytes.value


# ------------------------------------------------------------------ host-buffer pipelines
# Inputs / outputs in HOST memory: consecutive chunks go to a small ring of streams so that the
# H2D copy of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1 overlap (PCIe is full
# duplex).  Results land in pinned host tensors (torch's caching host allocator recycles them).

HOST_CHUNK = 1 << 22      # points per chunk
_RING = 3


def copy_rows(dst, src, stream):
    """dst[r, :] <- src[r, :] for 2-D views with unit inner stride (device <-> pinned host), ONE cudaMemcpy2DAsync."""
    assert dst.dim() == 2 and src.dim() == 2 and dst.shape == src.shape and dst.stride(1) == 1 and src.stride(1) == 1
    es = dst.element_size()
    rc = library().bspy_cuda_copy_2d(C.c_void_p(dst.data_ptr()), int(dst.stride(0)) * es, C.c_void_p(src.data_ptr()),
                                     int(src.stride(0)) * es, int(dst.shape[1]) * es, int(dst.shape[0]),
                                     C.c_void_p(stream.cuda_stream))
    _check(rc, "bspy_cuda_copy_2d")


def eval_points_host(ds: DeviceSpline, host, layout, *, check=True, chunk=None, aos=False, **request):
    """``host``: CPU float64 tensor, (N, nInd) for layout "points" or (nInd, N) for "variables".
    Returns (dict of pinned CPU tensors in SoA layout -- or {"records": (N, stride)} with ``aos`` --, index of the first
    out-of-domain point or -1).  Per chunk: one H2D copy, the kernels, one D2H copy per output array."""
    dev = ds.device
    chunk = int(chunk or HOST_CHUNK)
    N = host.shape[0] if layout == "points" else host.shape[1]
    D = ds.normal_dim

    def pinned(shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    if aos:
        _, rstride = record_layout(ds, request.get("jacobian"), request.get("normal"))
        res = {"records": pinned((N, rstride)), "spans": pinned((ds.nInd, N), torch.int32) if request.get("spans") else None}
    else:
        res = {
            "values": pinned((ds.nDep, N)) if request.get("values") else None,
            "derivative": pinned((ds.nDep, N)) if request.get("wrt") is not None else None,
            "jacobian": pinned((ds.nDep, ds.nInd, N)) if request.get("jacobian") else None,
            "normal": pinned((D, N)) if request.get("normal") else None,
            "spans": pinned((ds.nInd, N), torch.int32) if request.get("spans") else None,
        }
    starts = list(range(0, N, chunk))
    flags = torch.full((max(len(starts), 1),), -1, dtype=torch.int64, device=dev) if check else None
    streams = [torch.cuda.Stream(dev) for _ in range(min(_RING, max(1, len(starts))))]
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))
    for c, start in enumerate(starts):
        n = min(chunk, N - start)
        st = streams[c % len(streams)]
        st.wait_event(ready)
        with torch.cuda.stream(st):
            if layout == "points":
                d_pts = host[start:start + n].to(dev, non_blocking=True)
                ps, vs = ds.nInd, 1
            else:
                d_pts = torch.empty((ds.nInd, n), dtype=torch.float64, device=dev)
                copy_rows(d_pts, host[:, start:start + n], st)
                ps, vs = 1, n
            fl = None if flags is None else flags[c:c + 1]
            if aos:
                rec, sp = eval_points_aos(ds, d_pts, ps, vs, n, jacobian=bool(request.get("jacobian")), normal=bool(request.get("normal")),
                                          normalize=request.get("normalize", True), normal_mask=request.get("normal_mask", 0),
                                          spans=bool(request.get("spans")), flag=fl)
                res["records"][start:start + n].copy_(rec, non_blocking=True)
                out = {"spans": sp}
            else:
                out = eval_points(ds, d_pts, ps, vs, n, flag=fl, **request)
            for key, full in res.items():
                if full is None or key == "records":
                    continue
                copy_rows(full.reshape(-1, N)[:, start:start + n], out[key].reshape(-1, n), st)
    for st in streams:
        st.synchronize()
    first = -1
    if check and starts:
        f = flags.cpu()
        for c, start in enumerate(starts):
            if int(f[c]) >= 0:
                first = start + int(f[c])
                break
    return res, first


def eval_grid_batch_host(ds: DeviceSpline, n_splines, knot_strides, coef_stride, axes, *, group=None, check=True, **request):
    """Batch-of-surfaces grid evaluation with HOST outputs: groups of splines are evaluated into
    device buffers on alternating streams while the previous group's results stream back into pinned
    host tensors.  ``ds`` describes element 0; coefficient / knot pointers advance by the strides."""
    dev = ds.device
    S = int(n_splines)
    nU, nV = (int(a.numel()) for a in axes)
    D = ds.normal_dim
    per_spline = nU * nV * 8 * (ds.nDep * (1 if request.get("values", True) else 0) +
                                ds.nDep * 2 * (1 if request.get("jacobian") else 0) + D * (1 if request.get("normal") else 0))
    if group is None:
        group = max(1, min(S, int((1 << 30) // max(per_spline, 1))))      # ~1 GiB of results per group

    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=True)

    res = {
        "values": pinned((S, ds.nDep, nU, nV)) if request.get("values", True) else None,
        "jacobian": pinned((S, ds.nDep, 2, nU, nV)) if request.get("jacobian") else None,
        "normal": pinned((S, D, nU, nV)) if request.get("normal") else None,
    }
    starts = list(range(0, S, group))
    flags = torch.full((max(len(starts), 1),), -1, dtype=torch.int64, device=dev) if check else None
    streams = [torch.cuda.Stream(dev) for _ in range(min(2, max(1, len(starts))))]
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(dev))
    base_k = [int(ds.c.knots[i]) for i in range(2)]
    base_c = int(ds.c.coefs)
    sub = CSpline()
    C.memmove(C.byref(sub), C.byref(ds.c), C.sizeof(CSpline))
    view = DeviceSpline.__new__(DeviceSpline)
    view.__dict__.update(ds.__dict__)
    view.c = sub
    for g, start in enumerate(starts):
        n = min(group, S - start)
        st = streams[g % len(streams)]
        st.wait_event(ready)
        with torch.cuda.stream(st):
            for i in range(2):
                sub.knots[i] = base_k[i] + 8 * int(knot_strides[i]) * start
            sub.coefs = base_c + 8 * int(coef_stride) * start
            out = eval_grid_batch(view, n, knot_strides, coef_stride, axes, flag=None if flags is None else flags[g:g + 1],
                                  **request)
            for key, full in res.items():
                if full is not None:
                    full[start:start + n].copy_(out[key], non_blocking=True)
    for st in streams:
        st.synchronize()
    first = -1
    if check and starts:
        f = flags.cpu()
        for g, start in enumerate(starts):
            if int(f[g]) >= 0:
                first = start * nU * nV + int(f[g])
                break
    return res, first
