"""Oracle-backed stand-in for the bspy_b200._cuda entry points, so that the HOST logic of the
package (argument-form dispatch, shapes, dtypes, error messages, caching, sharding) can be tested
on a machine without a GPU.  Test infrastructure only: it is monkeypatched over the binding by
tests/test_host_logic.py; the product never imports it and has no CPU path of its own."""
import numpy as np
import torch

from oracle import bspy_oracle as O

CPU = torch.device("cpu")
launches = [0]


def device(index=None):
    return CPU


def new_flag(dev):
    return torch.full((1,), -1, dtype=torch.int64)


def launch_count():
    return launches[0]


def _ospline(ds):
    return O.OracleSpline(ds.nInd, ds.nDep, ds.order, ds.nCoef, [k.numpy() for k in ds.knots], ds.coefs.numpy(),
                          {"negateNormal": ds.normal_sign < 0})


def _mask_indices(mask, D):
    return [i for i in range(D) if (mask >> i) & 1] if mask else list(range(D))


def _evaluate(ds, pts, *, wrt=None, values=True, jacobian=False, normal=False, normalize=True, normal_mask=0, spans=False,
              flag=None):
    launches[0] += 1
    s = _ospline(ds)
    N = pts.shape[0]
    out = {"values": None, "derivative": None, "jacobian": None, "normal": None, "spans": None}
    with np.errstate(all="ignore"):
        if values:
            out["values"] = torch.from_numpy(np.ascontiguousarray(O.evaluate_vec(s, pts).T.reshape(ds.nDep, N)))
        if wrt is not None:
            out["derivative"] = torch.from_numpy(np.ascontiguousarray(O.derivative_vec(s, wrt, pts).T.reshape(ds.nDep, N)))
        if jacobian:
            out["jacobian"] = torch.from_numpy(np.ascontiguousarray(np.transpose(O.jacobian_vec(s, pts), (1, 2, 0))))
        if normal:
            D = max(ds.nInd, ds.nDep)
            raw = O.normal_vec(s, pts, False)
            if normalize:
                sel = _mask_indices(normal_mask, D)
                raw = raw / np.sqrt((raw[:, sel] ** 2).sum(axis=1))[:, None]
            out["normal"] = torch.from_numpy(np.ascontiguousarray(raw.T))
        if spans:
            sp = np.stack([O.span_vec(s.knots[i], s.order[i], pts[:, i]) for i in range(ds.nInd)]) if ds.nInd else np.empty((0, N), np.int32)
            out["spans"] = torch.from_numpy(sp.astype(np.int32))
    if flag is not None:
        bad = O.check_domain_vec(s, pts)
        if bad >= 0 and (int(flag[0]) < 0 or bad < int(flag[0])):
            flag[0] = bad
    return out


def eval_points(ds, uvw, point_stride, var_stride, N, **request):
    pts = torch.as_strided(uvw, (N, ds.nInd), (point_stride, var_stride)).numpy().astype(np.float64)
    return _evaluate(ds, pts, **request)


def record_layout(ds, jacobian, normal):
    length = ds.nDep + (ds.nDep * ds.nInd if (jacobian or normal) else 0) + (max(ds.nInd, ds.nDep) if normal else 0)
    return length, (length + 3) // 4 * 4


def _records(ds, pts, *, jacobian=False, normal=False, normalize=True, normal_mask=0, spans=False, flag=None):
    out = _evaluate(ds, pts, values=True, jacobian=jacobian or normal, normal=normal, normalize=normalize, normal_mask=normal_mask,
                    spans=spans, flag=flag)
    N = pts.shape[0]
    length, stride = record_layout(ds, jacobian, normal)
    rec = torch.zeros((N, stride), dtype=torch.float64)
    rec[:, :ds.nDep] = out["values"].T
    at = ds.nDep
    if jacobian or normal:
        rec[:, at:at + ds.nDep * ds.nInd] = out["jacobian"].reshape(ds.nDep * ds.nInd, N).T
        at += ds.nDep * ds.nInd
    if normal:
        rec[:, at:at + out["normal"].shape[0]] = out["normal"].T
    return rec, out["spans"]


def eval_points_aos(ds, uvw, point_stride, var_stride, N, **request):
    pts = torch.as_strided(uvw, (N, ds.nInd), (point_stride, var_stride)).numpy().astype(np.float64)
    return _records(ds, pts, **request)


def eval_points_host(ds, host, layout, *, check=True, chunk=None, aos=False, **request):
    pts = host.numpy() if layout == "points" else host.numpy().T
    flag = new_flag(CPU) if check else None
    if aos:
        rec, sp = _records(ds, np.ascontiguousarray(pts), jacobian=request.get("jacobian", False), normal=request.get("normal", False),
                           normalize=request.get("normalize", True), normal_mask=request.get("normal_mask", 0),
                           spans=request.get("spans", False), flag=flag)
        return {"records": rec, "spans": sp}, (int(flag[0]) if check else -1)
    out = _evaluate(ds, np.ascontiguousarray(pts), flag=flag, **request)
    return out, (int(flag[0]) if check else -1)


def eval_grid(ds, axes, *, values=True, jacobian=False, normal=False, normalize=True, normal_mask=0, flag=None, out_f32=False):
    shape = tuple(int(a.numel()) for a in axes)
    mesh = np.meshgrid(*[a.numpy() for a in axes], indexing="ij") if axes else []
    pts = np.stack([m.reshape(-1) for m in mesh], axis=1) if axes else np.empty((1, 0))
    r = _evaluate(ds, pts, values=values, jacobian=jacobian, normal=normal, normalize=normalize, normal_mask=normal_mask, flag=flag)
    cast = (lambda t: t.to(torch.float32)) if out_f32 else (lambda t: t)
    return {"values": None if r["values"] is None else cast(r["values"].reshape(ds.nDep, *shape)),
            "jacobian": None if r["jacobian"] is None else cast(r["jacobian"].reshape(ds.nDep, ds.nInd, *shape)),
            "normal": None if r["normal"] is None else cast(r["normal"].reshape(-1, *shape))}


def spans(knots, order, u):
    launches[0] += 1
    return torch.from_numpy(O.span_vec(knots.numpy(), order, u.numpy()))


def basis(knots, order, u, deriv=0, taylor=False, spans_in=None):
    launches[0] += 1
    ix, b = O.basis_vec(knots.numpy(), order, u.numpy(), deriv, taylor, None if spans_in is None else spans_in.numpy())
    return torch.from_numpy(ix), torch.from_numpy(b)


def curvature(nInd, nDep, graph, d1, d2, normal):
    """numpy restatement of curvature_kernel on the SoA derivative tensors (host-logic tests only)."""
    launches[0] += 1
    a, b = d1.numpy(), d2.numpy()
    N = a.shape[-1]
    with np.errstate(all="ignore"):
        if nInd == 1:
            if graph:
                a, b = np.vstack([np.ones((1, N)), a]), np.vstack([np.zeros((1, N)), b])
            pp, pq, qq = (a * a).sum(0), (a * b).sum(0), (b * b).sum(0)
            num = a[0] * b[1] - a[1] * b[0] if a.shape[0] == 2 else np.sqrt(qq * pp - pq ** 2)
            return torch.from_numpy(num / (pp * np.sqrt(pp)))
        if graph:
            z, o = np.zeros(N), np.ones(N)
            su, sv = np.stack([o, z, a[0]]), np.stack([z, o, a[1]])
            suu, suv, svv = (np.stack([z, z, b[i]]) for i in range(3))
            n = np.cross(su.T, sv.T).T
            n = n / np.sqrt((n * n).sum(0))
        else:
            J = a.reshape(3, 2, N)
            su, sv = J[:, 0], J[:, 1]
            suu, suv, svv = b.reshape(3, 3, N)
            n = normal.numpy()
        E, F, G = (su * su).sum(0), (su * sv).sum(0), (sv * sv).sum(0)
        L, M, Nn = (suu * n).sum(0), (suv * n).sum(0), (svv * n).sum(0)
        return torch.from_numpy((L * Nn - M ** 2) / (E * G - F ** 2))


CURVATURE_MAX_ORDER = 8


def curvature_points(ds, uvw, point_stride, var_stride, N, flag=None):
    launches[0] += 1
    pts = torch.as_strided(uvw, (N, ds.nInd), (point_stride, var_stride)).numpy().astype(np.float64)
    s = _ospline(ds)
    if flag is not None:
        bad = O.check_domain_vec(s, pts)
        if bad >= 0:
            flag[0] = bad
    with np.errstate(all="ignore"):
        return torch.from_numpy(np.ascontiguousarray(O.curvature_vec(s, pts)))


def contract_axis(coefs, axis, first, order, basis):
    launches[0] += 1
    c = np.moveaxis(coefs.numpy(), axis, -1)[..., first:first + order]
    return torch.from_numpy(np.ascontiguousarray(c @ basis.numpy()))


def block_accumulate(dst, src, dst_rows):
    launches[0] += 1
    for r, row in enumerate(dst_rows):
        dst[row] += src[r]


def normal_from_jacobian(jac, nDep, nInd, sign, normalize, mask):
    launches[0] += 1
    T = np.transpose(jac.numpy(), (2, 0, 1))                  # (N, nDep, nInd)
    if nInd > nDep:
        T = np.swapaxes(T, 1, 2)
    D = T.shape[1]
    n = np.empty((T.shape[0], D))
    with np.errstate(all="ignore"):
        for i in range(D):
            n[:, i] = sign * ((-1) ** i) * np.linalg.det(T[:, [j for j in range(D) if j != i], :])
        if normalize:
            sel = _mask_indices(mask, D)
            n = n / np.sqrt((n[:, sel] ** 2).sum(axis=1))[:, None]
    return torch.from_numpy(np.ascontiguousarray(n.T))


def collocation(knots, order, u, deriv_orders=None):
    launches[0] += 1
    kn, uu = knots.numpy(), u.numpy()
    nCoef = kn.shape[0] - order
    A = np.zeros((uu.shape[0], nCoef))
    sp = np.empty(uu.shape[0], np.int32)
    for r in range(uu.shape[0]):
        ix, b = O.basis_pt(None, kn, order, uu[r], 0 if deriv_orders is None else int(deriv_orders[r]))
        A[r, ix - order:ix] = b
        sp[r] = ix
    return torch.from_numpy(sp), torch.from_numpy(A)


def install(monkeypatch):
    from bspy_b200 import _cuda
    for name in ("device", "new_flag", "launch_count", "eval_points", "eval_points_host", "eval_points_aos", "record_layout", "eval_grid", "spans", "basis", "curvature", "curvature_points", "CURVATURE_MAX_ORDER",
                 "contract_axis", "block_accumulate", "normal_from_jacobian", "collocation"):
        monkeypatch.setattr(_cuda, name, globals()[name])
    launches[0] = 0
