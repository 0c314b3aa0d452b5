"""Host-side logic on a CPU-only machine: constructor semantics, JSON round trip, argument-form
dispatch and return conventions of the reference API, error messages, device-copy revalidation.
The CUDA entry points are replaced by an oracle-backed fake (tests/fake_cuda.py); the numbers are
compared with the goldens generated from the unmodified reference."""
import json

import numpy as np
import pytest
import torch

import bspy_b200 as bspy
from bspy_b200 import _cuda
import os
import fake_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from golden_io import close, load_cases, load_npz

CASES = {c.tag: c for c in load_cases()}


@pytest.fixture(autouse=True)
def _fake(monkeypatch):
    fake_cuda.install(monkeypatch)


def _spline(c):
    return bspy.Spline(c.nInd, c.nDep, c.order, c.nCoef, c.knots, c.coefs, c.metadata)


def test_no_cpu_fallback_without_fake(monkeypatch):
    monkeypatch.undo()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    s = _spline(CASES["curve_o4"])
    with pytest.raises(_cuda.CudaPathError, match="no CPU fallback"):
        s(0.5)
    with pytest.raises(_cuda.CudaPathError):
        s.evaluate_points(np.array([[0.5]]))


def test_constructor_layouts_match_reference():
    """bspy/spline.py:68-75: (nDep,*nCoef) kept; list of points (first variable fastest) and nDep
    flat arrays are reshaped the same way as the reference (goldens: mySurface is given as nDep
    flat lists, myCurve as a list of points)."""
    t = load_npz("ref_tables.npz")
    curve = bspy.Spline(1, 2, [4], [5], [[0, 0, 0, 0, 0.3, 1, 1, 1, 1]], [[0, 0], [0.3, 1], [0.5, 0.0], [0.7, -0.5], [1, 1]])
    assert curve.coefs.shape == (2, 5) and np.array_equal(curve.coefs, t["curve/coefs"])
    surf = bspy.Spline(2, 3, [3, 4], [4, 5], [[0, 0, 0, .5, 1, 1, 1], [0, 0, 0, 0, .5, 1, 1, 1, 1]],
                       [[0, 0, 0, 0, 0, .3, .3, .3, .3, .3, .7, .7, .7, .7, .7, 1, 1, 1, 1, 1],
                        [0, .25, .5, .75, 1, 0, .25, .5, .75, 1, 0, .25, .5, .75, 1, 0, .25, .5, .75, 1],
                        [0, 0, 0, 0, 0, 0, 1, 2, 1, 0, 0, 2, 1, 2, 0, 0, 0, 0, 0, 0]])
    assert surf.coefs.shape == (3, 4, 5) and np.array_equal(surf.coefs, t["surface/coefs"])
    pts = [[i + 10 * j, -(i + 10 * j)] for j in range(3) for i in range(4)]          # 12 points, first variable fastest
    s = bspy.Spline(2, 2, (2, 2), (4, 3), [[0, 0, .3, .6, 1, 1], [0, 0, .5, 1, 1]], pts)
    assert s.coefs.shape == (2, 4, 3) and s.coefs[0, 2, 1] == 12 and s.coefs[1, 3, 2] == -23
    assert isinstance(s.order, tuple) and isinstance(s.knots, tuple) and isinstance(s.knots[0], np.ndarray)


@pytest.mark.parametrize("bad, msg", [
    (dict(order=(4, 4)), "len\\(order\\) != nInd"),
    (dict(nCoef=(5, 5)), "len\\(nCoef\\) != nInd"),
    (dict(knots=[[0, 0, 0, 0, 1, 1, 1, 1]]), "should have length 9"),
    (dict(knots=[[0, 0, 0, 0, .6, .5, 1, 1, 1]]), "Improper knot order or multiplicity"),
    (dict(knots=[[0, 0, 0, 0, 0, 1, 1, 1, 1]]), "Improper knot order or multiplicity"),
    (dict(coefs=[[1, 2, 3]]), "Length of coefs should be 5 or 2"),
    (dict(nInd=-1), "nInd < 0"),
])
def test_constructor_errors(bad, msg):
    args = dict(nInd=1, nDep=2, order=(4,), nCoef=(5,), knots=[[0, 0, 0, 0, .3, 1, 1, 1, 1]], coefs=np.zeros((2, 5)))
    args.update(bad)
    with pytest.raises(ValueError, match=msg):
        bspy.Spline(args["nInd"], args["nDep"], args["order"], args["nCoef"], args["knots"], args["coefs"])


def test_json_round_trip_and_format(tmp_path):
    c = CASES["surf_44_neg"]
    s = _spline(c)
    f = tmp_path / "s.json"
    s.save(str(f))
    raw = json.load(open(f))
    assert list(raw) == ["type", "nInd", "nDep", "order", "nCoef", "knots", "coefs", "metadata"] and raw["type"] == "Spline"
    assert open(f).read().startswith('{\n    "type": "Spline",')          # indent=4
    [back] = bspy.Spline.load(str(f))
    assert (back.nInd, back.nDep, back.order, back.nCoef) == (s.nInd, s.nDep, s.order, s.nCoef)
    assert all(np.array_equal(a, b) for a, b in zip(back.knots, s.knots)) and np.array_equal(back.coefs, s.coefs)
    assert back.metadata == {"negateNormal": True}
    s.save(str(f), _spline(CASES["curve_o3"]))
    both = bspy.Spline.load(str(f))
    assert len(both) == 2 and both[1].order == (3,)
    old = dict(raw, metadata={"flipNormal": True})
    assert bspy.Spline.from_dict(old).metadata == {"negateNormal": True}
    assert bspy.Manifold.factory["Spline"] is bspy.Spline and isinstance(bspy.Manifold.from_dict(raw), bspy.Spline)
    assert bspy.Spline.from_dict({k: v for k, v in raw.items() if k != "metadata"}).metadata == {}


def test_legacy_npz_load(tmp_path):
    c = CASES["surf_34"]
    f = tmp_path / "old.npz"
    np.savez(str(f), order=np.array(c.order), knots0=c.knots[0], knots1=c.knots[1], coefficients=c.coefs)
    [s] = bspy.Spline.load(str(f))
    assert s.order == c.order and s.nCoef == c.nCoef and np.array_equal(s.coefs, c.coefs) and s.metadata["Name"] == "old"


@pytest.mark.parametrize("tag", ["curve_o4", "surf_34", "vol_444_d3", "man_3333_d6", "surf_25_d1"])
def test_single_point_conventions(tag):
    c = CASES[tag]
    s = _spline(c)
    p = 7
    uvw = c.uvw[p]
    for v in (s(*uvw), s(list(uvw)), s.evaluate(uvw), s.evaluate(*uvw)):
        assert isinstance(v, np.ndarray) and v.shape == (c.nDep,) and close(v, c["values"][p])
    J = s.jacobian(uvw)
    assert J.shape == (c.nDep, c.nInd) and close(J, c["jacobian"][p]) and close(s.tangent_space(uvw), J)
    w = c.meta["wrt"][-1]
    assert close(s.derivative(w, uvw), c["deriv_" + "_".join(map(str, w))][p])
    assert close(s.derivative(w, *uvw), c["deriv_" + "_".join(map(str, w))][p])
    if c.meta["normal"]:
        idx = c.meta["normal_indices"]
        assert close(s.normal(uvw), c["normal_unit"][p]) and close(s.normal(uvw, False), c["normal_raw"][p])
        n = s.normal(uvw, True, idx)
        assert n.shape == (len(idx),) and close(n, c["normal_idx_unit"][p])
        assert close(s.negate_normal().normal(uvw, False), -c["normal_raw"][p])
    assert np.array_equal(s.domain(), np.array([[k[o - 1], k[n]] for k, o, n in zip(c.knots, c.order, c.nCoef)]))
    ix, b = bspy.Spline.bspline_values(None, c.knots[0], c.order[0], uvw[0], 1)
    assert isinstance(ix, int) and ix == c["spans"][p, 0] and np.array_equal(b, c["basis0_d1"][p])
    ix, b = bspy.Spline.bspline_values(5, c.knots[0], c.order[0], uvw[0]) if c.nCoef[0] >= 5 else (5, None)
    assert ix == 5


def test_ufunc_style_dispatch():
    t, d = load_npz("ref_tables.npz"), load_npz("ref_dispatch.npz")
    curve = bspy.Spline(1, 2, t["curve/order"], (5,), [t["curve/knots0"]], t["curve/coefs"])
    surf = bspy.Spline(2, 3, t["surface/order"], (4, 5), [t["surface/knots0"], t["surface/knots1"]], t["surface/coefs"])
    r = curve(d["uu"])
    assert isinstance(r, tuple) and len(r) == 2 and all(a.shape == d["uu"].shape and a.dtype == curve.coefs.dtype for a in r)
    assert close(np.array(r), d["curve_ufunc"]) and close(np.array(curve.derivative([1], d["uu"])), d["curve_deriv_ufunc"])
    r = surf(d["U"], d["V"])
    assert isinstance(r, tuple) and len(r) == 3 and r[0].shape == d["U"].shape and close(np.array(r), d["surf_ufunc"])
    r = surf(d["U"][:, :1], d["V"][:1, :])                                   # broadcasting, as np.frompyfunc does
    assert close(np.array(r), d["surf_ufunc"])
    assert close(np.array(surf.derivative([1, 1], d["U"], d["V"])), d["surf_deriv_ufunc"])
    assert close(surf([0.25, 0.5]), d["surf_point_list"]) and close(surf(0.25, 0.5), d["surf_point_scalars"])
    scalar = bspy.Spline(1, 1, (3,), (5,), [d["scalar_knots"]], d["scalar_coefs"])
    r = scalar(d["uu"])
    assert isinstance(r, np.ndarray) and r.shape == d["scalar_ufunc"].shape and close(r, d["scalar_ufunc"])
    assert close(scalar(0.5), d["scalar_point"])
    r = scalar(d["uu2d"])
    assert r.shape == d["scalar_ufunc_2d"].shape and close(r, d["scalar_ufunc_2d"])
    # where= / out= of the reference's np.frompyfunc ufunc (bspy/spline.py:940-947), against the same construction around this
    # package's single-point evaluate: only selected points are evaluated, results land in the out arrays
    def as_reference(sp, args, **kw):
        uf = np.frompyfunc(lambda *p: tuple(sp.evaluate(*p)), sp.nInd, sp.nDep)
        if sp.nDep > 1:
            return tuple(a.astype(sp.coefs.dtype, copy=False) for a in uf(*args, **kw))
        return np.array([x[0] for x in uf(*args, **kw)], sp.coefs.dtype)
    mask = (np.add.outer(np.arange(d["U"].shape[0]), np.arange(d["U"].shape[1])) % 3) != 0
    outs_a = tuple(np.full(d["U"].shape, -5.0, dtype=object) for _ in range(3))
    outs_b = tuple(np.full(d["U"].shape, -5.0, dtype=object) for _ in range(3))
    ra = surf(d["U"], d["V"], where=mask, out=outs_a)
    rb = as_reference(surf, (d["U"], d["V"]), where=mask, out=outs_b)
    assert isinstance(ra, tuple) and len(ra) == 3
    for a, b in zip(ra, rb):
        assert a.dtype == b.dtype and a.shape == b.shape and close(a[mask], b[mask]) and np.all(a[~mask] == -5.0) and np.all(b[~mask] == -5.0)
    bad = d["U"].copy(); bad[~mask] = 7.0                                   # outside the domain, but never evaluated
    assert close(np.array(surf(bad, d["V"], where=mask, out=tuple(np.zeros(d["U"].shape, dtype=object) for _ in range(3))))[:, mask],
                 np.array(rb)[:, mask])
    ra, rb = surf(d["U"], d["V"], where=mask), as_reference(surf, (d["U"], d["V"]), where=mask)   # entries left out: None -> nan in the cast
    for a, b in zip(ra, rb):
        assert close(a[mask], b[mask]) and np.isnan(a[~mask]).all() and np.isnan(b[~mask]).all()
    with pytest.raises(TypeError):                                           # nDep == 1: None cannot be subscripted, there as here
        scalar(d["uu"], where=np.arange(d["uu"].shape[0]) % 2 == 0)
    with pytest.raises(TypeError):
        as_reference(scalar, (d["uu"],), where=np.arange(d["uu"].shape[0]) % 2 == 0)
    m1 = np.arange(d["uu"].shape[0]) % 2 == 0
    o1, o2 = np.full(d["uu"].shape, None, dtype=object), np.full(d["uu"].shape, None, dtype=object)
    o1[~m1] = 0.0
    for i in np.nonzero(~m1)[0]:
        o2[i] = (0.0,)                                                       # what the reference's ufunc leaves in its out array: 1-tuples
    assert close(scalar(d["uu"], where=m1, out=o1), as_reference(scalar, (d["uu"],), where=m1, out=o2))
    with pytest.raises(NotImplementedError):
        surf(d["U"], d["V"], casting="unsafe")
    with pytest.raises(ValueError, match="invalid number of arguments"):
        surf(d["uu"])                                                        # one array for two variables
    assert close(np.array(curve(d["uu"], out=None)), d["curve_ufunc"])         # numpy's default
    with pytest.raises(ValueError, match="outside domain"):
        curve(np.array([0.1, 1.5, 0.2]))


def test_zero_independent_variables():
    s = bspy.Spline(0, 3, (), (), (), [1.0, 2.0, 3.0])
    assert np.array_equal(s(), [1.0, 2.0, 3.0]) and np.array_equal(s.evaluate([]), [1.0, 2.0, 3.0])
    assert np.array_equal(s.derivative([]), np.zeros(3))
    assert fake_cuda.launch_count() == 0


def test_errors():
    s = _spline(CASES["surf_34"])
    with pytest.raises(ValueError, match="Incorrect number of parameter values: 1"):
        s.evaluate([0.5])
    with pytest.raises(ValueError, match="Spline evaluation outside domain: \\[0.5 1.5\\]"):
        s.evaluate([0.5, 1.5])
    with pytest.raises(ValueError, match="Spline evaluation outside domain"):
        s.jacobian([-0.1, 0.5])
    with pytest.raises(ValueError, match="must be one different"):
        _spline(CASES["curve_cfg1"]).normal([0.5])
    with pytest.raises(ValueError, match="Incorrect number of parameter values"):
        s.evaluate_points(np.zeros((4, 3)))
    pts = np.array([[0.1, 0.2], [0.3, 7.0]])
    with pytest.raises(ValueError, match="Spline evaluation outside domain: \\[0.3 7. \\]"):
        s.evaluate_points(pts)
    assert np.allclose(s.normal([0.5, 0.5], True, (0, 0)), np.sign(s.normal([0.5, 0.5], False)[0]) * np.sqrt(0.5))   # repeated indices: norm over both
    with pytest.raises(IndexError):
        s.normal([0.5, 0.5], True, (5,))
    nan = s(np.nan, 0.5)                                                     # NaN passes the domain test, as in the reference
    assert np.isnan(nan).all()


def test_evaluate_points_shapes_and_kinds():
    c = CASES["vol_444_d3"]
    s = _spline(c)
    r = s.evaluate_points(c.uvw, jacobian=True, spans=True, with_respect_to=[1, 0, 1])
    N = c.uvw.shape[0]
    assert r.values.shape == (3, N) and r.jacobian.shape == (3, 3, N) and r.spans.shape == (3, N) and r.spans.dtype == np.int32
    assert r.derivative.shape == (3, N) and r.normal is None and isinstance(r.values, np.ndarray)
    assert close(r.values.T, c["values"]) and np.array_equal(r.spans.T, c["spans"])
    r2 = s.evaluate_points(torch.from_numpy(c.uvw.T.copy()), layout="variables", values=False, jacobian=True)
    assert isinstance(r2.jacobian, torch.Tensor) and r2.values is None and close(r2.jacobian.numpy(), r.jacobian)
    r3 = s.evaluate_points(c.uvw.tolist())
    assert close(r3.values, r.values)
    curve = _spline(CASES["curve_o4"])
    flat = curve.evaluate_points(CASES["curve_o4"].uvw[:, 0])                # flat (N,) accepted for curves
    assert flat.values.shape == (curve.nDep, N := CASES["curve_o4"].uvw.shape[0]) and close(flat.values.T, CASES["curve_o4"]["values"])
    g = s.evaluate_grid(np.linspace(0, 1, 3), np.linspace(0, 1, 4), np.linspace(0, 1, 5), jacobian=True)
    assert g.values.shape == (3, 3, 4, 5) and g.jacobian.shape == (3, 3, 3, 4, 5)
    assert close(g.values[:, 1, 2, 3], s([0.5, 2 / 3, 0.75]))
    with pytest.raises(ValueError, match="outside domain"):
        s.evaluate_grid([0, 1], [0, 1], [0, 1.01])


def test_import_shim_and_repeated_normal_indices():
    """`import bspy` through shim/ resolves to this package (north star: "a new bspy/_cuda module"); repeated normal indices
    count once per repetition in the norm, like the reference (bspy/_spline_evaluation.py:234-244)."""
    import subprocess, sys
    code = ("import sys; sys.path[:0] = [%r, %r]; import bspy, bspy._cuda, bspy_b200; "
            "assert bspy.Spline is bspy_b200.Spline and bspy._cuda is bspy_b200._cuda and 'bspy_cuda_eval_points' in bspy._cuda.SYMBOLS; print('ok')"
            % (os.path.join(ROOT, "shim"), ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-800:]
    c = CASES["surf_34"]
    s = _spline(c)
    raw = s.evaluate_points(c.uvw[:7], values=False, normal=True, normalize=False).normal
    rep = s.evaluate_points(c.uvw[:7], values=False, normal=True, indices=(2, 0, 2)).normal
    want = raw[[2, 0, 2]] / np.sqrt((raw[[2, 0, 2]] ** 2).sum(axis=0))
    assert rep.shape == (3, 7) and np.allclose(rep, want, rtol=1e-15)
    one = s.normal(c.uvw[3], True, (2, 0, 2))
    assert np.allclose(one, want[:, 3], rtol=1e-15)
    assert np.array_equal(s.evaluate_points(c.uvw[:7], values=False, normal=True, normalize=False, indices=(1, 1)).normal, raw[[1, 1]])
    with pytest.raises(NotImplementedError):
        s.evaluate_grid([0.1, 0.2], [0.3], normal=True, indices=(0, 0))


def test_array_of_structs_layout_host_logic():
    """out_layout="aos": one record [values | jacobian (d, i) | normal] per point, stride rounded up to 4 doubles;
    values / jacobian / normal are views into the records with the struct-of-arrays shapes."""
    c = CASES["vol_444_d3"]
    s = _spline(c)
    N = c.uvw.shape[0]
    soa = s.evaluate_points(c.uvw, jacobian=True, spans=True)
    aos = s.evaluate_points(c.uvw, jacobian=True, spans=True, out_layout="aos")
    assert aos.records.shape == (N, 12) and isinstance(aos.records, np.ndarray)
    assert aos.values.shape == (3, N) and aos.jacobian.shape == (3, 3, N) and aos.normal is None
    assert np.shares_memory(aos.values, aos.records) and np.shares_memory(aos.jacobian, aos.records)
    assert np.array_equal(aos.values, soa.values) and np.array_equal(aos.jacobian, soa.jacobian) and np.array_equal(aos.spans, soa.spans)
    assert np.array_equal(aos.records[:, 3:].reshape(N, 3, 3), np.transpose(soa.jacobian, (2, 0, 1)))   # the reference's (nDep, nInd) per point
    only = s.evaluate_points(torch.from_numpy(c.uvw), out_layout="aos")
    assert only.records.shape == (N, 4) and isinstance(only.records, torch.Tensor) and only.jacobian is None
    surf = _spline(CASES["surf_44_d3"]) if "surf_44_d3" in CASES else None
    if surf is not None:
        cu = CASES["surf_44_d3"]
        r = surf.evaluate_points(cu.uvw, normal=True, indices=(2, 0), out_layout="aos")
        ref = surf.evaluate_points(cu.uvw, normal=True, indices=(2, 0))
        assert r.records.shape[1] == 12 and r.normal.shape == (2, cu.uvw.shape[0]) and close(r.normal, ref.normal) and r.jacobian is None
    with pytest.raises(ValueError, match="out_layout"):
        s.evaluate_points(c.uvw, out_layout="rows")
    with pytest.raises(ValueError, match="with_respect_to"):
        s.evaluate_points(c.uvw, with_respect_to=[1, 0, 0], out_layout="aos")
    with pytest.raises(ValueError, match="defer"):
        s.evaluate_points(c.uvw, check_domain="defer")
    with pytest.raises(ValueError, match="outside domain"):
        s.evaluate_points(np.array([[0.5, 0.5, 1.5]]), out_layout="aos")


def test_device_copy_revalidation_and_freeze():
    c = CASES["curve_o4"]
    s = _spline(c)
    u = c.uvw[:5]
    a = s.evaluate_points(u).values.copy()
    ds1 = bspy._spline_evaluation.device_spline(s)
    assert bspy._spline_evaluation.device_spline(s) is ds1                   # unchanged: cached copy reused
    orig = s.coefs.copy()
    s.coefs[0, :] += 2.0
    assert bspy._spline_evaluation.device_spline(s) is not ds1               # mutation seen: re-uploaded
    b = s.evaluate_points(u).values.copy()
    assert not np.array_equal(a, b)
    s.freeze()
    s.coefs[0, :] = orig[0]
    assert np.array_equal(s.evaluate_points(u).values, b)
    s.unfreeze()
    assert np.array_equal(s.evaluate_points(u).values, a)
    s.metadata["negateNormal"] = True
    assert bspy._spline_evaluation.device_spline(s).normal_sign == -1
    t = s.copy()
    assert t is not s and "_bspy_device_cache" not in t.__dict__ and np.array_equal(t.coefs, s.coefs)


def test_float32_and_integer_inputs():
    kn = np.array([0, 0, 0, 0, 1, 1, 1, 1], np.float32)
    co = np.arange(48, dtype=np.float32).reshape(3, 4, 4)
    s = bspy.Spline(2, 3, (4, 4), (4, 4), (kn, kn), co)                      # examples/teapot.py builds float32 splines
    v = s(0.25, 0.5)
    assert v.dtype == np.float32
    ref = bspy.Spline(2, 3, (4, 4), (4, 4), (kn.astype(float), kn.astype(float)), co.astype(float))(0.25, 0.5)
    assert np.allclose(v, ref, rtol=1e-6)
    i = bspy.Spline(1, 1, (2,), (3,), [[0, 0, 1, 2, 2]], [[0, 2, 4]])         # integer lists: computed in float64 here
    assert close(i(0.5), [1.0])


def test_curvature_api():
    ref = load_npz("ref_curvature.npz")
    for tag in ("curve_o4", "curve_o5", "curve_o3", "surf_34", "surf_25_d1"):
        c = CASES[tag]
        s = _spline(c)
        k = s.curvature_points(c.uvw)
        want = ref[tag]
        ok = np.isfinite(want)
        assert isinstance(k, np.ndarray) and k.shape == want.shape
        assert np.allclose(k[ok], want[ok], rtol=1e-8, atol=1e-8 * max(1.0, np.abs(want[ok]).max())), tag
        p = int(np.flatnonzero(ok)[3])
        one = s.curvature(c.uvw[p] if c.nInd > 1 else c.uvw[p, 0])
        assert isinstance(one, float) and np.isclose(one, want[p], rtol=1e-8, atol=1e-8)
    with pytest.raises(ValueError):
        _spline(CASES["vol_444_d3"]).curvature([0.5, 0.5, 0.5])
    with pytest.raises(ValueError, match="outside domain"):
        _spline(CASES["curve_o4"]).curvature(1.5)


def test_contract_block_and_collocation_host_logic():
    """Host side of SURVEY 8(f) rows 1 and 3 (maps, row sums, remapping after contract, run lengths of equal parameters)
    against the reference's golden outputs, with the oracle-backed fake binding."""
    from golden_io import _spline_from, block_members, load_npz
    a = load_npz("ref_block.npz")
    for name in a["contract/names"]:
        tag, j = str(name).split("/")
        s = bspy.Spline(*_spline_from(a, f"contract/{tag}"))
        uvw = [None if np.isnan(v) else float(v) for v in a[f"contract/{tag}/{j}/uvw"]]
        c = s.contract(uvw)
        nInd, nDep, order, nCoef, knots, coefs = _spline_from(a, f"contract/{tag}/{j}/result")
        assert (c.nInd, c.nDep, tuple(c.order), tuple(c.nCoef)) == (nInd, nDep, order, nCoef)
        assert close(np.asarray(c.coefs), coefs.reshape(np.asarray(c.coefs).shape))
    for tag in ("A", "B", "C"):
        rows = [[(m, bspy.Spline(*sp)) for m, sp in row] for row in block_members(a, tag)]
        b = bspy.SplineBlock(rows)
        assert [b.nInd, b.nDep] == list(a[f"block/{tag}/nIndnDep"])
        uvw = a[f"block/{tag}/uvw"]
        r = b.evaluate_points(uvw, jacobian=True, with_respect_to=list(a[f"block/{tag}/wrt"][1]))
        assert close(r.values.T, a[f"block/{tag}/values"])
        assert close(np.transpose(r.jacobian, (2, 0, 1)), a[f"block/{tag}/jacobian"])
        assert close(r.derivative.T, a[f"block/{tag}/deriv1"])
        assert close(b(uvw[3]), a[f"block/{tag}/values"][3]) and close(b.jacobian(uvw[3]), a[f"block/{tag}/jacobian"][3])
        if f"block/{tag}/normal_unit" in a:
            assert close(b.normal(uvw[3]), a[f"block/{tag}/normal_unit"][3])
            assert close(b.normal(uvw[3], False, (0, 2)), a[f"block/{tag}/normal_raw"][3][[0, 2]])
    rows = [[(m, bspy.Spline(*sp)) for m, sp in row] for row in block_members(a, "A")]
    cb = bspy.SplineBlock(rows).contract([None if np.isnan(v) else float(v) for v in a["block/A/contract_uvw"]])
    assert [cb.nInd, cb.nDep] == list(a["block/A/contract_nIndnDep"])
    assert close(cb.evaluate_points(a["block/A/contract_pts"]).values.T, a["block/A/contract_values"])
    for tag in ("o4", "o3", "o6"):
        sp, A = bspy.Spline.collocation_matrix(a[f"colloc/{tag}/knots"], int(a[f"colloc/{tag}/order"]), a[f"colloc/{tag}/u"])
        assert np.array_equal(A, a[f"colloc/{tag}/A"])


def test_spline_batch_bulk_loader(tmp_path):
    """SplineBatch.load: lists written by Spline.save, and splines nested as the manifolds of a Solid's boundaries
    (the layout of the reference's tests/teapots.json), grouped by shape in file order."""
    rng = np.random.default_rng(3)
    a = [_spline(CASES["surf_34"]), _spline(CASES["curve_o4"]), _spline(CASES["surf_34"]), _spline(CASES["surf_44_neg"])]
    a[2].coefs = a[2].coefs + 1.0
    path = tmp_path / "list.json"
    a[0].save(str(path), *a[1:])
    batches = bspy.SplineBatch.load(str(path))
    assert [b.indices for b in batches] == [[0, 2], [1], [3]]
    assert batches[0].nSplines == 2 and batches[0].order == (3, 4) and batches[2].metadata.get("negateNormal") is True
    assert np.array_equal(batches[0].coefs[1].numpy(), np.asarray(a[2].coefs)) and np.array_equal(batches[0].spline(0).knots[1], a[0].knots[1])
    solid = [{"type": "Solid", "dimension": 3, "containsInfinity": False, "metadata": {},
              "boundaries": [{"type": "Boundary", "manifold": s.to_dict(), "trim": {"type": "Solid", "boundaries": []}} for s in (a[0], a[2])]}]

    class Enc(json.JSONEncoder):
        def default(self, obj):
            return obj.tolist() if isinstance(obj, np.ndarray) else super().default(obj)
    nested = tmp_path / "solid.json"
    nested.write_text(json.dumps(solid, cls=Enc))
    (b,) = bspy.SplineBatch.load(str(nested))
    assert b.nSplines == 2 and np.array_equal(b.coefs[0].numpy(), np.asarray(a[0].coefs))
    import os
    ref = "/root/reference/tests/teapots.json"
    if os.path.exists(ref):                                       # build container only: teapot patches + their trim curves
        loaded = bspy.SplineBatch.load(ref)
        shapes = {}
        for t in loaded:                                          # batches are also split by metadata['negateNormal']
            key = (t.nInd, t.nDep, t.order, t.nCoef)
            shapes[key] = shapes.get(key, 0) + t.nSplines
        assert shapes == {(2, 3, (4, 4), (4, 4)): 82, (1, 2, (4,), (4,)): 24, (1, 2, (4,), (6,)): 4, (1, 2, (4,), (8,)): 4}


def test_grid_dtype_option_host_logic():
    """evaluate_grid(dtype=...): float32 for surfaces only, float64 default, anything else refused (fake binding)."""
    s = _spline(CASES["surf_34"])
    g = [np.linspace(*s.domain()[i], 5 + i) for i in range(2)]
    r64 = s.evaluate_grid(*g, jacobian=True)
    r32 = s.evaluate_grid(*g, jacobian=True, dtype=np.float32)
    assert r64.values.dtype == np.float64 and r32.values.dtype == np.float32 and r32.jacobian.dtype == np.float32
    assert np.array_equal(r32.values, r64.values.astype(np.float32))
    assert s.evaluate_grid(*g, dtype=np.float64).values.dtype == np.float64
    with pytest.raises(ValueError, match="dtype"):
        s.evaluate_grid(*g, dtype=np.int32)
    c = _spline(CASES["curve_o4"])
    with pytest.raises(NotImplementedError):
        c.evaluate_grid(np.linspace(*c.domain()[0], 7), dtype=np.float32)
