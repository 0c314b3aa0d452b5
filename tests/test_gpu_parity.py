"""Parity tests proper: the CUDA path (through the C ABI) against the goldens generated from the
unmodified reference and against the oracle.  Spans and basis values bit-exact; values, derivatives,
jacobians and normals within |x - ref| <= 1e-13 + 1e-12 |ref| (+ the condition term for the
reference's ill-conditioned fixtures, see oracle/bspy_oracle.py)."""
import numpy as np
import pytest
import torch

from golden_io import close, close_cond, load_cases, load_npz, nondegenerate_normal, well_conditioned_subset

pytestmark = pytest.mark.gpu

CASES = load_cases()
EPS = np.finfo(float).eps


def _mods():
    import bspy_b200
    from bspy_b200 import _cuda
    from oracle import bspy_oracle as O
    from oracle import c_oracle as CO
    return bspy_b200, _cuda, O, CO


def _spline(c):
    import bspy_b200
    return bspy_b200.Spline(c.nInd, c.nDep, c.order, c.nCoef, c.knots, c.coefs, c.metadata)


def _ospline(c):
    from oracle import bspy_oracle as O
    return O.OracleSpline(c.nInd, c.nDep, c.order, c.nCoef, c.knots, c.coefs, c.metadata)


def test_library_loaded_and_counts_launches():
    _, _cuda, _, _ = _mods()
    before = _cuda.launch_count()
    k = torch.tensor([0, 0, 0, 0.5, 1, 1, 1.0], dtype=torch.float64, device="cuda")
    u = torch.tensor([0.1, 0.5, 1.0], dtype=torch.float64, device="cuda")
    assert _cuda.spans(k, 3, u).tolist() == [3, 4, 4]
    assert _cuda.launch_count() == before + 1


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.tag)
def test_spans_and_basis_bit_exact(c):
    bspy, _cuda, _, _ = _mods()
    for i in range(c.nInd):
        k = torch.from_numpy(c.knots[i]).cuda()
        u = torch.from_numpy(np.ascontiguousarray(c.uvw[:, i])).cuda()
        assert np.array_equal(_cuda.spans(k, c.order[i], u).cpu().numpy(), c["spans"][:, i])
        for d in range(c.order[i] + 2):
            for taylor in (False, True):
                key = f"basis{i}_d{d}{'t' if taylor else ''}"
                if not c.has(key):
                    continue
                sp, B = _cuda.basis(k, c.order[i], u, d, taylor)
                assert np.array_equal(sp.cpu().numpy(), c["spans"][:, i])
                assert np.array_equal(B.cpu().numpy(), c[key], equal_nan=True), (c.tag, key)
        # given spans are used and returned unchanged (reference: `knot` argument)
        sp_in = torch.from_numpy(c["spans"][:, i].copy()).cuda()
        sp, B = _cuda.basis(k, c.order[i], u, 1, False, sp_in)
        assert np.array_equal(B.cpu().numpy(), c[f"basis{i}_d1"], equal_nan=True)
    # the single-parameter facade
    ix, b = bspy.Spline.bspline_values(None, c.knots[0], c.order[0], c.uvw[5, 0], 1)
    assert ix == c["spans"][5, 0] and np.array_equal(b, c["basis0_d1"][5], equal_nan=True)
    ix, b = bspy.Spline.bspline_values(int(c["spans"][7, 0]), c.knots[0], c.order[0], c.uvw[7, 0])
    assert ix == c["spans"][7, 0] and np.array_equal(b, c["basis0_d0"][7], equal_nan=True)


# The reference's own fixtures with nearly coincident knots (trim-issue.json: a knot gap of 1.9e-6 at order 7; the offset /
# reverse / Patterson / TomsNasty files) have derivative basis values up to 5e5 that cancel: ANY summation order other than
# numpy's differs from the reference by a few eps * sum|terms| there, the C restatement of the reference included.  Those
# files -- and only those -- are held to the condition-aware bar; the 23 synthetic cases to the strict bar
# |x - ref| <= 1e-13 + 1e-12 |ref|.  profiles/r02_parity_table.md lists, per fixture, how far the CUDA path actually is
# from the STRICT bar (tools/parity_table.py).
ILL_CONDITIONED = ("trim0", "reverse0", "offset0", "offset1", "offset2", "offset3", "offset4", "patterson0", "patterson1", "tomsnasty0")


def _bar(c, x, ref, scale, k=16.0):
    if c.tag in ILL_CONDITIONED:
        return close_cond(x, ref, scale, k=k)
    return close(x, ref)


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.tag)
def test_points_vs_reference(c):
    _, _, O, _ = _mods()
    s, so = _spline(c), _ospline(c)
    r = s.evaluate_points(c.uvw, values=True, jacobian=True, spans=True)
    assert np.array_equal(r.spans.T, c["spans"])
    S0 = O.derivative_abs_vec(so, [0] * c.nInd, c.uvw)
    assert _bar(c, r.values.T, c["values"], S0), c.tag
    SJ = O.jacobian_abs_vec(so, c.uvw)
    assert _bar(c, np.transpose(r.jacobian, (2, 0, 1)), c["jacobian"], SJ), c.tag
    for w in c.meta["wrt"]:
        ref = c["deriv_" + "_".join(map(str, w))]
        got = s.evaluate_points(c.uvw, values=False, with_respect_to=w).derivative.T
        assert _bar(c, got, ref, O.derivative_abs_vec(so, w, c.uvw)), (c.tag, w)
        if w[0] >= c.order[0]:
            assert not got.any()
    # values-only pass (different kernel instantiation) and CUDA-tensor input, (nInd, N) layout
    pts = torch.from_numpy(np.ascontiguousarray(c.uvw.T)).cuda()
    r2 = s.evaluate_points(pts, layout="variables")
    assert isinstance(r2.values, torch.Tensor) and r2.values.is_cuda
    assert _bar(c, r2.values.cpu().numpy().T, c["values"], S0)
    # array-of-structs records of the same request: the same numbers
    r3 = s.evaluate_points(c.uvw, jacobian=True, out_layout="aos")
    assert np.array_equal(r3.values, r.values, equal_nan=True) and np.array_equal(r3.jacobian, r.jacobian, equal_nan=True)


@pytest.mark.parametrize("c", [c for c in CASES if c.meta["normal"]], ids=lambda c: c.tag)
def test_normals_vs_reference(c):
    _, _, O, _ = _mods()
    s, so = _spline(c), _ospline(c)
    Sn = O.normal_abs_vec(so, c.uvw)
    with np.errstate(all="ignore"):
        Su = (Sn.max(axis=1) / np.sqrt((c["normal_raw"] ** 2).sum(axis=1)))[:, None]
    raw = s.evaluate_points(c.uvw, values=False, normal=True, normalize=False).normal.T
    assert close_cond(raw, c["normal_raw"], Sn, k=64), c.tag
    unit = s.evaluate_points(c.uvw, values=False, normal=True).normal.T
    assert close_cond(unit, c["normal_unit"], Su, k=64), c.tag
    idx = c.meta["normal_indices"]
    ok = well_conditioned_subset(c["normal_raw"], idx)
    sub = s.evaluate_points(c.uvw, values=False, normal=True, indices=idx).normal.T
    with np.errstate(all="ignore"):
        Si = (Sn[:, idx].max(axis=1) / np.sqrt((c["normal_idx_raw"] ** 2).sum(axis=1)))[:, None]
    assert sub.shape == c["normal_idx_unit"].shape
    assert close_cond(sub[ok], c["normal_idx_unit"][ok], Si[ok], k=64), c.tag
    subraw = s.evaluate_points(c.uvw, values=False, normal=True, normalize=False, indices=idx).normal.T
    assert close_cond(subraw, c["normal_idx_raw"], Sn[:, idx], k=64)
    # negate_normal() flips the sign exactly
    flipped = s.negate_normal().evaluate_points(c.uvw, values=False, normal=True, normalize=False).normal.T
    assert np.array_equal(flipped, -raw, equal_nan=True)


def test_reference_golden_tables():
    """truthCurve / truthSurface of the reference's tests (bspy_test.py:743-757)."""
    bspy, _, _, _ = _mods()
    t = load_npz("ref_tables.npz")
    curve = bspy.Spline(1, 2, t["curve/order"], (5,), [t["curve/knots0"]], t["curve/coefs"])
    tab = t["curve/table"]
    got = curve.evaluate_points(tab[:, 0]).values.T
    assert np.sqrt(((got - tab[:, 1:]) ** 2).sum(axis=1)).max() <= 4 * EPS
    for u, x, y in tab[::10]:
        assert np.hypot(*(curve.evaluate([u]) - (x, y))) <= 4 * EPS
    surf = bspy.Spline(2, 3, t["surface/order"], (4, 5), [t["surface/knots0"], t["surface/knots1"]], t["surface/coefs"])
    g = np.linspace(0, 1, 21)
    uv = np.array([(u, v) for v in g for u in g])
    got = surf.evaluate_points(uv).values.T
    assert np.sqrt(((got - t["surface/table"]) ** 2).sum(axis=1)).max() <= 4 * EPS
    grid = surf.evaluate_grid(g, g).values           # (3, nU, nV)
    assert np.sqrt(((grid.transpose(2, 1, 0).reshape(-1, 3) - t["surface/table"]) ** 2).sum(axis=1)).max() <= 4 * EPS
    # derivative of order >= spline order is exactly zero (bspy_test.py:759-761)
    assert not curve.derivative([4], [0.5]).any()


@pytest.mark.parametrize("c", [CASES[3], CASES[13], CASES[17], CASES[20], CASES[21]], ids=lambda c: c.tag)
def test_single_point_api(c):
    s = _spline(c)
    for p in range(0, c.uvw.shape[0], 11):
        uvw = c.uvw[p]
        v = s(uvw) if c.nInd > 1 else s(uvw[0])
        assert v.shape == (c.nDep,) and close(v, c["values"][p])
        assert close(s.evaluate(list(uvw)), c["values"][p])
        J = s.jacobian(uvw)
        assert J.shape == (c.nDep, c.nInd) and close(J, c["jacobian"][p])
        assert close(s.tangent_space(uvw), c["jacobian"][p])
        w = c.meta["wrt"][0]
        assert close(s.derivative(w, uvw), c["deriv_" + "_".join(map(str, w))][p])
        if c.meta["normal"]:
            assert close(s.normal(uvw), c["normal_unit"][p])
            assert close(s.normal(uvw, False), c["normal_raw"][p], atol=1e-13 * max(1, np.abs(c["normal_raw"][p]).max()))
            idx = c.meta["normal_indices"]
            n = s.normal(uvw, True, idx)
            assert n.shape == (len(idx),)
            if well_conditioned_subset(c["normal_raw"][p:p + 1], idx)[0]:
                assert close(n, c["normal_idx_unit"][p])


def test_dispatch_forms():
    """Argument forms of Spline.evaluate / derivative (bspy/spline.py:757-770, 936-949)."""
    bspy, _, _, _ = _mods()
    t, d = load_npz("ref_tables.npz"), load_npz("ref_dispatch.npz")
    curve = bspy.Spline(1, 2, t["curve/order"], (5,), [t["curve/knots0"]], t["curve/coefs"])
    surf = bspy.Spline(2, 3, t["surface/order"], (4, 5), [t["surface/knots0"], t["surface/knots1"]], t["surface/coefs"])
    r = curve(d["uu"])
    assert isinstance(r, tuple) and len(r) == 2 and r[0].shape == d["uu"].shape
    assert close(np.array(r), d["curve_ufunc"])
    assert close(np.array(curve.derivative([1], d["uu"])), d["curve_deriv_ufunc"])
    r = surf(d["U"], d["V"])
    assert isinstance(r, tuple) and len(r) == 3 and r[0].shape == d["U"].shape
    assert close(np.array(r), d["surf_ufunc"])
    assert close(np.array(surf.derivative([1, 1], d["U"], d["V"])), d["surf_deriv_ufunc"])
    assert close(surf([0.25, 0.5]), d["surf_point_list"]) and close(surf(0.25, 0.5), d["surf_point_scalars"])
    scalar = bspy.Spline(1, 1, (3,), (5,), [d["scalar_knots"]], d["scalar_coefs"])
    r = scalar(d["uu"])
    assert isinstance(r, np.ndarray) and r.shape == d["scalar_ufunc"].shape and close(r, d["scalar_ufunc"])
    assert close(scalar(0.5), d["scalar_point"])
    if "scalar_ufunc_2d" in d:
        r = scalar(d["uu2d"])
        assert r.shape == d["scalar_ufunc_2d"].shape and close(r, d["scalar_ufunc_2d"])
    # where= / out= of the reference's frompyfunc ufunc: only the selected points are evaluated (the others may lie outside)
    mask = (np.add.outer(np.arange(d["U"].shape[0]), np.arange(d["U"].shape[1])) % 3) != 0
    bad = d["U"].copy(); bad[~mask] = 7.0
    outs = tuple(np.full(d["U"].shape, -5.0, dtype=object) for _ in range(3))
    r = surf(bad, d["V"], where=mask, out=outs)
    assert close(np.array(r)[:, mask], d["surf_ufunc"][:, mask]) and np.all(np.array(r)[:, ~mask] == -5.0)
    r = surf(d["U"], d["V"], where=mask)
    assert close(np.array(r)[:, mask], d["surf_ufunc"][:, mask]) and np.isnan(np.array(r)[:, ~mask]).all()


def test_errors_match_reference():
    bspy, _, _, _ = _mods()
    c = CASES[13]
    s = _spline(c)
    with pytest.raises(ValueError, match="Spline evaluation outside domain"):
        s([0.5, 1.5])
    with pytest.raises(ValueError, match="Incorrect number of parameter values: 1"):
        s.evaluate([0.5])
    with pytest.raises(ValueError, match="invalid number of arguments"):   # ufunc path, as np.frompyfunc says it
        s([0.5, 0.5, 0.5])
    pts = c.uvw.copy()
    pts[17, 1] = 2.0
    with pytest.raises(ValueError, match="Spline evaluation outside domain"):
        s.evaluate_points(pts)
    with pytest.raises(ValueError, match="Spline evaluation outside domain"):
        s.evaluate_points(torch.from_numpy(pts).cuda())
    with pytest.raises(ValueError, match="Spline evaluation outside domain"):
        s.evaluate_grid(np.linspace(0, 1, 5), np.linspace(0, 1.1, 5))
    r = s.evaluate_points(pts, check_domain=False)       # no check requested: extrapolates silently
    assert np.isfinite(r.values).all()
    space_curve = bspy.Spline(1, 3, (3,), (4,), [[0, 0, 0, .5, 1, 1, 1]], np.arange(12.0).reshape(3, 4))
    with pytest.raises(ValueError, match="must be one different"):
        space_curve.normal([0.5])
    with pytest.raises(ValueError, match="must be one different"):
        space_curve.evaluate_points(np.array([[0.5]]), normal=True)
    # NaN parameter: inside the domain for the reference's comparisons, NaN result, last span
    r = s.evaluate_points(np.array([[np.nan, 0.5]]), spans=True)
    assert np.isnan(r.values).all() and r.spans[0, 0] == c.nCoef[0]


def test_teapot_grid_and_batch():
    """Values / derivatives everywhere; unit normals wherever the surface is regular.  At the singular
    points of the lid and bottom patches the reference returns NaN (56 of 2592 grid points) or an
    arbitrary unit vector, depending on rounding: there we only require NaN-or-unit-length."""
    bspy, _, O, _ = _mods()
    t = load_npz("teapot.npz")
    g = t["grid"]
    kn = t["knots"]
    uv = np.stack([m.reshape(-1) for m in np.meshgrid(g, g, indexing="ij")], axis=1)
    splines = [bspy.Spline(2, 3, (4, 4), (4, 4), (kn, kn), t["coefs"][p]) for p in range(32)]
    regular = np.empty((32, 9, 9), bool)
    for p in range(32):
        so = O.OracleSpline.of(splines[p])
        regular[p] = nondegenerate_normal(O.normal_vec(so, uv, False), O.normal_abs_vec(so, uv)).reshape(9, 9)
    assert regular.sum() >= 32 * 81 - 120 and not np.isnan(t["normal"][np.broadcast_to(regular[:, None], t["normal"].shape)]).any()

    def check_normals(n, p):
        m = np.broadcast_to(regular[p][None], n.shape)
        assert close(n[m], t["normal"][p][m]), p
        rest = n[:, ~regular[p]]
        length = np.sqrt((rest ** 2).sum(axis=0))
        assert np.all(np.isnan(length) | (np.abs(length - 1) < 1e-12)), p

    for p in (0, 5, 19, 20, 28, 31):     # 20.. include the degenerate lid / bottom patches
        r = splines[p].evaluate_grid(g, g, jacobian=True, normal=True)
        assert close(r.values, t["values"][p]) and close(r.jacobian[:, 0], t["du"][p]) and close(r.jacobian[:, 1], t["dv"][p])
        check_normals(r.normal, p)
        pts = splines[p].evaluate_points(uv, values=False, normal=True).normal.reshape(3, 9, 9)
        check_normals(pts, p)
    batch = bspy.SplineBatch.from_splines(splines)
    r = batch.evaluate_grid(g, g, jacobian=True, normal=True)
    assert r.values.shape == (32, 3, 9, 9)
    assert close(r.values, t["values"]) and close(r.jacobian[:, :, 0], t["du"]) and close(r.jacobian[:, :, 1], t["dv"])
    for p in range(32):
        check_normals(r.normal[p], p)


@pytest.mark.parametrize("c", [c for c in CASES if c.nInd in (2, 3)], ids=lambda c: c.tag)
def test_grid_vs_oracle(c):
    """Tensor-grid entry against the oracle on ragged (non-multiple-of-tile) axes that include every
    knot, both ends and their 1-ulp neighbours."""
    _, _, O, _ = _mods()
    s, so = _spline(c), _ospline(c)
    rng = np.random.default_rng(7)
    axes = []
    for i in range(c.nInd):
        a = np.unique(c.uvw[:, i])
        a = a[~np.isnan(a)]
        n = {0: 37, 1: 301 if c.nInd == 2 else 13, 2: 11}[i]
        axes.append(np.sort(rng.choice(a, size=min(n, a.size), replace=False)))
    mesh = np.meshgrid(*axes, indexing="ij")
    uvw = np.stack([m.reshape(-1) for m in mesh], axis=1)
    want_normal = abs(c.nInd - c.nDep) == 1
    r = s.evaluate_grid(*axes, jacobian=True, normal=want_normal, normalize=False)
    shape = tuple(len(a) for a in axes)
    assert r.values.shape == (c.nDep, *shape) and r.jacobian.shape == (c.nDep, c.nInd, *shape)
    assert close_cond(r.values.reshape(c.nDep, -1).T, O.evaluate_vec(so, uvw), O.derivative_abs_vec(so, [0] * c.nInd, uvw))
    assert close_cond(np.transpose(r.jacobian.reshape(c.nDep, c.nInd, -1), (2, 0, 1)), O.jacobian_vec(so, uvw),
                      O.jacobian_abs_vec(so, uvw))
    if want_normal:
        assert close_cond(r.normal.reshape(max(c.nInd, c.nDep), -1).T, O.normal_vec(so, uvw, False),
                          O.normal_abs_vec(so, uvw), k=64)
    # without normals volumes take the tensor-pipe kernel too (with normals: the scattered kernels in grid mode)
    r2 = s.evaluate_grid(*axes, jacobian=True)
    assert close_cond(r2.values.reshape(c.nDep, -1).T, O.evaluate_vec(so, uvw), O.derivative_abs_vec(so, [0] * c.nInd, uvw))
    assert close_cond(np.transpose(r2.jacobian.reshape(c.nDep, c.nInd, -1), (2, 0, 1)), O.jacobian_vec(so, uvw),
                      O.jacobian_abs_vec(so, uvw))
    r3 = s.evaluate_grid(*axes, values=True)
    assert close_cond(r3.values.reshape(c.nDep, -1).T, O.evaluate_vec(so, uvw), O.derivative_abs_vec(so, [0] * c.nInd, uvw))


def test_many_curves_vs_oracle():
    bspy, _, O, _ = _mods()
    rng = np.random.default_rng(1003)
    for order, nCoef, nDep, S, nPts in ((4, 32, 3, 257, 256), (3, 9, 2, 40, 77), (6, 11, 1, 9, 33), (2, 5, 3, 5, 1)):
        knots = np.empty((S, order + nCoef))
        for s in range(S):
            w = rng.uniform(0.25, 1.75, nCoef - order + 1)
            inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
            knots[s] = np.concatenate((np.zeros(order - 1), inner, np.ones(order - 1)))
        coefs = rng.standard_normal((S, nDep, nCoef))
        u = rng.uniform(0, 1, (S, nPts))
        u[0, 0], u[-1, -1] = 0.0, 1.0
        if nPts > 4:
            u[1, :3] = knots[1, order:order + 3]
        batch = bspy.SplineBatch(1, nDep, (order,), (nCoef,), [knots], coefs)
        r = batch.evaluate(u, derivative=True)
        assert r.values.shape == (S, nDep, nPts)
        for s in range(0, S, max(1, S // 9)):
            so = O.OracleSpline(1, nDep, (order,), (nCoef,), [knots[s]], coefs[s])
            assert close(r.values[s].T, O.evaluate_vec(so, u[s][:, None]))
            assert close_cond(r.derivative[s].T, O.derivative_vec(so, [1], u[s][:, None]),
                              O.derivative_abs_vec(so, [1], u[s][:, None]))
        with pytest.raises(ValueError, match="outside domain"):
            u2 = u.copy(); u2[S // 2, 0] = -0.5
            batch.evaluate(u2)
    # shared knots (1-D) broadcast to every curve
    shared = bspy.SplineBatch(1, 3, (4,), (32,), [np.concatenate((np.zeros(3), np.linspace(0, 1, 30), np.ones(3)))],
                              rng.standard_normal((6, 3, 32)))
    uu = rng.uniform(0, 1, (6, 50))
    r = shared.evaluate(uu)
    so = O.OracleSpline(1, 3, (4,), (32,), [shared.knots[0].cpu().numpy()], shared.coefs[4].cpu().numpy())
    assert close(r.values[4].T, O.evaluate_vec(so, uu[4][:, None]))


def test_many_curves_cached_images():
    """Value-only requests on a resident batch of curves read cached per-curve images (bucket table + per-span polynomial
    rows, fetched by TMA; bspy_cuda_many_table_build / bspy_cuda_eval_many_tab): values inside the strict bar of the
    recurrence kernel and of the oracle, the first offender outside the domain reported alike, ragged sizes (points per
    curve not a multiple of 32 and above the prefetched 256, curves not a multiple of the warps per block), NaN parameters,
    shared knots; a curve whose image is flagged invalid is evaluated by the recurrence inside the same kernel."""
    bspy, _cuda, O, _ = _mods()
    rng = np.random.default_rng(1703)
    for order, nCoef, nDep, S, nPts, shared in ((4, 32, 3, 3001, 256, False), (3, 9, 2, 1500, 77, False), (6, 11, 1, 1100, 333, False),
                                                (2, 5, 3, 1030, 40, True), (5, 40, 2, 1027, 300, False)):
        def one():
            w = rng.uniform(0.25, 1.75, nCoef - order + 1)
            inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
            return np.concatenate((np.zeros(order - 1), inner, np.ones(order - 1)))
        knots = one() if shared else np.stack([one() for _ in range(S)])
        coefs = rng.standard_normal((S, nDep, nCoef))
        u = rng.uniform(0, 1, (S, nPts))
        u[0, 0], u[-1, -1] = 0.0, 1.0
        kn1 = knots if shared else knots[1]
        u[1, :3] = kn1[order:order + 3]
        u[2, 5] = float("nan")
        batch = bspy.SplineBatch(1, nDep, (order,), (nCoef,), [knots], coefs)
        ut = torch.from_numpy(u).cuda()
        batch.cache_tables = False
        plain = batch.evaluate(ut).values
        plain_d = batch.evaluate(ut, derivative=True)
        batch.cache_tables = "now"
        cached = batch.evaluate(ut).values
        cached_d = batch.evaluate(ut, derivative=True)                  # Horner with derivative on the same rows
        assert _close_t(cached_d.values, plain_d.values), (order, nCoef, nDep)
        assert torch.equal(torch.nan_to_num(cached_d.values, nan=-7.0), torch.nan_to_num(cached, nan=-7.0))

        def close_der(x, y):
            # where a derivative cancels, two evaluation orders differ by a few eps times the terms they sum: the strict bar
            # widened by 32 eps times the curve's largest derivative
            scale = torch.nan_to_num(y).abs().amax(dim=(1, 2), keepdim=True)
            both = torch.isfinite(x) & torch.isfinite(y)
            return bool((torch.isfinite(x) == torch.isfinite(y)).all()) and \
                bool((((x - y).abs() <= 1e-13 + 1e-12 * y.abs() + 32 * np.finfo(float).eps * scale) | ~both).all())
        assert close_der(cached_d.derivative, plain_d.derivative), (order, nCoef, nDep)
        table = batch.__dict__["_curve_images_cache"][0]
        assert table is not None and table.numel() % S == 0
        assert _close_t(cached, plain), (order, nCoef, nDep)
        if order == 4:
            assert not torch.equal(torch.nan_to_num(cached), torch.nan_to_num(plain)), "cached images not in use"
        assert bool(torch.isnan(cached[2, :, 5]).all())
        for s_ in range(0, S, max(1, S // 7)):
            so = O.OracleSpline(1, nDep, (order,), (nCoef,), [knots if shared else knots[s_]], coefs[s_])
            keep = np.isfinite(u[s_])
            assert close(cached[s_].cpu().numpy().T[keep], O.evaluate_vec(so, u[s_][keep][:, None]))
            assert close_cond(cached_d.derivative[s_].cpu().numpy().T[keep], O.derivative_vec(so, [1], u[s_][keep][:, None]),
                              O.derivative_abs_vec(so, [1], u[s_][keep][:, None]))
        # curves 3 and S-1 flagged invalid in their images: the kernel evaluates them with the recurrence (tolerance-equal
        # to the plain kernel, which hoists the knot gaps into reciprocals)
        per = table.numel() // S
        for c in (3, S - 1):
            table[c * per + 8:c * per + 16] = 0
        again = batch.evaluate(ut).values
        again_d = batch.evaluate(ut, derivative=True)
        assert _close_t(again, plain) and close_der(again_d.derivative, plain_d.derivative)
        assert torch.equal(torch.nan_to_num(again[:3]), torch.nan_to_num(cached[:3]))
        # outside the domain: same first offender, same exception
        u2 = ut.clone(); u2[S // 2, 1] = -0.5; u2[S - 1, 0] = 1.5
        with pytest.raises(ValueError, match="outside domain"):
            batch.evaluate(u2)
        r = batch.evaluate(u2, check_domain="defer")
        assert int(r.first_outside.item()) == (S // 2) * nPts + 1
        # the images are made by the second call of a batch, never by the first
        lazy = bspy.SplineBatch(1, nDep, (order,), (nCoef,), [knots], coefs)
        lazy.evaluate(ut)
        assert "_curve_images_cache" not in lazy.__dict__
        lazy.evaluate(ut)
        assert "_curve_images_cache" in lazy.__dict__


def test_large_sample_vs_c_oracle_and_properties():
    """Config-4-shaped spline at a size the C oracle finishes in seconds, plus size-independent
    properties at a larger size: partition of unity, linearity in the coefficients, jacobian of an
    affine map."""
    bspy, _, _, CO = _mods()
    CO.build()
    rng = np.random.default_rng(1004)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    kn = [K(4, 32) for _ in range(3)]
    c1, c2 = rng.standard_normal((2, 3, 32, 32, 32))
    s1 = bspy.Spline(3, 3, (4, 4, 4), (32, 32, 32), kn, c1)
    N = 200_000
    uvw = rng.uniform(0, 1, (N, 3))
    r = s1.evaluate_points(uvw, jacobian=True, spans=True)
    ref = CO.evaluate(s1, uvw, jacobian=True, spans=True)
    assert np.array_equal(r.spans.T, ref["spans"])
    assert close(r.values.T, ref["values"])
    assert close(np.transpose(r.jacobian, (2, 0, 1)), ref["jacobian"], atol=1e-12)   # entries up to ~1e3: 1e-12 rel dominates
    # properties on the device at 4M points
    M = 4_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    pts = torch.rand((M, 3), dtype=torch.float64, device="cuda", generator=g)
    ones = bspy.Spline(3, 1, (4, 4, 4), (32, 32, 32), kn, np.ones((1, 32, 32, 32)))
    one = ones.evaluate_points(pts, jacobian=True)
    assert float((one.values - 1).abs().max()) <= 8 * EPS and float(one.jacobian.abs().max()) <= 1e-11
    s2 = bspy.Spline(3, 3, (4, 4, 4), (32, 32, 32), kn, c2)
    s12 = bspy.Spline(3, 3, (4, 4, 4), (32, 32, 32), kn, 2.0 * c1 - 0.5 * c2)
    a, b, ab = (s.evaluate_points(pts).values for s in (s1, s2, s12))
    assert float((ab - (2.0 * a - 0.5 * b)).abs().max()) <= 1e-13 * 40


def test_volume_grid_large_and_high_order():
    """nInd == 3 grids on the tensor pipe: a 32^3-coefficient tricubic volume on a ragged 45 x 37 x 301 grid and an
    order (5, 3, 6) volume (two K steps), against the oracle; out-of-domain axes raise."""
    bspy, _, O, _ = _mods()
    rng = np.random.default_rng(21)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    for order, nCoef, nDep, shape in (((4, 4, 4), (32, 32, 32), 3, (45, 37, 301)), ((5, 3, 6), (7, 5, 9), 2, (9, 11, 130)),
                                      ((2, 2, 2), (3, 4, 5), 4, (5, 4, 36))):
        s = bspy.Spline(3, nDep, order, nCoef, [K(o, n) for o, n in zip(order, nCoef)], rng.standard_normal((nDep, *nCoef)))
        axes = [np.sort(rng.uniform(0, 1, n)) for n in shape]
        axes[2][0], axes[2][-1], axes[0][0] = 0.0, 1.0, 0.0
        axes[1][3] = s.knots[1][order[1]] if nCoef[1] > order[1] else axes[1][3]
        r = s.evaluate_grid(*axes, jacobian=True)
        uvw = np.stack([m.reshape(-1) for m in np.meshgrid(*axes, indexing="ij")], axis=1)
        so = O.OracleSpline.of(s)
        assert r.values.shape == (nDep, *shape) and r.jacobian.shape == (nDep, 3, *shape)
        assert close(r.values.reshape(nDep, -1).T, O.evaluate_vec(so, uvw))
        assert close_cond(np.transpose(r.jacobian.reshape(nDep, 3, -1), (2, 0, 1)), O.jacobian_vec(so, uvw), O.jacobian_abs_vec(so, uvw))
        bad = [a.copy() for a in axes]
        bad[1][2] = 1.25
        with pytest.raises(ValueError, match="outside domain"):
            s.evaluate_grid(*bad)


def test_strided_input_and_zero_points():
    bspy, _, _, _ = _mods()
    c = CASES[17]
    s = _spline(c)
    wide = torch.zeros((c.uvw.shape[0], 7), dtype=torch.float64, device="cuda")
    wide[:, 1:6:2] = torch.from_numpy(c.uvw).cuda()
    view = wide[:, 1:6:2]                       # (N, 3) with strides (7, 2)
    assert close(s.evaluate_points(view).values.cpu().numpy().T, c["values"])
    r = s.evaluate_points(np.empty((0, c.nInd)))
    assert r.values.shape == (c.nDep, 0)


def test_mutation_is_seen_and_freeze_is_not():
    bspy, _, _, _ = _mods()
    c = CASES[3]
    s = _spline(c)
    u = c.uvw[:9]
    before = s.evaluate_points(u).values.copy()
    orig = s.coefs.copy()
    s.coefs[0, :] += 1.0
    after = s.evaluate_points(u).values
    assert not np.array_equal(before, after)
    s.freeze()
    s.coefs[0, :] = orig[0]
    assert np.array_equal(s.evaluate_points(u).values, after)     # frozen handle: device copy reused
    s.unfreeze()
    assert np.array_equal(s.evaluate_points(u).values, before)


def test_binned_path_is_bit_identical_to_unbinned():
    """Big scattered batches on a spline that lives in L2 are counting-sorted by knot-span cell and
    evaluated in cell order (bspy_cuda_eval_points_binned); per-point arithmetic is unchanged."""
    bspy, _cuda, _, _ = _mods()
    from bspy_b200._spline_evaluation import device_spline
    rng = np.random.default_rng(11)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    for shape in (dict(nInd=3, nDep=3, order=(4, 4, 4), nCoef=(32, 32, 32)), dict(nInd=4, nDep=6, order=(3,) * 4, nCoef=(16,) * 4)):
        s = bspy.Spline(shape["nInd"], shape["nDep"], shape["order"], shape["nCoef"],
                        [K(o, n) for o, n in zip(shape["order"], shape["nCoef"])], rng.standard_normal((shape["nDep"], *shape["nCoef"])))
        ds = device_spline(s)
        N = (1 << 19) + 12345                                  # more than one chunk, ragged tail
        assert _cuda.library().bspy_cuda_binned_workspace_bytes(ds.c, N) > 0
        g = torch.Generator(device="cuda").manual_seed(3)
        pts = torch.rand((N, s.nInd), dtype=torch.float64, device="cuda", generator=g)
        pts[5, 0], pts[N - 1, s.nInd - 1], pts[777, 1] = 0.0, 1.0, float(s.knots[1][7])
        for request in (dict(values=True, jacobian=True, spans=True), dict(values=True), dict(values=False, wrt=[1] + [0] * (s.nInd - 1))):
            a = _cuda.eval_points(ds, pts, s.nInd, 1, N, binned=True, **request)
            b = _cuda.eval_points(ds, pts, s.nInd, 1, N, binned=False, **request)
            for key in a:
                assert (a[key] is None) == (b[key] is None)
                if a[key] is not None:
                    assert torch.equal(a[key], b[key]), (shape, key)
        bad = pts.clone()
        bad[(1 << 19) + 5, 1] = 1.5
        bad[(1 << 19) + 900, 0] = -0.1
        flag = _cuda.new_flag(pts.device)
        _cuda.eval_points(ds, bad, s.nInd, 1, N, flag=flag, binned=True)
        assert int(flag.item()) == (1 << 19) + 5
    # sorted-record mode (>= 2 Mi points): same bits again, unit normals included
    s = bspy.Spline(2, 3, (4, 4), (70, 80), [K(4, 70), K(4, 80)], rng.standard_normal((3, 70, 80)))
    ds = device_spline(s)
    N = (1 << 22) + 4321
    pts = torch.rand((N, 2), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    for request in (dict(values=True, jacobian=True, normal=True, spans=True), dict(values=False, normal=True, normalize=False), dict(values=True)):
        a = _cuda.eval_points(ds, pts, 2, 1, N, binned=True, **request)
        b = _cuda.eval_points(ds, pts, 2, 1, N, binned=False, **request)
        for key in a:
            if a[key] is not None:
                assert torch.equal(a[key], b[key]), key
    # curves / small splines / small N keep the direct kernel
    small = bspy.Spline(2, 3, (4, 4), (8, 8), [K(4, 8), K(4, 8)], rng.standard_normal((3, 8, 8)))
    assert _cuda.library().bspy_cuda_binned_workspace_bytes(device_spline(small).c, 1 << 20) == 0


def test_record_mode_staged_windows_bit_identical(option):
    """Sorted-record mode (32-byte point records in cell order, per-span reciprocal records, warp-staged windows for
    the volume shapes, one dependent variable per pass for the 4-variate nDep-6 shape, cp.async un-permute): same
    bits as the direct thread-per-point kernel, for dense cells (staged kernel) and for a sparse tail."""
    bspy, _cuda, O, _ = _mods()
    from bspy_b200._spline_evaluation import device_spline
    rng = np.random.default_rng(31)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    option("BIN_MODE", 1)
    option("CELL_POLY", 0)                                         # the recurrence kernels: bit for bit (cell polynomials: next test)
    shapes = [((4, 4, 4), 3, (18, 18, 18)), ((3, 3, 3), 3, (20, 19, 18)), ((4, 4, 4), 1, (28, 27, 26)), ((4, 4, 4), 4, (17, 16, 18)),
              ((3, 3, 3, 3), 6, (10, 10, 9, 10))]
    for order, nDep, nCoef in shapes:
        nInd = len(order)
        s = bspy.Spline(nInd, nDep, order, nCoef, [K(o, n) for o, n in zip(order, nCoef)], rng.standard_normal((nDep, *nCoef)))
        ds = device_spline(s)
        cells = int(np.prod([n - o + 1 for o, n in zip(order, nCoef)]))
        for N in (64 * cells + 4321, 70_000):                       # dense cells (staged windows) / sparse cells
            assert _cuda.library().bspy_cuda_binned_workspace_bytes(ds.c, N) > 0, (order, nDep, N)
            g = torch.Generator(device="cuda").manual_seed(N)
            pts = torch.rand((N, nInd), dtype=torch.float64, device="cuda", generator=g)
            pts[7, 0], pts[N - 1, nInd - 1], pts[99, 1] = 0.0, 1.0, float(s.knots[1][order[1] + 2])
            requests = [dict(values=True, jacobian=True, spans=True), dict(values=True),
                        dict(values=False, wrt=[1] + [0] * (nInd - 2) + [2]), dict(values=True, wrt=[0] * (nInd - 1) + [1])]
            if abs(nInd - nDep) == 1:
                requests.append(dict(values=True, jacobian=True, normal=True))
            for request in requests:
                a = _cuda.eval_points(ds, pts, nInd, 1, N, binned=True, **request)
                b = _cuda.eval_points(ds, pts, nInd, 1, N, binned=False, **request)
                for key in a:
                    assert (a[key] is None) == (b[key] is None)
                    if a[key] is not None:
                        assert torch.equal(a[key], b[key]), (order, nDep, N, key)
            # and the direct kernel itself against the oracle on a sample
            idx = rng.integers(0, N, 3000)
            so = O.OracleSpline.of(s)
            ph = pts[idx].cpu().numpy()
            assert close(a["values"][:, idx].cpu().numpy().T, O.evaluate_vec(so, ph))


def _close_t(x, y):
    """strict bar |x - y| <= 1e-13 + 1e-12 |y| on device tensors, NaNs matching"""
    return bool(torch.isclose(x, y, rtol=1e-12, atol=1e-13, equal_nan=True).all())


def test_cell_kernel_tensor_pipe_and_aos_records(option):
    """The cell-sorted pipeline with the first contraction stage on the FP64 tensor pipe (eval_cell_mma_kernel) and
    array-of-structs records written straight to the points' original positions: against the direct thread-per-point
    kernel (strict bar on every point: the summation order differs in the last variable only), the oracle on a sample,
    spans bit-exact; struct-of-arrays and array-of-structs outputs of the same request are bit-identical; many small
    chunks (ragged, straddling tiles), dense and sparse cells, the staged / gather kernels behind the same records."""
    bspy, _cuda, O, _ = _mods()
    from bspy_b200._spline_evaluation import device_spline
    from oracle import c_oracle as CO
    rng = np.random.default_rng(131)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    shapes = [((4, 4, 4), 3, (18, 18, 18)), ((3, 3, 3), 3, (20, 19, 18)), ((4, 4, 4), 1, (28, 27, 26)), ((3, 3, 3, 3), 6, (10, 10, 9, 10)),
              ((4, 4, 4), 4, (17, 16, 18)), ((3, 4), 3, (150, 140)), ((5, 4), 3, (40, 30)), ((3, 3, 3), 2, (12, 21, 13)),
              ((4, 4, 4), 2, (9, 8, 30))]
    for order, nDep, nCoef in shapes:
        nInd = len(order)
        s = bspy.Spline(nInd, nDep, order, nCoef, [K(o, n) for o, n in zip(order, nCoef)], rng.standard_normal((nDep, *nCoef)))
        ds = device_spline(s)
        cells = int(np.prod([n - o + 1 for o, n in zip(order, nCoef)]))
        want_normal = abs(nInd - nDep) == 1
        for N, chunk_log2 in ((64 * cells + 4321, 22), (3 * (1 << 18) + 777, 18), (70_000, 22)):
            option("BIN_REC_CHUNK_LOG2", chunk_log2)
            g = torch.Generator(device="cuda").manual_seed(N)
            pts = torch.rand((N, nInd), dtype=torch.float64, device="cuda", generator=g)
            pts[7, 0], pts[N - 1, nInd - 1], pts[99, 1] = 0.0, 1.0, float(s.knots[1][order[1] + 2])
            direct = _cuda.eval_points(ds, pts, nInd, 1, N, binned=False, values=True, jacobian=True, normal=want_normal, spans=True)
            # tensor-pipe experiment, cell polynomials (default), recurrence, cell polynomials on staged images with two points per lane
            for cell_kernel, cell_poly in ((1, None), (0, None), (0, 0), (0, 2022), (0, 1023)):
                option("CELL_KERNEL", cell_kernel)
                option("CELL_POLY", cell_poly)
                rec, sp = _cuda.eval_points_aos(ds, pts, nInd, 1, N, jacobian=True, normal=want_normal, spans=True)
                assert rec.shape == (N, (nDep * (1 + nInd) + (max(nInd, nDep) if want_normal else 0) + 3) // 4 * 4)
                assert torch.equal(sp, direct["spans"]), (order, nDep, N, cell_kernel)
                assert _close_t(rec[:, :nDep].T, direct["values"]), (order, nDep, N, cell_kernel)
                assert _close_t(rec[:, nDep:nDep * (1 + nInd)].T.reshape(nDep, nInd, N), direct["jacobian"]), (order, nDep, N, cell_kernel)
                if want_normal:
                    assert _close_t(rec[:, nDep * (1 + nInd):nDep * (1 + nInd) + max(nInd, nDep)].T, direct["normal"])
                if N >= (1 << 16):                                  # struct-of-arrays through the same sorted records + un-permute
                    option("BIN_MODE", 1)
                    soa = _cuda.eval_points(ds, pts, nInd, 1, N, binned=True, values=True, jacobian=True, normal=want_normal)
                    option("BIN_MODE", None)
                    assert torch.equal(soa["values"], rec[:, :nDep].T) and \
                        torch.equal(soa["jacobian"], rec[:, nDep:nDep * (1 + nInd)].T.reshape(nDep, nInd, N)), (order, nDep, N, cell_kernel)
                values_only, _ = _cuda.eval_points_aos(ds, pts, nInd, 1, N)
                assert torch.equal(values_only[:, :nDep].T, _cuda.eval_points(ds, pts, nInd, 1, N, binned=False)["values"])
            option("CELL_KERNEL", None)
            option("CELL_POLY", None)
            idx = rng.integers(0, N, 4000)
            ref = CO.evaluate(s, pts[idx].cpu().numpy(), values=True, jacobian=True, spans=True)
            assert np.array_equal(sp[:, idx].cpu().numpy().T, ref["spans"])
            assert close(rec[idx, :nDep].cpu().numpy(), ref["values"])
            assert close(rec[idx, nDep:nDep * (1 + nInd)].cpu().numpy().reshape(-1, nDep, nInd), ref["jacobian"])
        # outside the domain: the first offender survives sorting and chunking
        bad = pts.clone()
        bad[60_123, nInd - 1] = 1.5
        bad[65_000, 0] = -0.25
        flag = _cuda.new_flag(pts.device)
        _cuda.eval_points_aos(ds, bad, nInd, 1, N, jacobian=True, flag=flag)
        assert int(flag.item()) == 60_123
    # the public API: views into the records, host inputs through the chunked pipeline
    s = bspy.Spline(3, 3, (4, 4, 4), (18, 18, 18), [K(4, 18) for _ in range(3)], rng.standard_normal((3, 18, 18, 18)))
    pts = torch.rand((150_000, 3), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(8))
    a = s.evaluate_points(pts, jacobian=True, out_layout="aos", check_domain="defer")
    b = s.evaluate_points(pts.cpu().numpy(), jacobian=True, out_layout="aos")
    assert a.records.shape == (150_000, 12) and int(a.first_outside.item()) == -1 and a.raise_if_outside() is a
    assert np.array_equal(a.records.cpu().numpy(), b.records) and np.array_equal(a.jacobian.cpu().numpy(), b.jacobian)
    c = s.evaluate_points(pts, jacobian=True)
    assert _close_t(a.values, c.values) and _close_t(a.jacobian, c.jacobian)


def test_curve_replicated_rows_bit_identical(option):
    """Big batches on one curve go through eval_curve_repl_kernel (bank-replicated span rows, bucket table + short
    advance instead of a bisection): spans bit-exact, values and derivatives bit-identical to the plain curve kernel,
    and both within tolerance of the oracle."""
    bspy, _cuda, O, _ = _mods()
    from bspy_b200._spline_evaluation import device_spline
    rng = np.random.default_rng(21)

    def K(o, n, clamp=True, cluster=False):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        if cluster and n - o + 1 > 6:
            w[2:5] = 1e-7                                         # several knots inside one bucket
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        if clamp:
            return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))
        return np.concatenate((-np.cumsum(rng.uniform(0.1, 0.3, o - 1))[::-1], inner, 1.0 + np.cumsum(rng.uniform(0.1, 0.3, o - 1))))

    shapes = [(4, 3, 64, True, False), (4, 3, 64, False, True), (2, 1, 5, True, False), (3, 2, 300, True, True), (6, 4, 17, True, False),
              (1, 2, 9, True, False), (5, 2, 40, False, False), (4, 3, 2000, True, False)]
    for order, nDep, nCoef, clamp, cluster in shapes:
        s = bspy.Spline(1, nDep, (order,), (nCoef,), [K(order, nCoef, clamp, cluster)], rng.standard_normal((nDep, nCoef)))
        ds = device_spline(s)
        lo, hi = s.knots[0][order - 1], s.knots[0][nCoef]
        N = 140_000 + 1234
        g = torch.Generator(device="cuda").manual_seed(order * 100 + nCoef)
        u = lo + (hi - lo) * torch.rand((N, 1), dtype=torch.float64, device="cuda", generator=g)
        u.clamp_(lo, hi)
        kn = torch.from_numpy(np.asarray(s.knots[0][order - 1:nCoef + 1], dtype=np.float64)).cuda()
        m = kn.numel()
        u[:m, 0] = kn                                             # every knot, both ends ...
        u[m:2 * m, 0] = torch.from_numpy(np.clip(np.nextafter(kn.cpu().numpy(), -np.inf), lo, hi)).cuda()   # ... and their neighbours
        u[2 * m:3 * m, 0] = torch.from_numpy(np.clip(np.nextafter(kn.cpu().numpy(), np.inf), lo, hi)).cuda()
        u[N - 1, 0] = float("nan")
        requests = [dict(values=True, spans=True), dict(values=True, jacobian=True), dict(values=False, wrt=[1])]
        if nDep == 2:
            requests.append(dict(values=True, jacobian=True, normal=True))
            requests.append(dict(values=False, normal=True, normalize=False))
        option("CURVE_POLY", 0)                                      # recurrence rows for value-only requests too: the bit-for-bit comparisons
        for request in requests:
            option("CURVE_REPL", 1)
            a = _cuda.eval_points(ds, u, 1, 1, N, **request)
            option("CURVE_REPL", 0)
            b = _cuda.eval_points(ds, u, 1, 1, N, **request)
            option("CURVE_REPL", 1)
            option("CURVE_TMA", 0)                                  # tables rebuilt in every thread block (no cached image)
            c2 = _cuda.eval_points(ds, u, 1, 1, N, **request)
            option("CURVE_TMA", None)
            for key in a:
                if a[key] is not None:
                    x, y = a[key], c2[key]
                    if x.dtype.is_floating_point:
                        x, y = torch.nan_to_num(x, nan=-7.0), torch.nan_to_num(y, nan=-7.0)
                    assert torch.equal(x, y), (order, nDep, nCoef, key, "cached tables vs per-block tables")
            for key in a:
                assert (a[key] is None) == (b[key] is None)
                if a[key] is not None:
                    x, y = a[key], b[key]
                    if x.dtype.is_floating_point:
                        x, y = torch.nan_to_num(x, nan=-7.0), torch.nan_to_num(y, nan=-7.0)
                    assert torch.equal(x, y), (order, nDep, nCoef, key)
        # value-only requests on the cached tables read per-span polynomial rows (validated against the recurrence when the
        # table is built): spans bit-exact, values inside the strict bar of the recurrence rows
        option("CURVE_REPL", None)
        rec_rows = _cuda.eval_points(ds, u, 1, 1, N, values=True, spans=True)
        option("CURVE_POLY", None)
        poly_rows = _cuda.eval_points(ds, u, 1, 1, N, values=True, spans=True)
        assert torch.equal(poly_rows["spans"], rec_rows["spans"]), (order, nDep, nCoef)
        assert _close_t(poly_rows["values"], rec_rows["values"]), (order, nDep, nCoef)
        if (order, nDep, nCoef, clamp, cluster) == (4, 3, 64, True, False):       # the shape of config 1: the rows must be in use
            assert not torch.equal(torch.nan_to_num(poly_rows["values"]), torch.nan_to_num(rec_rows["values"])), "polynomial rows not in use"
        # the same rows serve requests with the first derivative (Horner with derivative; validated separately at the build)
        option("CURVE_POLY", 0)
        rec_der = _cuda.eval_points(ds, u, 1, 1, N, values=True, jacobian=True)
        option("CURVE_POLY", None)
        poly_der = _cuda.eval_points(ds, u, 1, 1, N, values=True, jacobian=True)
        assert _close_t(poly_der["values"], rec_der["values"]), (order, nDep, nCoef)
        if not cluster:
            assert _close_t(poly_der["jacobian"], rec_der["jacobian"]), (order, nDep, nCoef)
        if (order, nDep, nCoef, clamp, cluster) == (4, 3, 64, True, False):
            assert not torch.equal(torch.nan_to_num(poly_der["jacobian"]), torch.nan_to_num(rec_der["jacobian"])), "polynomial rows not in use"
        if ds.curve_table is not None:
            # a table whose rows failed the validation of the build (flags in the image's 16-byte trailer cleared): the kernel
            # fetches the recurrence rows instead -- bit-identical to them
            saved = ds.curve_table[-16:].clone()
            ds.curve_table[-16:] = 0
            fallback = _cuda.eval_points(ds, u, 1, 1, N, values=True, spans=True)
            fallback_der = _cuda.eval_points(ds, u, 1, 1, N, values=True, jacobian=True)
            ds.curve_table[-16:] = saved
            assert torch.equal(fallback["spans"], rec_rows["spans"])
            assert torch.equal(torch.nan_to_num(fallback["values"], nan=-7.0), torch.nan_to_num(rec_rows["values"], nan=-7.0)), (order, nDep, nCoef)
            assert torch.equal(torch.nan_to_num(fallback_der["jacobian"], nan=-7.0), torch.nan_to_num(rec_der["jacobian"], nan=-7.0)), (order, nDep, nCoef)
        # coefficients of magnitude 1e9: where the value cancels, any two evaluation orders differ by a few eps * max|coef|;
        # the two row forms must agree to the strict bar widened by exactly that
        big = bspy.Spline(1, nDep, (order,), (nCoef,), [np.array(s.knots[0])], 1e9 * np.asarray(s.coefs) + 1.0)
        dbig = device_spline(big)
        pb = _cuda.eval_points(dbig, u, 1, 1, N, values=True)["values"]
        option("CURVE_POLY", 0)
        rb = _cuda.eval_points(dbig, u, 1, 1, N, values=True)["values"]
        option("CURVE_POLY", None)
        slack = 32 * np.finfo(float).eps * float(np.abs(big.coefs).max())
        both = torch.isfinite(pb) & torch.isfinite(rb)
        assert bool((torch.isfinite(pb) == torch.isfinite(rb)).all()), (order, nDep, nCoef)
        assert bool(((pb - rb).abs()[both] <= 1e-13 + 1e-12 * rb.abs()[both] + slack).all()), (order, nDep, nCoef, "magnitude 1e9")
        # default dispatch (replicated rows for this N) against the oracle: spans bit-exact, values within tolerance
        r = _cuda.eval_points(ds, u, 1, 1, N, values=True, jacobian=True, spans=True)
        idx = np.concatenate((np.arange(0, min(3 * m, 6000)), rng.integers(0, N - 1, 4000)))
        so = O.OracleSpline.of(s)
        uh = u[idx].cpu().numpy()
        ref, sp = O.evaluate_vec(so, uh, return_spans=True)
        assert np.array_equal(r["spans"][0, idx].cpu().numpy(), sp[:, 0]), (order, nDep, nCoef)
        if not cluster:                                            # 1e-7 gaps make the derivative scale 1e7: condition-aware bar elsewhere
            assert close(r["values"][:, idx].cpu().numpy().T, ref)
            assert close(poly_rows["values"][:, idx].cpu().numpy().T, ref)
            assert close(r["jacobian"][:, 0, idx].cpu().numpy().T, O.jacobian_vec(so, uh)[:, :, 0])
        # outside the domain: first offender reported, like the direct kernel
        bad = u.clone()
        bad[N - 1, 0] = lo
        bad[136_000, 0] = hi + 1.0
        bad[138_000, 0] = lo - 1.0
        flag = _cuda.new_flag(u.device)
        _cuda.eval_points(ds, bad, 1, 1, N, flag=flag)
        assert int(flag.item()) == 136_000


def test_record_mode_multi_chunk_overlap_bit_identical(option):
    """More than one 4 Mi-point chunk: the sort / un-permute passes of neighbouring chunks run on a second stream under
    the evaluation (double-buffered workspace halves, event fork / join).  Same bits with and without the overlap and
    as the direct kernel; the first out-of-domain index survives the chunking."""
    bspy, _cuda, O, _ = _mods()
    from bspy_b200._spline_evaluation import device_spline
    rng = np.random.default_rng(41)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    s = bspy.Spline(3, 3, (4, 4, 4), (18, 18, 18), [K(4, 18) for _ in range(3)], rng.standard_normal((3, 18, 18, 18)))
    ds = device_spline(s)
    N = 2 * (1 << 22) + 70_001                                     # three chunks, ragged sparse tail
    pts = torch.rand((N, 3), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    ref = _cuda.eval_points(ds, pts, 3, 1, N, binned=False, values=True, jacobian=True)
    # default dispatch: cell polynomials (value + jacobian requests) -- strict bar against the recurrence, every point
    a = _cuda.eval_points(ds, pts, 3, 1, N, binned=True, values=True, jacobian=True)
    assert _close_t(a["values"], ref["values"]) and _close_t(a["jacobian"], ref["jacobian"])
    assert not torch.equal(a["jacobian"], ref["jacobian"]), "cell polynomials not in use"
    # BIN_PERM=1: the kernel gathers its points through sorted (cell key, index) pairs instead of reading point records: the
    # same bits
    option("BIN_PERM", 1)
    b = _cuda.eval_points(ds, pts, 3, 1, N, binned=True, values=True, jacobian=True)
    option("BIN_PERM", None)
    assert torch.equal(a["values"], b["values"]) and torch.equal(a["jacobian"], b["jacobian"])
    ra, _ = _cuda.eval_points_aos(ds, pts, 3, 1, N, jacobian=True)
    assert torch.equal(ra[:, :3].T, a["values"]) and torch.equal(ra[:, 3:12].T.reshape(3, 3, N), a["jacobian"])
    del a, b, ra
    option("CELL_POLY", 0)
    for flag in ("1", "0", "v1"):
        option("BIN_OVERLAP", 1 if flag == "v1" else int(flag))
        option("EXP_A", 1 if flag == "v1" else None)               # v1: first-generation staged kernel (per-lane span records)
        a = _cuda.eval_points(ds, pts, 3, 1, N, binned=True, values=True, jacobian=True)
        torch.cuda.synchronize()
        assert torch.equal(a["values"], ref["values"]) and torch.equal(a["jacobian"], ref["jacobian"]), flag
        del a
    option("BIN_OVERLAP", None)
    # a spline whose cell polynomials fail the validation of their build (an empty LAST span in the first variable: the
    # recurrence gives inf / NaN there, which a polynomial piece cannot reproduce entry for entry): the device flag sends the
    # whole call to the recurrence kernels
    kz = np.array(s.knots[0]); kz[18:] = kz[18] + np.array([0.0, 0.1, 0.2, 0.3]); kz[17] = kz[18]   # unclamped right end, k[nCoef-1] == k[nCoef]
    rough = bspy.Spline(3, 3, (4, 4, 4), (18, 18, 18), [kz, np.array(s.knots[1]), np.array(s.knots[2])], np.array(s.coefs))
    dr = device_spline(rough)
    n2 = (1 << 21) + 333
    rec0, _ = _cuda.eval_points_aos(dr, pts[:n2], 3, 1, n2, jacobian=True)
    option("CELL_POLY", None)
    rec1, _ = _cuda.eval_points_aos(dr, pts[:n2], 3, 1, n2, jacobian=True)
    option("CELL_POLY", 0)
    assert torch.equal(rec0, rec1), "fallback to the recurrence kernels"
    # the same for the 4-variate nDep-6 shape, whose fallback is the two-points-per-thread recurrence kernel on even-padded
    # records (launched with a small grid behind the gate: it strides over the pairs)
    k4 = [K(3, 10), K(3, 10), K(3, 9), K(3, 10)]
    k4[0][10:] = k4[0][10] + np.array([0.0, 0.1, 0.2]); k4[0][9] = k4[0][10]
    man = bspy.Spline(4, 6, (3, 3, 3, 3), (10, 10, 9, 10), k4, rng.standard_normal((6, 10, 10, 9, 10)))
    dm = device_spline(man)
    p4 = torch.rand((n2, 4), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    m0, _ = _cuda.eval_points_aos(dm, p4, 4, 1, n2, jacobian=True)
    option("CELL_POLY", None)
    m1, _ = _cuda.eval_points_aos(dm, p4, 4, 1, n2, jacobian=True)
    option("CELL_POLY", 0)
    assert torch.equal(m0, m1), "fallback to the recurrence pair kernel"
    bad = pts.clone()
    bad[(1 << 22) + 17, 2] = 1.25
    bad[2 * (1 << 22) + 5, 0] = -0.5
    f = _cuda.new_flag(pts.device)
    _cuda.eval_points(ds, bad, 3, 1, N, flag=f, binned=True)
    assert int(f.item()) == (1 << 22) + 17


def test_batch_api_variants():
    """SplineBatch: per-spline knots on the grid path, indices on normals, shard(), spline(i), host and device
    inputs; bspline_values_batch; evaluate_grid for a curve."""
    bspy, _, O, _ = _mods()
    rng = np.random.default_rng(31)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    S = 7
    ku = np.stack([K(3, 6) for _ in range(S)])
    kv = K(4, 9)                                            # shared in v, per-spline in u
    coefs = rng.standard_normal((S, 3, 6, 9))
    batch = bspy.SplineBatch(2, 3, (3, 4), (6, 9), [ku, kv], coefs)
    ua, va = np.sort(rng.uniform(0, 1, 21)), np.sort(rng.uniform(0, 1, 150))
    r = batch.evaluate_grid(ua, va, jacobian=True, normal=True, indices=(2, 0))
    assert r.values.shape == (S, 3, 21, 150) and r.jacobian.shape == (S, 3, 2, 21, 150) and r.normal.shape == (S, 2, 21, 150)
    uv = np.stack([m.reshape(-1) for m in np.meshgrid(ua, va, indexing="ij")], axis=1)
    for i in (0, 3, 6):
        si = batch.spline(i)
        assert np.array_equal(si.knots[0], ku[i]) and np.array_equal(si.coefs, coefs[i])
        so = O.OracleSpline.of(si)
        assert close(r.values[i].reshape(3, -1).T, O.evaluate_vec(so, uv))
        assert close(np.transpose(r.jacobian[i].reshape(3, 2, -1), (2, 0, 1)), O.jacobian_vec(so, uv), atol=1e-12)
        assert close(r.normal[i].reshape(2, -1).T, O.normal_vec(so, uv, True, (2, 0)))
    half = batch.shard(1, 2)
    assert len(half) == 3 and torch.equal(half.coefs, batch.coefs[4:7])
    rd = half.evaluate_grid(torch.from_numpy(ua).cuda(), torch.from_numpy(va).cuda())
    assert rd.values.is_cuda and np.array_equal(rd.values.cpu().numpy(), r.values[4:7])
    with pytest.raises(ValueError, match="outside domain: spline"):
        batch.evaluate_grid(ua, np.append(va, 1.5))
    # from_splines detects shared knots and refuses mixed shapes
    b2 = bspy.SplineBatch.from_splines([batch.spline(0), batch.spline(1)])
    assert b2.knots[0].dim() == 2 and b2.knots[1].dim() == 1
    with pytest.raises(ValueError):
        bspy.SplineBatch.from_splines([batch.spline(0), bspy.Spline(2, 3, (3, 3), (6, 9), [ku[0], K(3, 9)], coefs[0])])
    # vectorised bspline_values: numpy in -> numpy out, CUDA in -> CUDA out, given spans
    kn = K(5, 12)
    u = rng.uniform(0, 1, 300)
    ix, B = bspy.Spline.bspline_values_batch(None, kn, 5, u, 2, True)
    ixo, Bo = O.basis_vec(kn, 5, u, 2, True)
    assert np.array_equal(ix, ixo) and np.array_equal(B, Bo)
    ix2, B2 = bspy.Spline.bspline_values_batch(ix, kn, 5, torch.from_numpy(u).cuda(), 2, True)
    assert B2.is_cuda and np.array_equal(B2.cpu().numpy(), Bo)
    # grid entry for a curve = the scattered path on the axis
    c = CASES[3]
    s = _spline(c)
    g = s.evaluate_grid(c.uvw[:, 0], jacobian=True)
    assert g.values.shape == (c.nDep, c.uvw.shape[0]) and close(g.values.T, c["values"]) and close(g.jacobian[:, 0].T, c["jacobian"][:, :, 0])


def _curvature_close(k, want, uncertainty):
    """|k - want| <= 1e-13 + 1e-12 |want| + the uncertainty of the reference's own value (oracle.curvature_uncertainty_vec:
    how far the formula moves when its derivative inputs move by their rounding uncertainty); NaN / inf must match"""
    k, want, uncertainty = np.asarray(k), np.asarray(want), np.asarray(uncertainty)
    fin = np.isfinite(want)
    if not np.array_equal(np.isnan(k), np.isnan(want)):
        return False
    with np.errstate(all="ignore"):
        return bool(np.all(np.abs(k - want)[fin] <= (1e-13 + 1e-12 * np.abs(want) + uncertainty)[fin]))


def test_curvature_vs_reference():
    """batched curvature (curves nDep 1/2/3, surfaces nDep 3 and graph of a scalar function) in one
    fused pass against Spline.curvature of the reference.  Bar: 1e-12 relative (+1e-13) plus the uncertainty the reference's
    own value has at the point; on the well-conditioned samples (uncertainty below a tenth of the strict bar: most of them)
    that IS the strict bar."""
    _, _cuda, O, _ = _mods()
    ref = load_npz("ref_curvature.npz")
    by_tag = {c.tag: c for c in CASES}
    strict = total = 0
    for tag, want in ref.items():
        c = by_tag[tag]
        s = _spline(c)
        unc = O.curvature_uncertainty_vec(_ospline(c), c.uvw)
        k = s.curvature_points(c.uvw)
        assert k.shape == want.shape
        assert _curvature_close(k, want, unc), tag
        fin = np.isfinite(want)
        strict += int((fin & (unc <= 0.1 * (1e-13 + 1e-12 * np.abs(want)))).sum())
        total += int(fin.sum())
        kd = s.curvature_points(torch.from_numpy(c.uvw).cuda())
        assert kd.is_cuda and np.array_equal(kd.cpu().numpy(), k, equal_nan=True)
        p = int(np.flatnonzero(fin)[5])
        assert _curvature_close([s.curvature(c.uvw[p] if c.nInd > 1 else c.uvw[p, 0])], [want[p]], [unc[p]])
        # the composed path (derivative passes + bspy_cuda_curvature) stays available for shapes the fused kernel does not take
        d1 = s.evaluate_points(c.uvw, values=False, with_respect_to=[1] + [0] * (c.nInd - 1)).derivative
        assert d1.shape == (c.nDep, c.uvw.shape[0])
    assert strict >= 0.6 * total, (strict, total)                   # most samples are effectively held to the strict bar


def test_curvature_known_answers():
    """The reference's own curvature test (tests/bspy_test.py:693-700) on the GPU path: the section curve has curvature 1 at
    u = 0 and 2 at u = 1 (2e-15), mySurface has Gaussian curvature 1.024 at (0.25, 0.5) (1e-14); plus the reference's values
    on the planar projection at 101 parameters and on a 7 x 7 grid of the surface."""
    bspy, _, O, _ = _mods()
    a = load_npz("ref_curvature_kat.npz")

    def spline(tag):
        nInd = int(a[f"{tag}/nInd"])
        return bspy.Spline(nInd, int(a[f"{tag}/nDep"]), tuple(a[f"{tag}/order"]), tuple(a[f"{tag}/nCoef"]),
                           [a[f"{tag}/knots{i}"] for i in range(nInd)], a[f"{tag}/coefs"])

    section = spline("section")
    assert abs(section.curvature(0.0) - 1.0) < 2.0e-15 and abs(section.curvature(1.0) - 2.0) < 2.0e-15
    k = section.curvature_points(a["section/u"])
    assert np.all(np.abs(k - a["section/curvature"]) < 2.0e-15)
    planar = spline("planar")
    kp = planar.curvature_points(a["planar/u"])
    so = O.OracleSpline.of(planar)
    assert _curvature_close(kp, a["planar/curvature"], O.curvature_uncertainty_vec(so, a["planar/u"][:, None]))
    surface = spline("surface")
    assert abs(surface.curvature([0.25, 0.5]) - 1.024) < 1.0e-14
    ks = surface.curvature_points(a["surface/uv"])
    assert abs(ks[0] - 1.024) < 1.0e-14
    assert _curvature_close(ks, a["surface/curvature"], O.curvature_uncertainty_vec(O.OracleSpline.of(surface), a["surface/uv"]))


def test_grid_edge_cases():
    """Unsorted axes (a 16-column step may mix any knot spans), order-1 / order-8 variables, empty axes, odd sizes
    that defeat the 32- and 16-byte store paths."""
    bspy, _, O, _ = _mods()
    rng = np.random.default_rng(41)

    def K(o, n):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    for order, nCoef, nDep, shape in (((4, 4), (12, 40), 3, (33, 301)), ((1, 8), (5, 11), 2, (17, 130)), ((8, 1), (9, 6), 1, (9, 18)),
                                      ((3, 3, 4), (6, 5, 30), 3, (5, 6, 203)), ((1, 2, 8), (3, 4, 10), 2, (3, 5, 66)),
                                      ((3, 2, 4), (4, 3, 90), 3, (4, 3, 300))):       # > 32 coefficient columns per chunk
        nInd = len(order)
        s = bspy.Spline(nInd, nDep, order, nCoef, [K(o, n) for o, n in zip(order, nCoef)], rng.standard_normal((nDep, *nCoef)))
        axes = [rng.uniform(0, 1, n) for n in shape]           # NOT sorted
        axes[-1][:3] = (1.0, 0.0, s.knots[-1][order[-1]] if nCoef[-1] > order[-1] else 0.5)
        r = s.evaluate_grid(*axes, jacobian=True)
        uvw = np.stack([m.reshape(-1) for m in np.meshgrid(*axes, indexing="ij")], axis=1)
        so = O.OracleSpline.of(s)
        assert close(r.values.reshape(nDep, -1).T, O.evaluate_vec(so, uvw)), (order, "values")
        assert close_cond(np.transpose(r.jacobian.reshape(nDep, nInd, -1), (2, 0, 1)), O.jacobian_vec(so, uvw),
                          O.jacobian_abs_vec(so, uvw)), (order, "jacobian")
        empty = s.evaluate_grid(*([axes[0][:0]] + axes[1:]), jacobian=True)
        assert empty.values.shape == (nDep, 0, *shape[1:]) and empty.jacobian.shape == (nDep, nInd, 0, *shape[1:])


def test_contract_vs_reference():
    """Spline.contract against the unmodified reference's results, and the reference's own test
    (tests/bspy_test.py:639-652: surface(.25, u) == contracted(u) to 2.5 eps) on its golden surface."""
    bspy, _, O, _ = _mods()
    from golden_io import _spline_from
    a = load_npz("ref_block.npz")
    for name in a["contract/names"]:
        tag, j = str(name).split("/")
        s = bspy.Spline(*_spline_from(a, f"contract/{tag}"))
        uvw = [None if np.isnan(v) else float(v) for v in a[f"contract/{tag}/{j}/uvw"]]
        c = s.contract(uvw)
        nInd, nDep, order, nCoef, knots, coefs = _spline_from(a, f"contract/{tag}/{j}/result")
        assert (c.nInd, c.nDep, tuple(c.order), tuple(c.nCoef)) == (nInd, nDep, order, nCoef)
        assert all(np.array_equal(x, y) for x, y in zip(c.knots, knots))
        assert close(np.asarray(c.coefs), coefs.reshape(np.asarray(c.coefs).shape)), name
    t = load_npz("ref_tables.npz")
    surf = bspy.Spline(2, 3, t["surface/order"], t["surface/coefs"].shape[1:], [t["surface/knots0"], t["surface/knots1"]], t["surface/coefs"])
    assert surf.contract([None, None]) is surf
    for fixed, at in (([.25, None], 0), ([None, .75], 1)):
        c = surf.contract(fixed)
        worst_err = 0.0
        for u in np.linspace(0, 1, 21):
            full = [.25, u] if at == 0 else [u, .75]
            d = surf(full) - c([u])
            worst_err = max(worst_err, float(np.sqrt(d @ d)))
        assert worst_err <= 2.5 * EPS
    with pytest.raises(ValueError, match="outside domain"):
        surf.contract([1.5, None])


def test_spline_block_vs_reference():
    """SplineBlock.evaluate / derivative / jacobian / normal / contract: single-point API and the
    batched evaluate_points against outputs of the unmodified reference (maps, row sums, nInd > nDep normal)."""
    bspy, _cuda, O, _ = _mods()
    from golden_io import block_members
    a = load_npz("ref_block.npz")
    for tag in ("A", "B", "C"):
        rows = [[(m, bspy.Spline(*sp)) for m, sp in row] for row in block_members(a, tag)]
        b = bspy.SplineBlock(rows)
        assert [b.nInd, b.nDep] == list(a[f"block/{tag}/nIndnDep"])
        assert np.array_equal(b.domain(), a[f"block/{tag}/domain"])
        uvw = a[f"block/{tag}/uvw"]
        has_normal = f"block/{tag}/normal_unit" in a
        r = b.evaluate_points(uvw, jacobian=True, normal=has_normal, with_respect_to=list(a[f"block/{tag}/wrt"][0]))
        assert close(r.values.T, a[f"block/{tag}/values"])
        assert close(np.transpose(r.jacobian, (2, 0, 1)), a[f"block/{tag}/jacobian"])
        assert close(r.derivative.T, a[f"block/{tag}/deriv0"])
        r2 = b.evaluate_points(torch.from_numpy(uvw).cuda(), values=False, with_respect_to=list(a[f"block/{tag}/wrt"][1]))
        assert r2.derivative.is_cuda and close(r2.derivative.cpu().numpy().T, a[f"block/{tag}/deriv1"])
        if has_normal:
            assert close(r.normal.T, a[f"block/{tag}/normal_unit"])
            assert close(b.evaluate_points(uvw, values=False, normal=True, normalize=False).normal.T, a[f"block/{tag}/normal_raw"])
            assert close(b.evaluate_points(uvw, values=False, normal=True, indices=(0, 2)).normal.T, a[f"block/{tag}/normal_idx"])
        for p in (0, 1, 7):
            assert close(b(uvw[p]), a[f"block/{tag}/values"][p])
            assert close(b.jacobian(uvw[p]), a[f"block/{tag}/jacobian"][p])
            assert close(b.derivative(list(a[f"block/{tag}/wrt"][0]), uvw[p]), a[f"block/{tag}/deriv0"][p])
            if has_normal:
                assert close(b.normal(uvw[p]), a[f"block/{tag}/normal_unit"][p])
                assert close(b.normal(uvw[p], False, (0, 2)), a[f"block/{tag}/normal_raw"][p][[0, 2]])
            else:
                with pytest.raises(ValueError, match="one different"):
                    b.normal(uvw[p])
        bad = uvw.copy(); bad[5, 0] = a[f"block/{tag}/domain"][0, 1] + 1.0
        with pytest.raises(ValueError, match="outside domain"):
            b.evaluate_points(bad)
    # contracted block
    rows = [[(m, bspy.Spline(*sp)) for m, sp in row] for row in block_members(a, "A")]
    cb = bspy.SplineBlock(rows).contract([None if np.isnan(v) else float(v) for v in a["block/A/contract_uvw"]])
    assert [cb.nInd, cb.nDep] == list(a["block/A/contract_nIndnDep"])
    assert close(cb.evaluate_points(a["block/A/contract_pts"]).values.T, a["block/A/contract_values"])
    # constructor errors of the reference (bspy/spline_block.py:83, 91, 96)
    F, G = rows[0][0][1], rows[0][1][1]
    with pytest.raises(ValueError, match="same nDep"):
        bspy.SplineBlock([[F, rows[1][0][1]]])
    with pytest.raises(ValueError, match="Multiple splines in the same row"):
        bspy.SplineBlock([[([0, 1, 2], F), ([2, 3], G)]])
    # a bigger batch against the oracle
    rng = np.random.default_rng(9)
    ob = O.OracleBlock([[(m, O.OracleSpline.of(sp)) for m, sp in row] for row in rows])
    dom = a["block/A/domain"]
    big = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.uniform(0, 1, (20_000, 5))
    r = bspy.SplineBlock(rows).evaluate_points(big, jacobian=True)
    assert close(r.values.T, ob.evaluate_vec(big)) and close(np.transpose(r.jacobian, (2, 0, 1)), ob.jacobian_vec(big))


def test_collocation_rows_and_normal_sampling_vs_reference():
    """the collocation matrix of Spline.least_squares (rows assembled from the reference's own
    bspline_values, repeated parameters raising the derivative order) bit-for-bit, and the un-normalised normal
    sampling of normal_spline through the grid kernels."""
    bspy, _cuda, O, _ = _mods()
    a = load_npz("ref_block.npz")
    for tag in ("o4", "o3", "o6"):
        knots, order, u = a[f"colloc/{tag}/knots"], int(a[f"colloc/{tag}/order"]), a[f"colloc/{tag}/u"]
        sp, A = bspy.Spline.collocation_matrix(knots, order, u)                     # derivative orders from the runs of equal u
        assert np.array_equal(A, a[f"colloc/{tag}/A"]), tag
        sp2, A2 = bspy.Spline.collocation_matrix(torch.from_numpy(knots).cuda(), order, torch.from_numpy(u).cuda(),
                                                 derivativeOrders=a[f"colloc/{tag}/derivs"])
        assert A2.is_cuda and np.array_equal(A2.cpu().numpy(), a[f"colloc/{tag}/A"])
        assert np.array_equal(sp, sp2.cpu().numpy())
        rows = np.arange(len(u))
        assert all(np.all(A[r, :sp[r] - order] == 0) and np.all(A[r, sp[r]:] == 0) for r in rows)
    from golden_io import _spline_from
    surf = bspy.Spline(*_spline_from(a, "nsample/surf"))
    g = surf.evaluate_grid(a["nsample/gu"], a["nsample/gv"], values=False, normal=True, normalize=False)
    assert close(np.transpose(g.normal, (1, 2, 0)), a["nsample/normals"])


def test_random_shapes_vs_oracle():
    """Seeded sweep over random shapes (nInd 1-4, orders 1-6, nDep 1-6, clamped / unclamped / repeated interior knots):
    spans bit-exact, values / jacobians / one mixed partial / normals against the oracle -- exercises the fixed-shape
    kernels where a shape is compiled and the any-shape kernel elsewhere, scattered and grid entry points."""
    bspy, _cuda, O, _ = _mods()
    rng = np.random.default_rng(4242)

    def knots(o, n, style):
        w = rng.uniform(0.25, 1.75, n - o + 1)
        inner = np.concatenate(([0.0], np.cumsum(w))); inner /= inner[-1]
        if style == 2 and n - o + 1 >= 3 and o >= 3:
            inner[2] = inner[1]                                    # double interior knot (legal for order >= 3)
        if style == 1:
            left = -np.cumsum(rng.uniform(0.05, 0.3, o - 1))[::-1]
            right = 1.0 + np.cumsum(rng.uniform(0.05, 0.3, o - 1))
            return np.concatenate((left, inner, right))
        return np.concatenate((np.zeros(o - 1), inner, np.ones(o - 1)))

    for trial in range(48):
        nInd = int(rng.integers(1, 5))
        nDep = int(rng.integers(1, 7))
        order = tuple(int(rng.integers(1, 7 if nInd < 3 else 5)) for _ in range(nInd))
        nCoef = tuple(int(o + rng.integers(0, 6)) for o in order)
        kk = [knots(o, n, int(rng.integers(0, 3))) for o, n in zip(order, nCoef)]
        coefs = rng.standard_normal((nDep, *nCoef))
        s = bspy.Spline(nInd, nDep, order, nCoef, kk, coefs, {"negateNormal": bool(trial % 2)})
        so = O.OracleSpline.of(s)
        dom = s.domain()
        N = 700
        pts = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.uniform(0, 1, (N, nInd))
        pts[0], pts[1] = dom[:, 0], dom[:, 1]
        for i in range(nInd):                                      # knots of every variable as parameters
            inside = np.unique(kk[i][(kk[i] >= dom[i, 0]) & (kk[i] <= dom[i, 1])])
            pts[2:2 + len(inside), i] = inside
        wrt = [int(rng.integers(0, 3)) for _ in range(nInd)]
        want_normal = abs(nInd - nDep) == 1
        r = s.evaluate_points(pts, jacobian=True, spans=True, with_respect_to=wrt, normal=want_normal, normalize=False)
        ref, sp = O.evaluate_vec(so, pts, return_spans=True)
        tag = (trial, nInd, nDep, order, nCoef)
        assert np.array_equal(r.spans.T, sp), tag
        assert close(r.values.T, ref), tag
        assert close_cond(np.transpose(r.jacobian, (2, 0, 1)), O.jacobian_vec(so, pts), O.jacobian_abs_vec(so, pts)), tag
        assert close_cond(r.derivative.T, O.derivative_vec(so, wrt, pts), O.derivative_abs_vec(so, wrt, pts)), tag
        if want_normal:
            assert close_cond(r.normal.T, O.normal_vec(so, pts, False), O.normal_abs_vec(so, pts)), tag
        if nInd in (2, 3):                                         # tensor-grid entry (DMMA kernels for these shapes)
            axes = [np.sort(pts[:17 + 3 * i, i]) for i in range(nInd)]
            g = s.evaluate_grid(*axes, jacobian=True)
            mesh = np.stack([m.reshape(-1) for m in np.meshgrid(*axes, indexing="ij")], axis=1)
            assert close(g.values.reshape(nDep, -1).T, O.evaluate_vec(so, mesh)), tag
            assert close_cond(np.transpose(g.jacobian.reshape(nDep, nInd, -1), (2, 0, 1)), O.jacobian_vec(so, mesh),
                              O.jacobian_abs_vec(so, mesh)), tag


def test_grid_float32_outputs():
    """float32 outputs of the surface grid kernels (the viewer's tessellation buffers) are the float64
    results rounded to nearest, element for element -- single surfaces and batches, ragged and unaligned sizes, orders
    above 4, all three output kinds; other shapes refuse the flag."""
    bspy, _cuda, O, _ = _mods()
    a = load_npz("teapot.npz")
    kn = a["knots"]
    patches = [bspy.Spline(2, 3, (4, 4), (4, 4), (kn, kn), a["coefs"][p]) for p in range(6)]
    batch = bspy.SplineBatch.from_splines(patches)
    for nU, nV in ((64, 256), (33, 77), (8, 4)):
        u = torch.linspace(0, 1, nU, dtype=torch.float64, device="cuda")
        v = torch.linspace(0, 1, nV, dtype=torch.float64, device="cuda")
        r64 = batch.evaluate_grid(u, v, jacobian=True, normal=True)
        r32 = batch.evaluate_grid(u, v, jacobian=True, normal=True, dtype=np.float32)
        for x64, x32 in ((r64.values, r32.values), (r64.jacobian, r32.jacobian), (r64.normal, r32.normal)):
            assert x32.dtype == torch.float32 and x32.shape == x64.shape
            assert torch.equal(torch.nan_to_num(x32, nan=-7.0), torch.nan_to_num(x64.to(torch.float32), nan=-7.0)), (nU, nV)
    c = [c for c in CASES if c.tag == "surf_25_d1"][0]                      # orders (2, 5), nDep 1 (nInd > nDep normal)
    s = _spline(c)
    g = [np.linspace(*s.domain()[i], 19 + 6 * i) for i in range(2)]
    r64 = s.evaluate_grid(*g, jacobian=True, normal=True)
    r32 = s.evaluate_grid(*g, jacobian=True, normal=True, dtype=np.float32)
    assert r32.values.dtype == np.float32
    assert np.array_equal(r32.values, r64.values.astype(np.float32)) and np.array_equal(r32.jacobian, r64.jacobian.astype(np.float32))
    assert np.array_equal(r32.normal, r64.normal.astype(np.float32), equal_nan=True)
    vol = _spline([c for c in CASES if c.tag == "vol_444_d3"][0])
    with pytest.raises(NotImplementedError):
        vol.evaluate_grid(*[np.linspace(0, 1, 5)] * 3, dtype=np.float32)


@pytest.mark.parametrize("cfg,scale", [("cfg2", 1.0), ("cfg1", 1.0), ("cfg3", 0.12), ("cfg4", 0.06), ("cfg4_soa", 0.06), ("cfg5", 0.03),
                                       ("cfg5_soa", 0.03), ("grid3", 0.5)])
def test_bench_shapes_at_bench_tile_paths(cfg, scale):
    """The BASELINE configurations through the public API and the DEFAULT dispatch at sizes that reach the same kernels and
    tile paths as the bench: the teapot at the full 2048 x 2048 grid per patch (border rows / columns in full + random interior
    of 8 patches incl. lid and bottom), 1.2e5 curves of config 3, 6e6 points of config 4 and 3.75e6 points of config 5
    (cell-sorted pipeline, array-of-structs and struct-of-arrays outputs, value + jacobian), the 256^3 volume grid -- each
    against the C restatement of the reference on a >= 1e5-point sample of the outputs: strict bar, spans ==.  This is the
    parity block of bench.py (same code), so a green test here is what the bench line's "parity" reports."""
    import bench
    wl = bench.CONFIGS[cfg]()
    wl.setup(torch.device("cuda", torch.cuda.current_device()), 0, scale)
    wl.step()
    torch.cuda.synchronize()
    assert wl.flags_ok()
    rep = wl.parity()
    assert rep["ok"] and rep["spans_equal"] and rep["nan_match"] and rep["worst_ratio"] <= 1.0 and rep["n"] >= 100_000, rep
    wl.teardown()
    torch.cuda.empty_cache()
