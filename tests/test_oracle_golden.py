"""Pins the oracle (oracle/bspy_oracle.py, both tiers) to the reference:
the reference's own golden tables and outputs generated from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_io import close, load_cases, load_npz
from oracle import bspy_oracle as O

CASES = load_cases()
EPS = np.finfo(float).eps


def _spline(c):
    return O.OracleSpline(c.nInd, c.nDep, c.order, c.nCoef, c.knots, c.coefs, c.metadata)


def test_truth_curve_table():
    """reference tests/bspy_test.py:743-748 (error <= eps)."""
    t = load_npz("ref_tables.npz")
    s = O.OracleSpline(1, 2, t["curve/order"], (5,), [t["curve/knots0"]], t["curve/coefs"])
    tab = t["curve/table"]
    vec = O.evaluate_vec(s, tab[:, :1])
    assert np.sqrt(((vec - tab[:, 1:]) ** 2).sum(axis=1)).max() <= EPS
    for u, x, y in tab[::7]:
        assert np.hypot(*(O.evaluate_pt(s, [u]) - (x, y))) <= EPS


def test_truth_surface_table():
    """reference tests/bspy_test.py:749-757 (error <= 2.5 eps; v outer, u inner)."""
    t = load_npz("ref_tables.npz")
    s = O.OracleSpline(2, 3, t["surface/order"], (4, 5), [t["surface/knots0"], t["surface/knots1"]], t["surface/coefs"])
    g = np.linspace(0, 1, 21)
    uv = np.array([(u, v) for v in g for u in g])
    err = np.sqrt(((O.evaluate_vec(s, uv) - t["surface/table"]) ** 2).sum(axis=1)).max()
    assert err <= 2.5 * EPS
    assert np.sqrt(((O.evaluate_pt(s, uv[200]) - t["surface/table"][200]) ** 2).sum()) <= 2.5 * EPS


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.tag)
def test_spans_and_basis_bit_exact(c):
    for i in range(c.nInd):
        u = c.uvw[:, i]
        assert np.array_equal(O.span_vec(c.knots[i], c.order[i], u), c["spans"][:, i])
        for d in range(c.order[i] + 2):
            for taylor in (False, True):
                key = f"basis{i}_d{d}{'t' if taylor else ''}"
                if not c.has(key):
                    continue
                ix, B = O.basis_vec(c.knots[i], c.order[i], u, d, taylor)
                assert np.array_equal(ix, c["spans"][:, i])
                assert np.array_equal(B, c[key], equal_nan=True), (c.tag, key)
        # scalar tier on a subsample
        for p in range(0, len(u), 9):
            ix, b = O.basis_pt(None, c.knots[i], c.order[i], u[p], 1, False)
            assert ix == c["spans"][p, i]
            assert np.array_equal(b, c[f"basis{i}_d1"][p], equal_nan=True)


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.tag)
def test_values_derivatives_jacobian(c):
    s = _spline(c)
    v, spans = O.evaluate_vec(s, c.uvw, return_spans=True)
    assert np.array_equal(spans, c["spans"])
    assert close(v, c["values"], rtol=1e-14, atol=1e-15)
    assert close(O.jacobian_vec(s, c.uvw), c["jacobian"], rtol=1e-14, atol=1e-14)
    for w in c.meta["wrt"]:
        ref = c["deriv_" + "_".join(map(str, w))]
        got = O.derivative_vec(s, w, c.uvw)
        assert close(got, ref, rtol=1e-13, atol=1e-13 * max(1.0, np.nanmax(np.abs(ref)))), (c.tag, w)
        if w[0] >= c.order[0]:
            assert not got.any()
    with np.errstate(all="ignore"):
        for p in range(0, c.uvw.shape[0], 17):
            assert close(O.evaluate_pt(s, c.uvw[p]), c["values"][p], rtol=1e-15, atol=0)
            assert close(O.jacobian_pt(s, c.uvw[p]), c["jacobian"][p], rtol=1e-15, atol=0)


@pytest.mark.parametrize("c", [c for c in CASES if c.meta["normal"]], ids=lambda c: c.tag)
def test_normals(c):
    s = _spline(c)
    idx = c.meta["normal_indices"]
    scale = max(1.0, float(np.nanmax(np.abs(c["normal_raw"]))))
    assert close(O.normal_vec(s, c.uvw, False), c["normal_raw"], rtol=1e-13, atol=1e-14 * scale)
    assert close(O.normal_vec(s, c.uvw, True), c["normal_unit"], rtol=1e-12, atol=1e-13)
    assert close(O.normal_vec(s, c.uvw, False, idx), c["normal_idx_raw"], rtol=1e-13, atol=1e-14 * scale)
    assert close(O.normal_vec(s, c.uvw, True, idx), c["normal_idx_unit"], rtol=1e-12, atol=1e-13)
    with np.errstate(all="ignore"):
        for p in range(0, c.uvw.shape[0], 23):
            assert close(O.normal_pt(s, c.uvw[p]), c["normal_unit"][p], rtol=1e-15, atol=0)
            assert close(O.normal_pt(s, c.uvw[p], False, idx), c["normal_idx_raw"][p], rtol=1e-15, atol=0)


def test_teapot_grid():
    t = load_npz("teapot.npz")
    g = t["grid"]
    uv = np.array([(u, v) for u in g for v in g])
    for p in range(0, 32, 5):
        s = O.OracleSpline(2, 3, (4, 4), (4, 4), (t["knots"], t["knots"]), t["coefs"][p])
        assert close(O.evaluate_vec(s, uv).T.reshape(3, 9, 9), t["values"][p], rtol=1e-14, atol=1e-15)
        J = O.jacobian_vec(s, uv)
        assert close(J[:, :, 0].T.reshape(3, 9, 9), t["du"][p], rtol=1e-14, atol=1e-14)
        assert close(O.normal_vec(s, uv).T.reshape(3, 9, 9), t["normal"][p])


def test_domain_errors():
    c = CASES[3]
    s = _spline(c)
    with pytest.raises(ValueError, match="outside domain"):
        O.evaluate_pt(s, [1.5])
    with pytest.raises(ValueError, match="Incorrect number"):
        O.evaluate_pt(s, [0.1, 0.2])
    assert O.check_domain_vec(s, np.array([[0.5], [1.5], [-1.0]])) == 1
    assert O.check_domain_vec(s, np.array([[0.5], [np.nan]])) == -1


def test_curvature_vs_reference():
    """oracle.curvature_vec against Spline.curvature of the reference (bspy/_spline_evaluation.py:80-107)."""
    ref = load_npz("ref_curvature.npz")
    by_tag = {c.tag: c for c in CASES}
    for tag, want in ref.items():
        c = by_tag[tag]
        got = O.curvature_vec(_spline(c), c.uvw)
        ok = np.isfinite(want)
        assert np.array_equal(np.isnan(got), np.isnan(want)) or tag == "tomsnasty0"
        # curvature divides by |f'|^3 or EG-F^2 and subtracts nearly equal products: compare on a relative scale
        assert np.allclose(got[ok], want[ok], rtol=1e-8, atol=1e-8 * max(1.0, np.nanmax(np.abs(want[ok])))), tag


def test_contract_and_block_oracle_vs_reference():
    """the oracle's contract / SplineBlock restatements against outputs of the unmodified reference
    (tests/golden/make_golden_block.py)."""
    from golden_io import _spline_from, block_members
    a = load_npz("ref_block.npz")
    for name in a["contract/names"]:
        tag, j = str(name).split("/")
        s = O.OracleSpline(*_spline_from(a, f"contract/{tag}"))
        uvw = [None if np.isnan(v) else float(v) for v in a[f"contract/{tag}/{j}/uvw"]]
        c = O.contract(s, uvw)
        nInd, nDep, order, nCoef, knots, coefs = _spline_from(a, f"contract/{tag}/{j}/result")
        assert (c.nInd, c.nDep, tuple(c.order), tuple(c.nCoef)) == (nInd, nDep, order, nCoef)
        assert all(np.array_equal(x, y) for x, y in zip(c.knots, knots))
        assert close(c.coefs, coefs.reshape(c.coefs.shape))
    for tag in ("A", "B", "C"):
        rows = [[(m, O.OracleSpline(*sp)) for m, sp in row] for row in block_members(a, tag)]
        b = O.OracleBlock(rows)
        assert [b.nInd, b.nDep] == list(a[f"block/{tag}/nIndnDep"])
        uvw = a[f"block/{tag}/uvw"]
        assert close(b.evaluate_vec(uvw), a[f"block/{tag}/values"])
        assert close(b.jacobian_vec(uvw), a[f"block/{tag}/jacobian"])
        for k in (0, 1):
            assert close(b.derivative_vec(list(a[f"block/{tag}/wrt"][k]), uvw), a[f"block/{tag}/deriv{k}"])
        if f"block/{tag}/normal_unit" in a:
            assert close(b.normal_vec(uvw), a[f"block/{tag}/normal_unit"])
            assert close(b.normal_vec(uvw, False), a[f"block/{tag}/normal_raw"])
            assert close(b.normal_vec(uvw, True, (0, 2)), a[f"block/{tag}/normal_idx"])
