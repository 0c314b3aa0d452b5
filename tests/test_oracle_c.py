"""Pins the C restatement (oracle/bspy_oracle.c) to the reference goldens.  CPU only."""
import numpy as np
import pytest

from golden_io import close, close_cond, load_cases, well_conditioned_subset
from oracle import bspy_oracle as O
from oracle import c_oracle as CO

CASES = load_cases()


@pytest.fixture(scope="module", autouse=True)
def _built():
    CO.build()


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.tag)
def test_c_spans_basis_bit_exact(c):
    for i in range(c.nInd):
        u = c.uvw[:, i]
        assert np.array_equal(CO.spans(c.knots[i], c.order[i], u), c["spans"][:, i])
        for d in range(c.order[i] + 2):
            for taylor in (False, True):
                key = f"basis{i}_d{d}{'t' if taylor else ''}"
                if c.has(key):
                    ix, B = CO.basis(c.knots[i], c.order[i], u, d, taylor)
                    assert np.array_equal(ix, c["spans"][:, i])
                    assert np.array_equal(B, c[key], equal_nan=True), (c.tag, key)


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.tag)
def test_c_eval(c):
    r = CO.evaluate(c, c.uvw, jacobian=True, spans=True, normal=c.meta["normal"], normalize=True)
    assert r["first_oob"] == -1
    assert np.array_equal(r["spans"], c["spans"])
    assert close(r["values"], c["values"], rtol=1e-14, atol=1e-15)
    s = O.OracleSpline(c.nInd, c.nDep, c.order, c.nCoef, c.knots, c.coefs, c.metadata)
    assert close_cond(r["jacobian"], c["jacobian"], O.jacobian_abs_vec(s, c.uvw), rtol=1e-14, atol=1e-14)
    for w in c.meta["wrt"]:
        ref = c["deriv_" + "_".join(map(str, w))]
        got = CO.evaluate(c, c.uvw, wrt=w, values=False)["deriv"]
        assert close_cond(got, ref, O.derivative_abs_vec(s, w, c.uvw), rtol=1e-14, atol=1e-14), w
    if c.meta["normal"]:
        Sn = O.normal_abs_vec(s, c.uvw)
        with np.errstate(all="ignore"):
            Su = (Sn.max(axis=1) / np.sqrt((c["normal_raw"] ** 2).sum(axis=1)))[:, None]
        assert close_cond(r["normal"], c["normal_unit"], Su, k=64)
        raw = CO.evaluate(c, c.uvw, values=False, normal=True, normalize=False)["normal"]
        assert close_cond(raw, c["normal_raw"], Sn, k=64)
        idx = c.meta["normal_indices"]
        sub = CO.evaluate(c, c.uvw, values=False, normal=True, normalize=True, indices=idx)["normal"][:, idx]
        ok = well_conditioned_subset(c["normal_raw"], idx)
        with np.errstate(all="ignore"):
            Si = (Sn[:, idx].max(axis=1) / np.sqrt((c["normal_idx_raw"] ** 2).sum(axis=1)))[:, None]
        assert close_cond(sub[ok], c["normal_idx_unit"][ok], Si[ok], k=64)


def test_c_oob_and_nan():
    c = CASES[3]
    r = CO.evaluate(c, np.array([[0.5], [2.0], [0.1]]))
    assert r["first_oob"] == 1
    r = CO.evaluate(c, np.array([[np.nan]]), spans=True)
    assert r["first_oob"] == -1 and r["spans"][0, 0] == c.nCoef[0] and np.isnan(r["values"]).all()
