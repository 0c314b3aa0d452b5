#!/usr/bin/env python
"""Golden fixtures for SURVEY 8(f) row 1 from the UNMODIFIED reference: ``Spline.contract``
(``bspy/_spline_operations.py:184-223``) and ``SplineBlock.evaluate / derivative / jacobian / normal / contract``
(``bspy/spline_block.py``), plus row 3: the collocation rows of ``Spline.least_squares`` and the normal sampling of
``normal_spline``.  Run in the build container only:

    python tests/golden/make_golden_block.py [--ref /root/reference]

Writes ``ref_block.npz`` (inputs and the reference's outputs; small, committed)."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, knots_nonuniform  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    bspy = import_reference(args.ref)
    from bspy.spline_block import SplineBlock
    Spline = bspy.Spline
    rng = np.random.default_rng(20261019)
    out = {}

    def synth(nInd, nDep, order, nCoef, knots=None):
        kk = knots if knots is not None else [knots_nonuniform(order[i], nCoef[i], rng) for i in range(nInd)]
        return Spline(nInd, nDep, order, nCoef, kk, rng.standard_normal((nDep, *nCoef)), {})

    def put_spline(tag, s):
        out[f"{tag}/shape"] = np.array([s.nInd, s.nDep, *s.order, *s.nCoef], np.int64)
        for i, k in enumerate(s.knots):
            out[f"{tag}/knots{i}"] = np.asarray(k, float)
        out[f"{tag}/coefs"] = np.ascontiguousarray(s.coefs, dtype=float)

    # ---- contract ------------------------------------------------------------------------------------------
    contract_cases = {
        "surf": (synth(2, 3, (3, 4), (6, 7)), [[0.25, None], [None, 0.75], [0.0, None], [None, 1.0], [0.3, 0.6]]),
        "vol": (synth(3, 3, (4, 4, 4), (6, 7, 5)), [[None, 0.4, None], [0.1, None, 0.9], [0.5, 0.5, None], [None, None, 0.2]]),
        "man": (synth(4, 2, (3, 2, 3, 2), (4, 3, 5, 3)), [[0.2, None, None, 0.7], [None, 0.5, 0.5, None], [0.1, 0.2, 0.3, None]]),
    }
    names = []
    for tag, (s, requests) in contract_cases.items():
        put_spline(f"contract/{tag}", s)
        for j, uvw in enumerate(requests):
            c = s.contract(uvw)
            out[f"contract/{tag}/{j}/uvw"] = np.array([np.nan if v is None else v for v in uvw])
            put_spline(f"contract/{tag}/{j}/result", c)
            names.append(f"{tag}/{j}")
    out["contract/names"] = np.array(names)

    # ---- blocks --------------------------------------------------------------------------------------------
    # A: F(u, v, w) + G(t, s) = 0 (2-D), h(u, t, w, s) = 0 (scalar), variables (s, t, u, v, w) = 0..4; same knots per
    #    variable so that the domains match (splines of one row may not share a variable: spline_block.py:90-91)
    kv = {i: None for i in range(5)}
    orders = {0: 3, 1: 2, 2: 4, 3: 3, 4: 4}
    ncoef = {0: 5, 1: 4, 2: 6, 3: 5, 4: 7}
    for i in range(5):
        kv[i] = knots_nonuniform(orders[i], ncoef[i], rng)

    def member(vars_, nDep):
        return Spline(len(vars_), nDep, [orders[i] for i in vars_], [ncoef[i] for i in vars_], [kv[i] for i in vars_],
                      rng.standard_normal((nDep, *[ncoef[i] for i in vars_])), {})
    F, G, h = member([2, 3, 4], 2), member([1, 0], 2), member([2, 1, 4, 0], 1)
    blockA = SplineBlock([[([2, 3, 4], F), ([1, 0], G)], [([2, 1, 4, 0], h)]])
    # B: nInd 3, nDep 2 (normal exists): a(u, v) + b(w), c(v, w); default maps for the first row
    a, b, c = member([0, 1], 1), member([2], 1), member([1, 2], 1)
    blockB = SplineBlock([[a, b], [([1, 2], c)]])
    # C: square system nInd 2 = nDep 2 built from one 2-D spline (a list of splines is one row)
    blockC = SplineBlock([member([0, 1], 2), member([2], 2)])
    for tag, block, members, maps in (("A", blockA, [F, G, h], [[2, 3, 4], [1, 0], [2, 1, 4, 0]]),
                                      ("B", blockB, [a, b, c], [[0, 1], [2], [1, 2]]),
                                      ("C", blockC, None, None)):
        if members is None:
            members = [s for row in block.block for _, s in row]
            maps = [m for row in block.block for m, _ in row]
        out[f"block/{tag}/rows"] = np.array([len(row) for row in block.block])
        for j, (s, m) in enumerate(zip(members, maps)):
            put_spline(f"block/{tag}/member{j}", s)
            out[f"block/{tag}/member{j}/map"] = np.array(m)
        out[f"block/{tag}/nIndnDep"] = np.array([block.nInd, block.nDep])
        out[f"block/{tag}/domain"] = np.asarray(block.domain(), float)
        dom = block.domain()
        N = 60
        pts = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.uniform(0, 1, (N, block.nInd))
        pts[0], pts[1] = dom[:, 0], dom[:, 1]
        out[f"block/{tag}/uvw"] = pts
        out[f"block/{tag}/values"] = np.array([block.evaluate(p) for p in pts])
        out[f"block/{tag}/jacobian"] = np.array([block.jacobian(p) for p in pts])
        wrt = [1] + [0] * (block.nInd - 1)
        wrt2 = [0] * block.nInd
        wrt2[-1] = 2
        out[f"block/{tag}/wrt"] = np.array([wrt, wrt2])
        out[f"block/{tag}/deriv0"] = np.array([block.derivative(wrt, p) for p in pts])
        out[f"block/{tag}/deriv1"] = np.array([block.derivative(wrt2, p) for p in pts])
        if abs(block.nInd - block.nDep) == 1:
            out[f"block/{tag}/normal_unit"] = np.array([block.normal(p) for p in pts])
            out[f"block/{tag}/normal_raw"] = np.array([block.normal(p, False) for p in pts])
            out[f"block/{tag}/normal_idx"] = np.array([block.normal(p, True, (0, 2)) for p in pts])
    # contracted block (A with s and w fixed): its values at points of the remaining variables
    cu = [0.3, None, None, None, 0.8]
    cb = blockA.contract(cu)
    out["block/A/contract_uvw"] = np.array([np.nan if v is None else v for v in cu])
    dom = cb.domain()
    pts = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.uniform(0, 1, (25, cb.nInd))
    out["block/A/contract_pts"] = pts
    out["block/A/contract_values"] = np.array([cb.evaluate(p) for p in pts])
    out["block/A/contract_nIndnDep"] = np.array([cb.nInd, cb.nDep])
    # ---- collocation rows: the assembly loop of Spline.least_squares, bspy/_spline_fitting.py:736-750,
    #      driven by the reference's own bspline_values; repeated parameters raise the derivative order ----------
    for tag, order, nCoef in (("o4", 4, 12), ("o3", 3, 7), ("o6", 6, 9)):
        knots = knots_nonuniform(order, nCoef, rng)
        uu = np.sort(rng.uniform(0, 1, 40))
        uu = np.concatenate(([0.0, 0.0], uu[:10], [uu[10]] * 3, uu[11:], [knots[order + 1]] * 2, [1.0, 1.0]))
        uu = np.sort(uu)
        A = np.zeros((len(uu), nCoef))
        derivs = np.zeros(len(uu), np.int32)
        u = -np.finfo(float).max
        for iRow in range(len(uu)):
            uNew = uu[iRow]
            if uNew != u:
                iDerivative = 0
                u = uNew
                ix = None
            else:
                iDerivative += 1
            ix, row = Spline.bspline_values(ix, knots, order, u, iDerivative)
            A[iRow, ix - order:ix] = row
            derivs[iRow] = iDerivative
        out[f"colloc/{tag}/knots"] = knots
        out[f"colloc/{tag}/order"] = np.array(order)
        out[f"colloc/{tag}/u"] = uu
        out[f"colloc/{tag}/derivs"] = derivs
        out[f"colloc/{tag}/A"] = A
    # ---- normal_spline sampling (bspy/_spline_operations.py:735-751): un-normalised normals on a tensor grid --------
    surf = synth(2, 3, (3, 4), (6, 7))
    put_spline("nsample/surf", surf)
    gu, gv = np.linspace(0, 1, 9), np.sort(rng.uniform(0, 1, 7))
    out["nsample/gu"], out["nsample/gv"] = gu, gv
    out["nsample/normals"] = np.array([[surf.normal([a, b], False) for b in gv] for a in gu])      # (9, 7, 3)
    np.savez_compressed(os.path.join(HERE, "ref_block.npz"), **out)
    print("ref_block.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
