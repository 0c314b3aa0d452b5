#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py [--ref /root/reference]

The reference package hard-imports its tkinter/OpenGL viewer (``bspy/__init__.py:28-29``),
which is not installable here, so inert stand-ins for those GUI modules are put on
``sys.modules`` before the import; no reference file is modified or copied.

Outputs (all small, committed):

* ``ref_tables.npz``   – the reference's own golden tables ``truthCurve`` (101 rows) and
                         ``truthSurface`` (441 rows) with the splines they belong to
                         (reference ``tests/bspy_test.py:15-564``).
* ``ref_cases.npz`` + ``ref_cases.json`` – outputs of the reference (spans, basis values,
                         values, mixed derivatives, jacobians, normals) for a battery of
                         synthetic splines and for the spline fixtures the reference ships
                         (``tests/*.json``, ``examples/TomsNasty.json``).
* ``teapot.npz``       – the 32 bicubic Utah-teapot patches of ``examples/teapot.py`` as
                         float64 coefficient blocks (read by parsing the two tuple
                         literals; the module itself is never imported because it opens the
                         viewer) plus reference values / derivatives / normals on a 9x9 grid.
* ``ref_curvature.npz`` – ``Spline.curvature`` (``bspy/_spline_evaluation.py:80-107``) at the sample points of ten of the
                         cases above (curves with nDep 1/2/3, surfaces with nDep 3 and 1).
* ``ref_dispatch.npz`` – results of the ufunc-style argument forms of ``Spline.evaluate`` /
                         ``Spline.derivative`` (``bspy/spline.py:757-770, 936-949``).
"""
import argparse
import ast
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference(root):
    class _Inert:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Inert()

        def __call__(self, *a, **k):
            return _Inert()

    for name in ("tkinter", "tkinter.ttk", "tkinter.colorchooser", "tkinter.filedialog",
                 "OpenGL", "OpenGL.GL", "OpenGL.GLU", "OpenGL.GL.shaders", "pyopengltk"):
        m = types.ModuleType(name)
        m.__all__ = []
        m.__path__ = []
        m.__getattr__ = lambda attr, _n=name: type(attr, (_Inert,), {})
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)
    sys.path.insert(0, root)
    import bspy  # noqa: the reference
    assert os.path.abspath(bspy.__file__).startswith(os.path.abspath(root)), bspy.__file__
    return bspy


def knots_nonuniform(order, n, rng, clamp=True):
    widths = rng.uniform(0.25, 1.75, n - order + 1)
    inner = np.concatenate(([0.0], np.cumsum(widths)))
    inner /= inner[-1]
    if clamp:
        return np.concatenate((np.zeros(order - 1), inner, np.ones(order - 1)))
    left = -np.cumsum(rng.uniform(0.05, 0.3, order - 1))[::-1]
    right = 1.0 + np.cumsum(rng.uniform(0.05, 0.3, order - 1))
    return np.concatenate((left, inner, right))


def adversarial_params(knots, order, nCoef, rng, nRandom):
    lo, hi = knots[order - 1], knots[nCoef]
    inside = np.unique(knots[(knots >= lo) & (knots <= hi)])
    pts = [rng.uniform(lo, hi, nRandom), inside]
    pts.append(np.nextafter(inside, hi))
    pts.append(np.nextafter(inside, lo))
    u = np.concatenate(pts)
    return u[(u >= lo) & (u <= hi)]


def sample_points(spline, rng, nRandom):
    cols = [adversarial_params(np.asarray(spline.knots[i], float), spline.order[i], spline.nCoef[i], rng, nRandom)
            for i in range(spline.nInd)]
    n = max(len(c) for c in cols)
    out = np.empty((n, spline.nInd))
    for i, c in enumerate(cols):
        reps = np.concatenate([c, rng.choice(c, n - len(c))]) if len(c) < n else c
        out[:, i] = rng.permutation(reps)
    return out


def wrt_list(spline):
    n = spline.nInd
    combos = []
    for i in range(n):
        e = [0] * n; e[i] = 1; combos.append(e)
        e = [0] * n; e[i] = 2; combos.append(e)
    if n >= 2:
        e = [0] * n; e[0] = 1; e[-1] = 1; combos.append(e)
        e = [1] * n; combos.append(e)
    e = [0] * n; e[0] = spline.order[0]; combos.append(e)          # >= order: exact zeros
    e = [0] * n; e[-1] = max(spline.order[-1] - 1, 0); combos.append(e)  # highest non-zero
    uniq = []
    for c in combos:
        if c not in uniq:
            uniq.append(c)
    return uniq


def run_case(bspy, spline, uvw, arrays, meta, tag, normals=True):
    N = uvw.shape[0]
    nInd, nDep = spline.nInd, spline.nDep
    arrays[f"{tag}/uvw"] = uvw
    for i in range(nInd):
        arrays[f"{tag}/knots{i}"] = np.asarray(spline.knots[i], float)
    arrays[f"{tag}/coefs"] = np.ascontiguousarray(spline.coefs, dtype=float)
    spans = np.empty((N, nInd), np.int32)
    with np.errstate(all="ignore"):
        for i in range(nInd):
            k = np.asarray(spline.knots[i], float)
            for d in range(spline.order[i] + 2):
                for taylor in (False, True):
                    if taylor and d == 0:
                        continue
                    B = np.empty((N, spline.order[i]))
                    for p in range(N):
                        ix, B[p] = bspy.Spline.bspline_values(None, k, spline.order[i], uvw[p, i], d, taylor)
                        spans[p, i] = ix
                    arrays[f"{tag}/basis{i}_d{d}{'t' if taylor else ''}"] = B
        arrays[f"{tag}/spans"] = spans
        arrays[f"{tag}/values"] = np.array([spline.evaluate(uvw[p]) for p in range(N)]).reshape(N, nDep)
        arrays[f"{tag}/jacobian"] = np.array([spline.jacobian(uvw[p]) for p in range(N)]).reshape(N, nDep, nInd)
        wrts = wrt_list(spline)
        for w in wrts:
            arrays[f"{tag}/deriv_" + "_".join(map(str, w))] = np.array(
                [spline.derivative(w, uvw[p]) for p in range(N)]).reshape(N, nDep)
        entry = dict(tag=tag, nInd=nInd, nDep=nDep, order=list(spline.order), nCoef=list(spline.nCoef),
                     metadata={k: v for k, v in spline.metadata.items() if isinstance(v, (bool, int, float, str))},
                     wrt=wrts, N=N, normal=False, normal_indices=None)
        if normals and abs(nInd - nDep) == 1:
            D = max(nInd, nDep)
            entry["normal"] = True
            arrays[f"{tag}/normal_unit"] = np.array([spline.normal(uvw[p]) for p in range(N)]).reshape(N, D)
            arrays[f"{tag}/normal_raw"] = np.array([spline.normal(uvw[p], False) for p in range(N)]).reshape(N, D)
            idx = (0, D - 1) if D > 2 else (1,)
            entry["normal_indices"] = list(idx)
            arrays[f"{tag}/normal_idx_unit"] = np.array([spline.normal(uvw[p], True, idx) for p in range(N)]).reshape(N, len(idx))
            arrays[f"{tag}/normal_idx_raw"] = np.array([spline.normal(uvw[p], False, idx) for p in range(N)]).reshape(N, len(idx))
    meta.append(entry)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    bspy = import_reference(args.ref)
    Spline = bspy.Spline

    # ---- 1. the reference's own golden tables -------------------------------------------
    cwd = os.getcwd()
    os.chdir(args.ref)
    sys.path.insert(0, os.path.join(args.ref, "tests"))
    import bspy_test as reftests  # noqa  (module level only defines data and test functions)
    os.chdir(cwd)
    tables = {}
    for name, spline, table in (("curve", reftests.myCurve, reftests.truthCurve),
                                ("surface", reftests.mySurface, reftests.truthSurface)):
        tables[f"{name}/table"] = np.array(table, float)
        tables[f"{name}/coefs"] = np.ascontiguousarray(spline.coefs, dtype=float)
        tables[f"{name}/order"] = np.array(spline.order)
        for i, k in enumerate(spline.knots):
            tables[f"{name}/knots{i}"] = np.asarray(k, float)
    np.savez_compressed(os.path.join(HERE, "ref_tables.npz"), **tables)

    # ---- 2. battery of synthetic + shipped splines --------------------------------------
    rng = np.random.default_rng(20261018)
    arrays, meta = {}, []

    def synth(nInd, nDep, order, nCoef, clamp=True, metadata=None, knots=None):
        kk = knots if knots is not None else [knots_nonuniform(order[i], nCoef[i], rng, clamp) for i in range(nInd)]
        c = rng.standard_normal((nDep, *nCoef))
        return Spline(nInd, nDep, order, nCoef, kk, c, metadata or {})

    cases = []
    for o in (1, 2, 3, 4, 5, 7, 10):
        cases.append((f"curve_o{o}", synth(1, 1 + o % 3, (o,), (o + 6,)), 60))
    cases.append(("curve_cfg1", synth(1, 3, (4,), (64,)), 200))
    cases.append(("curve_unclamped", synth(1, 2, (4,), (9,), clamp=False), 60))
    cases.append(("curve_doubleknot", synth(1, 2, (4,), (8,), knots=[np.array([0, 0, 0, 0, .2, .5, .5, .8, 1, 1, 1, 1.])]), 60))
    cases.append(("curve_c0knot", synth(1, 2, (3,), (7,), knots=[np.array([0, 0, 0, .25, .6, .6, .9, 1, 1, 1.])]), 60))
    # zero-width last span (k[nCoef-1] == k[nCoef], legal when the right end is unclamped): NaN at u = hi
    cases.append(("curve_zerolast", synth(1, 1, (3,), (5,), knots=[np.array([0, 0, 0, .4, .7, .7, 1.0, 1.2])]), 30))
    cases.append(("planar_neg", synth(1, 2, (4,), (7,), metadata={"negateNormal": True}), 60))
    cases.append(("surf_34", synth(2, 3, (3, 4), (6, 7)), 120))
    cases.append(("surf_44_neg", synth(2, 3, (4, 4), (5, 9), metadata={"negateNormal": True}), 120))
    cases.append(("surf_25_d1", synth(2, 1, (2, 5), (4, 8)), 100))         # nInd > nDep normal
    cases.append(("surf_52_d2", synth(2, 2, (5, 2), (7, 3), clamp=False), 100))
    cases.append(("vol_444_d3", synth(3, 3, (4, 4, 4), (6, 7, 5)), 150))   # cfg4-like
    cases.append(("vol_332_d2", synth(3, 2, (3, 3, 2), (5, 4, 4)), 100))   # nInd > nDep normal
    cases.append(("vol_343_d4", synth(3, 4, (3, 4, 3), (4, 6, 5)), 100))   # 3x3 cofactors
    cases.append(("man_3333_d6", synth(4, 6, (3, 3, 3, 3), (5, 4, 6, 5)), 150))  # cfg5-like
    cases.append(("man_3232_d5", synth(4, 5, (3, 2, 3, 2), (4, 3, 5, 3)), 80))   # 4x4 cofactors
    cases.append(("man_2222_d3", synth(4, 3, (2, 2, 2, 2), (3, 3, 3, 3)), 60))   # nInd > nDep

    shipped = [("tests/trim-issue.json", "trim"), ("tests/reverse-thing.json", "reverse"),
               ("tests/offset-issue.json", "offset"), ("tests/patterson001.json", "patterson"),
               ("examples/TomsNasty.json", "tomsnasty")]
    for rel, tag in shipped:
        for j, s in enumerate(Spline.load(os.path.join(args.ref, rel))):
            cases.append((f"{tag}{j}", s, 40))

    for tag, spline, nRandom in cases:
        uvw = sample_points(spline, rng, nRandom)
        if uvw.shape[0] > 400:
            uvw = uvw[rng.permutation(uvw.shape[0])[:400]]
        run_case(bspy, spline, uvw, arrays, meta, tag)
        print(tag, uvw.shape)
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **arrays)

    # ---- 2b. curvature (bspy/_spline_evaluation.py:80-107) for curves (nDep 1, 2, 3) and surfaces (nDep 3, 1) ----
    curv = {}
    with np.errstate(all="ignore"):
        for tag, spline, _ in cases:
            if tag in ("curve_o3", "curve_o4", "curve_o5", "curve_cfg1", "planar_neg", "curve_doubleknot", "surf_34",
                       "surf_44_neg", "surf_25_d1", "tomsnasty0"):
                uvw = arrays[f"{tag}/uvw"]
                curv[tag] = np.array([float(spline.curvature(uvw[p] if spline.nInd > 1 else uvw[p, 0])) for p in range(uvw.shape[0])])
    np.savez_compressed(os.path.join(HERE, "ref_curvature.npz"), **curv)
    with open(os.path.join(HERE, "ref_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)

    # ---- 3. teapot ------------------------------------------------------------------------
    src = open(os.path.join(args.ref, "examples", "teapot.py")).read()
    lits = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            if node.targets[0].id in ("teapotPatches", "teapotVertices"):
                lits[node.targets[0].id] = ast.literal_eval(node.value)
    patches, verts = lits["teapotPatches"], lits["teapotVertices"]
    coefs = np.empty((len(patches), 3, 4, 4))
    names = []
    for p, patch in enumerate(patches):
        names.append(str(patch[0]))
        for i in range(4):
            for j in range(4):
                v = verts[patch[4 * i + j + 1] - 1]
                coefs[p, :, i, j] = (v[0], v[2], v[1])      # y/z swap as in examples/teapot.py:356-358
    kn = np.array([0, 0, 0, 0, 1, 1, 1, 1.])
    g = np.linspace(0.0, 1.0, 9)
    val = np.empty((len(patches), 3, 9, 9)); du = np.empty_like(val); dv = np.empty_like(val); nrm = np.empty_like(val)
    with np.errstate(all="ignore"):
        for p in range(len(patches)):
            s = Spline(2, 3, (4, 4), (4, 4), (kn, kn), coefs[p])
            for a, u in enumerate(g):
                for b, v in enumerate(g):
                    val[p, :, a, b] = s.evaluate((u, v))
                    J = s.jacobian((u, v))
                    du[p, :, a, b], dv[p, :, a, b] = J[:, 0], J[:, 1]
                    nrm[p, :, a, b] = s.normal((u, v))
    np.savez_compressed(os.path.join(HERE, "teapot.npz"), coefs=coefs, names=np.array(names), knots=kn, grid=g,
                        values=val, du=du, dv=dv, normal=nrm)
    print("teapot", coefs.shape, "NaN normals:", int(np.isnan(nrm).any(axis=1).sum()))

    # ---- 4. ufunc-style dispatch ----------------------------------------------------------
    disp = {}
    curve, surf = reftests.myCurve, reftests.mySurface
    uu = np.linspace(0, 1, 11)
    r = curve(uu)
    disp["curve_ufunc"] = np.array(r, float)                     # tuple of nDep arrays
    disp["curve_deriv_ufunc"] = np.array(curve.derivative([1], uu), float)
    U, V = np.meshgrid(np.linspace(0, 1, 5), np.linspace(0, 1, 4), indexing="ij")
    disp["surf_ufunc"] = np.array(surf(U, V), float)             # (3, 5, 4)
    disp["surf_deriv_ufunc"] = np.array(surf.derivative([1, 1], U, V), float)
    disp["surf_point_list"] = np.array(surf([0.25, 0.5]), float)
    disp["surf_point_scalars"] = np.array(surf(0.25, 0.5), float)
    scalar = Spline(1, 1, (3,), (5,), [np.array([0, 0, 0, .3, .6, 1, 1, 1.])], np.array([[1., 2, 0, -1, 3]]))
    disp["scalar_ufunc"] = np.array(scalar(uu), float)           # nDep == 1: 1-D array
    disp["scalar_point"] = np.array(scalar(0.5), float)
    disp["uu2d"] = np.linspace(0, 1, 12).reshape(3, 4)
    disp["scalar_ufunc_2d"] = np.array(scalar(disp["uu2d"]), float)   # the reference's row-iteration quirk
    disp["scalar_knots"] = np.asarray(scalar.knots[0], float)
    disp["scalar_coefs"] = np.asarray(scalar.coefs, float)
    disp["uu"] = uu
    disp["U"], disp["V"] = U, V
    np.savez_compressed(os.path.join(HERE, "ref_dispatch.npz"), **disp)
    print("done")


if __name__ == "__main__":
    main()
