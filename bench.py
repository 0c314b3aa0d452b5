#!/usr/bin/env python
"""bench.py -- float64 spline point-evaluations per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config all|cfg1|cfg1_1e8|cfg2|cfg3|cfg4|cfg4_soa|cfg5|cfg5_soa|grid3]
                    [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline workload (`value`, `e2e`,
`roofline`, `cpu_baseline`, `parity` at the top level of the JSON line) is BASELINE.json configs[1] (cfg2): the 32
bicubic Utah-teapot patches evaluated with value, d/du, d/dv and unit normal on a 2048 x 2048 grid per patch (134 M
points, 12.9 GB of output per step per GPU).  With `--config all` (default) the same run then times every other
BASELINE config and reports them under `"configs"`: cfg1 (1 M points on a cubic 3-D curve; also at 1e8 points), cfg3
(1 M curves x 256 points), cfg4 (trivariate volume, 1e8 scattered points, value + jacobian), cfg5 (nInd 4 / nDep 6
manifold, 1.25e8 points per GPU, value + first derivatives) and the 512^3 volume grid -- each with its own roofline
(HBM bound for cfg1-3 / grids, FP64-pipe bound for cfg4 / cfg5, both fractions printed) and its own parity block.

`parity`: after the timed steps a seeded subsample (>= 1e5 points) of the TIMED outputs is copied back and compared with
the C restatement of the reference (oracle/, the checker): values / derivatives / normals by |x - ref| <= 1e-13 +
1e-12 |ref| with NaNs matching, knot spans `==`.  A failed parity block makes the exit code non-zero.

One JSON line is printed by rank 0 (contract in the task description).  Multi-GPU: one process per GPU (torchrun), every
rank evaluates its own shard of patches / points / curves (weak scaling, no data-path collective), time = max over
ranks; with N > 1 one strong-scaling job is added (`"strong_scaling"`): a single 1e8-point cfg4 batch split with
shard_points, evaluated, and re-assembled on every rank with an NCCL all-gather (kernel-only and gather-inclusive).

`--impl reference` times the reference's own CPU implementation of the path on all host cores: the UNMODIFIED reference
package when it can be imported (baseline/_ref, installed by __graft_entry__.build(); kind "reference"), else the
oracle's scalar port of it (kind "port"), on a bounded sample of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "float64 spline point-evals/sec"
UNIT = "points/s"
L2_BYTES = 126e6
RTOL, ATOL = 1e-12, 1e-13          # the parity bar of BASELINE.json north_star


# ----------------------------------------------------------------------------------- inputs

def knots_nonuniform(order, n, rng):
    """Clamped non-uniform knots: span widths U(0.25, 1.75), normalised to [0, 1]."""
    w = rng.uniform(0.25, 1.75, n - order + 1)
    inner = np.concatenate(([0.0], np.cumsum(w)))
    inner /= inner[-1]
    return np.concatenate((np.zeros(order - 1), inner, np.ones(order - 1)))


def teapot_patches():
    t = np.load(os.path.join(ROOT, "tests", "golden", "teapot.npz"))
    return t["coefs"], t["knots"]


class _Plain:
    """attribute bag with the Spline fields (payload for the CPU workers / the oracle)"""

    def __init__(self, nInd, nDep, order, nCoef, knots, coefs, metadata=None):
        self.nInd, self.nDep, self.order, self.nCoef = nInd, nDep, tuple(order), tuple(nCoef)
        self.knots, self.coefs, self.metadata = [np.asarray(k) for k in knots], np.asarray(coefs), dict(metadata or {})


def _payload(s):
    return dict(nInd=s.nInd, nDep=s.nDep, order=tuple(s.order), nCoef=tuple(s.nCoef), knots=[np.asarray(k) for k in s.knots],
                coefs=np.asarray(s.coefs), metadata=dict(s.metadata))


# ----------------------------------------------------------------------------------- parity

class Parity:
    """Accumulates comparisons of timed outputs with the oracle on a subsample."""

    def __init__(self):
        self.n = 0
        self.worst_ratio = 0.0          # max |x - ref| / (ATOL + RTOL |ref|): <= 1 is inside the bar
        self.worst_excess = -np.inf     # max |x - ref| - (ATOL + RTOL |ref|): <= 0 is inside the bar
        self.nan_match = True
        self.spans_equal = None
        self.notes = []

    def values(self, x, ref, what, mask=None):
        x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
        if mask is not None:
            x, ref = x[mask], ref[mask]
        if not np.array_equal(np.isnan(x), np.isnan(ref)):
            self.nan_match = False
            self.notes.append(f"{what}: NaN pattern differs ({int(np.isnan(x).sum())} vs {int(np.isnan(ref).sum())})")
        fin = np.isfinite(ref) & np.isfinite(x)
        inf = ~np.isnan(ref) & ~np.isfinite(ref)
        if inf.any() and not np.array_equal(x[inf], ref[inf]):
            self.nan_match = False
            self.notes.append(f"{what}: infinities differ")
        if fin.any():
            err = np.abs(x[fin] - ref[fin])
            bar = ATOL + RTOL * np.abs(ref[fin])
            self.worst_ratio = max(self.worst_ratio, float((err / bar).max()))
            self.worst_excess = max(self.worst_excess, float((err - bar).max()))

    def spans(self, x, ref):
        eq = bool(np.array_equal(np.asarray(x), np.asarray(ref)))
        self.spans_equal = eq if self.spans_equal is None else (self.spans_equal and eq)

    def report(self):
        ok = self.nan_match and self.worst_ratio <= 1.0 and self.spans_equal is not False and self.n > 0
        r = {"n": int(self.n), "spans_equal": self.spans_equal, "worst_excess": self.worst_excess if np.isfinite(self.worst_excess) else None,
             "worst_ratio": self.worst_ratio, "nan_match": self.nan_match, "ok": bool(ok),
             "bar": "|x-ref| <= 1e-13 + 1e-12|ref|, NaN/inf matching, spans ==", "checker": "oracle/bspy_oracle.c (C restatement of the reference)"}
        if self.notes:
            r["notes"] = self.notes[:6]
        return r


def _oracle():
    from oracle import c_oracle as CO
    CO.build()
    return CO


# ------------------------------------------------------------------------------- workloads

class Workload:
    name = ""
    kernel = ""
    bytes_per_point = 0.0     # algorithmic (compulsory) HBM bytes per point, SURVEY 8(d)
    flops_per_point = 0.0     # algorithmic flops per point, SURVEY 8(d)
    bound = "hbm"             # roofline the config is graded on: "hbm" or "fp64"
    points = 0                # per step per GPU
    note = ""
    cpu_calls = ("evaluate",)
    e2e_fraction = 1.0        # the e2e leg of a secondary config may run on a leading slice of the batch

    def setup(self, dev, rank, scale):
        raise NotImplementedError

    def step(self):           # device-resident inputs and outputs; domain check written on the device (deferred read)
        raise NotImplementedError

    def flags_ok(self):       # the deferred domain flags of the timed steps: nothing outside
        return True

    def e2e_step(self):       # host buffers in, host buffers out; returns (points, h2d_bytes, d2h_bytes)
        raise NotImplementedError

    def reference_task(self, n):  # (spline payload, points (n, nInd)) for the CPU arm
        raise NotImplementedError

    def parity(self):
        raise NotImplementedError

    def teardown(self):
        for k in list(self.__dict__):
            if k not in ("torch", "bspy", "dev"):
                self.__dict__.pop(k)


class Cfg2Teapot(Workload):
    name = "cfg2: 32 bicubic Utah-teapot patches, value+du+dv+unit normal on a 2048x2048 grid per patch"
    kernel = "grid2_dmma_kernel<3,4,false>"
    bytes_per_point, flops_per_point, bound = 96.0, 92.0, "hbm"
    cpu_calls = ("evaluate", "jacobian", "normal")

    def setup(self, dev, rank, scale):
        import torch
        import bspy_b200 as bspy
        self.torch, self.bspy, self.dev = torch, bspy, dev
        coefs, kn = teapot_patches()
        self.S = coefs.shape[0]
        self.n = max(64, int(round(2048 * scale)) // 8 * 8)
        self.coefs_host, self.kn = coefs, kn
        self.splines = [bspy.Spline(2, 3, (4, 4), (4, 4), (kn, kn), coefs[p]) for p in range(self.S)]
        self.batch = bspy.SplineBatch.from_splines(self.splines, device=dev)
        self.axis_host = np.linspace(0.0, 1.0, self.n)
        self.axis = torch.from_numpy(self.axis_host).to(dev)
        self.points = self.S * self.n * self.n
        shape = (self.S, 3, self.n, self.n)
        self.out = {"values": torch.empty(shape, dtype=torch.float64, device=dev),
                    "jacobian": torch.empty((self.S, 3, 2, self.n, self.n), dtype=torch.float64, device=dev),
                    "normal": torch.empty(shape, dtype=torch.float64, device=dev)}
        self.note = f"grid {self.n}x{self.n} per patch; outputs {self.points * 96 / 1e9:.1f} GB per step >> L2, inputs (axes, 32x48 coefficients) are KB-sized"
        self.working_set = self.points * 96
        self.flag = None

    def step(self):
        r = self.batch.evaluate_grid(self.axis, self.axis, jacobian=True, normal=True, check_domain="defer", out=self.out)
        self.flag = r.first_outside

    def flags_ok(self):
        return self.flag is None or int(self.flag.item()) < 0

    def e2e_step(self):
        # host splines + host axes in, host arrays out (public API: SplineBatch.from_splines + evaluate_grid)
        batch = self.bspy.SplineBatch.from_splines(self.splines, device=self.dev)
        r = batch.evaluate_grid(self.axis_host, self.axis_host, jacobian=True, normal=True)
        h2d = self.coefs_host.nbytes + self.kn.nbytes * 2 + self.axis_host.nbytes * 2
        d2h = r.values.nbytes + r.jacobian.nbytes + r.normal.nbytes
        del r                     # the pinned result buffers go back to torch's host allocator for the next step
        return self.points, h2d, d2h

    def reference_task(self, n):
        g = np.linspace(0.0, 1.0, getattr(self, "n", 2048))
        rng = np.random.default_rng(2)
        ia, ib = rng.integers(0, len(g), n), rng.integers(0, len(g), n)
        coefs, kn = teapot_patches()
        sp = dict(nInd=2, nDep=3, order=(4, 4), nCoef=(4, 4), knots=[kn, kn], coefs=coefs[3], metadata={})
        return sp, np.stack([g[ia], g[ib]], axis=1)

    def parity(self):
        """Timed outputs of 8 patches (rim, body, handle, spout, lid knob, lid, two bottoms): the four border rows /
        columns in full plus 12288 random interior nodes each.  Values and derivatives: strict bar everywhere.  Unit
        normals, by how well the cross product is conditioned at the node (|du x dv| against the sum of the magnitudes of
        the terms it is built from, oracle.normal_abs_vec):
          well-conditioned (ratio >= 1e-3): strict bar;
          ill-conditioned (next to the collapsed control rows of the lid / bottom patches the partials themselves cancel;
            two restatements of the reference already differ by 2.3x the strict bar there): the bar widened by
            16 eps sum|terms| / |du x dv|;
          degenerate (the collapsed rows themselves: the raw normal is an exact or inexact zero and the reference returns
            NaN or an arbitrary unit vector depending on rounding): NaN or unit length."""
        CO = _oracle()
        from oracle import bspy_oracle as O
        torch, n = self.torch, self.n
        P = Parity()
        rng = np.random.default_rng(22)
        classes = {"well": 0, "ill": 0, "degenerate": 0}
        eps = np.finfo(float).eps
        sp = self.bspy._cuda.spans(torch.from_numpy(self.kn).to(self.dev), 4, self.axis).cpu().numpy()
        for p in (0, 5, 13, 17, 20, 25, 28, 31):
            a = np.concatenate([np.zeros(n, np.int64), np.full(n, n - 1), np.arange(n), np.arange(n), rng.integers(0, n, 12288)])
            b = np.concatenate([np.arange(n), np.arange(n), np.zeros(n, np.int64), np.full(n, n - 1), rng.integers(0, n, 12288)])
            ia, ib = torch.from_numpy(a).to(self.dev), torch.from_numpy(b).to(self.dev)
            v = self.out["values"][p][:, ia, ib].cpu().numpy().T
            j = self.out["jacobian"][p][:, :, ia, ib].cpu().numpy().transpose(2, 0, 1)
            nr = self.out["normal"][p][:, ia, ib].cpu().numpy().T
            s = _Plain(2, 3, (4, 4), (4, 4), (self.kn, self.kn), self.coefs_host[p])
            uv = np.stack([self.axis_host[a], self.axis_host[b]], axis=1)
            ref = CO.evaluate(s, uv, values=True, jacobian=True, normal=True, normalize=True, spans=True)
            raw = CO.evaluate(s, uv, values=False, normal=True, normalize=False)["normal"]
            with np.errstate(all="ignore"):
                scale = O.normal_abs_vec(O.OracleSpline(**_payload(s)), uv).max(axis=1)
                mag = np.sqrt((raw ** 2).sum(axis=1))
                degenerate = ~(mag > 1e-9 * scale)
                well = (scale > 0) & (mag >= 1e-3 * scale)
                ill = ~well & ~degenerate
            P.values(v, ref["values"], f"patch {p} values")
            P.values(j, ref["jacobian"], f"patch {p} jacobian")
            P.values(nr, ref["normal"], f"patch {p} normal", mask=well)
            if ill.any():
                x, r = nr[ill], ref["normal"][ill]
                tol = ATOL + RTOL * np.abs(r) + 16 * eps * (scale[ill] / mag[ill])[:, None]
                if not (np.isfinite(x).all() and np.all(np.abs(x - r) <= tol)):
                    P.nan_match = False
                    P.notes.append(f"patch {p}: ill-conditioned normals outside the condition-aware bar")
            rest = nr[degenerate]
            if rest.size:
                length = np.sqrt((rest ** 2).sum(axis=1))
                if not np.all(np.isnan(length) | (np.abs(length - 1.0) < 1e-12)):
                    P.nan_match = False
                    P.notes.append(f"patch {p}: singular-point normal neither NaN nor unit length")
            classes["well"] += int(well.sum()); classes["ill"] += int(ill.sum()); classes["degenerate"] += int(degenerate.sum())
            # spans of the grid axes through the span kernel of the same library (the grid kernel does not output them)
            P.spans(sp[a], ref["spans"][:, 0])
            P.spans(sp[b], ref["spans"][:, 1])
            P.n += len(a)
        rep = P.report()
        rep["normal_classes"] = classes
        return rep


class ScatteredBase(Workload):
    jac = False
    seed = 0
    N = 0
    layout = "soa"             # "aos": one [values | jacobian] record per point (whole 32-byte sectors)

    def make_spline(self, rng, bspy):
        raise NotImplementedError

    def setup(self, dev, rank, scale):
        import torch
        import bspy_b200 as bspy
        self.torch, self.bspy, self.dev = torch, bspy, dev
        rng = np.random.default_rng(self.seed)
        self.spline = self.make_spline(rng, bspy)
        self.spline.freeze(dev)
        self.points = max(1024, int(self.N * scale))
        g = torch.Generator(device=dev).manual_seed(self.seed + 17 * rank)
        self.pts = torch.rand((self.points, self.spline.nInd), dtype=torch.float64, device=dev, generator=g)
        self.working_set = self.points * self.bytes_per_point
        self.note = f"{self.points} points per GPU; working set {self.working_set / 1e9:.2f} GB; output layout {self.layout}"
        self.host_pts = None
        self.last = None

    def step(self):
        self.last = self.spline.evaluate_points(self.pts, values=True, jacobian=self.jac, check_domain="defer",
                                                out_layout=self.layout)

    def flags_ok(self):
        return self.last is None or self.last.first_outside is None or int(self.last.first_outside.item()) < 0

    def e2e_step(self):
        torch = self.torch
        n = max(1024, int(self.points * self.e2e_fraction))
        if self.host_pts is None:
            self.host_pts = torch.empty((n, self.spline.nInd), dtype=torch.float64, pin_memory=True)
            self.host_pts.copy_(self.pts[:n])
        r = self.spline.evaluate_points(self.host_pts, values=True, jacobian=self.jac, out_layout=self.layout)
        d2h = r.values.numel() * 8 + (r.jacobian.numel() * 8 if r.jacobian is not None else 0)
        del r
        return n, self.host_pts.numel() * 8, d2h

    def reference_task(self, n):
        rng = np.random.default_rng(self.seed)
        import types
        s = self.make_spline(rng, types.SimpleNamespace(Spline=lambda *a: _Plain(*a)))
        pts = np.random.default_rng(self.seed + 1).uniform(0, 1, (n, s.nInd))
        return _payload(s), pts

    def parity(self, n=200_000):
        """A seeded sample of the TIMED outputs (values, jacobian) against the oracle; spans from a second call on the
        sampled points through the same dispatch (timed steps do not write spans: SURVEY 8(d))."""
        CO = _oracle()
        torch = self.torch
        n = min(n, self.points)
        g = torch.Generator(device=self.dev).manual_seed(99)
        idx = torch.randint(0, self.points, (n,), device=self.dev, generator=g)
        idx[0], idx[-1] = 0, self.points - 1
        sub = self.pts[idx].contiguous()
        uh = sub.cpu().numpy()
        s = self.spline
        ref = CO.evaluate(s, uh, values=True, jacobian=self.jac, spans=True)
        P = Parity()
        P.values(self.last.values[:, idx].cpu().numpy().T, ref["values"], "values")
        if self.jac:
            P.values(self.last.jacobian[:, :, idx].cpu().numpy().transpose(2, 0, 1), ref["jacobian"], "jacobian")
        again = s.evaluate_points(sub, values=True, spans=True, out_layout=self.layout)
        P.spans(again.spans.cpu().numpy().T, ref["spans"])
        P.values(again.values.cpu().numpy().T, ref["values"], "values (span call)")
        P.n = n
        return P.report()


class Cfg1Curve(ScatteredBase):
    name = "cfg1: cubic 3-D curve, 64 coefficients, non-uniform knots, 1M random parameters, values"
    kernel = "eval_curve_tab_kernel<4,3,false,2,true> (cached table image by TMA, per-span polynomial rows)"
    bytes_per_point, flops_per_point, seed, N = 32.0, 66.0, 1001, 1_000_000

    def make_spline(self, rng, bspy):
        return bspy.Spline(1, 3, (4,), (64,), [knots_nonuniform(4, 64, rng)], rng.standard_normal((3, 64)))


class Cfg1Curve1e8(Cfg1Curve):
    name = "cfg1 at 1e8 parameters (same curve; the asymptotic rate of the kernel)"
    N = 100_000_000
    e2e_fraction = 0.25


class Cfg4Volume(ScatteredBase):
    name = "cfg4: trivariate order-4 volume (nInd 3, nDep 3, 32^3 coefficients), 1e8 scattered points, value + jacobian, array-of-structs records"
    kernel = "cell-sorted pipeline: bin_keys, bin_scan, bin_scatter_records, eval_poly_kernel<3,4,4,4,0,3,3,5> (cell polynomials) writing [values|jacobian] records in place (whole step)"
    bytes_per_point, flops_per_point, seed, N, jac, bound = 120.0, 1320.0, 1004, 100_000_000, True, "fp64"
    layout = "aos"
    cpu_calls = ("evaluate", "jacobian")

    def make_spline(self, rng, bspy):
        return bspy.Spline(3, 3, (4, 4, 4), (32, 32, 32), [knots_nonuniform(4, 32, rng) for _ in range(3)],
                           rng.standard_normal((3, 32, 32, 32)))


class Cfg4VolumeSoA(Cfg4Volume):
    name = "cfg4 with struct-of-arrays outputs (values (3,N) + jacobian (3,3,N)): adds the un-permute pass"
    kernel = "cell-sorted pipeline + bin_unpermute (whole step)"
    layout = "soa"


class Cfg5Manifold(ScatteredBase):
    name = "cfg5: nInd 4 / nDep 6 order-3 manifold (16^4 coefficients), 1.25e8 scattered points per GPU, value + first derivatives, array-of-structs records"
    kernel = "cell-sorted pipeline: bin_keys, bin_scan, bin_pad, bin_scatter_records, eval_poly2_kernel<4,3,3,3,3,6,2,3> (cell polynomials, two points per thread) writing [values|jacobian] records in place (whole step)"
    bytes_per_point, flops_per_point, seed, N, jac, bound = 272.0, 3650.0, 1005, 125_000_000, True, "fp64"
    layout = "aos"
    cpu_calls = ("evaluate", "jacobian")
    e2e_fraction = 0.2

    def make_spline(self, rng, bspy):
        return bspy.Spline(4, 6, (3,) * 4, (16,) * 4, [knots_nonuniform(3, 16, rng) for _ in range(4)],
                           rng.standard_normal((6, 16, 16, 16, 16)))


class Cfg5ManifoldSoA(Cfg5Manifold):
    name = "cfg5 with struct-of-arrays outputs (values (6,N) + jacobian (6,4,N)): adds the un-permute pass"
    kernel = "cell-sorted pipeline + bin_unpermute (whole step)"
    layout = "soa"


class Cfg3Curves(Workload):
    name = "cfg3: 1M independent cubic 3-D curves (32 coefficients each), 256 points per curve"
    kernel = "many_tab_kernel<4,3> (cached per-curve images by TMA, per-span polynomial rows)"
    bytes_per_point, flops_per_point, bound = 9248.0 / 256.0, 66.0, "hbm"

    def setup(self, dev, rank, scale):
        import torch
        import bspy_b200 as bspy
        self.torch, self.bspy, self.dev = torch, bspy, dev
        self.S = max(256, int(1_000_000 * scale))
        g = torch.Generator(device=dev).manual_seed(1003 + 17 * rank)
        w = torch.rand((self.S, 29), dtype=torch.float64, device=dev, generator=g) * 1.5 + 0.25
        inner = torch.cat([torch.zeros((self.S, 1), dtype=torch.float64, device=dev), torch.cumsum(w, 1)], 1)
        inner = inner / inner[:, -1:]
        inner[:, -1] = 1.0
        self.knots = torch.cat([torch.zeros((self.S, 3), dtype=torch.float64, device=dev), inner,
                                torch.ones((self.S, 3), dtype=torch.float64, device=dev)], 1).contiguous()
        self.coefs = torch.randn((self.S, 3, 32), dtype=torch.float64, device=dev, generator=g)
        self.u = torch.rand((self.S, 256), dtype=torch.float64, device=dev, generator=g)
        self.batch = bspy.SplineBatch(1, 3, (4,), (32,), [self.knots], self.coefs)
        self.points = self.S * 256
        self.out = {"values": torch.empty((self.S, 3, 256), dtype=torch.float64, device=dev), "derivative": None}
        self.working_set = self.S * 9248
        self.note = (f"{self.S} curves per GPU; working set {self.working_set / 1e9:.2f} GB; the batch's cached per-curve images "
                     f"(3.6 KB per curve, built once by its second evaluation, i.e. during warm-up) are read instead of the raw "
                     f"1.06 KB of knots and coefficients: 11.8 KB of HBM traffic per curve for 9.25 KB algorithmic")
        self.host = None
        self.flag = None
        # a resident batch that is evaluated repeatedly: its cached per-curve images exist before the first timed (or captured)
        # step whatever --warmup is -- they are built by the batch's second value-only evaluation
        for _ in range(2):
            self.batch.evaluate(self.u, check_domain="defer", out=self.out)

    def step(self):
        r = self.batch.evaluate(self.u, check_domain="defer", out=self.out)
        self.flag = r.first_outside

    def flags_ok(self):
        return self.flag is None or int(self.flag.item()) < 0

    def e2e_step(self):
        torch = self.torch
        if self.host is None:
            self.host = [torch.empty(t.shape, dtype=torch.float64, pin_memory=True).copy_(t) for t in (self.knots, self.coefs, self.u)]
            self.host_out = torch.empty((self.S, 3, 256), dtype=torch.float64, pin_memory=True)
        k, c, u = (t.to(self.dev, non_blocking=True) for t in self.host)
        b = self.bspy.SplineBatch(1, 3, (4,), (32,), [k], c)
        r = b.evaluate(u)
        self.host_out.copy_(r.values)
        return self.points, sum(t.numel() * 8 for t in self.host), self.host_out.numel() * 8

    def reference_task(self, n):
        rng = np.random.default_rng(1003)
        sp = dict(nInd=1, nDep=3, order=(4,), nCoef=(32,), knots=[knots_nonuniform(4, 32, rng)],
                  coefs=rng.standard_normal((3, 32)), metadata={})
        return sp, rng.uniform(0, 1, (n, 1))

    def parity(self, curves=512):
        """>= 64 whole curves (here 512: first, last and random ones) of the timed output against the oracle; spans of
        their parameters through bspy_cuda_spans."""
        CO = _oracle()
        torch = self.torch
        rng = np.random.default_rng(33)
        pick = np.unique(np.concatenate(([0, self.S - 1], rng.integers(0, self.S, min(curves, self.S)))))
        it = torch.from_numpy(pick).to(self.dev)
        kn, cf, u = self.knots[it].cpu().numpy(), self.coefs[it].cpu().numpy(), self.u[it].cpu().numpy()
        got = self.out["values"][it].cpu().numpy()
        P = Parity()
        for r, s in enumerate(pick):
            sp = _Plain(1, 3, (4,), (32,), [kn[r]], cf[r])
            ref = CO.evaluate(sp, u[r][:, None], values=True, spans=True)
            P.values(got[r].T, ref["values"], f"curve {s}")
            if r < 64:
                mine = self.bspy._cuda.spans(self.knots[int(s)].contiguous(), 4, self.u[int(s)].contiguous()).cpu().numpy()
                P.spans(mine, ref["spans"][:, 0])
            P.n += u.shape[1]
        return P.report()


class Grid3Volume(Workload):
    name = "grid3: the cfg4 volume spline on a 512^3 tensor grid, value + jacobian (FP64 tensor pipe)"
    kernel = "grid3_dmma_kernel<3,4>"
    bytes_per_point, flops_per_point, bound = 96.0, 400.0, "hbm"
    cpu_calls = ("evaluate", "jacobian")

    def setup(self, dev, rank, scale):
        import torch
        import bspy_b200 as bspy
        self.torch, self.bspy, self.dev = torch, bspy, dev
        rng = np.random.default_rng(1004)
        self.spline = bspy.Spline(3, 3, (4, 4, 4), (32, 32, 32), [knots_nonuniform(4, 32, rng) for _ in range(3)],
                                  rng.standard_normal((3, 32, 32, 32)))
        self.spline.freeze(dev)
        self.n = max(32, int(round(512 * scale)) // 16 * 16)
        self.axes_host = [np.linspace(0.0, 1.0, self.n) for _ in range(3)]
        self.axes = [torch.from_numpy(a).to(dev) for a in self.axes_host]
        self.points = self.n ** 3
        self.working_set = self.points * 96
        self.note = f"grid {self.n}^3; outputs {self.working_set / 1e9:.1f} GB per step"
        self.last = None

    def step(self):
        self.last = self.spline.evaluate_grid(*self.axes, values=True, jacobian=True, check_domain="defer")

    def flags_ok(self):
        return self.last is None or self.last.first_outside is None or int(self.last.first_outside.item()) < 0

    def e2e_step(self):
        r = self.spline.evaluate_grid(*self.axes_host, values=True, jacobian=True)
        d2h = r.values.nbytes + r.jacobian.nbytes
        del r
        return self.points, sum(a.nbytes for a in self.axes_host), d2h

    def reference_task(self, n):
        rng = np.random.default_rng(1004)
        s = _Plain(3, 3, (4, 4, 4), (32, 32, 32), [knots_nonuniform(4, 32, rng) for _ in range(3)], rng.standard_normal((3, 32, 32, 32)))
        g = np.linspace(0.0, 1.0, 512)
        return _payload(s), g[np.random.default_rng(5).integers(0, 512, (n, 3))]

    def parity(self, n=150_000):
        CO = _oracle()
        torch, m = self.torch, self.n
        rng = np.random.default_rng(44)
        idx = rng.integers(0, m, (n, 3))
        idx[:8] = np.array([[a, b, c] for a in (0, m - 1) for b in (0, m - 1) for c in (0, m - 1)])
        ia, ib, ic = (torch.from_numpy(idx[:, k]).to(self.dev) for k in range(3))
        v = self.last.values[:, ia, ib, ic].cpu().numpy().T
        j = self.last.jacobian[:, :, ia, ib, ic].cpu().numpy().transpose(2, 0, 1)
        uvw = np.stack([self.axes_host[k][idx[:, k]] for k in range(3)], axis=1)
        ref = CO.evaluate(self.spline, uvw, values=True, jacobian=True, spans=True)
        P = Parity()
        P.values(v, ref["values"], "values")
        P.values(j, ref["jacobian"], "jacobian")
        for k in range(3):
            sp = self.bspy._cuda.spans(torch.from_numpy(np.ascontiguousarray(self.spline.knots[k])).to(self.dev), 4, self.axes[k]).cpu().numpy()
            P.spans(sp[idx[:, k]], ref["spans"][:, k])
        P.n = n
        return P.report()


CONFIGS = {"cfg1": Cfg1Curve, "cfg1_1e8": Cfg1Curve1e8, "cfg2": Cfg2Teapot, "cfg3": Cfg3Curves, "cfg4": Cfg4Volume,
           "cfg4_soa": Cfg4VolumeSoA, "cfg5": Cfg5Manifold, "cfg5_soa": Cfg5ManifoldSoA, "grid3": Grid3Volume}
HEADLINE = "cfg2"
SECONDARY = ["cfg1", "cfg1_1e8", "cfg3", "cfg4", "cfg4_soa", "cfg5", "cfg5_soa", "grid3"]


# ------------------------------------------------------------------------ CPU reference arm

_WORKER = {}


def _cpu_worker(args):
    """One share of the sample on one core.  kind "reference": the UNMODIFIED reference package (baseline/_ref or
    /root/reference), its own Spline.evaluate / jacobian / normal, one call per point as its API requires.  kind "port":
    the scalar tier of the oracle = the reference's algorithm and cost model restated."""
    kind, payload, pts, calls = args
    sys.path.insert(0, ROOT)
    with np.errstate(all="ignore"):
        if kind == "reference":
            if "bspy" not in _WORKER:
                from baseline import load_reference
                _WORKER["bspy"], _ = load_reference.load()
            b = _WORKER["bspy"]
            s = b.Spline(payload["nInd"], payload["nDep"], payload["order"], payload["nCoef"], payload["knots"], payload["coefs"],
                         payload["metadata"])
            fns = [getattr(s, c) for c in calls]
            t0 = time.perf_counter()
            for p in pts:
                for f in fns:
                    f(p)
            return time.perf_counter() - t0
        from oracle import bspy_oracle as O
        s = O.OracleSpline(**payload)
        fns = [{"evaluate": O.evaluate_pt, "jacobian": O.jacobian_pt, "normal": O.normal_pt}[c] for c in calls]
        t0 = time.perf_counter()
        for p in pts:
            for f in fns:
                f(s, p)
        return time.perf_counter() - t0


def reference_kind():
    """("reference", root) when the unmodified reference imports on this box, else ("port", why)."""
    try:
        from baseline import load_reference
        mod, where = load_reference.load()
        if mod is not None:
            for k in [k for k in sys.modules if k == "bspy" or k.startswith("bspy.")]:
                sys.modules.pop(k, None)          # the parent process never keeps the reference imported
            if where in sys.path:
                sys.path.remove(where)
            return "reference", where
        return "port", where
    except Exception as exc:  # pragma: no cover
        return "port", f"{type(exc).__name__}: {exc}"


def cpu_points_per_second(kind, workload, n_points, cores, pool):
    payload, pts = workload.reference_task(n_points)
    chunks = np.array_split(pts, cores)
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(kind, payload, c, workload.cpu_calls) for c in chunks])
    wall = time.perf_counter() - t0
    return n_points / wall, wall


def cpu_baseline(kind, where, workload, seconds, cores, pool, native=True):
    cpu_points_per_second(kind, workload, cores * 8, cores, pool)                     # warms the workers (imports)
    rate, _ = cpu_points_per_second(kind, workload, cores * 100, cores, pool)
    n = int(max(cores * 100, min(rate * seconds, 5e6)))
    rate, wall = cpu_points_per_second(kind, workload, n, cores, pool)
    what = "the unmodified reference package (" + os.path.relpath(where, ROOT) + ")" if kind == "reference" else \
        "scalar port of the reference (oracle/bspy_oracle.py *_pt; reference not importable: " + str(where)[:120] + ")"
    out = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"{n} points of the same workload, {'+'.join(workload.cpu_calls)} per point, {what} in {cores} processes, {wall:.1f} s"}
    if native:
        out["native_port_value"] = cpu_native_points_per_second(workload, 400_000)
    return out


def cpu_native_points_per_second(workload, n_points):
    """The C/OpenMP restatement on all cores (context only: far faster than the reference itself)."""
    try:
        CO = _oracle()
        payload, pts = workload.reference_task(n_points)
        s = _Plain(**payload)
        calls = workload.cpu_calls
        t0 = time.perf_counter()
        CO.evaluate(s, pts, values="evaluate" in calls, jacobian="jacobian" in calls, normal="normal" in calls)
        return n_points / (time.perf_counter() - t0)
    except Exception:
        return None


# ------------------------------------------------------------------------------- utilities

class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region: NVML in a background thread
    (~1 ms period, so that even a 10 ms region gets samples), `nvidia-smi -lms` as a fallback."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() else gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        it = 0
        while not self.stop_flag:
            try:
                clk = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                # the power query is the slow one (milliseconds): every 8th round only, so that a 20 ms timed
                # region still collects clock samples
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0 if it % 8 == 0 else None
                self.rows.append((time.perf_counter(), clk, pw, [k for k, b in bits.items() if mask & b]))
            except Exception:
                pass
            it += 1
            time.sleep(0.0005)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.smax = float(f[1])
                self.rows.append((time.perf_counter(), float(f[0]), float(f[2]) if f[2][:1].isdigit() else None,
                                  [nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except Exception:
                continue

    def stop(self, t0, t1):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        time.sleep(0.03)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        sm = [r[1] for r in inside]
        power = [r[2] for r in inside if r[2] is not None]
        reasons = sorted({x for r in inside for x in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": getattr(self, "smax", None), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(power) if power else None,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured copy bandwidth (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg, points_per_launch):
    """DRAM bytes per launch of the dominant kernel, scaled from the committed `ncu --set full` capture
    (profiles/traffic.json holds bytes per point and the capture it came from); None if there is none."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(path))[cfg]
        return float(e["dram_bytes_per_point"]) * points_per_launch, f"profiles/traffic.json ({e.get('capture', 'ncu --set full')}), scaled by points per launch; not measured in this run"
    except Exception:
        return None, None


def fp64_peaks(_cuda, torch, dev):
    """Best of 5 of the dependent-free DFMA probe and of the DMMA.8x8x4 probe (TFLOP/s); cfg4 / cfg5 are graded on the
    higher of the two (MEASURED_PEAKS.json carries no FP64 figure)."""
    out = {}
    for kind, name in ((0, "dfma"), (1, "dmma")):
        _cuda.probe_fp64(kind, 2000, dev)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            flops = _cuda.probe_fp64(kind, 20000, dev)
            b.record()
            torch.cuda.synchronize()
            best = max(best, flops / (a.elapsed_time(b) * 1e-3) / 1e12)
        out[name + "_tflops"] = best
    out["peak_tflops"] = max(out["dfma_tflops"], out["dmma_tflops"])
    out["source"] = "live probes in this run (bspy_cuda_probe_fp64: dependent-free DFMA chains / DMMA.8x8x4 chains on every SM), best of 5"
    return out


def hbm_probes(_cuda, torch, dev, flush, working_set):
    nd = 1 << 28
    src = torch.empty(nd, dtype=torch.float64, device=dev)
    dst = torch.empty(nd, dtype=torch.float64, device=dev)
    probes = {}
    for kind, label in ((0, "copy_gbs"), (1, "write_only_gbs")):
        best = 0.0
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nbytes = _cuda.probe_hbm(kind, src, dst)
            b.record()
            torch.cuda.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        probes[label] = best
    if flush is not None:
        # small steps never see the asymptotic bandwidth: a plain copy of the same number of bytes under the same
        # protocol (L2 flushed before each launch), and an 8 KB copy = what one launch costs between two events
        for label, nd2 in (("copy_of_step_bytes_us", max(1024, int(working_set) // 16)), ("launch_floor_us", 512)):
            ts = []
            for _ in range(6):
                flush.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _cuda.probe_hbm(0, src[:nd2], dst[:nd2])
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3)
            probes[label] = min(ts[1:]) * 1e6
    del src, dst
    return probes


# ------------------------------------------------------------------------------------ arms

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = HEADLINE if args.config == "all" else args.config
    wl = CONFIGS[cfg]()
    cores = os.cpu_count() or 1
    kind, where = reference_kind()
    with mp.get_context("spawn").Pool(cores) as pool:
        # calibrate so that one step is ~3 s of wall time on all cores (first map warms the workers)
        cpu_points_per_second(kind, wl, cores * 8, cores, pool)
        rate, _ = cpu_points_per_second(kind, wl, cores * 200, cores, pool)
        n = int(max(cores * 100, min(rate * 3.0, 5e6)))
        for _ in range(args.warmup):
            cpu_points_per_second(kind, wl, max(cores * 50, n // 10), cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_points_per_second(kind, wl, n, cores, pool)
        wall = time.perf_counter() - t0
    value = n * args.steps / wall
    what = f"the unmodified reference package ({os.path.relpath(where, ROOT)})" if kind == "reference" else \
        f"scalar port of the reference (oracle/bspy_oracle.py *_pt); reference not importable: {str(where)[:160]}"
    sample = f"{n} points per step of the same workload ({'+'.join(wl.cpu_calls)} per point, {what}, {cores} processes)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl.name, "sample_points_per_step": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


class Runner:
    """Times one workload on this rank (CUDA events per step on the launching stream, max over ranks) and builds its
    report: value, roofline (both fractions), parity of the timed outputs, e2e through the public API, clocks."""

    def __init__(self, args, rank, world, local, dev):
        import torch
        import torch.distributed as dist
        from bspy_b200 import _cuda
        self.args, self.rank, self.world, self.local, self.dev = args, rank, world, local, dev
        self.torch, self.dist, self._cuda = torch, dist, _cuda
        self.flush = None
        self.hbm_peak, self.hbm_how = measured_hbm_peak()
        self.fp64 = None
        self.total_launches = 0

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, flush):
        """per-step CUDA events on the launching stream; the L2 flush (when needed) sits between the event pairs,
        outside the timed regions"""
        torch = self.torch
        evs = []
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) * 1e-3 for a, b in evs]

    def run(self, cfg, headline):
        torch, _cuda, args = self.torch, self._cuda, self.args
        wl = CONFIGS[cfg]()
        wl.setup(self.dev, self.rank, args.scale)
        flush = None
        if wl.working_set < 2 * L2_BYTES:
            if self.flush is None:
                self.flush = torch.empty(int(2 * L2_BYTES) // 8, dtype=torch.float64, device=self.dev)
            flush = self.flush
        steps = args.steps if headline else max(3, min(args.steps, args.secondary_steps))
        if not headline and wl.working_set < 2 * L2_BYTES:
            steps = max(steps, 40)            # a 15 us step between L2 flushes: the mean of 5 moves by 10 % from run to run
        warmup = max(args.warmup, 3)
        for _ in range(warmup):
            wl.step()
        self.barrier()
        # Launch-bound steps (config 1: 5 us of HBM traffic; the cell-sorted path: dozens of launches on three streams) are
        # captured once in a CUDA graph and replayed, so that the timed region holds GPU work, not Python overhead.
        step_fn, graphed = wl.step, False
        if args.graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    wl.step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    wl.step()
                graph.replay()
                torch.cuda.synchronize()
                step_fn, graphed = graph.replay, True
            except Exception as exc:  # capture not possible: time eager launches
                sys.stderr.write(f"[{cfg}] CUDA graph capture failed, timing eager launches: {exc}\n")
                torch.cuda.synchronize()
        launches_per_step = None
        if graphed:
            c0 = _cuda.launch_count()
            wl.step()
            torch.cuda.synchronize()
            launches_per_step = _cuda.launch_count() - c0
            graph.replay()                      # the timed outputs (and flags) are the graph's
            torch.cuda.synchronize()
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        l0 = _cuda.launch_count()
        self.barrier()
        t0 = time.perf_counter()
        per_step = self.timed(step_fn, steps, flush)
        self.barrier()
        t1 = time.perf_counter()
        launches = launches_per_step * steps if graphed else _cuda.launch_count() - l0
        self.total_launches += launches
        clocks = sampler.stop(t0, t1) if sampler else None
        if sampler and clocks and clocks.get("samples", 0) < 3:
            # timed region shorter than the sampling period: sample during an extra, untimed 200 ms of the same step
            extra = ClockSampler(self.local)
            ta_ = time.perf_counter()
            while time.perf_counter() - ta_ < 0.2:
                step_fn()
            torch.cuda.synchronize()
            more = extra.stop(ta_, time.perf_counter())
            if more.get("samples", 0) > clocks.get("samples", 0):
                more["note"] = f"timed region of {1e3 * (t1 - t0):.1f} ms held {clocks.get('samples', 0)} samples; sampled during 200 ms more of the same step right after it"
                clocks = more
        seconds = self.max_over_ranks(sum(per_step))
        value = wl.points * self.world * steps / seconds
        domain_ok = wl.flags_ok()

        # ---- parity of the timed outputs (rank 0) ----
        parity = None
        if self.rank == 0 and not args.no_parity:
            try:
                parity = wl.parity()
            except Exception as exc:  # a crash of the checker is a failed check, not a skipped one
                parity = {"ok": False, "error": f"{type(exc).__name__}: {exc}"}
            parity["domain_flag_clear"] = bool(domain_ok)
            parity["ok"] = bool(parity.get("ok")) and bool(domain_ok)

        # ---- end to end through the public API with host buffers ----
        e2e = None
        if headline or not args.no_secondary_e2e:
            e2e_steps = max(1, min(steps, args.e2e_steps)) if headline else 1
            wl.e2e_step()                     # warm-up: pinned staging buffers are allocated here
            self.barrier()
            ta = time.perf_counter()
            for _ in range(e2e_steps):
                n_e2e, h2d, d2h = wl.e2e_step()
            torch.cuda.synchronize()
            tb = self.max_over_ranks(time.perf_counter() - ta)
            e2e = {"value": n_e2e * self.world * e2e_steps / tb, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "steps": e2e_steps}
            if n_e2e != wl.points:
                e2e["sample"] = f"leading {n_e2e} of {wl.points} points per GPU (bounded pinned-memory footprint)"

        report = None
        if self.rank == 0:
            per_launch = max(1, launches // steps)
            step_s = float(np.mean(per_step))
            if self.fp64 is None:
                try:
                    self.fp64 = fp64_peaks(_cuda, torch, self.dev)
                except Exception as exc:  # pragma: no cover
                    self.fp64 = {"error": str(exc), "peak_tflops": 37.0, "source": "nominal (probe failed)"}
            hbm_gbs = wl.bytes_per_point * wl.points / step_s / 1e9
            tflops = wl.flops_per_point * wl.points / step_s / 1e12
            traffic, traffic_src = ncu_traffic(cfg, wl.points)
            roof = {"bound": wl.bound, "kernel": wl.kernel, "step_ms": step_s * 1e3, "launches_per_step": per_launch,
                    "hbm_frac": hbm_gbs / self.hbm_peak, "fp64_frac": tflops / self.fp64["peak_tflops"],
                    "algorithmic_bytes_per_point": wl.bytes_per_point, "algorithmic_flops_per_point": wl.flops_per_point,
                    "traffic": traffic, "traffic_source": traffic_src}
            if wl.bound == "fp64":
                roof.update(achieved=tflops, peak=self.fp64["peak_tflops"], unit="TFLOP/s", frac=tflops / self.fp64["peak_tflops"],
                            peak_source=self.fp64.get("source"), fp64_probes={k: v for k, v in self.fp64.items() if k.endswith("_tflops")})
            else:
                roof.update(achieved=hbm_gbs, peak=self.hbm_peak, unit="GB/s", frac=hbm_gbs / self.hbm_peak, peak_source=self.hbm_how)
            if per_launch > 1:
                roof["note"] = "whole step (all launches of the pipeline) against the roofline of the path's algorithmic work"
            if headline or flush is not None:
                try:
                    roof["hbm_probe"] = hbm_probes(_cuda, torch, self.dev, flush, wl.working_set)
                except Exception as exc:  # pragma: no cover
                    roof["hbm_probe"] = {"error": str(exc)}
            report = {"workload": wl.name, "value": value, "unit": UNIT, "ms_per_step": seconds / steps * 1e3, "steps": steps,
                      "points_per_step_per_gpu": wl.points, "gpu_launches": launches, "cuda_graph": graphed, "note": wl.note,
                      "l2": "L2 flushed (write of 252 MB) between timed steps" if flush is not None else "per-step working set larger than L2",
                      "domain_check": "on the device in every timed step (flag read once after the timed region)",
                      "roofline": roof, "parity": parity, "e2e": e2e, "clocks": clocks}
        self.wl = wl
        return report


def strong_scaling_job(R):
    """ONE job through the public API split over the ranks: a single 1e8-point cfg4 batch, present on every
    rank, split with shard_points, evaluated, and re-assembled on every rank with an NCCL all-gather of the
    [values | jacobian] records.  Reports kernel-only and gather-inclusive points/s (max over ranks)."""
    torch, dist = R.torch, R.dist
    from bspy_b200.sharding import gather_records, shard_points
    wl = Cfg4Volume()
    import bspy_b200 as bspy
    rng = np.random.default_rng(wl.seed)
    spline = wl.make_spline(rng, bspy)
    spline.freeze(R.dev)
    N = max(8192, int(wl.N * R.args.scale))
    g = torch.Generator(device=R.dev).manual_seed(4242)                     # the same batch on every rank
    pts = torch.rand((N, 3), dtype=torch.float64, device=R.dev, generator=g)
    mine = shard_points(pts, R.rank, R.world)
    full = torch.empty((N, 12), dtype=torch.float64, device=R.dev)

    def job(with_gather):
        r = spline.evaluate_points(mine, values=True, jacobian=True, check_domain="defer", out_layout="aos")
        if with_gather:
            gather_records(r.records, N, out=full)
        return r

    for _ in range(3):
        job(True)
    out = {}
    for label, flag in (("kernel_only", False), ("gather_inclusive", True)):
        R.barrier()
        ts = R.timed(lambda: job(flag), 5, None)
        out[label] = N * 5 / R.max_over_ranks(sum(ts))
    # the gathered records are every rank's shard in order: compare two far-apart rows with a local evaluation
    probe = torch.tensor([0, N // 2, N - 1], device=R.dev)
    local = spline.evaluate_points(pts[probe].contiguous(), values=True, jacobian=True, out_layout="aos").records
    ok = bool(torch.allclose(full[probe][:, :12], local[:, :12], rtol=1e-12, atol=1e-13))
    ok_all = R.max_over_ranks(0.0 if ok else 1.0) == 0.0
    del pts, full
    return {"workload": f"one cfg4 batch of {N} points split over {R.world} ranks with shard_points; [values|jacobian] records "
                        f"all-gathered to every rank over NCCL ({N * 96 / 1e9:.1f} GB per rank)",
            "scaling": "strong", "kernel_only_points_per_s": out["kernel_only"], "gather_inclusive_points_per_s": out["gather_inclusive"],
            "gathered_equals_local": ok_all}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bspy_b200.sharding import bind_to_gpu_numa_node, init_from_env

    rank, world, local = init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    numa = bind_to_gpu_numa_node(dev.index)           # pinned host buffers land on the GPU's own NUMA node
    R = Runner(args, rank, world, local, dev)
    head_cfg = HEADLINE if args.config == "all" else args.config
    t_start = time.perf_counter()
    head = R.run(head_cfg, True)
    head_wl = R.wl

    cpu = None
    pool = None
    kind = where = None
    cores = os.cpu_count() or 1
    if rank == 0 and not args.no_cpu:
        kind, where = reference_kind()
        pool = mp.get_context("spawn").Pool(cores)
        cpu = cpu_baseline(kind, where, head_wl, args.cpu_seconds, cores, pool)
    head_wl.teardown()
    torch.cuda.empty_cache()

    configs = {}
    if args.config == "all":
        for cfg in SECONDARY:
            if args.only and cfg not in args.only:
                continue
            try:
                rep = R.run(cfg, False)
                if rank == 0:
                    if pool is not None:
                        rep["cpu_baseline"] = cpu_baseline(kind, where, R.wl, args.secondary_cpu_seconds, cores, pool, native=False)
                    configs[cfg] = rep
            except torch.cuda.OutOfMemoryError as exc:  # pragma: no cover
                if rank == 0:
                    configs[cfg] = {"error": f"out of memory: {exc}"[:200], "parity": {"ok": False}}
            R.wl.teardown()
            torch.cuda.empty_cache()
            try:
                torch._C._host_emptyCache()
            except Exception:
                pass
    strong = None
    if world > 1 and args.config == "all":
        try:
            strong = strong_scaling_job(R)
        except Exception as exc:  # pragma: no cover
            strong = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.empty_cache()
    if pool is not None:
        pool.close()
        pool.join()
    R.barrier()
    rc = 0
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": head["workload"], "points_per_step_per_gpu": head["points_per_step_per_gpu"], "note": head["note"],
                           "l2": head["l2"], "sharding": "each rank evaluates its own shard (no data-path collective)",
                           "cuda_graph": head["cuda_graph"], "domain_check": head["domain_check"], "host_numa_binding": numa},
                "clocks": head["clocks"], "gpu_launches": head["gpu_launches"], "e2e": head["e2e"], "roofline": head["roofline"],
                "cpu_baseline": cpu, "parity": head["parity"]}
        if configs:
            line["configs"] = configs
            line["gpu_launches_all_configs"] = R.total_launches
        if strong:
            line["strong_scaling"] = strong
        line["wall_s"] = time.perf_counter() - t_start
        bad = [k for k, v in [(head_cfg, head)] + list(configs.items()) if v.get("parity") is not None and not v["parity"].get("ok")]
        if bad:
            line["parity_failed"] = bad
            rc = 3
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="all", choices=["all"] + sorted(CONFIGS))
    ap.add_argument("--only", nargs="*", default=None, help="with --config all: restrict the secondary configs to these")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workloads (testing only; 1.0 = BASELINE size)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--secondary-steps", type=int, default=10)
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--secondary-cpu-seconds", type=float, default=1.5)
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
