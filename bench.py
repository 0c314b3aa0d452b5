#!/usr/bin/env python
"""bench.py -- float64 spline point-evaluations per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload
(--config cfg2, BASELINE.json configs[1]): the 32 bicubic Utah-teapot patches evaluated with
value, d/du, d/dv and unit normal on a 2048 x 2048 grid per patch (134 M points, 12.9 GB of
output per step per GPU).  Other BASELINE configs: cfg1 (1 M points on a cubic 3-D curve), cfg3
(1 M curves x 256 points), cfg4 (trivariate volume, 1e8 scattered points, value + jacobian), cfg5
(nInd 4 / nDep 6 manifold, 1.25e8 points per GPU, value + first derivatives).

One JSON line is printed by rank 0 (see the contract in the task description): `value` is
whole-job throughput with inputs resident in HBM, `e2e` the same metric through the public API
with host buffers (H2D and D2H inside the timed region), `roofline` the dominant kernel against
the measured HBM peak, `cpu_baseline` the reference's algorithm timed on this box's host cores.
Multi-GPU: one process per GPU (torchrun), every rank evaluates its own shard of patches / points
/ curves (weak scaling, no data-path collective), time = max over ranks.

`--impl reference` times the reference's own CPU algorithm (the oracle's scalar tier: one
interpreter pass per point through the same recurrence and numpy calls as
bspy/_spline_evaluation.py) on all host cores, on a bounded sample of the same workload.
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


class Workload:
    name = ""
    kernel = ""
    bytes_per_point = 0.0     # algorithmic (compulsory) HBM bytes per point, SURVEY 8(d)
    flops_per_point = 0.0
    points = 0                # per step per GPU
    note = ""

    def setup(self, dev, rank, scale):
        raise NotImplementedError

    def step(self):           # device-resident inputs and outputs
        raise NotImplementedError

    def e2e_step(self):       # host buffers in, host buffers out; returns (h2d_bytes, d2h_bytes)
        raise NotImplementedError

    def reference_task(self, n):  # (payload for cpu worker) describing n sample points
        raise NotImplementedError


def _spline_payload(s):
    return dict(nInd=s.nInd, nDep=s.nDep, order=s.order, nCoef=s.nCoef, knots=[np.asarray(k) for k in s.knots],
                coefs=np.asarray(s.coefs), metadata=dict(s.metadata))


class Cfg2Teapot(Workload):
    name = "cfg2: 32 bicubic Utah-teapot patches, value+du+dv+unit normal on a 2048x2048 grid per patch"
    kernel = "grid2_dmma_kernel<3,4,false>"
    bytes_per_point = 96.0
    flops_per_point = 92.0
    bound = "hbm"
    calls = "evaluate+jacobian+normal"

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

    def step(self):
        self.batch.evaluate_grid(self.axis, self.axis, jacobian=True, normal=True, check_domain=False, out=self.out)

    def e2e_step(self):
        # host splines + host axes in, host arrays out (public API: SplineBatch.from_splines + evaluate_grid)
        batch = self.bspy.SplineBatch.from_splines(self.splines, device=self.dev)
        r = batch.evaluate_grid(self.axis_host, self.axis_host, jacobian=True, normal=True)
        h2d = self.coefs_host.nbytes + self.kn.nbytes * 2 + self.axis_host.nbytes * 2
        d2h = r.values.nbytes + r.jacobian.nbytes + r.normal.nbytes
        del r                     # the pinned result buffers go back to torch's host allocator for the next step
        return h2d, d2h

    def reference_task(self, n):
        g = np.linspace(0.0, 1.0, self.n if hasattr(self, "n") else 2048)
        rng = np.random.default_rng(2)
        ia, ib = rng.integers(0, len(g), n), rng.integers(0, len(g), n)
        coefs, kn = teapot_patches()
        sp = dict(nInd=2, nDep=3, order=(4, 4), nCoef=(4, 4), knots=[kn, kn], coefs=coefs[3], metadata={})
        return sp, np.stack([g[ia], g[ib]], axis=1), ("evaluate", "jacobian", "normal")


class ScatteredBase(Workload):
    bound = "hbm"
    jac = False
    seed = 0
    N = 0

    def make_spline(self, rng):
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
        self.note = f"{self.points} points per GPU; working set {self.working_set / 1e9:.2f} GB"
        self.host_pts = None

    def step(self):
        self.last = self.spline.evaluate_points(self.pts, values=True, jacobian=self.jac, check_domain=False)

    def e2e_step(self):
        torch = self.torch
        if self.host_pts is None:
            self.host_pts = torch.empty(self.pts.shape, dtype=torch.float64, pin_memory=True)
            self.host_pts.copy_(self.pts)
        r = self.spline.evaluate_points(self.host_pts, values=True, jacobian=self.jac)
        d2h = r.values.numel() * 8 + (r.jacobian.numel() * 8 if r.jacobian is not None else 0)
        del r
        return self.host_pts.numel() * 8, d2h

    def reference_task(self, n):
        rng = np.random.default_rng(self.seed)
        import types
        s = self.make_spline(rng, types.SimpleNamespace(Spline=lambda *a: types.SimpleNamespace(
            nInd=a[0], nDep=a[1], order=tuple(a[2]), nCoef=tuple(a[3]), knots=a[4], coefs=a[5], metadata={})))
        pts = np.random.default_rng(self.seed + 1).uniform(0, 1, (n, s.nInd))
        return _spline_payload(s), pts, ("evaluate", "jacobian") if self.jac else ("evaluate",)


class Cfg1Curve(ScatteredBase):
    name = "cfg1: cubic 3-D curve, 64 coefficients, non-uniform knots, 1M random parameters, values"
    kernel = "eval_curve_repl_kernel<4,3,false>"
    bytes_per_point, flops_per_point, seed, N = 32.0, 66.0, 1001, 1_000_000

    def make_spline(self, rng, bspy):
        return bspy.Spline(1, 3, (4,), (64,), [knots_nonuniform(4, 64, rng)], rng.standard_normal((3, 64)))


class Cfg4Volume(ScatteredBase):
    name = "cfg4: trivariate order-4 volume (nInd 3, nDep 3, 32^3 coefficients), 1e8 scattered points, value + jacobian"
    kernel = "cell-binned pipeline: bin_keys, bin_scan, bin_scatter_records, eval_staged_kernel<3,4,4,4,0,3,true,3,4>, bin_unpermute (whole step)"
    bytes_per_point, flops_per_point, seed, N, jac = 120.0, 1320.0, 1004, 100_000_000, True

    def make_spline(self, rng, bspy):
        return bspy.Spline(3, 3, (4, 4, 4), (32, 32, 32), [knots_nonuniform(4, 32, rng) for _ in range(3)],
                           rng.standard_normal((3, 32, 32, 32)))


class Cfg5Manifold(ScatteredBase):
    name = "cfg5: nInd 4 / nDep 6 order-3 manifold (16^4 coefficients), 1.25e8 scattered points per GPU, value + first derivatives"
    kernel = "cell-binned pipeline: bin_keys, bin_scan, bin_scatter_records, eval_fixed_kernel<4,3,3,3,3,6,true,1,4>, bin_unpermute (whole step)"
    bytes_per_point, flops_per_point, seed, N, jac = 272.0, 3650.0, 1005, 125_000_000, True

    def make_spline(self, rng, bspy):
        return bspy.Spline(4, 6, (3,) * 4, (16,) * 4, [knots_nonuniform(3, 16, rng) for _ in range(4)],
                           rng.standard_normal((6, 16, 16, 16, 16)))


class Cfg3Curves(Workload):
    name = "cfg3: 1M independent cubic 3-D curves (32 coefficients each), 256 points per curve"
    kernel = "many_kernel<4,false>"
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
        self.note = f"{self.S} curves per GPU; working set {self.working_set / 1e9:.2f} GB"
        self.host = None

    def step(self):
        self.batch.evaluate(self.u, check_domain=False, out=self.out)

    def e2e_step(self):
        torch = self.torch
        if self.host is None:
            self.host = [torch.empty(t.shape, dtype=torch.float64, pin_memory=True).copy_(t) for t in (self.knots, self.coefs, self.u)]
            self.host_out = torch.empty((self.S, 3, 256), dtype=torch.float64, pin_memory=True)
        k, c, u = (t.to(self.dev, non_blocking=True) for t in self.host)
        b = self.bspy.SplineBatch(1, 3, (4,), (32,), [k], c)
        r = b.evaluate(u)
        self.host_out.copy_(r.values)
        return sum(t.numel() * 8 for t in self.host), self.host_out.numel() * 8

    def reference_task(self, n):
        rng = np.random.default_rng(1003)
        sp = dict(nInd=1, nDep=3, order=(4,), nCoef=(32,), knots=[knots_nonuniform(4, 32, rng)],
                  coefs=rng.standard_normal((3, 32)), metadata={})
        return sp, rng.uniform(0, 1, (n, 1)), ("evaluate",)


CONFIGS = {"cfg1": Cfg1Curve, "cfg2": Cfg2Teapot, "cfg3": Cfg3Curves, "cfg4": Cfg4Volume, "cfg5": Cfg5Manifold}


# ------------------------------------------------------------------------ CPU reference arm

def _cpu_worker(args):
    """Scalar tier of the oracle = the reference's algorithm and cost model: one Python pass per
    point and per call (evaluate / jacobian / normal), numpy float64 scalars."""
    payload, pts, calls = args
    sys.path.insert(0, ROOT)
    from oracle import bspy_oracle as O
    s = O.OracleSpline(**payload)
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        for p in pts:
            if "evaluate" in calls:
                O.evaluate_pt(s, p)
            if "jacobian" in calls:
                O.jacobian_pt(s, p)
            if "normal" in calls:
                O.normal_pt(s, p)
    return time.perf_counter() - t0


def cpu_points_per_second(workload, n_points, cores, pool):
    payload, pts, calls = workload.reference_task(n_points)
    chunks = np.array_split(pts, cores)
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(payload, c, calls) for c in chunks])
    wall = time.perf_counter() - t0
    return n_points / wall, wall, calls


def cpu_native_points_per_second(workload, n_points):
    """The C/OpenMP restatement on all cores (context only: far faster than the reference itself)."""
    try:
        from oracle import c_oracle as CO
        from oracle import bspy_oracle as O
        CO.build()
        payload, pts, calls = workload.reference_task(n_points)
        s = O.OracleSpline(**payload)
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
        self.last_inside = len(inside)
        power = [r[2] for r in inside if r[2] is not None]
        reasons = sorted({x for r in inside for x in r[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": getattr(self, "smax", None), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(power) if power else None,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg, points_per_launch):
    """DRAM bytes per launch of the dominant kernel, scaled from the committed `ncu --set full` capture
    (profiles/traffic.json holds bytes per point and the capture it came from); None if there is none."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return float(json.load(open(path))[cfg]["dram_bytes_per_point"]) * points_per_launch
    except Exception:
        return None


# ------------------------------------------------------------------------------------ arms

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = CONFIGS[args.config]()
    cores = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        # calibrate so that one step is ~3 s of wall time on all cores (first map warms the workers)
        cpu_points_per_second(wl, cores * 20, cores, pool)
        rate, _, calls = cpu_points_per_second(wl, cores * 400, cores, pool)
        n = int(max(cores * 200, min(rate * 3.0, 5e6)))
        for _ in range(args.warmup):
            cpu_points_per_second(wl, max(cores * 100, n // 10), cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_points_per_second(wl, n, cores, pool)
        wall = time.perf_counter() - t0
    value = n * args.steps / wall
    sample = f"{n} points per step of the same workload ({'+'.join(calls)} per point, scalar port of the reference, {cores} processes)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl.name, "sample_points_per_step": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bspy_b200 import _cuda
    from bspy_b200.sharding import init_from_env

    rank, world, local = init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    wl = CONFIGS[args.config]()
    wl.setup(dev, rank, args.scale)
    flush = torch.empty(int(2 * L2_BYTES) // 8, dtype=torch.float64, device=dev) if wl.working_set < 2 * L2_BYTES else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """per-step CUDA events on the launching stream; the L2 flush (when needed) sits between the
        event pairs, outside the timed regions"""
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

    for _ in range(max(args.warmup, 3)):
        wl.step()
    barrier()
    # Launch-bound steps (config 1: 5 us of HBM traffic; the binned path: hundreds of small launches) are
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
            sys.stderr.write(f"CUDA graph capture failed, timing eager launches: {exc}\n")
            torch.cuda.synchronize()
    launches_per_step = None
    if graphed:
        c0 = _cuda.launch_count()
        wl.step()
        torch.cuda.synchronize()
        launches_per_step = _cuda.launch_count() - c0
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _cuda.launch_count()
    t0 = time.perf_counter()
    per_step = timed(step_fn, args.steps)
    barrier()
    t1 = time.perf_counter()
    launches = launches_per_step * args.steps if graphed else _cuda.launch_count() - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    if sampler and clocks and clocks.get("samples", 0) < 3:
        # timed region shorter than the sampling period: sample during an extra, untimed 200 ms of the same step
        extra = ClockSampler(local)
        ta_ = time.perf_counter()
        while time.perf_counter() - ta_ < 0.2:
            step_fn()
        torch.cuda.synchronize()
        more = extra.stop(ta_, time.perf_counter())
        if more.get("samples", 0) > clocks.get("samples", 0):
            more["note"] = f"timed region of {1e3 * (t1 - t0):.1f} ms held {clocks.get('samples', 0)} samples; sampled during 200 ms more of the same step right after it"
            clocks = more
    total = torch.tensor([sum(per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    seconds = float(total.item())
    value = wl.points * world * args.steps / seconds

    # ---- end to end through the public API with host buffers ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    wl.e2e_step()                     # warm-up: pinned staging buffers are allocated here
    barrier()
    ta = time.perf_counter()
    for _ in range(e2e_steps):
        h2d, d2h = wl.e2e_step()
    torch.cuda.synchronize()
    tb = time.perf_counter() - ta
    te = torch.tensor([tb], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = wl.points * world * e2e_steps / float(te.item())

    if rank != 0:
        return 0
    peak, how = measured_peaks()
    launch_s = float(np.mean(per_step)) / max(1, launches // args.steps)
    achieved = wl.bytes_per_point * wl.points / max(1, launches // args.steps) / launch_s / 1e9
    roofline = {"bound": wl.bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.config, wl.points / max(1, launches // args.steps)), "kernel": wl.kernel, "peak_source": how,
                "algorithmic_bytes_per_point": wl.bytes_per_point, "launch_ms": launch_s * 1e3}
    # FP64 context: flops/point x points/s against a live FMA probe
    try:
        it = 4096
        _cuda.probe_fp64(0, it, dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        flops = _cuda.probe_fp64(0, it, dev)
        b.record()
        torch.cuda.synchronize()
        fp64_peak = flops / (a.elapsed_time(b) * 1e-3) / 1e12
        roofline["fp64"] = {"achieved_tflops": wl.flops_per_point * wl.points / launch_s / max(1, launches // args.steps) / 1e12,
                            "peak_tflops_probe": fp64_peak, "flops_per_point": wl.flops_per_point}
    except Exception as exc:  # pragma: no cover
        roofline["fp64"] = {"error": str(exc)}

    # live HBM probes on this device: copy (read+write bytes) and write-only fill of 2 GiB
    try:
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
            # small steps never see the asymptotic bandwidth: time a plain copy of the same number of bytes under
            # the same protocol (L2 flushed before each launch) as the reachable reference for this step size
            nd2 = max(1024, int(wl.working_set) // 16)
            ts = []
            for _ in range(6):
                flush.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                nbytes = _cuda.probe_hbm(0, src[:nd2], dst[:nd2])
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3)
            probes["copy_of_step_bytes_us"] = min(ts[1:]) * 1e6
            probes["copy_of_step_bytes_gbs"] = nbytes / min(ts[1:]) / 1e9
            # ... and of 8 KB: what one launch costs between two events after the flush, with nothing to do
            ts = []
            for _ in range(6):
                flush.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _cuda.probe_hbm(0, src[:512], dst[:512])
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3)
            probes["launch_floor_us"] = min(ts[1:]) * 1e6
        roofline["hbm_probe"] = probes
        del src, dst
    except Exception as exc:  # pragma: no cover
        roofline["hbm_probe"] = {"error": str(exc)}

    cpu = None
    if world == 1 or rank == 0:
        cores = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(cores) as pool:
            cpu_points_per_second(wl, cores * 20, cores, pool)
            rate, _, calls = cpu_points_per_second(wl, cores * 400, cores, pool)
            n = int(max(cores * 200, min(rate * args.cpu_seconds, 5e6)))
            rate, wall, calls = cpu_points_per_second(wl, n, cores, pool)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} points of the same workload, {'+'.join(calls)} per point, scalar port of the reference "
                         f"(oracle/bspy_oracle.py *_pt) in {cores} processes, {wall:.1f} s",
               "native_port_value": cpu_native_points_per_second(wl, min(2_000_000, wl.points))}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "points_per_step_per_gpu": wl.points, "note": wl.note,
                       "l2": "L2 flushed (write of 252 MB) between timed steps" if flush is not None else "per-step working set larger than L2",
                       "sharding": "each rank evaluates its own shard (no data-path collective)",
                       "cuda_graph": graphed},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
            "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only; 1.0 = BASELINE size)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
