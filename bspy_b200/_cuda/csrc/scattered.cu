// Scattered-point evaluation: bspy_cuda_eval_points.
//
// Replaces evaluate / derivative / jacobian / normal of bspy/_spline_evaluation.py:109-246 for N
// points per launch.  One thread per point:
//   1. knot span per variable (upper-bound bisection, bit-exact with np.searchsorted 'right'),
//   2. Cox-de Boor basis values (and first derivatives) in registers from the 2(o-1)-knot window,
//      the o-2 lower stages shared between values and derivatives,
//   3. sum-factorised tensor-product contraction, last variable first (the reference's order),
//      carrying one "derivative already taken" slot per variable so that the value and the whole
//      jacobian come out of ONE pass over the coefficient window (the reference makes nInd+1
//      passes and nInd^2 + nInd basis evaluations),
//   4. optional normal: signed cofactors of the jacobian, optional 2-norm division,
//   5. struct-of-arrays stores: consecutive threads write consecutive doubles.
// Two implementations: eval_fixed<> (compile-time nInd/orders/nDep, everything in registers) for
// the common shapes, eval_generic for any nInd <= 8, order <= 32 and any nDep.
#include <stdlib.h>

#include "scattered.cuh"

namespace bspy {

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, bool JAC, int NDT = NDEP, int MINB = 0>
__global__ void __launch_bounds__(128, MINB ? MINB : fixed_min_blocks(NIND, O0, O1, O2, O3, NDT, JAC)) eval_fixed_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                         const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    if (gate_closed(in)) return;
    const bool recs = in.records != nullptr;
    const bool pairs = !recs && in.recKI != nullptr;              // sorted (cell key, index) pairs: parameters gathered here
    const bool binned = in.perm != nullptr || recs || pairs;
    // sorted records of an even-padded sort: the slot count is on the device (<= N, the caller's upper bound), spare slots
    // hold dummy records (index -1) that are skipped
    const long long limit = ((recs || pairs) && in.sortedTotal) ? min(N, (long long)__ldg(in.sortedTotal)) : N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < limit; t += (long long)gridDim.x * blockDim.x) {
        FixedCtx<Ord, NDT, JAC> c;
        int ix[NIND];
        double u[NIND];
        bool outside = false;
        long long p = t, dest = t;                        // dest: where the result record goes (array-of-structs modes)
        if (binned) {
            int key;
            if (recs) {
                const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
                const double2 r0 = __ldcs(rp), r1 = __ldcs(rp + 1);
                u[0] = r0.x;
                if constexpr (NIND > 1) u[1] = r0.y;
                if constexpr (NIND > 2) u[2] = r1.x;
                if constexpr (NIND > 3) u[3] = r1.y;
                const long long ki = NIND > 3 ? __ldcs(reinterpret_cast<const long long *>(in.recKI) + t) : __double_as_longlong(r1.y);
                key = (int)ki;
                if ((ki >> 32) < 0) continue;                 // dummy slot
                dest = out.aosScatter ? out.aosBase + (ki >> 32) : t;
            } else if (pairs) {
                const int2 ki = __ldcs(in.recKI + t);
                key = ki.x;
                if (ki.y < 0) continue;                       // dummy slot
                p = in.base + ki.y;
                dest = out.aosScatter ? out.aosBase + ki.y : t;
            } else {
                p = in.base + __ldg(in.perm + t);
                key = __ldg(in.cellKey + t);
            }
#pragma unroll
            for (int iv = NIND - 1; iv >= 0; --iv) {
                const int m = s.nCoef[iv] - Ord::at(iv) + 1;
                ix[iv] = Ord::at(iv) + key % m;
                key /= m;
            }
        }
        if (!recs) {
            long long rem = p;
#pragma unroll
            for (int iv = NIND - 1; iv >= 0; --iv) u[iv] = fetch_param(in, p, iv, rem);
        }
        setup_variable<0, Ord, NDT, JAC>(s, u[0], wrt.d[0], c, ix, outside, binned, in.spanRec[0]);
        if constexpr (NIND > 1) setup_variable<1, Ord, NDT, JAC>(s, u[1], wrt.d[1], c, ix, outside, binned, in.spanRec[1]);
        if constexpr (NIND > 2) setup_variable<2, Ord, NDT, JAC>(s, u[2], wrt.d[2], c, ix, outside, binned, in.spanRec[2]);
        if constexpr (NIND > 3) setup_variable<3, Ord, NDT, JAC>(s, u[3], wrt.d[3], c, ix, outside, binned, in.spanRec[3]);
        if (outside && out.firstOutside) report_outside((int64_t *)out.firstOutside, p);
        long long off = 0;
#pragma unroll
        for (int iv = 0; iv < NIND; ++iv) {
            c.stride[iv] = s.stride[iv];
            off += (long long)(ix[iv] - Ord::at(iv)) * s.stride[iv];
        }
        c.depStride = s.depStride;
        if constexpr (NDT < NDEP) {
            // one pass per tile of dependent variables, results stored as they come (no normals here)
#pragma unroll
            for (int d0 = 0; d0 < NDEP; d0 += NDT) {
                double vt[NDT];
                double gt[NIND][NDT];
                Contract<0, Ord, NDT, JAC>::run(s.coefs + off + d0 * s.depStride, c, vt, gt);
                if (out.aos) {
                    store_result_tile<NIND, NDEP, NDT, JAC>(out.aos + dest * out.aosStride, d0, vt, gt);
                } else {
                    if (out.values) {
#pragma unroll
                        for (int d = 0; d < NDT; ++d) __stcs(out.values + (d0 + d) * out.ld + p, vt[d]);
                    }
                    if constexpr (JAC) {
                        if (out.jacobian) {
#pragma unroll
                            for (int d = 0; d < NDT; ++d)
#pragma unroll
                                for (int iv = 0; iv < NIND; ++iv) __stcs(out.jacobian + ((d0 + d) * NIND + iv) * out.ld + p, gt[iv][d]);
                        }
                    }
                }
            }
            if (!binned && out.spans) {
#pragma unroll
                for (int iv = 0; iv < NIND; ++iv) __stcs(out.spans + iv * out.ld + p, ix[iv]);
            }
        } else {
        double v[NDEP];
        double g[NIND][NDEP];
        Contract<0, Ord, NDEP, JAC>::run(s.coefs + off, c, v, g);
        if (out.aos) {
            // sorted-record mode: one contiguous, sector-aligned result record per point
            store_result_record<NIND, NDEP, JAC>(s, out, out.aos + dest * out.aosStride, v, g);
            if (!binned && out.spans) {
#pragma unroll
                for (int iv = 0; iv < NIND; ++iv) __stcs(out.spans + iv * out.ld + p, ix[iv]);
            }
            continue;
        }
        if (out.values) {
#pragma unroll
            for (int d = 0; d < NDEP; ++d) __stcs(out.values + d * out.ld + p, v[d]);
        }
        if (out.spans) {
#pragma unroll
            for (int iv = 0; iv < NIND; ++iv) __stcs(out.spans + iv * out.ld + p, ix[iv]);
        }
        if constexpr (JAC) {
            if (out.jacobian) {
#pragma unroll
                for (int d = 0; d < NDEP; ++d)
#pragma unroll
                    for (int iv = 0; iv < NIND; ++iv) __stcs(out.jacobian + (d * NIND + iv) * out.ld + p, g[iv][d]);
            }
            if constexpr (NIND - NDEP == 1 || NDEP - NIND == 1) {
                if (out.normal) {
                    constexpr int D = NIND > NDEP ? NIND : NDEP;
                    double J[NDEP * NIND];
#pragma unroll
                    for (int d = 0; d < NDEP; ++d)
#pragma unroll
                        for (int iv = 0; iv < NIND; ++iv) J[d * NIND + iv] = g[iv][d];
                    double n[D];
                    normal_from_jacobian<NIND, NDEP>(J, s.normalSign, out.normalize, out.normalMask, n);
#pragma unroll
                    for (int i = 0; i < D; ++i) __stcs(out.normal + i * out.ld + p, n[i]);
                }
            }
        }
        }   // NDT == NDEP
    }
}

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, bool JAC, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_staged_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                 const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using WS = WindowShape<Ord, NDEP>;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    if (gate_closed(in)) return;
    extern __shared__ __align__(16) double stagedWindows[];         // per warp: two window images
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *w0 = stagedWindows + warp * 2 * WS::size;
    int slotKey0 = -1, slotKey1 = -1;                               // cells held by the two slots (warp-uniform)
    // contiguous run of 32-point tiles for this warp
    const long long tiles = (N + 31) >> 5, nWarps = gridDim.x * 4LL;
    const long long per = (tiles + nWarps - 1) / nWarps;
    const long long firstTile = (blockIdx.x * 4LL + warp) * per;
    const long long endTile = firstTile + per < tiles ? firstTile + per : tiles;
    double2 r0 = make_double2(0.0, 0.0), r1 = r0;
    long long k4 = -1;
    auto fetch = [&](long long tile) {
        const long long t = tile * 32 + lane;
        if (tile < endTile && t < N) {
            const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
            r0 = __ldcs(rp);
            r1 = __ldcs(rp + 1);
            if constexpr (NIND > 3) k4 = __ldcs(reinterpret_cast<const long long *>(in.recKI) + t);
        }
    };
    fetch(firstTile);
    for (long long tile = firstTile; tile < endTile; ++tile) {
        const long long t = tile * 32 + lane;
        const bool live = t < N;
        FixedCtx<Ord, NDT, JAC> c;
        int ix[NIND];
        double u[NIND];
        u[0] = r0.x;
        if constexpr (NIND > 1) u[1] = r0.y;
        if constexpr (NIND > 2) u[2] = r1.x;
        if constexpr (NIND > 3) u[3] = r1.y;
        const long long ki = NIND > 3 ? k4 : __double_as_longlong(r1.y);
        const int key = live ? (int)ki : -1;
        const long long dest = out.aosScatter ? out.aosBase + (ki >> 32) : t;
        fetch(tile + 1);                                            // next tile's records arrive under this tile's arithmetic
        {
            int k = key < 0 ? 0 : key;
#pragma unroll
            for (int iv = NIND - 1; iv >= 0; --iv) {
                const int m = s.nCoef[iv] - Ord::at(iv) + 1;
                ix[iv] = Ord::at(iv) + k % m;
                k /= m;
            }
        }
        bool outside = false;
        setup_variable<0, Ord, NDT, JAC>(s, u[0], wrt.d[0], c, ix, outside, true, in.spanRec[0]);
        if constexpr (NIND > 1) setup_variable<1, Ord, NDT, JAC>(s, u[1], wrt.d[1], c, ix, outside, true, in.spanRec[1]);
        if constexpr (NIND > 2) setup_variable<2, Ord, NDT, JAC>(s, u[2], wrt.d[2], c, ix, outside, true, in.spanRec[2]);
        if constexpr (NIND > 3) setup_variable<3, Ord, NDT, JAC>(s, u[3], wrt.d[3], c, ix, outside, true, in.spanRec[3]);
        bool done = !live;
        while (true) {
            const unsigned pending = __ballot_sync(0xffffffffu, !done);
            if (!pending) break;
            // up to two distinct cells per pass, each in the slot that already holds it or freshly staged
            const int k0 = __shfl_sync(0xffffffffu, key, __ffs(pending) - 1);
            const bool in0 = !done && key == k0;
            const unsigned rest = __ballot_sync(0xffffffffu, !done && !in0);
            const int k1 = rest ? __shfl_sync(0xffffffffu, key, __ffs(rest) - 1) : -1;
            const bool in1 = !done && !in0 && key == k1;
            int s0, s1 = -1;
            bool staged = false;
            if (k0 == slotKey0) s0 = 0;
            else if (k0 == slotKey1) s0 = 1;
            else {
                s0 = (k1 >= 0 && k1 == slotKey0) ? 1 : 0;
                stage_window<Ord, NDEP>(s, k0, w0 + s0 * WS::size, lane);
                if (s0) slotKey1 = k0; else slotKey0 = k0;
                staged = true;
            }
            if (k1 >= 0) {
                s1 = 1 - s0;
                if ((s1 ? slotKey1 : slotKey0) != k1) {
                    stage_window<Ord, NDEP>(s, k1, w0 + s1 * WS::size, lane);
                    if (s1) slotKey1 = k1; else slotKey0 = k1;
                    staged = true;
                }
            }
            if (staged) {
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            if (in0 || in1) {
                const double *w = w0 + (in1 ? s1 : s0) * WS::size;
                double *rec = out.aos + dest * out.aosStride;
                if constexpr (NDT == NDEP) {
                    double v[NDEP];
                    double g[NIND][NDEP];
                    ContractS<0, Ord, NDEP, NDEP, JAC>::run(w, 0, c, v, g);
                    store_result_record<NIND, NDEP, JAC>(s, out, rec, v, g);
                } else {
                    // one pass per tile of dependent variables; the loop stays rolled (instruction-cache footprint)
#pragma unroll 1
                    for (int d0 = 0; d0 < NDEP; d0 += NDT) {
                        double vt[NDT];
                        double gt[NIND][NDT];
                        ContractS<0, Ord, NDEP, NDT, JAC>::run(w + d0 * WS::perDepPad, 0, c, vt, gt);
                        store_result_tile<NIND, NDEP, NDT, JAC>(rec, d0, vt, gt);
                    }
                }
                done = true;
            }
            __syncwarp();                                            // slots may be overwritten by the next pass / tile
        }
    }
}

// ---- warp-staged windows, second generation: the CELL's span records travel with its window ------------------------
// All points of a cell share their knot spans, hence their per-span records (left knots | reciprocal gaps).  v1 above
// still decodes the cell key per lane (three integer divisions by run-time radices) and pulls the records through L1 with
// per-lane addresses on every tile (ncu: 1.3 long-scoreboard stalls per issue).  Here the warp decodes the key once per
// cell CHANGE and stages the records next to the window image (cp.async, 16 bytes per lane); a point is then: its
// 32-byte record (requested one tile ahead), broadcast LDS.128 of the slot's span records, the recurrence, the
// contraction over the slot's window with immediate offsets.  Same arithmetic per point: bit-identical to v1.
template <int IV, class Ord, int NDT, bool JAC>
__device__ __forceinline__ void setup_variable_shared(const double *__restrict__ cellRec, double u, int d, FixedCtx<Ord, NDT, JAC> &c)
{
    constexpr int O = Ord::at(IV);
    double b0[O], b1[O];
    basis_from_shared_record<O, JAC>(cellRec + CellRecords<Ord>::offset(IV), u, d, b0, b1);
#pragma unroll
    for (int j = 0; j < O; ++j) {
        c.B[IV][j] = b0[j];
        if constexpr (JAC) c.dB[IV][j] = b1[j];
    }
}

// the warp copies window and span records of cell `key` into a slot: [window image | records of variable 0 | 1 | ..]
template <class Ord, int NDEP>
__device__ __forceinline__ void stage_cell(const SplineDev &s, const PointsDev &in, int key, double *dst, const int lane)
{
    using WS = WindowShape<Ord, NDEP>;
    using CR = CellRecords<Ord>;
    stage_window<Ord, NDEP>(s, key, dst, lane);
    const unsigned recAddr = (unsigned)__cvta_generic_to_shared(dst + WS::size);
    int k = key;
#pragma unroll
    for (int iv = Ord::n - 1; iv >= 0; --iv) {
        const int m = s.nCoef[iv] - Ord::at(iv) + 1;
        const int span = k % m;
        k /= m;
        constexpr int dummy = 0;
        (void)dummy;
        const int st = CR::stride(iv);
        if (lane < st / 2) {
            const double *src = in.spanRec[iv] + (long long)span * st + 2 * lane;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(recAddr + (unsigned)(CR::offset(iv) + 2 * lane) * 8u), "l"(src) : "memory");
        }
    }
}

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, bool JAC, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_staged2_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                  const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using WS = WindowShape<Ord, NDEP>;
    using CR = CellRecords<Ord>;
    constexpr int SLOT = WS::size + CR::size;                       // doubles per slot (both parts even: 16-byte aligned)
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    if (gate_closed(in)) return;
    extern __shared__ __align__(16) double stagedWindows[];         // per warp: two slots
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *w0 = stagedWindows + warp * 2 * SLOT;
    int slotKey0 = -1, slotKey1 = -1;                               // cells held by the two slots (warp-uniform)
    const long long tiles = (N + 31) >> 5, nWarps = gridDim.x * 4LL;
    const long long per = (tiles + nWarps - 1) / nWarps;
    const long long firstTile = (blockIdx.x * 4LL + warp) * per;
    const long long endTile = firstTile + per < tiles ? firstTile + per : tiles;
    double2 r0 = make_double2(0.0, 0.0), r1 = r0;
    long long k4 = -1;
    auto fetch = [&](long long tile) {
        const long long t = tile * 32 + lane;
        if (tile < endTile && t < N) {
            const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
            r0 = __ldcs(rp);
            r1 = __ldcs(rp + 1);
            if constexpr (NIND > 3) k4 = __ldcs(reinterpret_cast<const long long *>(in.recKI) + t);
        }
    };
    fetch(firstTile);
    for (long long tile = firstTile; tile < endTile; ++tile) {
        const long long t = tile * 32 + lane;
        const bool live = t < N;
        double u[NIND];
        u[0] = r0.x;
        if constexpr (NIND > 1) u[1] = r0.y;
        if constexpr (NIND > 2) u[2] = r1.x;
        if constexpr (NIND > 3) u[3] = r1.y;
        const long long ki = NIND > 3 ? k4 : __double_as_longlong(r1.y);
        const int key = live ? (int)ki : -1;
        const long long dest = out.aosScatter ? out.aosBase + (ki >> 32) : t;
        fetch(tile + 1);                                            // next tile's records arrive under this tile's arithmetic
        bool done = !live;
        while (true) {
            const unsigned pending = __ballot_sync(0xffffffffu, !done);
            if (!pending) break;
            // up to two distinct cells per pass, each in the slot that already holds it or freshly staged
            const int k0 = __shfl_sync(0xffffffffu, key, __ffs(pending) - 1);
            const bool in0 = !done && key == k0;
            const unsigned rest = __ballot_sync(0xffffffffu, !done && !in0);
            const int k1 = rest ? __shfl_sync(0xffffffffu, key, __ffs(rest) - 1) : -1;
            const bool in1 = !done && !in0 && key == k1;
            int s0, s1 = -1;
            bool staged = false;
            if (k0 == slotKey0) s0 = 0;
            else if (k0 == slotKey1) s0 = 1;
            else {
                s0 = (k1 >= 0 && k1 == slotKey0) ? 1 : 0;
                stage_cell<Ord, NDEP>(s, in, k0, w0 + s0 * SLOT, lane);
                if (s0) slotKey1 = k0; else slotKey0 = k0;
                staged = true;
            }
            if (k1 >= 0) {
                s1 = 1 - s0;
                if ((s1 ? slotKey1 : slotKey0) != k1) {
                    stage_cell<Ord, NDEP>(s, in, k1, w0 + s1 * SLOT, lane);
                    if (s1) slotKey1 = k1; else slotKey0 = k1;
                    staged = true;
                }
            }
            if (staged) {
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            if (in0 || in1) {
                const double *w = w0 + (in1 ? s1 : s0) * SLOT;
                FixedCtx<Ord, NDT, JAC> c;
                setup_variable_shared<0, Ord, NDT, JAC>(w + WS::size, u[0], wrt.d[0], c);
                if constexpr (NIND > 1) setup_variable_shared<1, Ord, NDT, JAC>(w + WS::size, u[1], wrt.d[1], c);
                if constexpr (NIND > 2) setup_variable_shared<2, Ord, NDT, JAC>(w + WS::size, u[2], wrt.d[2], c);
                if constexpr (NIND > 3) setup_variable_shared<3, Ord, NDT, JAC>(w + WS::size, u[3], wrt.d[3], c);
                double *rec = out.aos + dest * out.aosStride;
                if constexpr (NDT == NDEP) {
                    double v[NDEP];
                    double g[NIND][NDEP];
                    ContractS<0, Ord, NDEP, NDEP, JAC>::run(w, 0, c, v, g);
                    store_result_record<NIND, NDEP, JAC>(s, out, rec, v, g);
                } else {
#pragma unroll 1
                    for (int d0 = 0; d0 < NDEP; d0 += NDT) {
                        double vt[NDT];
                        double gt[NIND][NDT];
                        ContractS<0, Ord, NDEP, NDT, JAC>::run(w + d0 * WS::perDepPad, 0, c, vt, gt);
                        store_result_tile<NIND, NDEP, NDT, JAC>(rec, d0, vt, gt);
                    }
                }
                done = true;
            }
            __syncwarp();                                            // slots may be overwritten by the next pass / tile
        }
    }
}

// ---- warp-staged windows, two points of the same cell per lane ---------------------------------------------------------
// The load-return path of the SM (128 B/clk) has to deliver every window coefficient to every lane: 1.5 KB per point of a
// tricubic nDep-3 spline, as many cycles as the FP64 pipe needs for the point (ncu: l1tex 68 %, FP64 57 % -- neither full,
// the warps wait on both).  Two points per lane use every LDS.128 twice and carry twice the independent FMA chains.  The
// sort pads every cell's segment to an even length (bin_scan_kernel, evenPad): the aligned pairs of the sorted sequence
// never straddle a cell, spare slots hold dummy records (index -1: evaluated, never stored).  A tile is 64 sorted slots,
// lane l takes slots 2l and 2l+1; the two-slot window logic of eval_staged2_kernel applies per pair.  One dependent variable
// per pass (NDT), results collected in a [slot][lane] tile of shared memory and stored as whole 32-byte sectors.
template <int L, class Ord, int NDEP, int NDT>
struct ContractS2 {
    using WS = WindowShape<Ord, NDEP>;
    using Ctx = FixedCtx<Ord, NDT, true>;
    __device__ __forceinline__ static void run(const double *__restrict__ w, const int off, const Ctx &c0, const Ctx &c1,
                                               double (&v0)[NDT], double (&g0)[Ord::n][NDT], double (&v1)[NDT], double (&g1)[Ord::n][NDT])
    {
        constexpr int O = Ord::at(L);
#pragma unroll
        for (int d = 0; d < NDT; ++d) { v0[d] = 0.0; v1[d] = 0.0; }
#pragma unroll
        for (int m = L; m < Ord::n; ++m)
#pragma unroll
            for (int d = 0; d < NDT; ++d) { g0[m][d] = 0.0; g1[m][d] = 0.0; }
        if constexpr (L == Ord::n - 1) {
#pragma unroll
            for (int d = 0; d < NDT; ++d) {
                double x[O];
                load_run<O>(w, off + d * WS::perDepPad, x);
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    v0[d] = fma(x[i], c0.B[L][i], v0[d]);
                    v1[d] = fma(x[i], c1.B[L][i], v1[d]);
                    g0[L][d] = fma(x[i], c0.dB[L][i], g0[L][d]);
                    g1[L][d] = fma(x[i], c1.dB[L][i], g1[L][d]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < O; ++i) {
                double cv0[NDT], cv1[NDT];
                double cg0[Ord::n][NDT], cg1[Ord::n][NDT];
                ContractS2<L + 1, Ord, NDEP, NDT>::run(w, off + i * WS::stride(L), c0, c1, cv0, cg0, cv1, cg1);
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    v0[d] = fma(cv0[d], c0.B[L][i], v0[d]);
                    v1[d] = fma(cv1[d], c1.B[L][i], v1[d]);
                    g0[L][d] = fma(cv0[d], c0.dB[L][i], g0[L][d]);
                    g1[L][d] = fma(cv1[d], c1.dB[L][i], g1[L][d]);
#pragma unroll
                    for (int m = L + 1; m < Ord::n; ++m) {
                        g0[m][d] = fma(cg0[m][d], c0.B[L][i], g0[m][d]);
                        g1[m][d] = fma(cg1[m][d], c1.B[L][i], g1[m][d]);
                    }
                }
            }
        }
    }
};

template <int IV, class Ord, int NDT>
__device__ __forceinline__ void setup_variable_shared_pair(const double *__restrict__ cellRec, double ua, double ub,
                                                           FixedCtx<Ord, NDT, true> &ca, FixedCtx<Ord, NDT, true> &cb)
{
    constexpr int O = Ord::at(IV);
    using R = SpanRec<O>;
    const double *rec = cellRec + CellRecords<Ord>::offset(IV);
    double r[R::stride > 0 ? R::stride : 1];
#pragma unroll
    for (int j = 0; j < R::stride / 2; ++j) {
        const double2 x = *reinterpret_cast<const double2 *>(rec + 2 * j);
        r[2 * j] = x.x;
        r[2 * j + 1] = x.y;
    }
    double rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
    for (int j = 0; j < O * (O - 1) / 2; ++j) rc[j] = r[O - 1 + j];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const double u = p ? ub : ua;
        double dl[O > 1 ? O - 1 : 1], b0[O], b1[O];
#pragma unroll
        for (int j = 0; j < O - 1; ++j) dl[j] = u - r[j];
        basis_core<O, true>(dl, rc, 0, b0, b1);
        FixedCtx<Ord, NDT, true> &c = p ? cb : ca;
#pragma unroll
        for (int j = 0; j < O; ++j) { c.B[IV][j] = b0[j]; c.dB[IV][j] = b1[j]; }
    }
}

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_staged_pair_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                      const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using WS = WindowShape<Ord, NDEP>;
    using CR = CellRecords<Ord>;
    constexpr int SLOT = WS::size + CR::size;
    constexpr int R = NDEP * (1 + NIND), RP = (R + 3) & ~3;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    extern __shared__ __align__(16) double stagedWindows[];         // per warp: two slots | result tile [2][R][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *w0 = stagedWindows + warp * (2 * SLOT + 2 * R * 32);
    double *mine = w0 + 2 * SLOT + lane;                            // results of this lane: mine[(p * R + slot) * 32]
    int slotKey0 = -1, slotKey1 = -1;
    const long long total = __ldg(in.sortedTotal);                  // slots of the sorted sequence (segments padded to even)
    const long long tiles = (total + 63) >> 6, nWarps = gridDim.x * 4LL;
    const long long per = (tiles + nWarps - 1) / nWarps;
    const long long firstTile = (blockIdx.x * 4LL + warp) * per;
    const long long endTile = firstTile + per < tiles ? firstTile + per : tiles;
    double2 ra0 = make_double2(0.0, 0.0), ra1 = ra0, rb0 = ra0, rb1 = ra0;
    longlong2 kk = make_longlong2(-1, -1);
    auto fetch = [&](long long tile) {
        const long long t = tile * 64 + 2 * lane;
        if (tile < endTile && t < total) {
            const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
            ra0 = __ldcs(rp); ra1 = __ldcs(rp + 1); rb0 = __ldcs(rp + 2); rb1 = __ldcs(rp + 3);
            if constexpr (NIND > 3) kk = __ldcs(reinterpret_cast<const longlong2 *>(in.recKI + t));
        }
    };
    fetch(firstTile);
    for (long long tile = firstTile; tile < endTile; ++tile) {
        const long long t = tile * 64 + 2 * lane;
        const bool live = t < total;
        double ua[NIND], ub[NIND];
        ua[0] = ra0.x; ub[0] = rb0.x;
        if constexpr (NIND > 1) { ua[1] = ra0.y; ub[1] = rb0.y; }
        if constexpr (NIND > 2) { ua[2] = ra1.x; ub[2] = rb1.x; }
        if constexpr (NIND > 3) { ua[3] = ra1.y; ub[3] = rb1.y; }
        const long long kia = NIND > 3 ? kk.x : __double_as_longlong(ra1.y), kib = NIND > 3 ? kk.y : __double_as_longlong(rb1.y);
        const int key = live ? (int)kia : -1;
        fetch(tile + 1);
        bool done = !live;
        while (true) {
            const unsigned pending = __ballot_sync(0xffffffffu, !done);
            if (!pending) break;
            const int k0 = __shfl_sync(0xffffffffu, key, __ffs(pending) - 1);
            const bool in0 = !done && key == k0;
            const unsigned rest = __ballot_sync(0xffffffffu, !done && !in0);
            const int k1 = rest ? __shfl_sync(0xffffffffu, key, __ffs(rest) - 1) : -1;
            const bool in1 = !done && !in0 && key == k1;
            int s0, s1 = -1;
            bool staged = false;
            if (k0 == slotKey0) s0 = 0;
            else if (k0 == slotKey1) s0 = 1;
            else {
                s0 = (k1 >= 0 && k1 == slotKey0) ? 1 : 0;
                stage_cell<Ord, NDEP>(s, in, k0, w0 + s0 * SLOT, lane);
                if (s0) slotKey1 = k0; else slotKey0 = k0;
                staged = true;
            }
            if (k1 >= 0) {
                s1 = 1 - s0;
                if ((s1 ? slotKey1 : slotKey0) != k1) {
                    stage_cell<Ord, NDEP>(s, in, k1, w0 + s1 * SLOT, lane);
                    if (s1) slotKey1 = k1; else slotKey0 = k1;
                    staged = true;
                }
            }
            if (staged) {
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            if (in0 || in1) {
                const double *w = w0 + (in1 ? s1 : s0) * SLOT;
                FixedCtx<Ord, NDT, true> c0, c1;
                setup_variable_shared_pair<0, Ord, NDT>(w + WS::size, ua[0], ub[0], c0, c1);
                if constexpr (NIND > 1) setup_variable_shared_pair<1, Ord, NDT>(w + WS::size, ua[1], ub[1], c0, c1);
                if constexpr (NIND > 2) setup_variable_shared_pair<2, Ord, NDT>(w + WS::size, ua[2], ub[2], c0, c1);
                if constexpr (NIND > 3) setup_variable_shared_pair<3, Ord, NDT>(w + WS::size, ua[3], ub[3], c0, c1);
#pragma unroll 1
                for (int d0 = 0; d0 < NDEP; d0 += NDT) {
                    double v0[NDT], v1[NDT];
                    double g0[NIND][NDT], g1[NIND][NDT];
                    ContractS2<0, Ord, NDEP, NDT>::run(w + d0 * WS::perDepPad, 0, c0, c1, v0, g0, v1, g1);
#pragma unroll
                    for (int d = 0; d < NDT; ++d) {
                        mine[(d0 + d) * 32] = v0[d];
                        mine[(R + d0 + d) * 32] = v1[d];
#pragma unroll
                        for (int iv = 0; iv < NIND; ++iv) {
                            mine[(NDEP + (d0 + d) * NIND + iv) * 32] = g0[iv][d];
                            mine[(R + NDEP + (d0 + d) * NIND + iv) * 32] = g1[iv][d];
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const long long ki = p ? kib : kia;
                    const long long idx = ki >> 32;
                    if (idx < 0) continue;                          // dummy slot of an odd cell
                    const long long dest = out.aosScatter ? out.aosBase + idx : t + p;
                    double *rec = out.aos + dest * out.aosStride;
                    const double *src = mine + p * R * 32;
#pragma unroll
                    for (int j = 0; j < RP / 4; ++j) {
                        if (4 * j < out.aosStride) {
                            double x[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) x[e] = 4 * j + e < R ? src[(4 * j + e) * 32] : 0.0;
                            if (out.aosWide)
                                asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(rec + 4 * j), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
                            else {
                                __stcs(reinterpret_cast<double2 *>(rec + 4 * j), make_double2(x[0], x[1]));
                                __stcs(reinterpret_cast<double2 *>(rec + 4 * j) + 1, make_double2(x[2], x[3]));
                            }
                        }
                    }
                }
                done = true;
            }
            __syncwarp();
        }
    }
}

// ---- any-shape kernel -------------------------------------------------------------------------
// Runtime nInd / orders / nDep.  Per-thread basis rows live in shared memory ([slot][thread]);
// the window is walked with an odometer over all variables but the last, the last variable is
// contracted in the inner loop; one pass per dependent variable.
__device__ double det_lu_runtime(double *a, int n)
{
    double det = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = fabs(a[c * n + c]);
        for (int r = c + 1; r < n; ++r)
            if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); piv = r; }
        if (piv != c) {
            for (int k = 0; k < n; ++k) { const double t = a[c * n + k]; a[c * n + k] = a[piv * n + k]; a[piv * n + k] = t; }
            det = -det;
        }
        const double pv = a[c * n + c];
        if (pv == 0.0) return 0.0;
        for (int r = c + 1; r < n; ++r) {
            const double l = a[r * n + c] / pv;
            for (int k = c + 1; k < n; ++k) a[r * n + k] -= l * a[c * n + k];
        }
        det *= pv;
    }
    return det;
}

struct SmemCol {
    double *base;
    int stride;
    __device__ __forceinline__ double &operator()(int j) const { return base[j * stride]; }
};

// Relaxed (FMA allowed) runtime-order recurrence into shared-memory columns.
__device__ __forceinline__ void basis_runtime(const double *__restrict__ knots, int order, int ix, double u, int deriv,
                                              SmemCol b)
{
    for (int j = 0; j < order; ++j) b(j) = 0.0;
    if (deriv >= order) return;
    b(order - 1) = 1.0;
    const int nValue = order - deriv;
    for (int deg = 1; deg < order; ++deg) {
        int slot = order - deg;
        for (int i = ix - deg; i < ix; ++i, ++slot) {
            const double ki = __ldg(knots + i);
            const double gap = __ldg(knots + i + deg) - ki;
            if (deg < nValue) {
                const double a = (u - ki) / gap;
                b(slot - 1) += (1.0 - a) * b(slot);
                b(slot) *= a;
            } else {
                const double a = (double)deg / gap;
                b(slot - 1) -= a * b(slot);
                b(slot) *= a;
            }
        }
    }
}

__global__ void __launch_bounds__(128) eval_generic_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                           const WrtDev wrt, const OutDev out, const int jac,
                                                           const int rowLen /* sum of orders */)
{
    extern __shared__ double scratch[];
    const int T = blockDim.x;
    // column layout: [B rows (rowLen)] [dB rows (rowLen) if jac], each slot strided by T
    double *colB = scratch + threadIdx.x;
    double *colD = colB + (long long)rowLen * T;
    const int nInd = s.nInd, nDep = s.nDep;
    const int D = nInd > nDep ? nInd : nDep;
    for (long long p = blockIdx.x * (long long)T + threadIdx.x; p < N; p += (long long)gridDim.x * T) {
        int ix[BSPY_MAX_IND];
        int rowAt[BSPY_MAX_IND];
        bool outside = false;
        long long rem = p;
        double uu[BSPY_MAX_IND];
        for (int iv = nInd - 1; iv >= 0; --iv) uu[iv] = fetch_param(in, p, iv, rem);
        int at = 0;
        long long off = 0;
        for (int iv = 0; iv < nInd; ++iv) {
            const double *k = s.knots[iv];
            const int o = s.order[iv];
            const double u = uu[iv];
            outside |= (u < __ldg(k + o - 1)) | (u > __ldg(k + s.nCoef[iv]));
            ix[iv] = span_search(k, o + s.nCoef[iv], o, u);
            rowAt[iv] = at;
            basis_runtime(k, o, ix[iv], u, jac ? 0 : wrt.d[iv], SmemCol{colB + (long long)at * T, T});
            if (jac) basis_runtime(k, o, ix[iv], u, 1, SmemCol{colD + (long long)at * T, T});
            at += o;
            off += (long long)(ix[iv] - o) * s.stride[iv];
            if (out.spans) out.spans[iv * out.ld + p] = ix[iv];
        }
        if (outside && out.firstOutside) report_outside((int64_t *)out.firstOutside, p);
        double J[(BSPY_MAX_IND + 1) * BSPY_MAX_IND];  // (nDep, nInd) only when a normal is requested (nDep <= nInd+1)
        const int last = nInd - 1;
        for (int d = 0; d < nDep; ++d) {
            const double *cd = s.coefs + d * s.depStride + off;
            double val = 0.0;
            double der[BSPY_MAX_IND];
            for (int iv = 0; iv < nInd; ++iv) der[iv] = 0.0;
            if (nInd == 0) {
                val = __ldg(cd);
            } else {
                int idx[BSPY_MAX_IND];
                for (int iv = 0; iv < nInd; ++iv) idx[iv] = 0;
                bool more = true;
                while (more) {
                    long long o2 = 0;
                    for (int iv = 0; iv < last; ++iv) o2 += idx[iv] * s.stride[iv];
                    double s0 = 0.0, s1 = 0.0;
                    const int ol = s.order[last];
                    for (int k = 0; k < ol; ++k) {
                        const double x = __ldg(cd + o2 + k);
                        s0 = fma(x, colB[(long long)(rowAt[last] + k) * T], s0);
                        if (jac) s1 = fma(x, colD[(long long)(rowAt[last] + k) * T], s1);
                    }
                    double w = 1.0;
                    for (int iv = 0; iv < last; ++iv) w *= colB[(long long)(rowAt[iv] + idx[iv]) * T];
                    val = fma(w, s0, val);
                    if (jac) {
                        der[last] = fma(w, s1, der[last]);
                        for (int m = 0; m < last; ++m) {
                            double wm = 1.0;
                            for (int iv = 0; iv < last; ++iv)
                                wm *= (iv == m ? colD : colB)[(long long)(rowAt[iv] + idx[iv]) * T];
                            der[m] = fma(wm, s0, der[m]);
                        }
                    }
                    // odometer over variables 0 .. last-1 (last-1 fastest)
                    int iv = last - 1;
                    for (; iv >= 0; --iv) {
                        if (++idx[iv] < s.order[iv]) break;
                        idx[iv] = 0;
                    }
                    more = iv >= 0;
                }
            }
            if (out.aos) {                                  // array-of-structs record [values | jacobian (d, iv) | normal]
                double *rec = out.aos + p * out.aosStride;
                rec[d] = val;
                if (jac)
                    for (int iv = 0; iv < nInd; ++iv) rec[nDep + d * nInd + iv] = der[iv];
            }
            if (out.values) out.values[d * out.ld + p] = val;
            if (jac) {
                if (out.jacobian)
                    for (int iv = 0; iv < nInd; ++iv) out.jacobian[((long long)d * nInd + iv) * out.ld + p] = der[iv];
                if (out.normal || out.aosNormal)
                    for (int iv = 0; iv < nInd; ++iv) J[d * nInd + iv] = der[iv];
            }
        }
        if (jac && (out.normal || out.aosNormal)) {
            double minor[BSPY_MAX_IND * BSPY_MAX_IND];
            double n[BSPY_MAX_IND + 1];
            const int M = D - 1;
            double sq = 0.0;
            for (int i = 0; i < D; ++i) {
                int rr = 0;
                for (int r = 0; r < D; ++r) {
                    if (r == i) continue;
                    for (int c = 0; c < M; ++c) minor[rr * M + c] = (nInd > nDep) ? J[c * nInd + r] : J[r * nInd + c];
                    ++rr;
                }
                const double det = det_lu_runtime(minor, M);
                n[i] = ((i & 1) ? -det : det) * (double)s.normalSign;
                if (out.normalMask & (1u << i)) sq += n[i] * n[i];
            }
            const double len = sqrt(sq);
            if (out.aos)
                for (int i = 0; i < D; ++i) out.aos[p * out.aosStride + nDep + nDep * nInd + i] = out.normalize ? n[i] / len : n[i];
            else
                for (int i = 0; i < D; ++i) out.normal[(long long)i * out.ld + p] = out.normalize ? n[i] / len : n[i];
        }
    }
}

// ---- host dispatch ----------------------------------------------------------------------------
int launch_curve(const SplineDev &s, const PointsDev &in, long long N, const WrtDev &wrt, const OutDev &out, int jac,
                 cudaStream_t stream);

struct FixedEntry {
    int nInd, o[4], nDep, jac;
    FixedFn fn;
};

#define BSPY_FIXED(NI, A, B, C, D_, ND)                                        \
    {NI, {A, B, C, D_}, ND, 0, eval_fixed_kernel<NI, A, B, C, D_, ND, false>}, \
    {NI, {A, B, C, D_}, ND, 1, eval_fixed_kernel<NI, A, B, C, D_, ND, true>}

static const FixedEntry kFixed[] = {
    // curves
    BSPY_FIXED(1, 2, 0, 0, 0, 1), BSPY_FIXED(1, 2, 0, 0, 0, 2), BSPY_FIXED(1, 2, 0, 0, 0, 3),
    BSPY_FIXED(1, 3, 0, 0, 0, 1), BSPY_FIXED(1, 3, 0, 0, 0, 2), BSPY_FIXED(1, 3, 0, 0, 0, 3),
    BSPY_FIXED(1, 4, 0, 0, 0, 1), BSPY_FIXED(1, 4, 0, 0, 0, 2), BSPY_FIXED(1, 4, 0, 0, 0, 3),
    BSPY_FIXED(1, 5, 0, 0, 0, 1), BSPY_FIXED(1, 5, 0, 0, 0, 2), BSPY_FIXED(1, 5, 0, 0, 0, 3),
    BSPY_FIXED(1, 6, 0, 0, 0, 2), BSPY_FIXED(1, 6, 0, 0, 0, 3),
    // surfaces
    BSPY_FIXED(2, 2, 2, 0, 0, 1), BSPY_FIXED(2, 2, 2, 0, 0, 3),
    BSPY_FIXED(2, 3, 3, 0, 0, 1), BSPY_FIXED(2, 3, 3, 0, 0, 2), BSPY_FIXED(2, 3, 3, 0, 0, 3),
    BSPY_FIXED(2, 4, 4, 0, 0, 1), BSPY_FIXED(2, 4, 4, 0, 0, 2), BSPY_FIXED(2, 4, 4, 0, 0, 3),
    BSPY_FIXED(2, 3, 4, 0, 0, 3), BSPY_FIXED(2, 4, 3, 0, 0, 3), BSPY_FIXED(2, 4, 5, 0, 0, 3),
    BSPY_FIXED(2, 5, 5, 0, 0, 3),
    // volumes
    BSPY_FIXED(3, 2, 2, 2, 0, 3), BSPY_FIXED(3, 3, 3, 3, 0, 1), BSPY_FIXED(3, 3, 3, 3, 0, 3),
    BSPY_FIXED(3, 4, 4, 4, 0, 1), BSPY_FIXED(3, 4, 4, 4, 0, 3), BSPY_FIXED(3, 4, 4, 4, 0, 4),
    // 4-variate manifolds
    BSPY_FIXED(4, 2, 2, 2, 2, 3), BSPY_FIXED(4, 3, 3, 3, 3, 3), BSPY_FIXED(4, 3, 3, 3, 3, 5),
    BSPY_FIXED(4, 3, 3, 3, 3, 6),
};

// dependent-variable tiles for the two big-window shapes of the north-star configs (value + jacobian, no normals)
FixedFn find_fixed_tiled(const SplineDev &s, int jac, int code)
{
    // code = 10 * (dependent variables per pass) + (CTAs per SM the variant is compiled for).  Measured on the
    // 4-variate nDep-6 manifold (config 5), sorted records: 14 -> 2.52 Gpts/s, 24 -> 2.42, one pass (168 registers,
    // spills) -> 2.32, 33/34 -> 1.9; on the tricubic nDep-3 volume the single pass wins (5.8 vs 5.1-5.5).
    if (!jac) return nullptr;
    if (s.nInd == 4 && s.nDep == 6 && s.order[0] == 3 && s.order[1] == 3 && s.order[2] == 3 && s.order[3] == 3) {
        switch (code) {
            case 14: return eval_fixed_kernel<4, 3, 3, 3, 3, 6, true, 1, 4>;
            case 24: return eval_fixed_kernel<4, 3, 3, 3, 3, 6, true, 2, 4>;
        }
    }
    return nullptr;
}

#define BSPY_STAGED_(NI, A, B, C, D_, ND, J, NDT, MB, OPT)                                                        \
    {NI, {A, B, C, D_}, ND, J, 10 * NDT + MB, eval_staged_kernel<NI, A, B, C, D_, ND, J != 0, NDT, MB>,       \
     WindowShape<Orders<NI, A, B, C, D_>, ND>::size,                                                          \
     eval_staged2_kernel<NI, A, B, C, D_, ND, J != 0, NDT, MB>,                                               \
     WindowShape<Orders<NI, A, B, C, D_>, ND>::size + CellRecords<Orders<NI, A, B, C, D_>>::size, OPT}
#define BSPY_STAGED(NI, A, B, C, D_, ND, J, NDT, MB) BSPY_STAGED_(NI, A, B, C, D_, ND, J, NDT, MB, 0)
static const StagedEntry kStaged[] = {
    // volumes (measured: tricubic nDep 3, value + jacobian, 379 -> 326 us per 4 Mi points, FP64 pipe 37 -> 50 %);
    // the 4-variate nDep-6 window gained nothing here (register spills once the pass over the dependent variables is
    // rolled) and stays on eval_fixed_kernel with one dependent variable per pass
    BSPY_STAGED(3, 3, 3, 3, 0, 3, 0, 3, 4), BSPY_STAGED(3, 3, 3, 3, 0, 3, 1, 3, 4),
    BSPY_STAGED(3, 4, 4, 4, 0, 1, 0, 1, 4), BSPY_STAGED(3, 4, 4, 4, 0, 1, 1, 1, 4),
    BSPY_STAGED(3, 4, 4, 4, 0, 3, 0, 3, 4), BSPY_STAGED(3, 4, 4, 4, 0, 3, 1, 3, 4),
    BSPY_STAGED(3, 4, 4, 4, 0, 4, 0, 4, 4), BSPY_STAGED(3, 4, 4, 4, 0, 4, 1, 4, 4),
    // variants selected with STAGED=<10 * deps per pass + CTAs per SM> only
    BSPY_STAGED_(3, 4, 4, 4, 0, 3, 1, 3, 3, 1),
    BSPY_STAGED_(4, 3, 3, 3, 3, 6, 1, 1, 4, 1), BSPY_STAGED_(4, 3, 3, 3, 3, 6, 1, 2, 3, 1), BSPY_STAGED_(4, 3, 3, 3, 3, 6, 1, 6, 2, 1),
};

// two points per lane (value + jacobian).  code = 10 * (dependent variables per pass) + CTAs per SM; windowDoubles holds the
// per-warp shared memory in doubles: two slots (window + cell records) and the result tile of the lane pairs
#define BSPY_STAGED_PAIR(NI, A, B, C, D_, ND, NDT, MB, OPT)                                                     \
    {NI, {A, B, C, D_}, ND, 1, 10 * NDT + MB, eval_staged_pair_kernel<NI, A, B, C, D_, ND, NDT, MB>,           \
     2 * (WindowShape<Orders<NI, A, B, C, D_>, ND>::size + CellRecords<Orders<NI, A, B, C, D_>>::size) + 2 * ND * (1 + NI) * 32, nullptr, 0, OPT}
static const StagedEntry kStagedPair[] = {
    BSPY_STAGED_PAIR(3, 4, 4, 4, 0, 3, 1, 3, 0), BSPY_STAGED_PAIR(3, 4, 4, 4, 0, 3, 1, 2, 1), BSPY_STAGED_PAIR(3, 4, 4, 4, 0, 3, 3, 2, 1),
    BSPY_STAGED_PAIR(3, 4, 4, 4, 0, 3, 1, 4, 1),
};

const StagedEntry *find_staged_pair(const SplineDev &s, int code)
{
    for (const StagedEntry &e : kStagedPair) {
        if (e.nInd != s.nInd || e.nDep != s.nDep) continue;
        bool same = true;
        for (int i = 0; i < s.nInd; ++i) same &= e.o[i] == s.order[i];
        if (same && (code == 0 ? !e.optIn : code == e.code)) return &e;
    }
    return nullptr;
}

const StagedEntry *find_staged(const SplineDev &s, int jac, int code)
{
    for (const StagedEntry &e : kStaged) {
        if (e.nInd != s.nInd || e.nDep != s.nDep || e.jac != jac) continue;
        bool same = true;
        for (int i = 0; i < s.nInd; ++i) same &= e.o[i] == s.order[i];
        if (same && (code == 0 ? !e.optIn : code == e.code)) return &e;
    }
    return nullptr;
}

FixedFn find_fixed(const SplineDev &s, int jac)
{
    if (s.nInd < 1 || s.nInd > 4) return nullptr;
    for (const FixedEntry &e : kFixed) {
        if (e.nInd != s.nInd || e.nDep != s.nDep || e.jac != jac) continue;
        bool same = true;
        for (int i = 0; i < s.nInd; ++i) same &= e.o[i] == s.order[i];
        if (same) return e.fn;
    }
    return nullptr;
}

int make_spline_dev(const bspy_spline *sp, SplineDev &s, const char *who)
{
    if (!sp) { set_error("%s: spline is NULL", who); return BSPY_E_ARG; }
    if (sp->nInd < 0 || sp->nDep < 0 || !sp->coefs) { set_error("%s: bad spline header", who); return BSPY_E_ARG; }
    if (sp->nInd > BSPY_MAX_IND) { set_error("%s: nInd %d > BSPY_MAX_IND", who, sp->nInd); return BSPY_E_UNSUPPORTED; }
    s.nInd = sp->nInd;
    s.nDep = sp->nDep;
    s.coefs = sp->coefs;
    s.normalSign = sp->normalSign < 0 ? -1 : 1;
    s.curveTable = sp->curveTable;
    s.curveTableBytes = sp->curveTableBytes;
    long long stride = 1;
    for (int i = sp->nInd - 1; i >= 0; --i) {
        if (sp->order[i] < 1 || sp->nCoef[i] < sp->order[i] || !sp->knots[i]) {
            set_error("%s: bad order/nCoef/knots for variable %d", who, i);
            return BSPY_E_ARG;
        }
        if (sp->order[i] > BSPY_MAX_ORDER) { set_error("%s: order %d > BSPY_MAX_ORDER", who, sp->order[i]); return BSPY_E_UNSUPPORTED; }
        s.stride[i] = stride;
        stride *= sp->nCoef[i];
    }
    s.depStride = stride;
    for (int i = 0; i < BSPY_MAX_IND; ++i) {
        s.order[i] = i < sp->nInd ? sp->order[i] : 0;
        s.nCoef[i] = i < sp->nInd ? sp->nCoef[i] : 0;
        s.knots[i] = i < sp->nInd ? sp->knots[i] : nullptr;
        if (i >= sp->nInd) s.stride[i] = 0;
    }
    return 0;
}

// Launch one pass (values-or-wrt pass when jac == 0, value+jacobian(+normal) pass when jac == 1).
int launch_eval(const SplineDev &s, const PointsDev &in, long long N, const WrtDev &wrt, const OutDev &out, int jac,
                cudaStream_t stream)
{
    if (N <= 0) return 0;
    if (!in.perm && !out.aos) {
        const int rc = launch_curve(s, in, N, wrt, out, jac, stream);   // lean single-curve path (curve.cu)
        if (rc != -1000) return rc;
    }
    FixedFn fn = find_fixed(s, jac);
    const int threads = 128;
    long long blocks = (N + threads - 1) / threads;
    if (fn) {
        const long long cap = (long long)num_sms() * 32;
        if (blocks > cap) blocks = cap;
        fn<<<(unsigned)blocks, threads, 0, stream>>>(s, in, N, wrt, out);
    } else {
        int rowLen = 0;
        for (int i = 0; i < s.nInd; ++i) rowLen += s.order[i];
        int t = threads;
        size_t smem = (size_t)rowLen * (jac ? 2 : 1) * t * sizeof(double);
        while (smem > 160 * 1024 && t > 32) { t >>= 1; smem >>= 1; }
        if (smem > 200 * 1024) { set_error("eval_generic: basis scratch too large"); return BSPY_E_UNSUPPORTED; }
        if (int rc = allow_dynamic_smem(eval_generic_kernel, smem)) return rc;
        blocks = (N + t - 1) / t;
        const long long cap = (long long)num_sms() * 16;
        if (blocks > cap) blocks = cap;
        eval_generic_kernel<<<(unsigned)blocks, t, smem, stream>>>(s, in, N, wrt, out, jac, rowLen);
    }
    count_launch();
    return check_launch("bspy_cuda_eval");
}

int eval_common(const bspy_spline *spline, const PointsDev &in, long long N, const int32_t *wrt, uint32_t flags,
                uint32_t normalMask, double *values, double *deriv, double *jacobian, double *normal, int32_t *spans,
                int64_t *firstOutside, void *stream, const char *who)
{
    SplineDev s;
    int rc = make_spline_dev(spline, s, who);
    if (rc) return rc;
    if (N < 0) { set_error("%s: N < 0", who); return BSPY_E_ARG; }
    if ((deriv != nullptr) != (wrt != nullptr)) { set_error("%s: wrt and deriv must be given together", who); return BSPY_E_ARG; }
    const int D = s.nInd > s.nDep ? s.nInd : s.nDep;
    if (normal && (s.nInd - s.nDep != 1 && s.nDep - s.nInd != 1)) {
        set_error("The number of independent variables must be one different than the number of dependent variables.");
        return BSPY_E_NORMAL_DIMS;
    }
    if (normalMask == 0 || D >= 32) normalMask = 0xffffffffu;
    WrtDev zero{};
    OutDev out{};
    out.ld = N;
    out.firstOutside = (long long *)firstOutside;
    out.normalize = (flags & BSPY_NORMALIZE) ? 1u : 0u;
    out.normalMask = normalMask;
    bool spansDone = false, oobDone = false;
    if (jacobian || normal) {
        out.values = values;
        out.jacobian = jacobian;
        out.normal = normal;
        out.spans = spans;
        rc = launch_eval(s, in, N, zero, out, 1, (cudaStream_t)stream);
        if (rc) return rc;
        spansDone = oobDone = true;
    } else if (values || spans || (firstOutside && !deriv)) {
        out.values = values;
        out.spans = spans;
        rc = launch_eval(s, in, N, zero, out, 0, (cudaStream_t)stream);
        if (rc) return rc;
        spansDone = oobDone = true;
    }
    if (deriv) {
        WrtDev w{};
        for (int i = 0; i < s.nInd; ++i) {
            if (wrt[i] < 0) { set_error("%s: negative derivative order", who); return BSPY_E_ARG; }
            w.d[i] = wrt[i];
        }
        OutDev o2{};
        o2.ld = N;
        o2.values = deriv;
        o2.spans = spansDone ? nullptr : spans;
        o2.firstOutside = oobDone ? nullptr : (long long *)firstOutside;
        rc = launch_eval(s, in, N, w, o2, 0, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace bspy

using namespace bspy;

extern "C" int bspy_cuda_eval_points(const bspy_spline *spline, const double *uvw, int64_t pointStride, int64_t varStride,
                                     int64_t N, const int32_t *wrt, uint32_t flags, uint32_t normalMask, double *values,
                                     double *deriv, double *jacobian, double *normal, int32_t *spans,
                                     int64_t *firstOutside, void *stream)
{
    if (!uvw && N > 0 && spline && spline->nInd > 0) {
        set_error("bspy_cuda_eval_points: uvw is NULL");
        return BSPY_E_ARG;
    }
    PointsDev in{};
    in.uvw = uvw;
    in.pointStride = pointStride;
    in.varStride = varStride;
    in.grid = 0;
    return eval_common(spline, in, N, wrt, flags, normalMask, values, deriv, jacobian, normal, spans, firstOutside, stream,
                       "bspy_cuda_eval_points");
}
