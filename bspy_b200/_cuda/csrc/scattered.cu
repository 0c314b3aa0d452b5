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

#include <mutex>

#include "common.cuh"

namespace bspy {

// ---- cofactor normals -------------------------------------------------------------------------
// T is D x (D-1) (row r = dependent/independent index r of the larger dimension), n[i] =
// sign * (-1)^i * det(T without row i).  Closed forms for D <= 4, LU with partial pivoting above.
template <int D>
__device__ __forceinline__ double det_small(const double (&m)[(D > 0 ? D : 1) * (D > 0 ? D : 1)])
{
    if constexpr (D == 0) {
        return 1.0;
    } else if constexpr (D == 1) {
        return m[0];
    } else if constexpr (D == 2) {
        return m[0] * m[3] - m[1] * m[2];
    } else if constexpr (D == 3) {
        return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    } else {
        double a[D * D];
#pragma unroll
        for (int i = 0; i < D * D; ++i) a[i] = m[i];
        double det = 1.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            int piv = c;
            double best = fabs(a[c * D + c]);
#pragma unroll
            for (int r = c + 1; r < D; ++r) {
                const double v = fabs(a[r * D + c]);
                if (v > best) { best = v; piv = r; }
            }
            if (piv != c) {
#pragma unroll
                for (int r = c + 1; r < D; ++r)
                    if (r == piv) {
#pragma unroll
                        for (int k = 0; k < D; ++k) { const double t = a[c * D + k]; a[c * D + k] = a[r * D + k]; a[r * D + k] = t; }
                    }
                det = -det;
            }
            const double pv = a[c * D + c];
            if (pv == 0.0) return 0.0;
#pragma unroll
            for (int r = c + 1; r < D; ++r) {
                const double l = a[r * D + c] / pv;
#pragma unroll
                for (int k = c + 1; k < D; ++k) a[r * D + k] -= l * a[c * D + k];
            }
            det *= pv;
        }
        return det;
    }
}

// J is (NDEP, NIND) row-major in registers; writes D = max(NIND,NDEP) components.
template <int NIND, int NDEP>
__device__ __forceinline__ void normal_from_jacobian(const double (&J)[NDEP * NIND], int sign, unsigned normalize,
                                                      unsigned mask, double (&n)[(NIND > NDEP ? NIND : NDEP)])
{
    constexpr int D = NIND > NDEP ? NIND : NDEP;
    constexpr int M = D - 1;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double minor[(M > 0 ? M : 1) * (M > 0 ? M : 1)];
        int rr = 0;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            if (r == i) continue;
#pragma unroll
            for (int c = 0; c < M; ++c) minor[rr * M + c] = (NIND > NDEP) ? J[c * NIND + r] : J[r * NIND + c];
            ++rr;
        }
        const double det = det_small<M>(minor);
        n[i] = ((i & 1) ? -det : det) * (double)sign;
    }
    if (normalize) {
        double sq = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i)
            if (mask & (1u << i)) sq += n[i] * n[i];
        const double len = sqrt(sq);
#pragma unroll
        for (int i = 0; i < D; ++i) n[i] = n[i] / len;
    }
}

// ---- result records (sorted-record mode): [values | jacobian (d, iv) | normal], a multiple of 4 doubles ------
template <int NIND, int NDEP, bool JAC>
__device__ __forceinline__ void store_result_record(const SplineDev &s, const OutDev &out, double *__restrict__ recOut,
                                                    const double (&v)[NDEP], const double (&g)[NIND][NDEP])
{
    constexpr int DN = (NIND - NDEP == 1 || NDEP - NIND == 1) ? (NIND > NDEP ? NIND : NDEP) : 0;
    constexpr int R = JAC ? NDEP + NDEP * NIND + DN : NDEP;
    constexpr int RP = (R + 3) & ~3;
    double rec[RP];
#pragma unroll
    for (int j = 0; j < RP; ++j) rec[j] = 0.0;
#pragma unroll
    for (int d = 0; d < NDEP; ++d) rec[d] = v[d];
    if constexpr (JAC) {
#pragma unroll
        for (int d = 0; d < NDEP; ++d)
#pragma unroll
            for (int iv = 0; iv < NIND; ++iv) rec[NDEP + d * NIND + iv] = g[iv][d];
        if constexpr (DN > 0) {
            if (out.normal) {
                double J[NDEP * NIND];
#pragma unroll
                for (int d = 0; d < NDEP; ++d)
#pragma unroll
                    for (int iv = 0; iv < NIND; ++iv) J[d * NIND + iv] = g[iv][d];
                double n[DN];
                normal_from_jacobian<NIND, NDEP>(J, s.normalSign, out.normalize, out.normalMask, n);
#pragma unroll
                for (int i = 0; i < DN; ++i) rec[NDEP + NDEP * NIND + i] = n[i];
            }
        }
    }
    double2 *q = reinterpret_cast<double2 *>(recOut);
#pragma unroll
    for (int j = 0; j < RP / 2; ++j)
        if (2 * j < out.aosStride) __stcs(q + j, make_double2(rec[2 * j], rec[2 * j + 1]));
}

// one tile of NDT dependent variables starting at d0 (no normals)
template <int NIND, int NDEP, int NDT, bool JAC>
__device__ __forceinline__ void store_result_tile(double *__restrict__ rec, const int d0, const double (&vt)[NDT],
                                                  const double (&gt)[NIND][NDT])
{
#pragma unroll
    for (int d = 0; d < NDT; ++d) __stcs(rec + d0 + d, vt[d]);
    if constexpr (JAC) {
        constexpr int run = NDT * NIND;
        double jr[run];
#pragma unroll
        for (int d = 0; d < NDT; ++d)
#pragma unroll
            for (int iv = 0; iv < NIND; ++iv) jr[d * NIND + iv] = gt[iv][d];
        const int start = NDEP + d0 * NIND;
        if constexpr (NDEP % 2 == 0 && (NDT * NIND) % 2 == 0) {
#pragma unroll
            for (int j = 0; j < run / 2; ++j)
                __stcs(reinterpret_cast<double2 *>(rec + start) + j, make_double2(jr[2 * j], jr[2 * j + 1]));
        } else {
#pragma unroll
            for (int j = 0; j < run; ++j) __stcs(rec + start + j, jr[j]);
        }
    }
}

// ---- compile-time shape kernel ----------------------------------------------------------------
template <int NIND, int O0, int O1, int O2, int O3>
struct Orders {
    static constexpr int n = NIND;
    __host__ __device__ static constexpr int at(int i) { return i == 0 ? O0 : i == 1 ? O1 : i == 2 ? O2 : O3; }
    static constexpr int omax = (O0 > O1 ? O0 : O1) > (O2 > O3 ? O2 : O3) ? (O0 > O1 ? O0 : O1) : (O2 > O3 ? O2 : O3);
};

template <class Ord, int NDEP, bool JAC>
struct FixedCtx {
    double B[Ord::n][Ord::omax];   // values (or wrt-derivative) basis
    double dB[Ord::n][Ord::omax];  // first-derivative basis (JAC only)
    long long stride[Ord::n];
    long long depStride;
};

// Contract variables L .. NIND-1 of the window whose corner (for variables >= L the corner, for
// variables < L the fixed index) is at `cp`.  v[d]: value part, g[m][d]: derivative w.r.t. m >= L.
template <int L, class Ord, int NDEP, bool JAC>
struct Contract {
    __device__ __forceinline__ static void run(const double *__restrict__ cp, const FixedCtx<Ord, NDEP, JAC> &c,
                                               double (&v)[NDEP], double (&g)[Ord::n][NDEP])
    {
        constexpr int O = Ord::at(L);
#pragma unroll
        for (int d = 0; d < NDEP; ++d) v[d] = 0.0;
        if constexpr (JAC) {
#pragma unroll
            for (int m = L; m < Ord::n; ++m)
#pragma unroll
                for (int d = 0; d < NDEP; ++d) g[m][d] = 0.0;
        }
        if constexpr (L == Ord::n - 1) {
            // innermost variable: contiguous coefficients
#pragma unroll
            for (int d = 0; d < NDEP; ++d) {
                const double *row = cp + d * c.depStride;
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    const double x = __ldg(row + i);
                    v[d] = fma(x, c.B[L][i], v[d]);
                    if constexpr (JAC) g[L][d] = fma(x, c.dB[L][i], g[L][d]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < O; ++i) {
                double cv[NDEP];
                double cg[Ord::n][NDEP];
                Contract<L + 1, Ord, NDEP, JAC>::run(cp + i * c.stride[L], c, cv, cg);
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    v[d] = fma(cv[d], c.B[L][i], v[d]);
                    if constexpr (JAC) {
                        g[L][d] = fma(cv[d], c.dB[L][i], g[L][d]);
#pragma unroll
                        for (int m = L + 1; m < Ord::n; ++m) g[m][d] = fma(cg[m][d], c.B[L][i], g[m][d]);
                    }
                }
            }
        }
    }
};

// CTAs per SM the fixed-shape kernel is compiled for: 4 (128 registers) unless the window is so large that the
// accumulators spill (the 4-variate nDep-6 manifold: measured 1.83 vs 1.50 Gpts/s with 168 registers)
constexpr int fixed_min_blocks(int nInd, int o0, int o1, int o2, int o3, int nDep, bool jac)
{
    int w = nDep * o0;
    if (nInd > 1) w *= o1;
    if (nInd > 2) w *= o2;
    if (nInd > 3) w *= o3;
    return (jac && w > 256) ? 3 : 4;
}

template <int IV, class Ord, int NDEP, bool JAC>
__device__ __forceinline__ void setup_variable(const SplineDev &s, double u, int d, FixedCtx<Ord, NDEP, JAC> &c,
                                               int (&ix)[Ord::n], bool &outside, bool given, const double *__restrict__ rec)
{
    constexpr int O = Ord::at(IV);
    const double *k = s.knots[IV];
    const int nKnots = O + s.nCoef[IV];
    int span = ix[IV];
    if (!given) {   // binned mode hands the span in (searched and domain-checked by bin_keys_kernel)
        outside |= (u < __ldg(k + O - 1)) | (u > __ldg(k + s.nCoef[IV]));
        span = span_search_inner(k, nKnots, O, u);
        ix[IV] = span;
    }
    double b0[O], b1[O];
    if (rec) {
        basis_from_span_record<O, JAC>(rec + (long long)(span - O) * SpanRec<O>::stride, u, d, b0, b1);
    } else {
        double kw[2 * (O - 1) > 0 ? 2 * (O - 1) : 1];
        load_knot_window<O>(k, span, kw);
        basis_regs<O, JAC>(kw, u, d, b0, b1);
    }
#pragma unroll
    for (int j = 0; j < O; ++j) {
        c.B[IV][j] = b0[j];
        if constexpr (JAC) c.dB[IV][j] = b1[j];
    }
}

// NDT = dependent variables contracted per pass over the window (NDEP: one pass; fewer: smaller accumulator set,
// more resident warps; the basis is computed once either way).  Normals need the whole jacobian: NDT == NDEP.
template <int NIND, int O0, int O1, int O2, int O3, int NDEP, bool JAC, int NDT = NDEP, int MINB = 0>
__global__ void __launch_bounds__(128, MINB ? MINB : fixed_min_blocks(NIND, O0, O1, O2, O3, NDT, JAC)) eval_fixed_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                         const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    const bool recs = in.records != nullptr;
    const bool binned = in.perm != nullptr || recs;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
        FixedCtx<Ord, NDT, JAC> c;
        int ix[NIND];
        double u[NIND];
        bool outside = false;
        long long p = t;
        if (binned) {
            int key;
            if (recs) {
                const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
                const double2 r0 = __ldcs(rp), r1 = __ldcs(rp + 1);
                u[0] = r0.x;
                if constexpr (NIND > 1) u[1] = r0.y;
                if constexpr (NIND > 2) u[2] = r1.x;
                if constexpr (NIND > 3) u[3] = r1.y;
                key = NIND > 3 ? __ldcs(in.recKey + t) : (int)__double_as_longlong(r1.y);
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
                    store_result_tile<NIND, NDEP, NDT, JAC>(out.aos + t * out.aosStride, d0, vt, gt);
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
            if (!out.aos && out.spans) {
#pragma unroll
                for (int iv = 0; iv < NIND; ++iv) __stcs(out.spans + iv * out.ld + p, ix[iv]);
            }
        } else {
        double v[NDEP];
        double g[NIND][NDEP];
        Contract<0, Ord, NDEP, JAC>::run(s.coefs + off, c, v, g);
        if (out.aos) {
            // sorted-record mode: one contiguous, sector-aligned result record per point
            store_result_record<NIND, NDEP, JAC>(s, out, out.aos + t * out.aosStride, v, g);
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

// ---- warp-staged windows (sorted-record mode) -------------------------------------------------------------------
// In cell order the 32 points of a warp share one coefficient window (two where the warp straddles a cell boundary),
// yet the thread-per-point kernel above still walks it through L1 with run-time strides (~2 address instructions per
// load, 212 to 540 loads per point) and every warp pays the whole latency chain record -> spans -> window once per
// 32 points (ncu: 31 % of the warp time in long-scoreboard stalls, FP64 pipe 36-45 % active).  Here
//   * warps are persistent and walk CONTIGUOUS runs of sorted tiles, so the window of the previous tile is normally
//     the window of this one (110-170 points per cell): it stays in shared memory, two slots per warp;
//   * a new window is copied once by the warp into a compact image with compile-time strides (cp.async, lanes along
//     the elements), and the contraction reads it with immediate offsets and 16-byte broadcast loads: no address
//     arithmetic, half the load instructions;
//   * the next tile's point records are requested before the current tile is contracted.
// Per-point arithmetic and summation order are those of Contract<> above: results are bit-identical.
template <class Ord, int NDEP>
struct WindowShape {
    static constexpr int last = Ord::at(Ord::n - 1);                  // innermost variable: contiguous in the spline
    __host__ __device__ static constexpr int stride(int iv)          // compact stride of variable iv (doubles)
    {
        int st = 1;
        for (int m = Ord::n - 1; m > iv; --m) st *= Ord::at(m);
        return st;
    }
    static constexpr int perDep = stride(0) * Ord::at(0);
    static constexpr int perDepPad = (perDep + 1) & ~1;              // every dependent variable starts 16-byte aligned
    static constexpr int size = perDepPad * NDEP;
};

// load n consecutive doubles at the compile-time-foldable offset `off` of a 16-byte aligned shared-memory image
template <int N>
__device__ __forceinline__ void load_run(const double *__restrict__ w, const int off, double (&x)[N])
{
    const int head = off & 1;                       // folds to a constant once the recursion is unrolled
    if (head) x[0] = w[off];
#pragma unroll
    for (int j = 0; j < N / 2 + 1; ++j) {
        const int idx = head + 2 * j;
        if (idx + 1 < N) {
            const double2 t = *reinterpret_cast<const double2 *>(w + off + idx);
            x[idx] = t.x;
            x[idx + 1] = t.y;
        }
    }
    if ((N - head) & 1) x[N - 1] = w[off + N - 1];
}

// Contract variables L .. NIND-1 of the compact window image; same recursion and summation order as Contract<>.
template <int L, class Ord, int NDEP, int NDT, bool JAC>
struct ContractS {
    using WS = WindowShape<Ord, NDEP>;
    __device__ __forceinline__ static void run(const double *__restrict__ w, const int off, const FixedCtx<Ord, NDT, JAC> &c,
                                               double (&v)[NDT], double (&g)[Ord::n][NDT])
    {
        constexpr int O = Ord::at(L);
#pragma unroll
        for (int d = 0; d < NDT; ++d) v[d] = 0.0;
        if constexpr (JAC) {
#pragma unroll
            for (int m = L; m < Ord::n; ++m)
#pragma unroll
                for (int d = 0; d < NDT; ++d) g[m][d] = 0.0;
        }
        if constexpr (L == Ord::n - 1) {
#pragma unroll
            for (int d = 0; d < NDT; ++d) {
                double x[O];
                load_run<O>(w, off + d * WS::perDepPad, x);
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    v[d] = fma(x[i], c.B[L][i], v[d]);
                    if constexpr (JAC) g[L][d] = fma(x[i], c.dB[L][i], g[L][d]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < O; ++i) {
                double cv[NDT];
                double cg[Ord::n][NDT];
                ContractS<L + 1, Ord, NDEP, NDT, JAC>::run(w, off + i * WS::stride(L), c, cv, cg);
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    v[d] = fma(cv[d], c.B[L][i], v[d]);
                    if constexpr (JAC) {
                        g[L][d] = fma(cv[d], c.dB[L][i], g[L][d]);
#pragma unroll
                        for (int m = L + 1; m < Ord::n; ++m) g[m][d] = fma(cg[m][d], c.B[L][i], g[m][d]);
                    }
                }
            }
        }
    }
};

// the warp copies the window of cell `key` (packed spans, bin_keys_kernel) into the compact image `dst`;
// lanes run along the elements: consecutive 8-byte words in shared memory, whole rows of the spline in global memory
template <class Ord, int NDEP>
__device__ __forceinline__ void stage_window(const SplineDev &s, int key, double *dst, const int lane)
{
    using WS = WindowShape<Ord, NDEP>;
    long long base = 0;
#pragma unroll
    for (int iv = Ord::n - 1; iv >= 0; --iv) {
        const int m = s.nCoef[iv] - Ord::at(iv) + 1;
        base += (long long)(key % m) * s.stride[iv];
        key /= m;
    }
    const unsigned dstAddr = (unsigned)__cvta_generic_to_shared(dst);
    constexpr int total = WS::perDep * NDEP;
#pragma unroll
    for (int e0 = 0; e0 < total; e0 += 32) {
        const int e = e0 + lane;
        if (total % 32 == 0 || e < total) {
            const int d = e / WS::perDep;
            int q = e - d * WS::perDep;
            long long src = base + (long long)d * s.depStride;
#pragma unroll
            for (int iv = Ord::n - 1; iv >= 0; --iv) {
                src += (long long)(q % Ord::at(iv)) * s.stride[iv];
                q /= Ord::at(iv);
            }
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dstAddr + (unsigned)(d * WS::perDepPad + e - d * WS::perDep) * 8u), "l"(s.coefs + src) : "memory");
        }
    }
}

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, bool JAC, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_staged_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                 const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using WS = WindowShape<Ord, NDEP>;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
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
    int k4 = -1;
    auto fetch = [&](long long tile) {
        const long long t = tile * 32 + lane;
        if (tile < endTile && t < N) {
            const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
            r0 = __ldcs(rp);
            r1 = __ldcs(rp + 1);
            if constexpr (NIND > 3) k4 = __ldcs(in.recKey + t);
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
        const int key = live ? (NIND > 3 ? k4 : (int)__double_as_longlong(r1.y)) : -1;
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
                double *rec = out.aos + t * out.aosStride;
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
            if (out.values) out.values[d * out.ld + p] = val;
            if (jac) {
                if (out.jacobian)
                    for (int iv = 0; iv < nInd; ++iv) out.jacobian[((long long)d * nInd + iv) * out.ld + p] = der[iv];
                if (out.normal)
                    for (int iv = 0; iv < nInd; ++iv) J[d * nInd + iv] = der[iv];
            }
        }
        if (jac && out.normal) {
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
            for (int i = 0; i < D; ++i) out.normal[(long long)i * out.ld + p] = out.normalize ? n[i] / len : n[i];
        }
    }
}

// ---- cell binning --------------------------------------------------------------------------------
// Scattered points on a spline whose coefficients do not fit in L1 gather a window of prod(order)*nDep doubles
// per point from L2 (1.5 KB for a tricubic volume, 3.9 KB for the 4-variate manifold): measured, the
// thread-per-point kernel is then bound by L2->SM traffic (~7 TB/s) at 10% of the FP64 roofline.  Binning makes
// the lanes of a warp share their window: the points of a chunk (small enough that its outputs stay in L2) are
// counting-sorted by knot-span cell, evaluated in cell order (window loads become L1 broadcasts) and written
// straight back to their original positions.  Same arithmetic per point, so results are bit-identical to the
// unbinned kernel.
// per-span records of one variable (runtime order): rec[s] = { knots[ix-(o-1)..ix-1] | 1/(knots[ix+t]-knots[ix-deg+t]) }
// with ix = o + s, the layout basis_from_span_record<O> reads; one thread per span
__global__ void __launch_bounds__(128) span_records_kernel(const double *__restrict__ kn, const int o, const int nCoef,
                                                           double *__restrict__ rec, const int stride)
{
    const int sp = blockIdx.x * blockDim.x + threadIdx.x;
    if (sp > nCoef - o) return;
    const int ix = o + sp;
    double *r = rec + (long long)sp * stride;
    for (int j = 0; j < o - 1; ++j) r[j] = kn[ix - (o - 1) + j];
    int at = o - 1;
    for (int deg = 1; deg < o; ++deg)
        for (int t = 0; t < deg; ++t) r[at++] = 1.0 / (kn[ix + t] - kn[ix - deg + t]);
    for (; at < stride; ++at) r[at] = 0.0;
}

static int span_rec_stride(int o) { return ((o - 1 + o * (o - 1) / 2) + 1) & ~1; }

__global__ void __launch_bounds__(256) bin_keys_kernel(const SplineDev s, const PointsDev in, const long long base, const int n,
                                                       int *__restrict__ keys, int *__restrict__ hist, int *__restrict__ rank,
                                                       const OutDev out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned active = __ballot_sync(0xffffffffu, t < n);
    if (t >= n) return;
    const long long p = base + t;
    int key = 0;
    bool outside = false;
    for (int iv = 0; iv < s.nInd; ++iv) {
        const double *k = s.knots[iv];
        const int o = s.order[iv];
        const double u = __ldg(in.uvw + p * in.pointStride + iv * in.varStride);
        outside |= (u < __ldg(k + o - 1)) | (u > __ldg(k + s.nCoef[iv]));
        const int ix = span_search_inner(k, o + s.nCoef[iv], o, u);
        if (out.spans) __stcs(out.spans + iv * out.ld + p, ix);
        key = key * (s.nCoef[iv] - o + 1) + (ix - o);
    }
    if (outside && out.firstOutside) report_outside((int64_t *)out.firstOutside, p);
    keys[t] = key;
    // one atomic per distinct cell in the warp (coherent inputs would otherwise serialise on one counter)
    const unsigned peers = __match_any_sync(active, key);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int first = 0;
    if (lane == leader) first = atomicAdd(hist + key, __popc(peers));
    if (rank) {   // position inside the cell: the scatter pass then needs no second round of atomics
        first = __shfl_sync(peers, first, leader);
        rank[t] = first + __popc(peers & ((1u << lane) - 1));
    }
}

// The same pass with P points per thread in flight (coordinates of all P points requested first, the P bisections
// advance together, P atomics outstanding), plus the rank of every point inside its cell; used by the sorted-record
// pipeline.  The pass is latency-bound (halving its occupancy doubles its time), and a quarter of the threads with
// four points each keeps more in flight than one point per thread.
template <int P>
__global__ void __launch_bounds__(128) bin_keys_batched_kernel(const SplineDev s, const PointsDev in, const long long base,
                                                               const int n, int *__restrict__ keys, int *__restrict__ hist,
                                                               int *__restrict__ rank, const OutDev out)
{
    const int t0 = blockIdx.x * (blockDim.x * P) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int key[P];
    bool outside[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { key[j] = 0; outside[j] = false; }
    for (int iv = 0; iv < s.nInd; ++iv) {
        const double *k = s.knots[iv];
        const int o = s.order[iv], nKnots = o + s.nCoef[iv];
        const double lo = __ldg(k + o - 1), hi = __ldg(k + s.nCoef[iv]);
        double u[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int t = t0 + j * blockDim.x;
            u[j] = t < n ? __ldcs(in.uvw + (base + t) * in.pointStride + iv * in.varStride) : lo;
        }
        int at[P], cnt[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            outside[j] |= (u[j] < lo) | (u[j] > hi);
            at[j] = o;
            cnt[j] = (u[j] != u[j]) ? 0 : nKnots - 2 * o;
        }
        bool more = true;
        while (more) {                                    // P upper-bound bisections side by side (span_search_inner)
            more = false;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                if (cnt[j] > 0) {
                    const int half = cnt[j] >> 1, mid = at[j] + half;
                    const bool le = __ldg(k + mid) <= u[j];
                    at[j] = le ? mid + 1 : at[j];
                    cnt[j] = le ? cnt[j] - half - 1 : half;
                    more |= cnt[j] > 0;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int ix = (u[j] != u[j]) ? nKnots - o : at[j];
            const int t = t0 + j * blockDim.x;
            if (out.spans && t < n) __stcs(out.spans + iv * out.ld + base + t, ix);
            key[j] = key[j] * (s.nCoef[iv] - o + 1) + (ix - o);
        }
    }
    int first[P];
    unsigned peers[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        const unsigned active = __ballot_sync(0xffffffffu, t < n);
        first[j] = 0;
        peers[j] = 0;
        if (t < n) {
            if (outside[j] && out.firstOutside) report_outside((int64_t *)out.firstOutside, base + t);
            keys[t] = key[j];
            peers[j] = __match_any_sync(active, key[j]);
            if (lane == __ffs(peers[j]) - 1) first[j] = atomicAdd(hist + key[j], __popc(peers[j]));
        }
    }
    if (rank) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int t = t0 + j * blockDim.x;
            if (t < n) {
                const int f = __shfl_sync(peers[j], first[j], __ffs(peers[j]) - 1);
                rank[t] = f + __popc(peers[j] & ((1u << lane) - 1));
            }
        }
    }
}

// ---- sorted-record variant (large chunks; no L2-residency assumption) --------------------------------------
// scatter: 32-byte point records in cell order (a full sector per point, so the scattered write needs no
// read-modify-write) at offset[cell] + rank (no atomics) + the inverse permutation (coalesced); P points per thread
template <int P>
__global__ void __launch_bounds__(128) bin_scatter_records_batched_kernel(const SplineDev s, const PointsDev in, const long long base,
                                                                          const int n, const int *__restrict__ keys,
                                                                          const int *__restrict__ offset, double *__restrict__ records,
                                                                          int *__restrict__ recKey, int *__restrict__ inv)
{
    const int t0 = blockIdx.x * (blockDim.x * P) + threadIdx.x;
    int key[P], pos[P];
    double r[P][4];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        key[j] = t < n ? keys[t] : 0;
        pos[j] = t < n ? inv[t] : 0;
#pragma unroll
        for (int iv = 0; iv < 4; ++iv)
            r[j][iv] = (t < n && iv < s.nInd) ? __ldcs(in.uvw + (base + t) * in.pointStride + iv * in.varStride) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < P; ++j) pos[j] += __ldg(offset + key[j]);
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        if (t < n) {
            if (s.nInd <= 3) r[j][3] = __longlong_as_double((long long)key[j]);
            else recKey[pos[j]] = key[j];
            double2 *q = reinterpret_cast<double2 *>(records + 4LL * pos[j]);
            q[0] = make_double2(r[j][0], r[j][1]);
            q[1] = make_double2(r[j][2], r[j][3]);
            inv[t] = pos[j];
        }
    }
}

// warp-aggregated slot claim: the lanes of a warp that share a cell take consecutive slots with one atomic
__device__ __forceinline__ int claim_slot(int *cursor, int key, unsigned active)
{
    const unsigned peers = __match_any_sync(active, key);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int first = 0;
    if (lane == leader) first = atomicAdd(cursor + key, __popc(peers));
    first = __shfl_sync(peers, first, leader);
    return first + __popc(peers & ((1u << lane) - 1));
}

// exclusive scan of hist[0..cells) in place, one CTA: 4096 counters per round (coalesced 16-byte loads, warp
// shuffles, one shared-memory hop between the warps), running total carried from round to round
__global__ void __launch_bounds__(1024) bin_scan_kernel(int *__restrict__ hist, const int cells)
{
    __shared__ int warpSum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < cells; base += 4096) {
        const int i = base + 4 * threadIdx.x;
        int4 c = make_int4(0, 0, 0, 0);
        if (i + 3 < cells) c = *reinterpret_cast<const int4 *>(hist + i);
        else {
            if (i < cells) c.x = hist[i];
            if (i + 1 < cells) c.y = hist[i + 1];
            if (i + 2 < cells) c.z = hist[i + 2];
        }
        const int mine = c.x + c.y + c.z + c.w;
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = warpSum[lane];
            int wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += v;
            }
            warpSum[lane] = wi - w;                       // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int start = carry + warpSum[warp] + incl - mine;
        const int4 o = make_int4(start, start + c.x, start + c.x + c.y, start + c.x + c.y + c.z);
        if (i + 3 < cells) *reinterpret_cast<int4 *>(hist + i) = o;
        else {
            if (i < cells) hist[i] = o.x;
            if (i + 1 < cells) hist[i + 1] = o.y;
            if (i + 2 < cells) hist[i + 2] = o.z;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = start + mine;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const int *__restrict__ keys, int *__restrict__ cursor, const int n,
                                                          int *__restrict__ perm, int *__restrict__ sortedKey)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned active = __ballot_sync(0xffffffffu, t < n);
    if (t >= n) return;
    const int key = keys[t];
    const int pos = claim_slot(cursor, key, active);
    perm[pos] = t;
    sortedKey[pos] = key;
}

// un-permute: a warp takes 32 consecutive points and pulls their 32 result records (each a run of whole sectors
// somewhere in the sorted array) into shared memory with 16-byte cp.async copies -- lanes run along the records,
// every sector is requested once, and all of a warp's copies are in flight together -- then every lane reads its
// own record and the warp writes the struct-of-arrays outputs coalesced (8 warps side by side: 2 KB per plane).
constexpr int UNPERM_WARPS = 8;
template <int S_>
__global__ void __launch_bounds__(UNPERM_WARPS * 32) bin_unpermute_kernel(const double *__restrict__ aos, const int aosStride,
                                                                          const int *__restrict__ inv, const long long base,
                                                                          const int n, const int nDep, const int nJ,
                                                                          const int nNormal, const OutDev out)
{
    extern __shared__ __align__(16) double tile[];         // per warp: 32 rows of (stride + 2) doubles
    const int S = S_ ? S_ : aosStride;
    const int pitch = S + 2, S2 = S >> 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *tw = tile + (long long)warp * 32 * pitch;
    const int first = (blockIdx.x * UNPERM_WARPS + warp) * 32;
    if (first >= n) return;
    const int t = first + lane;
    const int myRec = t < n ? __ldg(inv + t) : -1;
    const unsigned twAddr = (unsigned)__cvta_generic_to_shared(tw);
    const int total = 32 * S2;
#pragma unroll 8
    for (int idx = lane; idx < total; idx += 32) {
        const int r = idx / S2, c = idx - r * S2;
        const int rec = __shfl_sync(0xffffffffu, myRec, r);
        if (rec >= 0) {
            const double *src = aos + (long long)rec * S + 2 * c;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(twAddr + (unsigned)(r * pitch + 2 * c) * 8u), "l"(src) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (t >= n) return;
    const long long p = base + t;
    const double *mine = tw + lane * pitch;
    if (out.values) {
        for (int k = 0; k < nDep; ++k) __stcs(out.values + k * out.ld + p, mine[k]);
    }
    if (out.jacobian) {
        for (int k = 0; k < nJ; ++k) __stcs(out.jacobian + k * out.ld + p, mine[nDep + k]);
    }
    if (out.normal) {
        for (int k = 0; k < nNormal; ++k) __stcs(out.normal + k * out.ld + p, mine[nDep + nJ + k]);
    }
}

template <int S_>
static int launch_unpermute(const double *aos, int stride, const int *inv, long long base, int n, int nDep, int nJ, int nN,
                            const OutDev &out, cudaStream_t st)
{
    const size_t smem = sizeof(double) * UNPERM_WARPS * 32 * (stride + 2);
    if (int rc = allow_dynamic_smem(bin_unpermute_kernel<S_>, smem)) return rc;
    bin_unpermute_kernel<S_><<<(n + UNPERM_WARPS * 32 - 1) / (UNPERM_WARPS * 32), UNPERM_WARPS * 32, smem, st>>>(
        aos, stride, inv, base, n, nDep, nJ, nN, out);
    return 0;
}

// points per chunk in sorted-record mode (BSPY_BIN_REC_CHUNK_LOG2 overrides for experiments)
static long long bin_rec_chunk()
{
    const long long lg = option(OPT_BIN_REC_CHUNK_LOG2, 22);
    return 1LL << (lg < 16 ? 16 : (lg > 26 ? 26 : lg));
}
#define BIN_REC_CHUNK bin_rec_chunk()
constexpr long long BIN_CHUNK_MAX = 1 << 20;  // workspace is sized for this many points per chunk

// Points per chunk: the outputs of a chunk are scattered back to their original positions 8 bytes at a time, which
// is only cheap while the chunk's output region stays in L2 (measured optimum: ~56 MB of outputs per chunk; larger
// chunks give more points per cell but turn the scatter into DRAM read-modify-writes and are 2x slower).
static long long bin_chunk(long long outBytesPerPoint)
{
    long long v = option(OPT_BIN_CHUNK, (56LL << 20) / (outBytesPerPoint > 0 ? outBytesPerPoint : 8));
    v = v / 1024 * 1024;
    if (v < 65536) v = 65536;
    if (v > BIN_CHUNK_MAX) v = BIN_CHUNK_MAX;
    return v;
}
constexpr long long BIN_MAX_CELLS = 1 << 18;  // histogram / scan size limit

// ---- host dispatch ----------------------------------------------------------------------------
int launch_curve(const SplineDev &s, const PointsDev &in, long long N, const WrtDev &wrt, const OutDev &out, int jac,
                 cudaStream_t stream);

typedef void (*FixedFn)(const SplineDev, const PointsDev, const long long, const WrtDev, const OutDev);

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
static FixedFn find_fixed_tiled(const SplineDev &s, int jac, int code)
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

// warp-staged variants (sorted-record mode).  code = 10 * (dependent variables per pass) + CTAs per SM; 0 = default
struct StagedEntry {
    int nInd, o[4], nDep, jac, code;
    FixedFn fn;
    int windowDoubles;
};
#define BSPY_STAGED(NI, A, B, C, D_, ND, J, NDT, MB)                                                          \
    {NI, {A, B, C, D_}, ND, J, 10 * NDT + MB, eval_staged_kernel<NI, A, B, C, D_, ND, J != 0, NDT, MB>,       \
     WindowShape<Orders<NI, A, B, C, D_>, ND>::size}
static const StagedEntry kStaged[] = {
    // volumes (measured: tricubic nDep 3, value + jacobian, 379 -> 326 us per 4 Mi points, FP64 pipe 37 -> 50 %);
    // the 4-variate nDep-6 window gained nothing here (register spills once the pass over the dependent variables is
    // rolled) and stays on eval_fixed_kernel with one dependent variable per pass
    BSPY_STAGED(3, 3, 3, 3, 0, 3, 0, 3, 4), BSPY_STAGED(3, 3, 3, 3, 0, 3, 1, 3, 4),
    BSPY_STAGED(3, 4, 4, 4, 0, 1, 0, 1, 4), BSPY_STAGED(3, 4, 4, 4, 0, 1, 1, 1, 4),
    BSPY_STAGED(3, 4, 4, 4, 0, 3, 0, 3, 4), BSPY_STAGED(3, 4, 4, 4, 0, 3, 1, 3, 4),
    BSPY_STAGED(3, 4, 4, 4, 0, 4, 0, 4, 4), BSPY_STAGED(3, 4, 4, 4, 0, 4, 1, 4, 4),
};

static const StagedEntry *find_staged(const SplineDev &s, int jac, int code)
{
    for (const StagedEntry &e : kStaged) {
        if (e.nInd != s.nInd || e.nDep != s.nDep || e.jac != jac) continue;
        bool same = true;
        for (int i = 0; i < s.nInd; ++i) same &= e.o[i] == s.order[i];
        if (same && (code == 0 || code == e.code)) return &e;
    }
    return nullptr;
}

static FixedFn find_fixed(const SplineDev &s, int jac)
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
    if (!in.perm) {
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

static long long binned_cells(const SplineDev &s)
{
    long long cells = 1;
    for (int i = 0; i < s.nInd; ++i) {
        cells *= (s.nCoef[i] - s.order[i] + 1);
        if (cells > BIN_MAX_CELLS) return 0;
    }
    return cells;
}

static int bin_mode(long long N)
{
    // 0: scatter results 8 bytes at a time within L2-sized chunks; 1: sorted 32-byte point records, array-of-structs
    // results and an un-permute pass over 4 Mi-point chunks (every scattered access is a whole sector)
    return option(OPT_BIN_MODE, N >= (1 << 21) ? 1 : 0) ? 1 : 0;
}

static int aos_stride(const SplineDev &s)
{
    const int D = (s.nInd - s.nDep == 1 || s.nDep - s.nInd == 1) ? (s.nInd > s.nDep ? s.nInd : s.nDep) : 0;
    return (s.nDep + s.nDep * s.nInd + D + 3) & ~3;
}

static long long pad64(long long n) { return (n + 63) / 64 * 64; }

static long long span_records_bytes(const SplineDev &s)
{
    long long doubles = 0;
    for (int i = 0; i < s.nInd; ++i) doubles += (long long)(s.nCoef[i] - s.order[i] + 1) * span_rec_stride(s.order[i]);
    return 8 * pad64(doubles);
}

// bytes of workspace for the binned path, 0 when binning does not apply to this spline
long long binned_workspace(const SplineDev &s, long long N)
{
    if (s.nInd < 2 || N < 65536) return 0;
    long long window = s.nDep;
    for (int i = 0; i < s.nInd; ++i) window *= s.order[i];
    if (window * 8 < 512) return 0;                         // small windows: the gather is cheap anyway
    if (s.depStride * s.nDep * 8 < 128 * 1024) return 0;    // the whole spline fits in L1
    const long long cells = binned_cells(s);
    if (!cells) return 0;
    if (!find_fixed(s, 0)) return 0;
    if (bin_mode(N) == 1 && s.nInd <= 4) {
        const long long chunk = N < BIN_REC_CHUNK ? N : BIN_REC_CHUNK;
        return 2 * (4 * (3 * pad64(chunk) + pad64(cells + 1)) + 8 * (4 * pad64(chunk) + (long long)aos_stride(s) * pad64(chunk))) +
               span_records_bytes(s);
    }
    const long long chunk = N < BIN_CHUNK_MAX ? N : BIN_CHUNK_MAX;
    return 3 * 4 * pad64(chunk) + 4 * pad64(cells + 1);
}

// Two internal helper streams (+ the events that fork from / join to the caller's stream) for the sorted-record
// pipeline: the sort / un-permute passes are memory-bound, the evaluation FP64-bound, so chunk c+1 is sorted
// (high-priority stream) while chunk c is evaluated (low-priority stream).  Capturable in a CUDA graph.
// Sets are pooled per device behind a mutex: a call owns its set exclusively while it ENQUEUES (event record / wait
// pairs of two host threads can therefore never interleave) and returns it when it is done enqueueing; work already
// enqueued keeps the dependencies it captured, so the next owner may reuse streams and events right away.
struct BinStreams {
    cudaStream_t sort = nullptr, eval = nullptr;
    cudaEvent_t fork = nullptr, sorted[2] = {nullptr, nullptr}, evaluated[2] = {nullptr, nullptr}, joinSort = nullptr,
                joinEval = nullptr;
    int device = -1;
    BinStreams *next = nullptr;
};

static std::mutex g_binStreamsMutex;
static BinStreams *g_binStreamsFree[64] = {};

static BinStreams *acquire_bin_streams()
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    {
        std::lock_guard<std::mutex> lock(g_binStreamsMutex);
        if (BinStreams *b = g_binStreamsFree[dev]) {
            g_binStreamsFree[dev] = b->next;
            b->next = nullptr;
            return b;
        }
    }
    BinStreams *b = new BinStreams();
    b->device = dev;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = lowest priority (numerically largest)
    bool ok = cudaStreamCreateWithPriority(&b->sort, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&b->eval, cudaStreamNonBlocking, lo) == cudaSuccess;
    cudaEvent_t *evs[] = {&b->fork, &b->sorted[0], &b->sorted[1], &b->evaluated[0], &b->evaluated[1], &b->joinSort, &b->joinEval};
    for (cudaEvent_t *e : evs) ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {   // leave the partially built set to the process; the caller runs without overlap
        cudaGetLastError();
        delete b;
        return nullptr;
    }
    return b;
}

static void release_bin_streams(BinStreams *b)
{
    std::lock_guard<std::mutex> lock(g_binStreamsMutex);
    b->next = g_binStreamsFree[b->device];
    g_binStreamsFree[b->device] = b;
}

static long long records_half_bytes(const SplineDev &s, long long chunk)
{
    const long long cells = binned_cells(s);
    return 4 * (3 * pad64(chunk) + pad64(cells + 1)) + 8 * (4 * pad64(chunk) + (long long)aos_stride(s) * pad64(chunk));
}

static int eval_binned_records(const SplineDev &s, PointsDev in, long long N, const WrtDev &wrt, OutDev out, int jac,
                               void *workspace, cudaStream_t stream)
{
    const long long cells = binned_cells(s);
    const long long chunk = N < BIN_REC_CHUNK ? N : BIN_REC_CHUNK;
    const long long cpad = pad64(chunk);
    const long long half = records_half_bytes(s, chunk);
    const int D = (s.nInd - s.nDep == 1 || s.nDep - s.nInd == 1) ? (s.nInd > s.nDep ? s.nInd : s.nDep) : 0;
    const int nJ = jac ? s.nDep * s.nInd : 0;
    const int nN = (jac && out.normal) ? D : 0;
    const int stride = (s.nDep + nJ + nN + 3) & ~3;
    FixedFn fn = find_fixed(s, jac);
    {
        const int ndt = (int)option(OPT_DEP_TILE, 14);
        FixedFn tiled = (ndt > 0 && !out.normal) ? find_fixed_tiled(s, jac, ndt) : nullptr;
        if (tiled) fn = tiled;
    }
    // warp-staged windows where the shape is compiled (no normals there: they need the whole jacobian in one pass)
    const StagedEntry *staged = nullptr;
    {
        const int code = (int)option(OPT_STAGED, 0);
        if (code >= 0) staged = find_staged(s, jac, code);
        if (staged) {
            if (int rc = allow_dynamic_smem(staged->fn, sizeof(double) * 4 * 2 * staged->windowDoubles)) return rc;
        }
    }
    // per-span records (left knots | reciprocal gaps) for every variable: no divisions in the evaluation kernel
    const double *spanRec[BSPY_MAX_IND] = {};
    {
        if (option(OPT_SPAN_RECORDS, 1)) {
            double *at = (double *)((char *)workspace + 2 * half);
            for (int i = 0; i < s.nInd; ++i) {
                const int spans = s.nCoef[i] - s.order[i] + 1, st = span_rec_stride(s.order[i]);
                span_records_kernel<<<(spans + 127) / 128, 128, 0, stream>>>(s.knots[i], s.order[i], s.nCoef[i], at, st);
                spanRec[i] = at;
                at += (long long)spans * st;
            }
            count_launch(s.nInd);
        }
    }
    // Sort / un-permute of the neighbouring chunks on a second stream under the evaluation of this one.  Measured: +9 %
    // on config 4 with the persistent staged kernel (6.40 -> 7.00 Gpts/s; its tail and the memory-bound passes fill
    // each other's gaps), nothing with the one-tile-per-CTA gather kernel (config 5: 2.47 -> 2.46), and a loss when
    // the evaluation is launched with fewer CTAs per SM to make room (3 CTAs: 6.6, 2 CTAs: 5.9).  BSPY_BIN_OVERLAP=0/1
    // overrides.
    const bool wantOverlap = option(OPT_BIN_OVERLAP, staged != nullptr ? 1 : 0) != 0;
    BinStreams *bs = wantOverlap ? acquire_bin_streams() : nullptr;
    struct Release { BinStreams *b; ~Release() { if (b) release_bin_streams(b); } } releaseOnExit{bs};
    const long long nChunks = (N + chunk - 1) / chunk;
    const bool overlap = bs != nullptr && nChunks > 1;
    cudaStream_t sSort = overlap ? bs->sort : stream, sEval = overlap ? bs->eval : stream;
    if (overlap) {
        cudaEventRecord(bs->fork, stream);
        cudaStreamWaitEvent(sSort, bs->fork, 0);
        cudaStreamWaitEvent(sEval, bs->fork, 0);
    }
    struct Buf { int *keys, *inv, *recKey, *hist; double *records, *aos; } buf[2];
    for (int h = 0; h < 2; ++h) {
        char *base = (char *)workspace + h * half;
        buf[h].keys = (int *)base; buf[h].inv = buf[h].keys + cpad; buf[h].recKey = buf[h].inv + cpad;
        buf[h].hist = buf[h].recKey + cpad;
        buf[h].records = (double *)(buf[h].hist + pad64(cells + 1));
        buf[h].aos = buf[h].records + 4 * cpad;
    }
    auto sortChunk = [&](long long c) -> int {
        const Buf &B = buf[c & 1];
        const long long base = c * chunk;
        const int n = (int)(N - base < chunk ? N - base : chunk);
        cudaError_t e = cudaMemsetAsync(B.hist, 0, sizeof(int) * (cells + 1), sSort);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        OutDev o1{};
        o1.ld = out.ld; o1.spans = out.spans; o1.firstOutside = out.firstOutside;
        // four points per thread in flight (measured against one point per thread at full occupancy: keys 106 -> 95 us,
        // scatter 88 -> 71 us per 4 Mi points)
        bin_keys_batched_kernel<4><<<(n + 511) / 512, 128, 0, sSort>>>(s, in, base, n, B.keys, B.hist, B.inv, o1);
        bin_scan_kernel<<<1, 1024, 0, sSort>>>(B.hist, (int)cells);
        bin_scatter_records_batched_kernel<4><<<(n + 511) / 512, 128, 0, sSort>>>(s, in, base, n, B.keys, B.hist, B.records, B.recKey, B.inv);
        if (overlap) cudaEventRecord(bs->sorted[c & 1], sSort);
        count_launch(3);
        return check_launch("bspy_cuda_eval_points_binned(sort)");
    };
    auto evalChunk = [&](long long c) -> int {
        const Buf &B = buf[c & 1];
        const long long base = c * chunk;
        const int n = (int)(N - base < chunk ? N - base : chunk);
        if (overlap) cudaStreamWaitEvent(sEval, bs->sorted[c & 1], 0);
        PointsDev pin{};
        pin.records = B.records; pin.recKey = B.recKey;
        for (int i = 0; i < s.nInd; ++i) pin.spanRec[i] = spanRec[i];
        OutDev o2 = out;
        o2.spans = nullptr; o2.firstOutside = nullptr;
        o2.aos = B.aos; o2.aosStride = stride;
        // the staged kernel lives on window reuse: it needs cells that hold a few tiles' worth of points (a sparse
        // tail chunk makes every tile straddle several cells); below that the L1-gather kernel is the faster one
        if (staged && n >= 48 * cells) {
            // persistent warps over contiguous runs of tiles (window reuse between consecutive tiles)
            long long blocks = (long long)num_sms() * (staged->code % 10);
            if (blocks > (n + 127) / 128) blocks = (n + 127) / 128;
            staged->fn<<<(unsigned)blocks, 128, sizeof(double) * 4 * 2 * staged->windowDoubles, sEval>>>(s, pin, n, wrt, o2);
        }
        else
            fn<<<(unsigned)((n + 127) / 128), 128, 0, sEval>>>(s, pin, n, wrt, o2);
        if (overlap) cudaEventRecord(bs->evaluated[c & 1], sEval);
        count_launch(1);
        return check_launch("bspy_cuda_eval_points_binned(eval)");
    };
    auto unpermChunk = [&](long long c) -> int {
        const Buf &B = buf[c & 1];
        const long long base = c * chunk;
        const int n = (int)(N - base < chunk ? N - base : chunk);
        if (overlap) cudaStreamWaitEvent(sSort, bs->evaluated[c & 1], 0);
        int urc;
        switch (stride) {
            case 4: urc = launch_unpermute<4>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 8: urc = launch_unpermute<8>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 12: urc = launch_unpermute<12>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 16: urc = launch_unpermute<16>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 32: urc = launch_unpermute<32>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            default: urc = launch_unpermute<0>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
        }
        if (urc) return urc;
        count_launch(1);
        return check_launch("bspy_cuda_eval_points_binned(unpermute)");
    };
    // software pipeline: sort(c+1) is enqueued before unpermute(c) so that it runs under eval(c)
    int rc = sortChunk(0);
    for (long long c = 0; c < nChunks && !rc; ++c) {
        rc = evalChunk(c);
        if (!rc && c + 1 < nChunks) rc = sortChunk(c + 1);
        if (!rc) rc = unpermChunk(c);
    }
    if (overlap) {
        cudaEventRecord(bs->joinSort, sSort);
        cudaEventRecord(bs->joinEval, sEval);
        cudaStreamWaitEvent(stream, bs->joinSort, 0);
        cudaStreamWaitEvent(stream, bs->joinEval, 0);
    }
    return rc;
}

int eval_binned(const SplineDev &s, PointsDev in, long long N, const WrtDev &wrt, OutDev out, int jac, void *workspace,
                cudaStream_t stream)
{
    if (bin_mode(N) == 1 && s.nInd <= 4) return eval_binned_records(s, in, N, wrt, out, jac, workspace, stream);
    const long long cells = binned_cells(s);
    const int D = s.nInd > s.nDep ? s.nInd : s.nDep;
    const long long outBytes = 8LL * ((out.values ? s.nDep : 0) + (out.jacobian ? s.nDep * s.nInd : 0) + (out.normal ? D : 0));
    const long long cap = N < BIN_CHUNK_MAX ? N : BIN_CHUNK_MAX;      // what the workspace was sized for
    const long long chunk = bin_chunk(outBytes) < cap ? bin_chunk(outBytes) : cap;
    const long long cpad = pad64(cap);
    int *keys = (int *)workspace, *perm = keys + cpad, *skey = perm + cpad, *hist = skey + cpad;
    FixedFn fn = find_fixed(s, jac);
    int32_t *spans = out.spans;
    long long *flag = out.firstOutside;
    for (long long base = 0; base < N; base += chunk) {
        const int n = (int)(N - base < chunk ? N - base : chunk);
        cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(int) * (cells + 1), stream);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        OutDev o1{};
        o1.ld = out.ld; o1.spans = spans; o1.firstOutside = flag;
        bin_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(s, in, base, n, keys, hist, nullptr, o1);
        bin_scan_kernel<<<1, 1024, 0, stream>>>(hist, (int)cells);
        bin_scatter_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, hist, n, perm, skey);
        PointsDev pin = in;
        pin.perm = perm; pin.cellKey = skey; pin.base = base;
        OutDev o2 = out;
        o2.spans = nullptr; o2.firstOutside = nullptr;
        long long blocks = (n + 127) / 128;
        fn<<<(unsigned)blocks, 128, 0, stream>>>(s, pin, n, wrt, o2);
        count_launch(4);
        int rc = check_launch("bspy_cuda_eval_points_binned");
        if (rc) return rc;
    }
    return 0;
}

}  // namespace bspy

using namespace bspy;

extern "C" int64_t bspy_cuda_binned_workspace_bytes(const bspy_spline *spline, int64_t N)
{
    SplineDev s;
    if (make_spline_dev(spline, s, "bspy_cuda_binned_workspace_bytes")) return 0;
    return binned_workspace(s, N);
}

extern "C" int bspy_cuda_eval_points_binned(const bspy_spline *spline, const double *uvw, int64_t pointStride, int64_t varStride,
                                            int64_t N, const int32_t *wrt, uint32_t flags, uint32_t normalMask, double *values,
                                            double *deriv, double *jacobian, double *normal, int32_t *spans,
                                            int64_t *firstOutside, void *workspace, int64_t workspaceBytes, void *stream)
{
    const char *who = "bspy_cuda_eval_points_binned";
    SplineDev s;
    int rc = make_spline_dev(spline, s, who);
    if (rc) return rc;
    const long long need = binned_workspace(s, N);
    const bool onePass = !(deriv && (values || jacobian || normal));
    if (!need || !workspace || workspaceBytes < need || !onePass || !uvw)
        return bspy_cuda_eval_points(spline, uvw, pointStride, varStride, N, wrt, flags, normalMask, values, deriv, jacobian,
                                     normal, spans, firstOutside, stream);
    if ((deriv != nullptr) != (wrt != nullptr)) { set_error("%s: wrt and deriv must be given together", who); return BSPY_E_ARG; }
    const int D = s.nInd > s.nDep ? s.nInd : s.nDep;
    if (normal && (s.nInd - s.nDep != 1 && s.nDep - s.nInd != 1)) {
        set_error("The number of independent variables must be one different than the number of dependent variables.");
        return BSPY_E_NORMAL_DIMS;
    }
    if (normalMask == 0 || D >= 32) normalMask = 0xffffffffu;
    PointsDev in{};
    in.uvw = uvw; in.pointStride = pointStride; in.varStride = varStride;
    OutDev out{};
    out.ld = N;
    out.firstOutside = (long long *)firstOutside;
    out.spans = spans;
    out.normalize = (flags & BSPY_NORMALIZE) ? 1u : 0u;
    out.normalMask = normalMask;
    WrtDev w{};
    int jac = 0;
    if (jacobian || normal) {
        jac = 1;
        out.values = values; out.jacobian = jacobian; out.normal = normal;
    } else if (deriv) {
        for (int i = 0; i < s.nInd; ++i) {
            if (wrt[i] < 0) { set_error("%s: negative derivative order", who); return BSPY_E_ARG; }
            w.d[i] = wrt[i];
        }
        out.values = deriv;
    } else {
        out.values = values;
    }
    return eval_binned(s, in, N, w, out, jac, workspace, (cudaStream_t)stream);
}

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
