// Device templates and host declarations shared by scattered.cu (thread-per-point kernels, host dispatch) and cells.cu
// (cell-sorted pipelines for big scattered batches): cofactor normals, result records, the register-resident
// sum-factorised contraction, warp-staged windows.
#pragma once
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
            if (out.normal || out.aosNormal) {
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
    if (out.aosWide) {
        // one 32-byte sector per store instruction (STG.256): half the requests of 16-byte stores
#pragma unroll
        for (int j = 0; j < RP / 4; ++j)
            if (4 * j < out.aosStride)
                asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(recOut + 4 * j), "d"(rec[4 * j]), "d"(rec[4 * j + 1]),
                             "d"(rec[4 * j + 2]), "d"(rec[4 * j + 3]) : "memory");
        return;
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

// per-span records of all variables of one cell, back to back (each part an even number of doubles)
template <class Ord>
struct CellRecords {
    __host__ __device__ static constexpr int stride(int iv)
    {
        const int o = Ord::at(iv);
        return ((o - 1 + o * (o - 1) / 2) + 1) & ~1;
    }
    __host__ __device__ static constexpr int offset(int iv)
    {
        int at = 0;
        for (int m = 0; m < iv; ++m) at += stride(m);
        return at;
    }
    static constexpr int size = offset(Ord::n);
};

// basis values and first derivatives from a span record held in shared memory (16-byte aligned)
template <int O, bool DER>
__device__ __forceinline__ void basis_from_shared_record(const double *__restrict__ rec, double u, int d, double (&b0)[O], double (&b1)[O])
{
    using R = SpanRec<O>;
    double r[R::stride > 0 ? R::stride : 1];
#pragma unroll
    for (int j = 0; j < R::stride / 2; ++j) {
        const double2 x = *reinterpret_cast<const double2 *>(rec + 2 * j);
        r[2 * j] = x.x;
        r[2 * j + 1] = x.y;
    }
    double dl[O > 1 ? O - 1 : 1], rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
    for (int j = 0; j < O - 1; ++j) dl[j] = u - r[j];
#pragma unroll
    for (int j = 0; j < O * (O - 1) / 2; ++j) rc[j] = r[O - 1 + j];
    basis_core<O, DER>(dl, rc, d, b0, b1);
}

// ---- host side (defined in scattered.cu) ------------------------------------------------------------------------
typedef void (*FixedFn)(const SplineDev, const PointsDev, const long long, const WrtDev, const OutDev);

// warp-staged variants (sorted-record mode).  code = 10 * (dependent variables per pass) + CTAs per SM; 0 = default
struct StagedEntry {
    int nInd, o[4], nDep, jac, code;
    FixedFn fn;
    int windowDoubles;
    FixedFn fn2;          // second generation: the cell's span records are staged with its window
    int slotDoubles;
    int optIn;            // only selected by an explicit STAGED=<code> (variants kept for measurement)
};

FixedFn find_fixed(const SplineDev &s, int jac);
FixedFn find_fixed_tiled(const SplineDev &s, int jac, int code);
const StagedEntry *find_staged(const SplineDev &s, int jac, int code);
const StagedEntry *find_staged_pair(const SplineDev &s, int code);   // windowDoubles = per-warp HALF: slot + result tile of one point
int make_spline_dev(const bspy_spline *sp, SplineDev &s, const char *who);
int launch_eval(const SplineDev &s, const PointsDev &in, long long N, const WrtDev &wrt, const OutDev &out, int jac,
                cudaStream_t stream);

// cells.cu
long long binned_workspace(const SplineDev &s, long long N);
long long binned_workspace(const SplineDev &s, long long N, bool aosOut);
int eval_binned(const SplineDev &s, PointsDev in, long long N, const WrtDev &wrt, OutDev out, int jac, void *workspace,
                cudaStream_t stream);

}  // namespace bspy
