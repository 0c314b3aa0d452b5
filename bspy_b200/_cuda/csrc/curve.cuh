// Lean curve evaluation (nInd == 1) from per-span records in shared memory.
//
// A scattered curve point at the HBM roofline has a budget of ~300 instructions and ~290 bytes of on-chip
// gathers (config 1: 32 B/pt of HBM traffic, config 3: 36 B/pt).  IEEE divisions in the recurrence (6 per cubic
// point, ~25 instructions each) and a bisection over the knots blow that budget, so everything that depends only
// on the knots is hoisted into a per-span record, built once per curve (per CTA for a single curve):
//     rec[s] = { knots[ix-(O-1)] .. knots[ix-1],  1/(knots[ix+t]-knots[ix-deg+t]) for deg = 1..O-1, t < deg }
// with ix = O + s.  The recurrence is then subtract / multiply / fma only (26 FP64 instructions for a cubic).
// alpha = (u-k)*r instead of (u-k)/gap differs from the reference by one rounding: covered by the 1e-12 bar
// (bit-exact basis values are the business of bspy_cuda_basis).  Zero-width spans give r = inf and the same
// inf/NaN results as the division.
#pragma once
#include "common.cuh"

namespace bspy {

// cooperative build by `nthreads` threads (a warp or a CTA); kn may be shared or global memory
template <int O>
__device__ __forceinline__ void build_span_records(const double *kn, int nCoef, double *rec, int tid, int nthreads)
{
    using R = SpanRec<O>;
    for (int s = tid; s <= nCoef - O; s += nthreads) {
        const int ix = O + s;
        double *r = rec + s * R::stride;
#pragma unroll
        for (int j = 0; j < O - 1; ++j) r[j] = kn[ix - (O - 1) + j];
        int at = O - 1;
#pragma unroll
        for (int deg = 1; deg < O; ++deg)
#pragma unroll
            for (int t = 0; t < deg; ++t) r[at++] = 1.0 / (kn[ix + t] - kn[ix - deg + t]);
        if (R::stride > R::used) r[R::used] = 0.0;
    }
}

// basis values b0 (and first derivatives b1 when DER) from a record held in registers
template <int O, bool DER>
__device__ __forceinline__ void basis_from_record(const double (&r)[SpanRec<O>::stride > 0 ? SpanRec<O>::stride : 1], double u,
                                                  double (&b0)[O], double (&b1)[O])
{
#pragma unroll
    for (int j = 0; j < O; ++j) { b0[j] = 0.0; b1[j] = 0.0; }
    b0[O - 1] = 1.0;
    double dl[O > 1 ? O - 1 : 1];   // u - left knots: dl[j] = u - knots[ix-(O-1)+j]
#pragma unroll
    for (int j = 0; j < O - 1; ++j) dl[j] = u - r[j];
    int at = O - 1;
#pragma unroll
    for (int deg = 1; deg < O; ++deg) {
        if (DER && deg == O - 1) {
#pragma unroll
            for (int j = 0; j < O; ++j) b1[j] = b0[j];
        }
#pragma unroll
        for (int t = 0; t < deg; ++t) {
            const int slot = O - deg + t;
            const double rc = r[at + t];
            const double a = dl[O - 1 - deg + t] * rc;      // knots[ix-deg+t] is left knot number O-1-deg+t
            if (DER && deg == O - 1) {
                const double g = (double)deg * rc;
                b1[slot - 1] = fma(-g, b1[slot], b1[slot - 1]);
                b1[slot] *= g;
            }
            b0[slot - 1] = fma(1.0 - a, b0[slot], b0[slot - 1]);
            b0[slot] *= a;
        }
        at += deg;
    }
}

// span of u among knots kn[0 .. O+nCoef): number of knots <= u clamped to [O, nCoef]; NaN -> nCoef
__device__ __forceinline__ int curve_span(const double *kn, int O, int nCoef, double u)
{
    int ix = O, n = nCoef - O;
    if (u != u) return nCoef;
    while (n > 0) {
        const int half = n >> 1;
        const bool le = kn[ix + half] <= u;
        ix = le ? ix + half + 1 : ix;
        n = le ? n - half - 1 : half;
    }
    return ix;
}

// one point: value v[NDEP] (and first derivative g[NDEP]) from shared-memory tables
//   kn   knots, rec per-span records, cf coefficients interleaved as cf[i * NDEP + d]
template <int O, int NDEP, bool DER>
__device__ __forceinline__ int curve_point(const double *kn, const double *rec, const double *cf, int nCoef, double u,
                                           double (&v)[NDEP], double (&g)[NDEP])
{
    using R = SpanRec<O>;
    const int ix = curve_span(kn, O, nCoef, u);
    double r[R::stride > 0 ? R::stride : 1];
    if constexpr (R::stride > 0) {
        const double2 *rp = reinterpret_cast<const double2 *>(rec + (ix - O) * R::stride);
#pragma unroll
        for (int j = 0; j < R::stride / 2; ++j) {
            const double2 x = rp[j];
            r[2 * j] = x.x;
            r[2 * j + 1] = x.y;
        }
    }
    double b0[O], b1[O];
    basis_from_record<O, DER>(r, u, b0, b1);
    const double *c = cf + (ix - O) * NDEP;
#pragma unroll
    for (int d = 0; d < NDEP; ++d) { v[d] = 0.0; g[d] = 0.0; }
#pragma unroll
    for (int j = 0; j < O; ++j)
#pragma unroll
        for (int d = 0; d < NDEP; ++d) {
            const double x = c[j * NDEP + d];
            v[d] = fma(x, b0[j], v[d]);
            if (DER) g[d] = fma(x, b1[j], g[d]);
        }
    return ix;
}

}  // namespace bspy
