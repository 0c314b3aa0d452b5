// Many independent curves of one shape: bspy_cuda_eval_many.
//
// Replaces a Python loop of `spline(uArray)` over a list of splines (reference ufunc-style
// evaluate, bspy/spline.py:940-947, one interpreter pass per point) for S curves that share
// order / nCoef / nDep but have their own knots, coefficients and parameters.
// One warp per curve, no block-level synchronisation:
//   1. the curve's knots and coefficients (a few hundred bytes) are staged in the warp's slice of shared memory
//      with coalesced loads, coefficients interleaved as cf[i][dep];
//   2. the lanes build the per-span records (left knots + reciprocal knot gaps, curve.cuh) once per curve --
//      the only divisions of the whole evaluation;
//   3. each lane evaluates every 32nd parameter: bisection over the staged knots, one record fetch (16-byte
//      loads), subtract/multiply/fma recurrence, window dot products;
//   4. stores are coalesced along the point index.
#include "curve.cuh"

namespace bspy {

struct ManyParams {
    int nCoef, nDep, nPts;
    long long nSplines;
    const double *knots, *coefs, *u;
    long long knotStride, coefStride;
    double *values, *deriv1;
    long long *firstOutside;
    int slice;  // doubles of shared memory per warp
};

// NDEP == 0: runtime nDep (one dot product per dependent variable)
template <int O, int NDEP, bool DER>
__global__ void __launch_bounds__(256) many_kernel(const ManyParams P)
{
    using R = SpanRec<O>;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const int nKnots = O + P.nCoef;
    const int nDep = NDEP ? NDEP : P.nDep;
    double *kn = sm + (long long)warp * P.slice;                 // knots
    double *rec = kn + ((nKnots + 1) & ~1);                      // (nCoef-O+1) records
    double *cf = rec + (P.nCoef - O + 1) * R::stride;            // cf[i * nDep + d]
    const int nC = nDep * P.nCoef;
    for (long long s = (long long)blockIdx.x * warpsPerBlock + warp; s < P.nSplines; s += (long long)gridDim.x * warpsPerBlock) {
        const double *gk = P.knots + s * P.knotStride;
        const double *gc = P.coefs + s * P.coefStride;
        __syncwarp();
        for (int i = lane; i < nKnots; i += 32) kn[i] = __ldcs(gk + i);
        for (int i = lane; i < nC; i += 32) {
            const int d = i / P.nCoef, c = i - d * P.nCoef;      // global layout (nDep, nCoef)
            cf[c * nDep + d] = __ldcs(gc + i);
        }
        __syncwarp();
        build_span_records<O>(kn, P.nCoef, rec, lane, 32);
        __syncwarp();
        const double lo = kn[O - 1], hi = kn[P.nCoef];
        const double *gu = P.u + s * P.nPts;
        double *ov = P.values + s * nDep * P.nPts;
        double *og = DER ? P.deriv1 + s * nDep * P.nPts : nullptr;
#pragma unroll 2
        for (int p = lane; p < P.nPts; p += 32) {
            const double u = __ldcs(gu + p);
            if (((u < lo) | (u > hi)) && P.firstOutside) report_outside((int64_t *)P.firstOutside, s * P.nPts + p);
            if constexpr (NDEP > 0) {
                double v[NDEP], g[NDEP];
                curve_point<O, NDEP, DER>(kn, rec, cf, P.nCoef, u, v, g);
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    __stcs(ov + d * P.nPts + p, v[d]);
                    if (DER) __stcs(og + d * P.nPts + p, g[d]);
                }
            } else {
                const int ix = curve_span(kn, O, P.nCoef, u);
                double r[R::stride > 0 ? R::stride : 1];
#pragma unroll
                for (int j = 0; j < R::stride; ++j) r[j] = rec[(ix - O) * R::stride + j];
                double b0[O], b1[O];
                basis_from_record<O, DER>(r, u, b0, b1);
                for (int d = 0; d < nDep; ++d) {
                    double v = 0.0, g = 0.0;
#pragma unroll
                    for (int j = 0; j < O; ++j) {
                        const double x = cf[(ix - O + j) * nDep + d];
                        v = fma(x, b0[j], v);
                        if (DER) g = fma(x, b1[j], g);
                    }
                    __stcs(ov + d * P.nPts + p, v);
                    if (DER) __stcs(og + d * P.nPts + p, g);
                }
            }
        }
    }
}

// ---- the same with the latencies taken out of the dependency chains ----------------------------------------------------
// ncu (source view) on many_kernel: 28 % of the warp time waits for the parameter u at the first instruction that uses it
// (the point loop requests it from HBM and needs it at once), 15 % on the five dependent probes of the bisection, 8 % on
// the curve's knots / coefficients at the stores that stage them -- the kernel is bound by these three latencies, not by
// bandwidth.  Here, per curve:
//   * the lane's first PF parameters are requested BEFORE the tables are built, and the NEXT curve's knots and coefficients
//     are requested into registers before the point loop and written to shared memory after it;
//   * points are processed G at a time: the G bisections advance together (independent probe chains), then the G
//     record / window fetches, then the arithmetic.
// Same per-point arithmetic as many_kernel (curve.cuh): bit-identical results.
constexpr int MANY_RAW = 6;     // staged elements per lane that travel through registers (6 * 32 = 192 doubles)
constexpr int MANY_PF = 8;      // parameters per lane requested up front (8 * 32 = 256 points per curve)
constexpr int MANY_G = 4;       // points per lane processed together

template <int O, int NDEP, bool DER>
__global__ void __launch_bounds__(256, 4) many2_kernel(const ManyParams P)
{
    using R = SpanRec<O>;
    extern __shared__ __align__(16) double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const int nKnots = O + P.nCoef;
    double *kn = sm + (long long)warp * P.slice;                 // knots
    double *rec = kn + ((nKnots + 1) & ~1);                      // (nCoef-O+1) records
    double *cf = rec + (P.nCoef - O + 1) * R::stride;            // cf[i * NDEP + d]
    const int nC = NDEP * P.nCoef, nRaw = nKnots + nC;
    const long long step = (long long)gridDim.x * warpsPerBlock;
    long long s = (long long)blockIdx.x * warpsPerBlock + warp;
    double raw[MANY_RAW];
    // element e of a curve's staged data: e < nKnots -> knots[e], else coefficient e - nKnots in (nDep, nCoef) order
    auto request = [&](long long curve) {
        const double *gk = P.knots + curve * P.knotStride, *gc = P.coefs + curve * P.coefStride;
#pragma unroll
        for (int k = 0; k < MANY_RAW; ++k) {
            const int e = lane + 32 * k;
            if (e < nRaw) raw[k] = __ldcs(e < nKnots ? gk + e : gc + (e - nKnots));
        }
    };
    auto put = [&](int e, double x) {
        if (e < nKnots) kn[e] = x;
        else {
            const int i = e - nKnots, d = i / P.nCoef, c = i - d * P.nCoef;
            cf[c * NDEP + d] = x;
        }
    };
    if (s < P.nSplines) request(s);
    for (; s < P.nSplines; s += step) {
        const double *gu = P.u + s * P.nPts;
        // this curve's parameters leave HBM now, before its tables exist
        double up[MANY_PF];
#pragma unroll
        for (int k = 0; k < MANY_PF; ++k) up[k] = lane + 32 * k < P.nPts ? __ldcs(gu + lane + 32 * k) : 0.0;
        __syncwarp();                                            // the previous curve's point loop is done with the tables
#pragma unroll
        for (int k = 0; k < MANY_RAW; ++k)
            if (lane + 32 * k < nRaw) put(lane + 32 * k, raw[k]);
        if (nRaw > 32 * MANY_RAW) {                              // long curves: the rest is staged directly
            const double *gk = P.knots + s * P.knotStride, *gc = P.coefs + s * P.coefStride;
            for (int e = lane + 32 * MANY_RAW; e < nRaw; e += 32) put(e, __ldcs(e < nKnots ? gk + e : gc + (e - nKnots)));
        }
        __syncwarp();
        build_span_records<O>(kn, P.nCoef, rec, lane, 32);
        if (s + step < P.nSplines) request(s + step);            // next curve's data travels under this curve's points
        __syncwarp();
        const double lo = kn[O - 1], hi = kn[P.nCoef];
        double *ov = P.values + s * NDEP * P.nPts;
        double *og = DER ? P.deriv1 + s * NDEP * P.nPts : nullptr;
        auto point = [&](int p, double u, int ix) {
            if (((u < lo) | (u > hi)) && P.firstOutside) report_outside((int64_t *)P.firstOutside, s * P.nPts + p);
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
            double v[NDEP], g[NDEP];
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
#pragma unroll
            for (int d = 0; d < NDEP; ++d) {
                __stcs(ov + d * P.nPts + p, v[d]);
                if (DER) __stcs(og + d * P.nPts + p, g[d]);
            }
        };
#pragma unroll
        for (int k0 = 0; k0 < MANY_PF; k0 += MANY_G) {
            if (lane + 32 * k0 >= P.nPts) break;
            // MANY_G upper-bound bisections side by side (curve_span: number of knots <= u clamped to [O, nCoef])
            int ix[MANY_G], cnt[MANY_G];
#pragma unroll
            for (int j = 0; j < MANY_G; ++j) { ix[j] = O; cnt[j] = (up[k0 + j] != up[k0 + j]) ? 0 : P.nCoef - O; }
            for (bool more = true; more;) {
                more = false;
#pragma unroll
                for (int j = 0; j < MANY_G; ++j) {
                    if (cnt[j] > 0) {
                        const int half = cnt[j] >> 1;
                        const bool le = kn[ix[j] + half] <= up[k0 + j];
                        ix[j] = le ? ix[j] + half + 1 : ix[j];
                        cnt[j] = le ? cnt[j] - half - 1 : half;
                        more |= cnt[j] > 0;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < MANY_G; ++j) {
                const int p = lane + 32 * (k0 + j);
                if (p < P.nPts) point(p, up[k0 + j], (up[k0 + j] != up[k0 + j]) ? P.nCoef : ix[j]);
            }
        }
        for (int p = lane + 32 * MANY_PF; p < P.nPts; p += 32) {   // curves with more than 32 * MANY_PF points
            const double u = __ldcs(gu + p);
            point(p, u, curve_span(kn, O, P.nCoef, u));
        }
    }
}

// ---- values only: per-span polynomial rows ------------------------------------------------------------------------------
// many_kernel is bound by the load-return path of the SM (ncu: l1tex 94 % busy): per point 5 bisection probes, an 80-byte
// span record and a 96-byte coefficient window gathered from the lanes' 32 different spans.  For value-only requests the
// lanes instead build, once per curve, one row per span with the span's polynomial in powers of (u - mid-span):
//     c_k[d] = S^(k)(m)[d] / k!  (Cox-de Boor with k derivative stages at m, from the same reciprocal-gap record, no extra
//     divisions), row = { c_0[d], c_1[d], .., c_{O-1}[d], m }
// and a point is: bisection, ONE aligned row of O*nDep+1 doubles (7 LDS.128 for the cubic 3-D curve instead of 5 LDS.128 +
// 12 LDS.64), t = u - m, Horner.  Centred at mid-span |t| <= h/2 and every gap in the derivative stages contains the span
// itself, so |c_k t^k| stays of the order of the coefficients: the evaluation error is a few eps * max|coef| like the
// recurrence's (left-knot Taylor rows, whose terms are 2^k larger, were measured at 0.08x of the tolerance on values).
// Spans of zero width give inf / NaN rows like the recurrence.  Derivative requests keep the Cox-de Boor rows.
template <int O, int NDEP>
__global__ void __launch_bounds__(256, 5) many_poly_kernel(const ManyParams P)
{
    using R = SpanRec<O>;
    constexpr int RS = (O * NDEP + 2) & ~1;                       // doubles per row (even: 16-byte aligned rows)
    extern __shared__ __align__(16) double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const int nKnots = O + P.nCoef, spans = P.nCoef - O + 1;
    double *kn = sm + (long long)warp * P.slice;                 // knots
    double *rows = kn + ((nKnots + 1) & ~1);                     // spans * RS
    for (long long s = (long long)blockIdx.x * warpsPerBlock + warp; s < P.nSplines; s += (long long)gridDim.x * warpsPerBlock) {
        const double *gk = P.knots + s * P.knotStride;
        const double *gc = P.coefs + s * P.coefStride;
        __syncwarp();
        for (int i = lane; i < nKnots; i += 32) kn[i] = __ldcs(gk + i);
        __syncwarp();
        for (int sp = lane; sp < spans; sp += 32) {
            const int ix = O + sp;
            double left[O > 1 ? O - 1 : 1], rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
            for (int j = 0; j < O - 1; ++j) left[j] = kn[ix - (O - 1) + j];
            int at = 0;
#pragma unroll
            for (int deg = 1; deg < O; ++deg)
#pragma unroll
                for (int t = 0; t < deg; ++t) rc[at++] = 1.0 / (kn[ix + t] - kn[ix - deg + t]);
            const double m = 0.5 * (kn[ix - 1] + kn[ix]);
            double dl[O > 1 ? O - 1 : 1];
#pragma unroll
            for (int j = 0; j < O - 1; ++j) dl[j] = m - left[j];
            double *row = rows + sp * RS;
            double invFact = 1.0;
#pragma unroll
            for (int k = 0; k < O; ++k) {
                double bk[O], unused[O];
                basis_core<O, false>(dl, rc, k, bk, unused);
                if (k > 1) invFact /= (double)k;
                // the window coefficients are re-read per derivative order (L1 hits, once per curve) to keep the
                // register footprint of this build phase inside the point loop's
#pragma unroll 1
                for (int d = 0; d < NDEP; ++d) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < O; ++j) acc = fma(__ldg(gc + d * P.nCoef + sp + j), bk[j], acc);
                    row[k * NDEP + d] = acc * invFact;
                }
            }
            row[O * NDEP] = m;
            if (RS > O * NDEP + 1) row[O * NDEP + 1] = 0.0;
        }
        __syncwarp();
        const double lo = kn[O - 1], hi = kn[P.nCoef];
        const double *gu = P.u + s * P.nPts;
        double *ov = P.values + s * NDEP * P.nPts;
#pragma unroll 2
        for (int p = lane; p < P.nPts; p += 32) {
            const double u = __ldcs(gu + p);
            if (((u < lo) | (u > hi)) && P.firstOutside) report_outside((int64_t *)P.firstOutside, s * P.nPts + p);
            const int ix = curve_span(kn, O, P.nCoef, u);
            double r[RS];
            const double2 *rp = reinterpret_cast<const double2 *>(rows + (ix - O) * RS);
#pragma unroll
            for (int j = 0; j < RS / 2; ++j) {
                const double2 x = rp[j];
                r[2 * j] = x.x;
                r[2 * j + 1] = x.y;
            }
            const double t = u - r[O * NDEP];
#pragma unroll
            for (int d = 0; d < NDEP; ++d) {
                double v = r[(O - 1) * NDEP + d];
#pragma unroll
                for (int k = O - 2; k >= 0; --k) v = fma(v, t, r[k * NDEP + d]);
                __stcs(ov + d * P.nPts + p, v);
            }
        }
    }
}

template <int O, int NDEP>
static int launch_many_poly(ManyParams P, cudaStream_t stream)
{
    const int threads = 256, warps = threads / 32;
    constexpr int RS = (O * NDEP + 2) & ~1;
    P.slice = ((O + P.nCoef + 1) & ~1) + (P.nCoef - O + 1) * RS;
    const size_t smem = (size_t)P.slice * warps * sizeof(double);
    if (smem > 200 * 1024) return -1000;                          // caller falls back to the record kernel
    long long blocks = (P.nSplines + warps - 1) / warps;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (int rc = allow_dynamic_smem(many_poly_kernel<O, NDEP>, smem)) return rc;
    many_poly_kernel<O, NDEP><<<(unsigned)blocks, threads, smem, stream>>>(P);
    count_launch();
    return check_launch("bspy_cuda_eval_many");
}

template <int O, int NDEP, bool DER>
static int launch_many3(const ManyParams &P, cudaStream_t stream)
{
    const int threads = 256, warps = threads / 32;
    const size_t smem = (size_t)P.slice * warps * sizeof(double);
    if (smem > 200 * 1024) { set_error("bspy_cuda_eval_many: curve too large for shared memory"); return BSPY_E_UNSUPPORTED; }
    long long blocks = (P.nSplines + warps - 1) / warps;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(many_kernel<O, NDEP, DER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    if constexpr (NDEP > 0) {
        if (option(OPT_MANY_MODE, 2) == 2) {                          // default: latency-restructured kernel
            if (int rc = allow_dynamic_smem(many2_kernel<O, NDEP, DER>, smem)) return rc;
            many2_kernel<O, NDEP, DER><<<(unsigned)blocks, threads, smem, stream>>>(P);
            count_launch();
            return check_launch("bspy_cuda_eval_many");
        }
    }
    many_kernel<O, NDEP, DER><<<(unsigned)blocks, threads, smem, stream>>>(P);
    count_launch();
    return check_launch("bspy_cuda_eval_many");
}

template <int O>
static int launch_many(const ManyParams &P, cudaStream_t stream)
{
    const bool der = P.deriv1 != nullptr;
    // value-only requests: polynomial rows
    // (MANY_MODE=1; measured on config 3: 71.7 Gpts/s against 74.3 for the record rows -- the kernel is latency-bound, not
    // bound by the row bytes -- so this stays an opt-in experiment)
    if (!der && O >= 2 && O <= 6 && P.nDep >= 1 && P.nDep <= 3 && option(OPT_MANY_MODE, 2) == 1) {
        int rc = -1000;
        switch (P.nDep) {
            case 1: rc = launch_many_poly<O, 1>(P, stream); break;
            case 2: rc = launch_many_poly<O, 2>(P, stream); break;
            default: rc = launch_many_poly<O, 3>(P, stream); break;
        }
        if (rc != -1000) return rc;
    }
    switch (P.nDep) {
        case 1: return der ? launch_many3<O, 1, true>(P, stream) : launch_many3<O, 1, false>(P, stream);
        case 2: return der ? launch_many3<O, 2, true>(P, stream) : launch_many3<O, 2, false>(P, stream);
        case 3: return der ? launch_many3<O, 3, true>(P, stream) : launch_many3<O, 3, false>(P, stream);
        default: return der ? launch_many3<O, 0, true>(P, stream) : launch_many3<O, 0, false>(P, stream);
    }
}

}  // namespace bspy

using namespace bspy;

extern "C" int bspy_cuda_eval_many(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines, const double *knots,
                                   int64_t knotStride, const double *coefs, int64_t coefStride, const double *u,
                                   int32_t nPts, double *values, double *deriv1, int64_t *firstOutside, void *stream)
{
    if (!knots || !coefs || !u || !values || order < 1 || nCoef < order || nDep < 1 || nSplines < 0 || nPts < 0) {
        set_error("bspy_cuda_eval_many: bad argument");
        return BSPY_E_ARG;
    }
    if (nSplines == 0 || nPts == 0) return 0;
    if (order > 8) {
        set_error("bspy_cuda_eval_many: order %d not supported (1..8)", order);
        return BSPY_E_UNSUPPORTED;
    }
    ManyParams P{};
    P.nCoef = nCoef; P.nDep = nDep; P.nPts = nPts; P.nSplines = nSplines;
    P.knots = knots; P.coefs = coefs; P.u = u;
    P.knotStride = knotStride; P.coefStride = coefStride;
    P.values = values; P.deriv1 = deriv1;
    P.firstOutside = (long long *)firstOutside;
    const int recStride = ((order - 1 + order * (order - 1) / 2) + 1) & ~1;
    P.slice = ((order + nCoef + 1) & ~1) + (nCoef - order + 1) * recStride + ((nDep * nCoef + 1) & ~1);
    cudaStream_t st = (cudaStream_t)stream;
    switch (order) {
        case 1: return launch_many<1>(P, st);
        case 2: return launch_many<2>(P, st);
        case 3: return launch_many<3>(P, st);
        case 4: return launch_many<4>(P, st);
        case 5: return launch_many<5>(P, st);
        case 6: return launch_many<6>(P, st);
        case 7: return launch_many<7>(P, st);
        default: return launch_many<8>(P, st);
    }
}
