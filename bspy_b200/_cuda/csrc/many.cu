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

// ---- cached per-curve images: polynomial rows built once per batch, fetched by TMA ---------------------------------------
// many_kernel spends a fifth of its instructions rebuilding every curve's tables for 256 points and is bound by the conflicted
// gathers of an 80-byte record and a 96-byte window per point (ncu: l1tex 94 %).  A batch of curves is a resident object
// (SplineBatch), so -- like the tables of a single curve (curve.cu) -- everything that depends on the curves only is built ONCE
// per batch into a caller-owned image, one per curve:
//     header { scale of the bucket function, flag: rows validated } | knots | bucket table of the span search | one row per span
//     with the span's polynomial in powers of (u - mid-span): { c_0[d], .., c_{O-1}[d], m }
// (values, or values + first derivative by Horner's rule with derivative; the rows are validated against the recurrence for
// both when they are built, exactly as in curve.cu).  The
// evaluation kernel is then, per curve and warp: one bulk asynchronous copy of the image into the warp's slice of shared
// memory (cp.async.bulk + mbarrier, double-buffered: the next curve's image and parameters travel under this curve's points),
// and per point a bucket look-up + short advance (bit-exact span), ONE row of O*nDep+1 doubles and a Horner evaluation.
// Costs (O*nDep+2)*8 bytes per span of memory and HBM traffic per curve instead of the raw knots and coefficients.
struct ManyTableLayout {
    int buckets, knotDoubles, tabBytes, rowDoubles, spans;
    long long image;        // bytes per curve, a multiple of 16; 0: shape without tables
};

static ManyTableLayout many_table_layout(int O, int nCoef, int nDep)
{
    ManyTableLayout L{};
    if (O < 2 || O > 6 || nDep < 1 || nDep > 3 || nCoef < O || nCoef > 4000) return L;
    L.spans = nCoef - O + 1;
    L.buckets = 32;
    while (L.buckets < 2 * L.spans && L.buckets < 1024) L.buckets <<= 1;
    L.knotDoubles = (O + nCoef + 1) & ~1;
    L.tabBytes = (2 * L.buckets + 15) & ~15;
    L.rowDoubles = (O * nDep + 2) & ~1;
    L.image = 16 + 8LL * L.knotDoubles + L.tabBytes + 8LL * L.spans * L.rowDoubles;
    if (L.image > 12 * 1024) L.image = 0;                         // two images per warp must fit beside three more CTAs
    return L;
}

__device__ __forceinline__ int many_bucket_of(double x, double lo, double scale, int lastBucket)
{
    return min(max(__double2int_rz((x - lo) * scale), 0), lastBucket);   // monotone in x; NaN -> 0
}

struct ManyTableParams {
    int nCoef;
    long long nSplines;
    const double *knots, *coefs;
    long long knotStride, coefStride;
    unsigned char *table;
    ManyTableLayout L;
};

template <int O, int NDEP>
__global__ void __launch_bounds__(256) many_table_kernel(const ManyTableParams P)
{
    extern __shared__ __align__(16) double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warpsPerBlock = blockDim.x >> 5;
    const ManyTableLayout &L = P.L;
    const int nKnots = O + P.nCoef, spans = L.spans, buckets = L.buckets;
    constexpr int PROW = (O * NDEP + 2) & ~1;
    double *kn = sm + (long long)warp * (L.knotDoubles + (buckets + 2) / 2 + 1);
    int *cnt = reinterpret_cast<int *>(kn + L.knotDoubles);      // buckets + 1 counters
    for (long long s = (long long)blockIdx.x * warpsPerBlock + warp; s < P.nSplines; s += (long long)gridDim.x * warpsPerBlock) {
        const double *gk = P.knots + s * P.knotStride, *gc = P.coefs + s * P.coefStride;
        unsigned char *img = P.table + s * L.image;
        double *iKn = reinterpret_cast<double *>(img + 16);
        unsigned short *iTab = reinterpret_cast<unsigned short *>(img + 16 + 8 * L.knotDoubles);
        double *iRows = reinterpret_cast<double *>(img + 16 + 8 * L.knotDoubles + L.tabBytes);
        __syncwarp();
        for (int i = lane; i < L.knotDoubles; i += 32) {
            const double k = i < nKnots ? __ldg(gk + i) : 0.0;
            kn[i] = k;
            iKn[i] = k;
        }
        for (int i = lane; i <= buckets; i += 32) cnt[i] = 0;
        __syncwarp();
        const double lo = kn[O - 1], hi = kn[P.nCoef];
        const double scale = (double)buckets / (hi - lo);
        for (int i = O + lane; i < P.nCoef; i += 32) atomicAdd(cnt + many_bucket_of(kn[i], lo, scale, buckets - 1) + 1, 1);
        __syncwarp();
        // tab[b] = O + number of interior knots in buckets < b
        {
            int run = 0;
            for (int base = 0; base < buckets; base += 32) {
                const int c = cnt[base + lane];
                int incl = c;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += v;
                }
                iTab[base + lane] = (unsigned short)(O + run + incl);
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        bool ok = true, okDer = true;
        for (int sp = lane; sp < spans; sp += 32) {
            const int ix = O + sp;
            double left[O - 1], rc[O * (O - 1) / 2];
#pragma unroll
            for (int j = 0; j < O - 1; ++j) left[j] = kn[ix - (O - 1) + j];
            int at = 0;
#pragma unroll
            for (int deg = 1; deg < O; ++deg)
#pragma unroll
                for (int t = 0; t < deg; ++t) rc[at++] = 1.0 / (kn[ix + t] - kn[ix - deg + t]);
            const double k0 = kn[ix - 1], k1 = kn[ix], m = 0.5 * (k0 + k1), h = k1 - k0;
            double cw[O][NDEP];
#pragma unroll
            for (int j = 0; j < O; ++j)
#pragma unroll
                for (int d = 0; d < NDEP; ++d) cw[j][d] = __ldg(gc + d * P.nCoef + sp + j);
            double dl[O - 1];
#pragma unroll
            for (int j = 0; j < O - 1; ++j) dl[j] = m - left[j];
            double row[PROW];
            double invFact = 1.0;
#pragma unroll
            for (int k = 0; k < O; ++k) {
                double bk[O], unused[O];
                basis_core<O, false>(dl, rc, k, bk, unused);
                if (k > 1) invFact /= (double)k;
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < O; ++j) acc = fma(cw[j][d], bk[j], acc);
                    row[k * NDEP + d] = acc * invFact;
                }
            }
            row[O * NDEP] = m;
            if (PROW > O * NDEP + 1) row[O * NDEP + 1] = 0.0;
#pragma unroll
            for (int j = 0; j < PROW / 2; ++j)
                *reinterpret_cast<double2 *>(iRows + sp * PROW + 2 * j) = make_double2(row[2 * j], row[2 * j + 1]);
            // validation (see curve.cu): conditioning of the Horner terms and agreement with the recurrence at nine parameters
            if (h > 0.0) {
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    double cmax = 0.0, terms = 0.0, tk = 1.0;
#pragma unroll
                    for (int j = 0; j < O; ++j) cmax = fmax(cmax, fabs(cw[j][d]));
#pragma unroll
                    for (int k = 0; k < O; ++k) { terms = fma(fabs(row[k * NDEP + d]), tk, terms); tk *= 0.5 * h; }
                    if (!(terms <= 16.0 * cmax)) ok = false;
                    for (int sidx = 0; sidx <= 8; ++sidx) {
                        const double u = sidx == 8 ? k1 : k0 + h * (0.125 * sidx);
                        double du[O - 1], b0[O], b1[O], unused[O];
#pragma unroll
                        for (int j = 0; j < O - 1; ++j) du[j] = u - left[j];
                        basis_core<O, false>(du, rc, 0, b0, unused);
                        basis_core<O, false>(du, rc, 1, b1, unused);
                        double ref = 0.0, dref = 0.0, dterms = 0.0;
#pragma unroll
                        for (int j = 0; j < O; ++j) {
                            ref = fma(cw[j][d], b0[j], ref);
                            dref = fma(cw[j][d], b1[j], dref);
                            dterms = fma(fabs(cw[j][d]), fabs(b1[j]), dterms);
                        }
                        const double t = u - m;
                        double hv = row[(O - 1) * NDEP + d], dv = 0.0;
#pragma unroll
                        for (int k = O - 2; k >= 0; --k) {
                            dv = (k == O - 2) ? hv : fma(dv, t, hv);
                            hv = fma(hv, t, row[k * NDEP + d]);
                        }
                        if (!(fabs(hv - ref) <= 0.25 * (1e-13 + 1e-12 * fabs(ref)) + 8.9e-16 * cmax)) ok = false;
                        if (!(fabs(dv - dref) <= 0.25 * (1e-13 + 1e-12 * fabs(dref)) + 8.9e-16 * dterms)) okDer = false;
                    }
                }
            } else if (sp == 0 || sp == spans - 1 || !(h == 0.0)) {
                ok = false;
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        okDer = __all_sync(0xffffffffu, ok && okDer);
        if (lane == 0) {
            *reinterpret_cast<double *>(img) = scale;
            reinterpret_cast<int *>(img)[2] = ok ? 1 : 0;          // rows validated for values ...
            reinterpret_cast<int *>(img)[3] = okDer ? 1 : 0;       // ... and for values + first derivative
        }
    }
}

struct ManyTabParams {
    ManyParams P;
    const unsigned char *table;
    ManyTableLayout L;
};

constexpr int MANY_TAB_WARPS = 8;

template <int O, int NDEP, bool DER>
__global__ void __launch_bounds__(MANY_TAB_WARPS * 32, 3) many_tab_kernel(const ManyTabParams Q)
{
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long mbar[MANY_TAB_WARPS][2];
    const ManyParams &P = Q.P;
    const ManyTableLayout &L = Q.L;
    constexpr int PROW = (O * NDEP + 2) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *slot = smraw + (size_t)warp * 2 * L.image;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&mbar[warp][0]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long step = (long long)gridDim.x * MANY_TAB_WARPS;
    long long s = (long long)blockIdx.x * MANY_TAB_WARPS + warp;
    auto fetch = [&](long long curve, int b) {   // lane 0: arm the barrier of buffer b and start the bulk copy of the curve's image
        const unsigned bar = bar0 + 8u * b;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)L.image) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(slot + (size_t)b * L.image)),
                     "l"(Q.table + curve * L.image), "r"((unsigned)L.image), "r"(bar)
                     : "memory");
    };
    if (s < P.nSplines && lane == 0) fetch(s, 0);
    unsigned phase0 = 0, phase1 = 0;
    int b = 0;
    for (; s < P.nSplines; s += step, b ^= 1) {
        const double *gu = P.u + s * P.nPts;
        // this curve's parameters leave HBM before its image is awaited
        double up[MANY_PF];
#pragma unroll
        for (int k = 0; k < MANY_PF; ++k) up[k] = lane + 32 * k < P.nPts ? __ldcs(gu + lane + 32 * k) : 0.0;
        // the next curve's image goes to the other buffer: every lane is done with it (previous iteration) by this barrier
        __syncwarp();
        if (s + step < P.nSplines && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fetch(s + step, b ^ 1);
        }
        {
            const unsigned bar = bar0 + 8u * b, parity = b ? phase1 : phase0;
            unsigned done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            if (b) phase1 ^= 1; else phase0 ^= 1;
        }
        const unsigned char *img = slot + (size_t)b * L.image;
        const double scale = *reinterpret_cast<const double *>(img);
        const bool valid = reinterpret_cast<const int *>(img)[DER ? 3 : 2] != 0;
        const double *kn = reinterpret_cast<const double *>(img + 16);
        const unsigned short *tab = reinterpret_cast<const unsigned short *>(img + 16 + 8 * L.knotDoubles);
        const double *rows = reinterpret_cast<const double *>(img + 16 + 8 * L.knotDoubles + L.tabBytes) - O * PROW;
        const double lo = kn[O - 1], hi = kn[P.nCoef];
        double *ov = P.values + s * NDEP * P.nPts;
        double *og = DER ? P.deriv1 + s * NDEP * P.nPts : nullptr;
        auto point = [&](int p, double u) {
            if (((u < lo) | (u > hi)) && P.firstOutside) report_outside((int64_t *)P.firstOutside, s * P.nPts + p);
            int ix = tab[many_bucket_of(u, lo, scale, L.buckets - 1)];
            while (ix < P.nCoef && kn[ix] <= u) ++ix;
            if (u != u) ix = P.nCoef;
            double v[NDEP], g[NDEP];
            if (valid) {
                double r[PROW];
                const double2 *rp = reinterpret_cast<const double2 *>(rows + ix * PROW);
#pragma unroll
                for (int j = 0; j < PROW / 2; ++j) {
                    const double2 x = rp[j];
                    r[2 * j] = x.x;
                    r[2 * j + 1] = x.y;
                }
                const double t = u - r[O * NDEP];
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    double h = r[(O - 1) * NDEP + d], dh = 0.0;
#pragma unroll
                    for (int k = O - 2; k >= 0; --k) {
                        if (DER) dh = (k == O - 2) ? h : fma(dh, t, h);
                        h = fma(h, t, r[k * NDEP + d]);
                    }
                    v[d] = h;
                    g[d] = dh;
                }
            } else {
                // rows that failed their validation (rare): the recurrence on the curve's own knots and coefficients
                double kw[2 * (O - 1)], b0[O], b1[O];
#pragma unroll
                for (int j = 0; j < 2 * (O - 1); ++j) kw[j] = kn[ix - (O - 1) + j];
                basis_regs<O, DER>(kw, u, 0, b0, b1);
                const double *gc = P.coefs + s * P.coefStride;
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    double acc = 0.0, dacc = 0.0;
#pragma unroll
                    for (int j = 0; j < O; ++j) {
                        const double cj = __ldg(gc + d * P.nCoef + ix - O + j);
                        acc = fma(cj, b0[j], acc);
                        if (DER) dacc = fma(cj, b1[j], dacc);
                    }
                    v[d] = acc;
                    g[d] = dacc;
                }
            }
#pragma unroll
            for (int d = 0; d < NDEP; ++d) {
                __stcs(ov + d * P.nPts + p, v[d]);
                if (DER) __stcs(og + d * P.nPts + p, g[d]);
            }
        };
#pragma unroll
        for (int k = 0; k < MANY_PF; ++k)
            if (lane + 32 * k < P.nPts) point(lane + 32 * k, up[k]);
        for (int p = lane + 32 * MANY_PF; p < P.nPts; p += 32) point(p, __ldcs(gu + p));
    }
}

template <int O, int NDEP>
static int launch_many_table(const ManyTableParams &T, cudaStream_t stream)
{
    const int warps = 8;
    const size_t smem = (size_t)warps * 8 * (T.L.knotDoubles + (T.L.buckets + 2) / 2 + 1);
    if (int rc = allow_dynamic_smem(many_table_kernel<O, NDEP>, smem)) return rc;
    long long blocks = (T.nSplines + warps - 1) / warps;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    many_table_kernel<O, NDEP><<<(unsigned)blocks, warps * 32, smem, stream>>>(T);
    count_launch();
    return check_launch("bspy_cuda_many_table_build");
}

template <int O, int NDEP, bool DER>
static int launch_many_tab2(const ManyTabParams &Q, cudaStream_t stream)
{
    const size_t smem = (size_t)MANY_TAB_WARPS * 2 * Q.L.image;
    if (int rc = allow_dynamic_smem(many_tab_kernel<O, NDEP, DER>, smem)) return rc;
    long long blocks = (Q.P.nSplines + MANY_TAB_WARPS - 1) / MANY_TAB_WARPS;
    const long long perSm = (220 * 1024) / (long long)(smem + 1024);
    const long long cap = (long long)num_sms() * (perSm < 1 ? 1 : perSm > 3 ? 3 : perSm);
    if (blocks > cap) blocks = cap;
    many_tab_kernel<O, NDEP, DER><<<(unsigned)blocks, MANY_TAB_WARPS * 32, smem, stream>>>(Q);
    count_launch();
    return check_launch("bspy_cuda_eval_many_tab");
}

template <int O, int NDEP>
static int launch_many_tab(const ManyTabParams &Q, cudaStream_t stream)
{
    return Q.P.deriv1 ? launch_many_tab2<O, NDEP, true>(Q, stream) : launch_many_tab2<O, NDEP, false>(Q, stream);
}

#define BSPY_MANY_TAB_DISPATCH(FN, ARG)                                                             \
    switch (order * 10 + nDep) {                                                                    \
        case 21: return FN<2, 1>(ARG, st); case 22: return FN<2, 2>(ARG, st); case 23: return FN<2, 3>(ARG, st); \
        case 31: return FN<3, 1>(ARG, st); case 32: return FN<3, 2>(ARG, st); case 33: return FN<3, 3>(ARG, st); \
        case 41: return FN<4, 1>(ARG, st); case 42: return FN<4, 2>(ARG, st); case 43: return FN<4, 3>(ARG, st); \
        case 51: return FN<5, 1>(ARG, st); case 52: return FN<5, 2>(ARG, st); case 53: return FN<5, 3>(ARG, st); \
        case 61: return FN<6, 1>(ARG, st); case 62: return FN<6, 2>(ARG, st); case 63: return FN<6, 3>(ARG, st); \
        default: break;                                                                             \
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

extern "C" int64_t bspy_cuda_many_table_bytes(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines)
{
    if (nSplines <= 0) return 0;
    const ManyTableLayout L = many_table_layout(order, nCoef, nDep);
    return L.image * nSplines;
}

extern "C" int bspy_cuda_many_table_build(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines, const double *knots,
                                          int64_t knotStride, const double *coefs, int64_t coefStride, void *table,
                                          int64_t tableBytes, void *stream)
{
    const ManyTableLayout L = many_table_layout(order, nCoef, nDep);
    if (!L.image || !knots || !coefs || !table || nSplines <= 0 || tableBytes < L.image * nSplines ||
        (reinterpret_cast<uintptr_t>(table) & 15)) {
        set_error("bspy_cuda_many_table_build: shape without tables, or bad / too small / unaligned buffer");
        return BSPY_E_ARG;
    }
    ManyTableParams T{};
    T.nCoef = nCoef; T.nSplines = nSplines; T.knots = knots; T.coefs = coefs;
    T.knotStride = knotStride; T.coefStride = coefStride; T.table = (unsigned char *)table; T.L = L;
    cudaStream_t st = (cudaStream_t)stream;
    BSPY_MANY_TAB_DISPATCH(launch_many_table, T)
    return BSPY_E_UNSUPPORTED;
}

extern "C" int bspy_cuda_eval_many_tab(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines, const double *knots,
                                       int64_t knotStride, const double *coefs, int64_t coefStride, const void *table,
                                       int64_t tableBytes, const double *u, int32_t nPts, double *values, double *deriv1,
                                       int64_t *firstOutside, void *stream)
{
    const ManyTableLayout L = many_table_layout(order, nCoef, nDep);
    if (!L.image || !knots || !coefs || !table || !u || !values || nSplines < 0 || nPts < 0 || tableBytes != L.image * nSplines ||
        (reinterpret_cast<uintptr_t>(table) & 15)) {
        set_error("bspy_cuda_eval_many_tab: bad argument (the table must be the one bspy_cuda_many_table_build wrote for this batch)");
        return BSPY_E_ARG;
    }
    if (nSplines == 0 || nPts == 0) return 0;
    ManyTabParams Q{};
    Q.P.nCoef = nCoef; Q.P.nDep = nDep; Q.P.nPts = nPts; Q.P.nSplines = nSplines;
    Q.P.knots = knots; Q.P.coefs = coefs; Q.P.u = u;
    Q.P.knotStride = knotStride; Q.P.coefStride = coefStride;
    Q.P.values = values; Q.P.deriv1 = deriv1;
    Q.P.firstOutside = (long long *)firstOutside;
    Q.table = (const unsigned char *)table; Q.L = L;
    cudaStream_t st = (cudaStream_t)stream;
    BSPY_MANY_TAB_DISPATCH(launch_many_tab, Q)
    return BSPY_E_UNSUPPORTED;
}
