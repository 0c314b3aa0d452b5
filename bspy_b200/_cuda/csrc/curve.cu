// Scattered points on ONE curve (nInd == 1): the lean path of bspy_cuda_eval_points.
// Every CTA stages the curve (knots, per-span records, interleaved coefficients; a few KB) in shared memory
// once and then streams points: 8 B in, 8*nDep B out per point, no divisions and no global gathers in the loop.
#include <stdlib.h>

#include "curve.cuh"

namespace bspy {

struct CurveParams {
    const double *knots, *coefs;
    int nCoef;
    int normalSign;
    PointsDev in;
    OutDev out;
    long long N;
    int derivAsValue;   // derivative([1], u): write the first derivative into out.values
    const void *table;  // caller-owned image of the replicated tables (bspy_cuda_curve_table_build), or nullptr
};

// results of one point: spans, values (or the first derivative when derivative([1], u) was asked for), jacobian,
// normal of a planar curve
template <int NDEP, bool DER>
__device__ __forceinline__ void store_curve_point(const CurveParams &P, const long long p, const int ix, const double (&v)[NDEP],
                                                  const double (&g)[NDEP])
{
    const OutDev &out = P.out;
    if (out.spans) __stcs(out.spans + p, ix);
    if (out.values) {
#pragma unroll
        for (int d = 0; d < NDEP; ++d) __stcs(out.values + d * out.ld + p, (DER && P.derivAsValue) ? g[d] : v[d]);
    }
    if constexpr (DER) {
        if (out.jacobian) {
#pragma unroll
            for (int d = 0; d < NDEP; ++d) __stcs(out.jacobian + d * out.ld + p, g[d]);
        }
        if constexpr (NDEP == 2) {
            if (out.normal) {
                // planar curve: T = J is 2x1, n = sign * (t_y, -t_x)
                double n0 = g[1] * P.normalSign, n1 = -g[0] * P.normalSign;
                if (out.normalize) {
                    double sq = 0.0;
                    if (out.normalMask & 1u) sq = fma(n0, n0, sq);
                    if (out.normalMask & 2u) sq = fma(n1, n1, sq);
                    const double len = sqrt(sq);
                    n0 = n0 / len;
                    n1 = n1 / len;
                }
                __stcs(out.normal + p, n0);
                __stcs(out.normal + out.ld + p, n1);
            }
        }
    }
}

template <int O, int NDEP, bool DER>
__global__ void __launch_bounds__(256) eval_curve_kernel(const CurveParams P)
{
    using R = SpanRec<O>;
    extern __shared__ double sm[];
    const int nKnots = O + P.nCoef;
    double *kn = sm;
    double *rec = kn + ((nKnots + 1) & ~1);
    double *cf = rec + (P.nCoef - O + 1) * R::stride;
    for (int i = threadIdx.x; i < nKnots; i += blockDim.x) kn[i] = __ldg(P.knots + i);
    for (int i = threadIdx.x; i < NDEP * P.nCoef; i += blockDim.x) {
        const int d = i / P.nCoef, c = i - d * P.nCoef;
        cf[c * NDEP + d] = __ldg(P.coefs + i);
    }
    __syncthreads();
    build_span_records<O>(kn, P.nCoef, rec, threadIdx.x, blockDim.x);
    __syncthreads();
    const double lo = kn[O - 1], hi = kn[P.nCoef];
    const OutDev &out = P.out;
#pragma unroll 2
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P.N; p += (long long)gridDim.x * blockDim.x) {
        long long rem = p;
        const double u = fetch_param(P.in, p, 0, rem);
        if (((u < lo) | (u > hi)) && out.firstOutside) report_outside((int64_t *)out.firstOutside, p);
        double v[NDEP], g[NDEP];
        const int ix = curve_point<O, NDEP, DER>(kn, rec, cf, P.nCoef, u, v, g);
        store_curve_point<NDEP, DER>(P, p, ix, v, g);
    }
}

// ---- bank-replicated span rows (big batches) ------------------------------------------------------------------
// The kernel above is bound by shared-memory wavefronts, not by HBM: per point it gathers 6 bisection probes, an
// 80-byte span record and a 96-byte coefficient window (cubic, nDep 3) from random places; lanes collide on the banks
// (~2.5 wavefronts where one would do), ~140 wavefronts per warp of 32 points = 68 Gpts/s on 148 SMs -- what is
// measured -- against 204 Gpts/s at the HBM roofline.  For big batches each CTA therefore builds
//   * one ROW per span holding everything a point of that span needs (left knots | reciprocal gaps | the O x nDep
//     coefficient window), replicated once per lane position of a quarter warp: chunk j (16 bytes) of span s for
//     copy q lives at 16-byte slot ((s * CH + j) * COPIES + q), so the 8 lanes that share a 128-byte wavefront of an
//     LDS.128 always hit 8 different bank groups -- whatever their spans -- and a row costs CH * 4 wavefronts per warp;
//   * a bucket table over the domain: tab[b] = first span whose knots may still be <= a parameter of bucket b
//     (exactly: order + the number of interior knots whose own bucket is < b, with the same floating-point bucket
//     function, which is monotone -- so the answer is bit-exact), followed by a short advance over the knots of the
//     bucket itself (knots replicated per lane position of a half warp: LDS.64 conflict-free).
// 44 + ~8 wavefronts per warp for the cubic nDep-3 curve instead of ~140.  Arithmetic per point is basis_core<> and the
// same dot products as curve_point<>: bit-identical results.
constexpr int REPL_COPIES = 8, REPL_KNOT_COPIES = 16;   // measured at 1e8 points: 8 / 4 / 2 row copies -> 132 / 87 / 77 Gpts/s
constexpr int REPL_U = 2;
constexpr int repl_threads(int O, int nDep) { return ((O - 1) + O * (O - 1) / 2 + O * nDep) <= 24 ? 512 : 256; }

// doubles of a polynomial row: O x nDep coefficients + the mid-span, rounded up to whole 16-byte chunks
__host__ __device__ constexpr int poly_row_doubles(int O, int nDep) { return (O * nDep + 2) & ~1; }

struct ReplLayout {
    int buckets;         // power of two
    size_t bytes;        // 0: the replicated tables do not fit
};

static ReplLayout repl_layout(int O, int nDep, int nCoef, size_t budget)
{
    ReplLayout L{};
    const int rowDoubles = ((O - 1) + O * (O - 1) / 2 + O * nDep + 1) & ~1;
    const int spans = nCoef - O + 1;
    L.buckets = 64;
    while (L.buckets < 4 * spans && L.buckets < 4096) L.buckets <<= 1;
    L.bytes = sizeof(double) * ((size_t)spans * rowDoubles * REPL_COPIES + (size_t)(O + nCoef) * REPL_KNOT_COPIES +
                               (size_t)nDep * nCoef) +
              sizeof(int) * (L.buckets + 2) + sizeof(unsigned short) * L.buckets + 64;
    if (L.bytes > budget) L.bytes = 0;
    return L;
}

// Measured (ncu, config 1 at 2e7 points): the LSU data pipe is 97 % busy -- 51 shared-memory wavefronts per warp of
// 32 points (44 for the rows, all conflict-free) plus ~22 for the 32 bytes/point of global traffic -- at 132 Gpts/s =
// 65 % of the HBM roofline; 16-byte global accesses (two consecutive points per thread) change nothing, the pipe
// charges global traffic by the byte.
// bucket function of the span search: monotone in x (subtraction and multiplication by a positive constant are, so is the
// saturating conversion); NaN -> 0
__device__ __forceinline__ int repl_bucket_of(double x, double lo, double scale, int lastBucket)
{
    return min(max(__double2int_rz((x - lo) * scale), 0), lastBucket);
}

// Builds the replicated tables of one curve in shared memory: rows (spans * ROW * CP doubles) | knots (nKnots * KC) |
// raw coefficients (scratch) | bucket counters (scratch) | tab (buckets unsigned shorts).  Called by every CTA of
// eval_curve_repl_kernel and once per spline by curve_table_kernel.
template <int O, int NDEP>
__device__ __forceinline__ void build_repl_tables(const double *__restrict__ knots, const double *__restrict__ coefs, const int nCoef,
                                                  const int buckets, double *rows, double *kn, double *raw, int *cnt,
                                                  unsigned short *tab)
{
    using R = SpanRec<O>;
    constexpr int ROW = ((O - 1) + O * (O - 1) / 2 + O * NDEP + 1) & ~1, CH = ROW / 2;
    constexpr int CP = REPL_COPIES, KC = REPL_KNOT_COPIES;
    const int nKnots = O + nCoef, spans = nCoef - O + 1;
    const int lane = threadIdx.x & 31;
    // knots (every copy) and coefficients in one DRAM round trip
    for (int i = threadIdx.x; i < nKnots * KC; i += blockDim.x) kn[i] = __ldg(knots + i / KC);
    for (int i = threadIdx.x; i < NDEP * nCoef; i += blockDim.x) raw[i] = __ldg(coefs + i);
    for (int i = threadIdx.x; i <= buckets; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const double lo = kn[(O - 1) * KC], hi = kn[nCoef * KC];
    const double scale = (double)buckets / (hi - lo);
    const int lastBucket = buckets - 1;
    // rows (copy 0 is written first, then replicated) and the bucket histogram of the interior knots
    for (int sp = threadIdx.x; sp < spans; sp += blockDim.x) {
        const int ix = O + sp;
        double r[ROW];
#pragma unroll
        for (int j = 0; j < O - 1; ++j) r[j] = kn[(ix - (O - 1) + j) * KC];
        int at = O - 1;
#pragma unroll
        for (int deg = 1; deg < O; ++deg)
#pragma unroll
            for (int t = 0; t < deg; ++t) r[at++] = 1.0 / (kn[(ix + t) * KC] - kn[(ix - deg + t) * KC]);
#pragma unroll
        for (int j = 0; j < O; ++j)
#pragma unroll
            for (int d = 0; d < NDEP; ++d) r[R::used + j * NDEP + d] = raw[d * nCoef + (sp + j)];
        if (ROW > R::used + O * NDEP) r[ROW - 1] = 0.0;
#pragma unroll
        for (int j = 0; j < CH; ++j)
            *reinterpret_cast<double2 *>(rows + 2 * ((sp * CH + j) * CP)) = make_double2(r[2 * j], r[2 * j + 1]);
    }
    // interior knots O .. nCoef-1 (the candidates of the span search) into their buckets
    for (int i = O + threadIdx.x; i < nCoef; i += blockDim.x) atomicAdd(cnt + repl_bucket_of(kn[i * KC], lo, scale, lastBucket) + 1, 1);
    __syncthreads();
    // replicate the rows
    {
        const int total = spans * CH * CP;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int q = i & (CP - 1);
            if (q) *reinterpret_cast<double2 *>(rows + 2 * i) = *reinterpret_cast<const double2 *>(rows + 2 * (i - q));
        }
    }
    // tab[b] = O + number of interior knots in buckets < b: inclusive scan of cnt (one warp)
    if (threadIdx.x < 32) {
        int run = 0;
        for (int base = 0; base <= buckets; base += 32) {
            const int i = base + lane;
            const int c = i <= buckets ? cnt[i] : 0;
            int incl = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            if (i < buckets) tab[i] = (unsigned short)(O + run + incl);
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
}

// The point loop of the replicated-row kernels: REPL_U points per thread and round; the parameters of the next round
// are requested before this round's arithmetic (the loop is otherwise a chain DRAM load -> table -> row -> arithmetic ->
// store per point).  un[] holds the first round's parameters, requested by the caller before the tables were ready.
// POLY (validated table image, see curve_table_kernel): the row of a span is its polynomial in powers of (u - mid-span) --
// { c_0[d], .., c_{O-1}[d], m } -- and a point is one row fetch and a Horner evaluation (with its derivative when DER).
template <int O, int NDEP, bool DER, int REPL_U, bool POLY = false>
__device__ __forceinline__ void repl_point_loop(const CurveParams &P, const int buckets, const double *rows, const double *kn,
                                                const unsigned short *tab, double (&un)[REPL_U], const double *up,
                                                const long long pfirst, const long long pstep, const long long rstep)
{
    using R = SpanRec<O>;
    constexpr int ROW = POLY ? poly_row_doubles(O, NDEP) : ((O - 1) + O * (O - 1) / 2 + O * NDEP + 1) & ~1, CH = ROW / 2;
    constexpr int CP = REPL_COPIES, KC = REPL_KNOT_COPIES;
    const int lane = threadIdx.x & 31;
    const long long ustep = pstep * P.in.pointStride, urstep = rstep * P.in.pointStride;
    const double lo = kn[(O - 1) * KC], hi = kn[P.nCoef * KC];
    const double scale = (double)buckets / (hi - lo);
    const int lastBucket = buckets - 1;
    // rows of this lane's copy start at myRows, O rows before span 0 so that the span index addresses them directly
    const double2 *myRows = reinterpret_cast<const double2 *>(rows) + (lane & (CP - 1)) - O * CH * CP;
    const double *myKnots = kn + (lane & (KC - 1));
    const OutDev &out = P.out;
    const bool report = out.firstOutside != nullptr;
    const int nCoef = P.nCoef;
    for (long long p0 = pfirst; p0 < P.N; p0 += rstep) {
        double u[REPL_U];
#pragma unroll
        for (int k = 0; k < REPL_U; ++k) u[k] = un[k];
        up += urstep;
#pragma unroll
        for (int k = 0; k < REPL_U; ++k)
            if (p0 + rstep + k * pstep < P.N) un[k] = __ldcs(up + k * ustep);
#pragma unroll
        for (int k = 0; k < REPL_U; ++k) {
            const long long p = p0 + k * pstep;
            if (p >= P.N) break;
            const double uk = u[k];
            if (report && ((uk < lo) | (uk > hi))) report_outside((int64_t *)out.firstOutside, p);
            int ix = tab[repl_bucket_of(uk, lo, scale, lastBucket)];
            while (ix < nCoef && myKnots[ix * KC] <= uk) ++ix;
            if (uk != uk) ix = nCoef;
            double r[ROW];
            const double2 *rp = myRows + ix * (CH * CP);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const double2 x = rp[j * CP];
                r[2 * j] = x.x;
                r[2 * j + 1] = x.y;
            }
            double v[NDEP], g[NDEP];
            if constexpr (POLY) {
                const double t = uk - r[O * NDEP];
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
                double dl[O > 1 ? O - 1 : 1], rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
                for (int j = 0; j < O - 1; ++j) dl[j] = uk - r[j];
#pragma unroll
                for (int j = 0; j < O * (O - 1) / 2; ++j) rc[j] = r[O - 1 + j];
                double b0[O], b1[O];
                basis_core<O, DER>(dl, rc, 0, b0, b1);
#pragma unroll
                for (int d = 0; d < NDEP; ++d) { v[d] = 0.0; g[d] = 0.0; }
#pragma unroll
                for (int j = 0; j < O; ++j)
#pragma unroll
                    for (int d = 0; d < NDEP; ++d) {
                        const double x = r[R::used + j * NDEP + d];
                        v[d] = fma(x, b0[j], v[d]);
                        if (DER) g[d] = fma(x, b1[j], g[d]);
                    }
            }
            store_curve_point<NDEP, DER>(P, p, ix, v, g);
        }
    }
}

// Measured (ncu, config 1 at 2e7 points): the LSU data pipe is 97 % busy -- 51 shared-memory wavefronts per warp of
// 32 points (44 for the rows, all conflict-free) plus ~22 for the 32 bytes/point of global traffic -- at 132 Gpts/s =
// 65 % of the HBM roofline; 16-byte global accesses (two consecutive points per thread) change nothing, the pipe
// charges global traffic by the byte.
template <int O, int NDEP, bool DER>
__global__ void __launch_bounds__(repl_threads(O, NDEP), 2) eval_curve_repl_kernel(const CurveParams P, const int buckets)
{
    constexpr int ROW = ((O - 1) + O * (O - 1) / 2 + O * NDEP + 1) & ~1;
    constexpr int CP = REPL_COPIES, KC = REPL_KNOT_COPIES;
    extern __shared__ __align__(16) double sm[];
    const int nKnots = O + P.nCoef, spans = P.nCoef - O + 1;
    double *rows = sm;                                             // spans * ROW * CP
    double *kn = rows + spans * ROW * CP;                          // nKnots * KC
    double *raw = kn + nKnots * KC;                                // NDEP * nCoef: the coefficients as they come
    int *cnt = reinterpret_cast<int *>(raw + NDEP * P.nCoef);      // buckets + 1 (scan scratch)
    unsigned short *tab = reinterpret_cast<unsigned short *>(cnt + buckets + 2);
    // the first round of parameters travels from HBM while the tables are built
    const long long stride = (long long)gridDim.x * blockDim.x;          // threads of the grid
    const long long pfirst = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // point of (round, k) = first + (2 * round + k) * stride
    const long long pstep = stride, rstep = REPL_U * stride;
    const double *up = P.in.uvw + pfirst * P.in.pointStride;
    double un[REPL_U];
#pragma unroll
    for (int k = 0; k < REPL_U; ++k) un[k] = pfirst + k * pstep < P.N ? __ldcs(up + k * pstep * P.in.pointStride) : 0.0;
    build_repl_tables<O, NDEP>(P.knots, P.coefs, P.nCoef, buckets, rows, kn, raw, cnt, tab);
    repl_point_loop<O, NDEP, DER, REPL_U>(P, buckets, rows, kn, tab, un, up, pfirst, pstep, rstep);
}

// ---- the tables built ONCE per spline and pulled in by TMA ------------------------------------------------------------
// At the specified size of config 1 (1e6 points, 32 MB of traffic = 4.9 us at the HBM roofline) the kernel above spends
// about half of its 21 us building the 95 KB of replicated tables in each of its 296 CTAs.  The tables depend on the
// spline only, so curve_table_kernel writes them once into a caller-owned image (rows | knots | tab; Python keeps it on the
// DeviceSpline next to the knots and coefficients) and eval_curve_tab_kernel -- one CTA of 1024 threads per SM -- fetches
// the image with bulk asynchronous copies (cp.async.bulk global -> shared, completion on an mbarrier: SASS UBLKCP +
// SYNCS) while its first round of parameters is already travelling from HBM.
constexpr int TAB_THREADS = 1024;

// image = recurrence rows | knots | bucket table | polynomial rows | 16-byte trailer (int: polynomial rows validated).
// Derivative requests fetch the first three parts (cdbBytes), value-only requests the last four (polyFetch bytes from
// offset rowsBytes) and fall back to the first three when the trailer says the polynomial rows did not pass.
struct TableLayout {
    int buckets;
    size_t rowsBytes, knotBytes, tabBytes, polyBytes, cdbBytes, polyFetch, bytes;   // bytes: the whole image, a multiple of 16
};

static TableLayout table_layout(int O, int nDep, int nCoef)
{
    TableLayout T{};
    const ReplLayout L = repl_layout(O, nDep, nCoef, 200 * 1024);
    if (!L.bytes) return T;
    const int rowDoubles = ((O - 1) + O * (O - 1) / 2 + O * nDep + 1) & ~1;
    T.buckets = L.buckets;
    T.rowsBytes = sizeof(double) * (size_t)(nCoef - O + 1) * rowDoubles * REPL_COPIES;
    T.knotBytes = sizeof(double) * (size_t)(O + nCoef) * REPL_KNOT_COPIES;
    T.tabBytes = (sizeof(unsigned short) * (size_t)L.buckets + 15) & ~(size_t)15;
    T.polyBytes = sizeof(double) * (size_t)(nCoef - O + 1) * poly_row_doubles(O, nDep) * REPL_COPIES;
    T.cdbBytes = T.rowsBytes + T.knotBytes + T.tabBytes;
    T.polyFetch = T.knotBytes + T.tabBytes + T.polyBytes + 16;
    T.bytes = T.cdbBytes + T.polyBytes + 16;
    return T;
}

template <int O, int NDEP>
__global__ void __launch_bounds__(512) curve_table_kernel(const double *__restrict__ knots, const double *__restrict__ coefs,
                                                          const int nCoef, const int buckets, unsigned char *__restrict__ image,
                                                          const TableLayout T)
{
    constexpr int ROW = ((O - 1) + O * (O - 1) / 2 + O * NDEP + 1) & ~1;
    constexpr int CP = REPL_COPIES, KC = REPL_KNOT_COPIES;
    extern __shared__ __align__(16) double sm[];
    const int nKnots = O + nCoef, spans = nCoef - O + 1;
    double *rows = sm;
    double *kn = rows + spans * ROW * CP;
    double *raw = kn + nKnots * KC;
    int *cnt = reinterpret_cast<int *>(raw + NDEP * nCoef);
    unsigned short *tab = reinterpret_cast<unsigned short *>(cnt + buckets + 2);
    build_repl_tables<O, NDEP>(knots, coefs, nCoef, buckets, rows, kn, raw, cnt, tab);
    double *gRows = reinterpret_cast<double *>(image);
    double *gKn = reinterpret_cast<double *>(image + T.rowsBytes);
    unsigned short *gTab = reinterpret_cast<unsigned short *>(image + T.rowsBytes + T.knotBytes);
    for (int i = threadIdx.x; i < spans * ROW * CP; i += blockDim.x) gRows[i] = rows[i];
    for (int i = threadIdx.x; i < nKnots * KC; i += blockDim.x) gKn[i] = kn[i];
    for (int i = threadIdx.x; i < (int)(T.tabBytes / 2); i += blockDim.x) gTab[i] = i < buckets ? tab[i] : (unsigned short)0;
    // polynomial rows: c_k[d] = S^(k)(m)[d] / k! at the mid-span m (Cox-de Boor with k derivative stages, from the span's own
    // reciprocal gaps), and their validation, span by span:
    //   * conditioning: sum_k |c_k| (h/2)^k <= 16 max_j |coef_j| over the span's window -- the Horner terms are then at most
    //     16 times the terms of the recurrence's convex combination, and so is the rounding error (a few eps times the terms
    //     either way: ~1e-14 for coefficients of order one).  Every knot gap of the derivative stages contains the span itself,
    //     so the ratio is bounded by ~3^(O-1) whatever the knots; cubics pass, rough data of order 5 and up may not;
    //   * agreement: at nine parameters of the span the Horner value matches the recurrence to a quarter of the parity bar
    //     |x - ref| <= 1e-13 + 1e-12 |ref|, widened by 4 eps max|coef| (where the value cancels, both forms carry that much);
    //   * an empty first / last span (the recurrence gives inf / NaN there, which a row would not reproduce entry for entry).
    // One failing span clears the flag and value-only requests keep the recurrence rows.
    constexpr int PROW = poly_row_doubles(O, NDEP), PCH = PROW / 2;
    double *gPoly = reinterpret_cast<double *>(image + T.cdbBytes);
    __shared__ int polyOk, polyDerOk;
    if (threadIdx.x == 0) { polyOk = 1; polyDerOk = 1; }
    __syncthreads();
    for (int sp = threadIdx.x; sp < spans; sp += blockDim.x) {
        const int ix = O + sp;
        double left[O > 1 ? O - 1 : 1], rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
        for (int j = 0; j < O - 1; ++j) left[j] = kn[(ix - (O - 1) + j) * KC];
        int at = 0;
#pragma unroll
        for (int deg = 1; deg < O; ++deg)
#pragma unroll
            for (int t = 0; t < deg; ++t) rc[at++] = 1.0 / (kn[(ix + t) * KC] - kn[(ix - deg + t) * KC]);
        const double k0 = kn[(ix - 1) * KC], k1 = kn[ix * KC], m = 0.5 * (k0 + k1), h = k1 - k0;
        double dl[O > 1 ? O - 1 : 1];
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
                for (int j = 0; j < O; ++j) acc = fma(raw[d * nCoef + sp + j], bk[j], acc);
                row[k * NDEP + d] = acc * invFact;
            }
        }
        row[O * NDEP] = m;
        if (PROW > O * NDEP + 1) row[O * NDEP + 1] = 0.0;
#pragma unroll
        for (int j = 0; j < PCH; ++j)
#pragma unroll
            for (int q = 0; q < CP; ++q)
                *reinterpret_cast<double2 *>(gPoly + 2 * ((sp * PCH + j) * CP + q)) = make_double2(row[2 * j], row[2 * j + 1]);
        bool ok = true, okDer = true;
        double cmax[NDEP];
#pragma unroll
        for (int d = 0; d < NDEP; ++d) {
            cmax[d] = 0.0;
#pragma unroll
            for (int j = 0; j < O; ++j) cmax[d] = fmax(cmax[d], fabs(raw[d * nCoef + sp + j]));
        }
        if (h > 0.0) {
#pragma unroll
            for (int d = 0; d < NDEP; ++d) {
                double terms = 0.0, tk = 1.0;
#pragma unroll
                for (int k = 0; k < O; ++k) { terms = fma(fabs(row[k * NDEP + d]), tk, terms); tk *= 0.5 * h; }
                if (!(terms <= 16.0 * cmax[d])) ok = false;
            }
            for (int sidx = 0; sidx <= 8; ++sidx) {
                const double u = sidx == 8 ? k1 : k0 + h * (0.125 * sidx);
                double du[O > 1 ? O - 1 : 1], b0[O], unused[O];
#pragma unroll
                for (int j = 0; j < O - 1; ++j) du[j] = u - left[j];
                basis_core<O, false>(du, rc, 0, b0, unused);
                const double t = u - m;
                double b1[O];
                basis_core<O, false>(du, rc, 1, b1, unused);        // first-derivative basis of the recurrence
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    double ref = 0.0, dref = 0.0, dterms = 0.0;
#pragma unroll
                    for (int j = 0; j < O; ++j) {
                        const double cj = raw[d * nCoef + sp + j];
                        ref = fma(cj, b0[j], ref);
                        dref = fma(cj, b1[j], dref);
                        dterms = fma(fabs(cj), fabs(b1[j]), dterms);
                    }
                    double hv = row[(O - 1) * NDEP + d], dv = 0.0;
#pragma unroll
                    for (int k = O - 2; k >= 0; --k) {
                        dv = (k == O - 2) ? hv : fma(dv, t, hv);
                        hv = fma(hv, t, row[k * NDEP + d]);
                    }
                    if (!(fabs(hv - ref) <= 0.25 * (1e-13 + 1e-12 * fabs(ref)) + 8.9e-16 * cmax[d])) ok = false;   // NaN / inf fail too
                    // derivative: the same bar, widened by 4 eps times the terms the recurrence itself sums
                    if (!(fabs(dv - dref) <= 0.25 * (1e-13 + 1e-12 * fabs(dref)) + 8.9e-16 * dterms)) okDer = false;
                }
            }
        } else if (sp == 0 || sp == spans - 1 || !(h == 0.0)) {
            ok = false;                                   // empty end span, or knots that are not ascending / not finite
        }
        if (!ok) polyOk = 0;
        if (!ok || !okDer) polyDerOk = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int *flag = reinterpret_cast<int *>(image + T.cdbBytes + T.polyBytes);
        flag[0] = polyOk; flag[1] = polyDerOk; flag[2] = 0; flag[3] = 0;   // values / values + first derivative
    }
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// U = points per thread and round: 2 for long streams (the next round's parameters travel under this round's
// arithmetic); 8 when the whole batch is a round or two -- config 1 as specified: 6.6 points per thread, all of them
// requested before the image has even arrived
// one elected thread arms the barrier with the byte count and issues the bulk copies (64 KB pieces)
__device__ __forceinline__ void tma_fetch(unsigned char *dst, const unsigned char *src, const size_t bytes, const unsigned bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)bytes) : "memory");
    for (size_t at = 0; at < bytes; at += 65536) {
        const unsigned n = (unsigned)(bytes - at < 65536 ? bytes - at : 65536);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + at)),
                     "l"(src + at), "r"(n), "r"(bar)
                     : "memory");
    }
}

__device__ __forceinline__ void mbar_wait(const unsigned bar, const unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <int O, int NDEP, bool DER, int U = REPL_U, bool POLY = false>
__global__ void __launch_bounds__(TAB_THREADS, 1) eval_curve_tab_kernel(const CurveParams P, const TableLayout T)
{
    extern __shared__ __align__(128) unsigned char image[];
    __shared__ __align__(8) unsigned long long mbar;
    const unsigned bar = smem_u32(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char *src = reinterpret_cast<const unsigned char *>(P.table);
    if (threadIdx.x == 0) {
        if constexpr (POLY) tma_fetch(image, src + T.rowsBytes, T.polyFetch, bar);   // knots | bucket table | polynomial rows | flag
        else tma_fetch(image, src, T.cdbBytes, bar);
    }
    // the first round of parameters travels from HBM while the image arrives
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long pfirst = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long pstep = stride, rstep = U * stride;
    const double *up = P.in.uvw + pfirst * P.in.pointStride;
    double un[U];
#pragma unroll
    for (int k = 0; k < U; ++k) un[k] = pfirst + k * pstep < P.N ? __ldcs(up + k * pstep * P.in.pointStride) : 0.0;
    mbar_wait(bar, 0);
    if constexpr (POLY) {
        const bool valid = reinterpret_cast<const int *>(image + T.polyFetch - 16)[DER ? 1 : 0] != 0;
        if (valid) {
            const double *kn = reinterpret_cast<const double *>(image);
            const unsigned short *tab = reinterpret_cast<const unsigned short *>(image + T.knotBytes);
            const double *rows = reinterpret_cast<const double *>(image + T.knotBytes + T.tabBytes);
            repl_point_loop<O, NDEP, DER, U, true>(P, T.buckets, rows, kn, tab, un, up, pfirst, pstep, rstep);
            return;
        }
        // the polynomial rows of this spline did not pass their validation: fetch the recurrence rows instead (rare)
        __syncthreads();                                              // everybody has read the flag
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tma_fetch(image, src, T.cdbBytes, bar);
        }
        mbar_wait(bar, 1);
    }
    const double *rows = reinterpret_cast<const double *>(image);
    const double *kn = reinterpret_cast<const double *>(image + T.rowsBytes);
    const unsigned short *tab = reinterpret_cast<const unsigned short *>(image + T.rowsBytes + T.knotBytes);
    repl_point_loop<O, NDEP, DER, U>(P, T.buckets, rows, kn, tab, un, up, pfirst, pstep, rstep);
}

template <int O, int NDEP, bool DER>
static int launch_curve3(const CurveParams &P, size_t smem, cudaStream_t stream)
{
    if (int rc = allow_dynamic_smem(eval_curve_kernel<O, NDEP, DER>, smem)) return rc;
    {
        // big batches: bank-replicated span rows (when 8 copies of the rows fit beside a second CTA)
        const long long re = option(OPT_CURVE_REPL, -1);
        const long long minN = re >= 0 ? (re ? 0 : (1LL << 62)) : 131072;
        const ReplLayout L = repl_layout(O, NDEP, P.nCoef, 100 * 1024);
        if (P.N >= minN && L.bytes && !P.in.grid && P.nCoef < 65535 && P.table && option(OPT_CURVE_TMA, 1)) {
            // tables built once per spline (caller-owned image), one 1024-thread CTA per SM, image fetched by TMA
            const TableLayout T = table_layout(O, NDEP, P.nCoef);
            long long blocks = (P.N + TAB_THREADS * 4 - 1) / (TAB_THREADS * 4);
            const long long cap = num_sms();
            if (blocks > cap) blocks = cap;
            // (eight points per thread requested up front for short batches were measured slower at 1e6 points:
            // 19.0 against 16.7 us)
            if constexpr (O >= 2) {
                // polynomial rows: values, and values + first derivative (CURVE_POLY=0: recurrence rows, bit-identical to the other
                // curve kernels; CURVE_POLY=2: value-only requests only)
                const long long cp = option(OPT_CURVE_POLY, 1);
                if (cp == 1 || (cp == 2 && !DER)) {
                    const size_t smem = T.polyFetch > T.cdbBytes ? T.polyFetch : T.cdbBytes;   // room for the fallback
                    if (int rc = allow_dynamic_smem(eval_curve_tab_kernel<O, NDEP, DER, REPL_U, true>, smem)) return rc;
                    eval_curve_tab_kernel<O, NDEP, DER, REPL_U, true><<<(unsigned)blocks, TAB_THREADS, smem, stream>>>(P, T);
                    count_launch();
                    return check_launch("bspy_cuda_eval_points(curve, cached polynomial rows)");
                }
            }
            if (int rc = allow_dynamic_smem(eval_curve_tab_kernel<O, NDEP, DER>, T.cdbBytes)) return rc;
            eval_curve_tab_kernel<O, NDEP, DER><<<(unsigned)blocks, TAB_THREADS, T.cdbBytes, stream>>>(P, T);
            count_launch();
            return check_launch("bspy_cuda_eval_points(curve, cached tables)");
        }
        if (P.N >= minN && L.bytes && !P.in.grid && P.nCoef < 65535) {
            if (int rc = allow_dynamic_smem(eval_curve_repl_kernel<O, NDEP, DER>, L.bytes)) return rc;
            const int threads = repl_threads(O, NDEP);
            long long blocks = (P.N + threads * 4 - 1) / (threads * 4);
            const long long cap = (long long)num_sms() * 2;
            if (blocks > cap) blocks = cap;
            eval_curve_repl_kernel<O, NDEP, DER><<<(unsigned)blocks, threads, L.bytes, stream>>>(P, L.buckets);
            count_launch();
            return check_launch("bspy_cuda_eval_points(curve, replicated rows)");
        }
    }
    const int threads = 256;
    const int ppt = (int)option(OPT_CURVE_PPT, 8);
    long long blocks = (P.N + threads * ppt - 1) / (threads * ppt);  // points per thread amortise the table build
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    eval_curve_kernel<O, NDEP, DER><<<(unsigned)blocks, threads, smem, stream>>>(P);
    count_launch();
    return check_launch("bspy_cuda_eval_points(curve)");
}

template <int O>
static int launch_curve2(const CurveParams &P, int nDep, bool der, size_t smem, cudaStream_t stream)
{
    switch (nDep) {
        case 1: return der ? launch_curve3<O, 1, true>(P, smem, stream) : launch_curve3<O, 1, false>(P, smem, stream);
        case 2: return der ? launch_curve3<O, 2, true>(P, smem, stream) : launch_curve3<O, 2, false>(P, smem, stream);
        case 3: return der ? launch_curve3<O, 3, true>(P, smem, stream) : launch_curve3<O, 3, false>(P, smem, stream);
        default: return der ? launch_curve3<O, 4, true>(P, smem, stream) : launch_curve3<O, 4, false>(P, smem, stream);
    }
}

// returns -1000 when the lean path does not apply (caller falls back to the general kernels)
int launch_curve(const SplineDev &s, const PointsDev &in, long long N, const WrtDev &wrt, const OutDev &out, int jac,
                 cudaStream_t stream)
{
    if (s.nInd != 1 || s.order[0] > 6 || s.nDep < 1 || s.nDep > 4) return -1000;
    if (!jac && wrt.d[0] > 1) return -1000;
    if (out.normal && s.nDep != 2) return -1000;
    const int O = s.order[0], nCoef = s.nCoef[0];
    const int recStride = ((O - 1 + O * (O - 1) / 2) + 1) & ~1;
    const size_t smem = sizeof(double) * (((O + nCoef + 1) & ~1) + (size_t)(nCoef - O + 1) * recStride + (size_t)s.nDep * nCoef);
    if (smem > 96 * 1024) return -1000;
    CurveParams P{};
    P.knots = s.knots[0]; P.coefs = s.coefs; P.nCoef = nCoef; P.normalSign = s.normalSign;
    P.in = in; P.out = out; P.N = N;
    // caller-owned table image: trusted only when its size is exactly what this shape needs
    P.table = (s.curveTable && s.curveTableBytes == (long long)table_layout(O, s.nDep, nCoef).bytes && s.curveTableBytes > 0) ? s.curveTable : nullptr;
    P.derivAsValue = (!jac && wrt.d[0] == 1) ? 1 : 0;
    const bool der = jac || wrt.d[0] == 1;
    switch (O) {
        case 1: return launch_curve2<1>(P, s.nDep, der, smem, stream);
        case 2: return launch_curve2<2>(P, s.nDep, der, smem, stream);
        case 3: return launch_curve2<3>(P, s.nDep, der, smem, stream);
        case 4: return launch_curve2<4>(P, s.nDep, der, smem, stream);
        case 5: return launch_curve2<5>(P, s.nDep, der, smem, stream);
        default: return launch_curve2<6>(P, s.nDep, der, smem, stream);
    }
}

template <int O>
static int build_table2(const bspy_spline *sp, const TableLayout &T, void *table, cudaStream_t stream)
{
    const int nCoef = sp->nCoef[0];
    const ReplLayout L = repl_layout(O, sp->nDep, nCoef, 200 * 1024);
    const double *kn = sp->knots[0], *cf = sp->coefs;
    unsigned char *img = (unsigned char *)table;
#define BSPY_TABLE_CASE(ND)                                                                                          \
    case ND:                                                                                                        \
        if (int rc = allow_dynamic_smem(curve_table_kernel<O, ND>, L.bytes)) return rc;                             \
        curve_table_kernel<O, ND><<<1, 512, L.bytes, stream>>>(kn, cf, nCoef, T.buckets, img, T);                   \
        break;
    switch (sp->nDep) {
        BSPY_TABLE_CASE(1) BSPY_TABLE_CASE(2) BSPY_TABLE_CASE(3) BSPY_TABLE_CASE(4)
        default: return BSPY_E_UNSUPPORTED;
    }
#undef BSPY_TABLE_CASE
    count_launch();
    return check_launch("bspy_cuda_curve_table_build");
}

static bool table_shape_ok(const bspy_spline *sp)
{
    return sp && sp->nInd == 1 && sp->order[0] >= 1 && sp->order[0] <= 6 && sp->nDep >= 1 && sp->nDep <= 4 &&
           sp->nCoef[0] >= sp->order[0] && sp->nCoef[0] < 65535 && sp->knots[0] && sp->coefs;
}

}  // namespace bspy

using namespace bspy;

extern "C" int64_t bspy_cuda_curve_table_bytes(const bspy_spline *spline)
{
    if (!table_shape_ok(spline)) return 0;
    const TableLayout T = table_layout(spline->order[0], spline->nDep, spline->nCoef[0]);
    return (T.bytes && T.cdbBytes <= 200 * 1024) ? (int64_t)T.bytes : 0;
}

extern "C" int bspy_cuda_curve_table_build(const bspy_spline *spline, void *table, int64_t tableBytes, void *stream)
{
    const int64_t need = bspy_cuda_curve_table_bytes(spline);
    if (!need || !table || tableBytes < need || (reinterpret_cast<uintptr_t>(table) & 15)) {
        set_error("bspy_cuda_curve_table_build: not a curve shape with tables, or table NULL / too small / not 16-byte aligned");
        return BSPY_E_ARG;
    }
    const TableLayout T = table_layout(spline->order[0], spline->nDep, spline->nCoef[0]);
    switch (spline->order[0]) {
        case 1: return build_table2<1>(spline, T, table, (cudaStream_t)stream);
        case 2: return build_table2<2>(spline, T, table, (cudaStream_t)stream);
        case 3: return build_table2<3>(spline, T, table, (cudaStream_t)stream);
        case 4: return build_table2<4>(spline, T, table, (cudaStream_t)stream);
        case 5: return build_table2<5>(spline, T, table, (cudaStream_t)stream);
        default: return build_table2<6>(spline, T, table, (cudaStream_t)stream);
    }
}
