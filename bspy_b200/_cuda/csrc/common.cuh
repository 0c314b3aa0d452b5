// Shared device/host helpers of libbspy_cuda.so (sm_100a).
// Algorithm provenance: the semantics restated here are those of the reference's
// bspy/_spline_evaluation.py:4-27 (span + basis recurrence); the code is written for the GPU
// (registers, unrolled triangular recurrence, shared stages for values and derivatives).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/bspy_cuda.h"

namespace bspy {

// ---- host side ------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);   // cudaGetLastError -> status code (+ error string)

// SM count of the CURRENT device (cached per device ordinal; benign race: every writer stores the same value)
static inline int num_sms()
{
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per function AND per device, so it is
// set on every launch that needs it (a driver call of well under a microsecond) instead of being cached per process.
template <class K>
static inline int allow_dynamic_smem(K kernel, size_t smem)
{
    if (smem <= 48 * 1024) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%zu bytes of shared memory): %s", smem, cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// Experiment switches (bspy_cuda_set_option / environment BSPY_<NAME> read ONCE when the library is first used; the
// launch path never calls getenv).  option() returns `unset` when the switch was neither set nor in the environment.
enum Option {
    OPT_CURVE_REPL, OPT_CURVE_PPT, OPT_BIN_MODE, OPT_BIN_OVERLAP, OPT_STAGED, OPT_DEP_TILE, OPT_SPAN_RECORDS, OPT_BIN_CHUNK,
    OPT_BIN_REC_CHUNK_LOG2, OPT_GRID_CHUNK, OPT_GRID_ROWS, OPT_GRID_GROUP, OPT_CELL_KERNEL, OPT_CURVE_TMA, OPT_MANY_MODE,
    OPT_GRID3_ROWS, OPT_GRID3_CHUNK, OPT_EXP_A, OPT_EXP_B, OPT_IMAGE, OPT_STAGED_PAIR, OPT_STAGED_WAVES, OPT_CURVE_POLY, OPT_CELL_POLY, OPT_BIN_PERM, OPT_COUNT
};
long long option(Option o, long long unset);

// ---- kernel parameter blocks shared by scattered.cu and grid.cu ------------------------------------
struct SplineDev {
    int nInd, nDep;
    int order[BSPY_MAX_IND];
    int nCoef[BSPY_MAX_IND];
    const double *knots[BSPY_MAX_IND];
    const double *coefs;
    long long stride[BSPY_MAX_IND];  // coefficient stride of variable i (elements)
    long long depStride;
    int normalSign;
    const void *curveTable;          // bspy_spline.curveTable / curveTableBytes (curves: replicated tables, built once)
    long long curveTableBytes;
};

struct PointsDev {
    const double *uvw;      // scattered: parameter i of point p at uvw[p*pointStride + i*varStride]
    long long pointStride, varStride;
    const double *axes[BSPY_MAX_IND];  // grid mode: axes[i][idx_i]
    long long nAxis[BSPY_MAX_IND];
    int grid;               // 0 = scattered, 1 = tensor grid (last variable fastest)
    // cell-binned mode (perm != nullptr): thread t evaluates point base + perm[t]; its knot spans are packed in
    // cellKey[t] = ((ix0-o0) * m1 + (ix1-o1)) * m2 + ...  with m_i = nCoef_i - o_i + 1 spans in variable i
    const int *perm;
    const int *cellKey;
    long long base;
    // sorted-record mode (records != nullptr): thread t evaluates the point whose parameters are
    // records[4t .. 4t+nInd) (32-byte records written in cell order by bin_scatter_records_kernel); the packed
    // span key (cell) and the point's index inside its chunk are the low / high 32 bits of the bit pattern of
    // records[4t+3] for nInd <= 3, recKI[t] = (key, index) for nInd == 4
    const double *records;
    const int2 *recKI;         // records == nullptr && recKI != nullptr: sorted (cell key, index) pairs only -- the kernel gathers
                               // the parameters of slot t from uvw[(base + index) * pointStride ..] itself (no record scatter pass)
    // per-span records of variable i (left knots | reciprocal knot gaps, SpanRec<order>::stride doubles per span), built
    // once per call by span_records_kernel for the binned path; nullptr: gaps are divided per point
    const double *spanRec[BSPY_MAX_IND];
    // cell images (eval_image_kernel): images[cell * size ..] = padded window of the cell followed by its span records
    const double *images;
    int prefetchImages;       // eval_image2_kernel: request all lines of the cell image up front
    const int *sortedTotal;   // number of slots of the sorted sequence (cell segments rounded up to even lengths)
    // gate (device int) of a kernel pair that shares one decision taken on the device: the kernel runs only when *gate ==
    // gateWant (cell polynomial images that passed / failed their validation); nullptr: always
    const int *gate;
    int gateWant;
};

struct OutDev {
    long long ld;      // leading dimension of the outputs = total number of points
    double *values;    // (nDep, N)
    double *jacobian;  // (nDep, nInd, N)
    double *normal;    // (D, N)
    int32_t *spans;    // (nInd, N)
    long long *firstOutside;
    unsigned normalize, normalMask;
    // array-of-structs results (sorted-record mode): point t writes [values | jacobian (d, iv) | normal] to
    // aos[t * aosStride ..]; aosStride is a multiple of 4 doubles so that records are whole 32-byte sectors
    // aosScatter != 0: the record of sorted point t goes to aos[(aosBase + index inside the chunk) * aosStride ..], i.e.
    // straight to the point's ORIGINAL position in a caller-visible array of records (no un-permute pass)
    double *aos;
    int aosStride;
    int aosScatter;
    int aosNormal;      // array-of-structs records carry the normal after the jacobian (out.normal itself stays NULL)
    int aosWide;        // records are 32-byte aligned with a stride that is a multiple of 4 doubles: 256-bit stores
    long long aosBase;
};

struct WrtDev {
    int d[BSPY_MAX_IND];
};

// ---- device side ----------------------------------------------------------------------------

// Number of knots <= u clamped to [order, nKnots - order]  (reference: np.searchsorted(.., 'right')
// + clamp).  NaN compares false against everything; numpy sorts it after every knot, so it
// selects the last span.  Branch-free bisection over a power-of-two ladder.
__device__ __forceinline__ int span_search(const double *__restrict__ knots, int nKnots, int order, double u)
{
    int lo = 0;           // invariant: knots[0..lo) <= u
    int n = nKnots;
    if (u != u) return nKnots - order;
    while (n > 0) {
        int half = n >> 1;
        int mid = lo + half;
        bool le = __ldg(knots + mid) <= u;
        lo = le ? mid + 1 : lo;
        n = le ? n - half - 1 : half;
    }
    lo = max(lo, order);
    lo = min(lo, nKnots - order);
    return lo;
}

// Restricted search: the caller guarantees the answer lies in [order, nKnots-order]; only the
// knots strictly inside that range are probed (saves the steps spent on the clamped ends).
__device__ __forceinline__ int span_search_inner(const double *__restrict__ knots, int nKnots, int order, double u)
{
    if (u != u) return nKnots - order;
    int lo = order;                  // answer in [lo, lo+n]
    int n = nKnots - 2 * order;      // candidates knots[order .. nKnots-order)
    while (n > 0) {
        int half = n >> 1;
        int mid = lo + half;
        bool le = __ldg(knots + mid) <= u;
        lo = le ? mid + 1 : lo;
        n = le ? n - half - 1 : half;
    }
    return lo;
}

// Strict (bit-exact) recurrence: IEEE operations in the reference's order, no FMA contraction.
// b has `order` entries (any addressable memory through the accessor).  Used by bspy_cuda_basis.
template <class Acc>
__device__ __forceinline__ void basis_strict(const double *__restrict__ knots, int order, int ix, double u,
                                             int deriv, bool taylor, Acc b)
{
    for (int j = 0; j < order; ++j) b(j) = 0.0;
    if (deriv >= order) return;
    b(order - 1) = 1.0;
    const int nValue = order - deriv;
    for (int deg = 1; deg < order; ++deg) {
        int slot = order - deg;
        if (deg < nValue) {
            for (int i = ix - deg; i < ix; ++i, ++slot) {
                const double ki = __ldg(knots + i);
                const double a = __ddiv_rn(__dsub_rn(u, ki), __dsub_rn(__ldg(knots + i + deg), ki));
                b(slot - 1) = __dadd_rn(b(slot - 1), __dmul_rn(__dsub_rn(1.0, a), b(slot)));
                b(slot) = __dmul_rn(b(slot), a);
            }
        } else {
            const double scale = __ddiv_rn((double)deg, taylor ? (double)(order - deg) : 1.0);
            for (int i = ix - deg; i < ix; ++i, ++slot) {
                const double ki = __ldg(knots + i);
                const double a = __ddiv_rn(scale, __dsub_rn(__ldg(knots + i + deg), ki));
                b(slot - 1) = __dadd_rn(b(slot - 1), __dmul_rn(-a, b(slot)));
                b(slot) = __dmul_rn(b(slot), a);
            }
        }
    }
}

// Register-resident recurrence for a compile-time order, written on reciprocal knot gaps so that the same code
// serves the per-point path (gaps divided here) and the per-span records of the binned path (gaps divided once per
// span by span_records_kernel): both give the same bits.
//   dl[j] = u - knots[ix-(O-1)+j]            (j < O-1: the left knots)
//   rc[at(deg)+t] = 1 / (knots[ix+t] - knots[ix-deg+t]),  at(deg) = deg(deg-1)/2, t < deg
//   DER == false : b0 = basis of derivative order `d` (runtime, 0 = values)
//   DER == true  : b0 = values, b1 = first derivatives; the O-2 lower stages are shared.
template <int O>
struct SpanRec {
    static constexpr int left = O - 1;                       // left knots
    static constexpr int recips = O * (O - 1) / 2;           // reciprocal gaps
    static constexpr int used = left + recips;
    static constexpr int stride = (used + 1) & ~1;           // doubles, even -> records are 16-byte aligned
};

template <int O, bool DER>
__device__ __forceinline__ void basis_core(const double (&dl)[O > 1 ? O - 1 : 1], const double (&rc)[O > 1 ? O * (O - 1) / 2 : 1],
                                           int d, double (&b0)[O], double (&b1)[O])
{
#pragma unroll
    for (int j = 0; j < O; ++j) { b0[j] = 0.0; b1[j] = 0.0; }
    if (!DER && d >= O) return;
    b0[O - 1] = 1.0;
    const int nValue = DER ? O - 1 : O - d;   // value stages: deg < nValue
    int at = 0;
#pragma unroll
    for (int deg = 1; deg < O; ++deg) {
        if (DER && deg == O - 1) {
            // last stage twice: once as a value stage into b0, once as a derivative stage into b1
#pragma unroll
            for (int j = 0; j < O; ++j) b1[j] = b0[j];
#pragma unroll
            for (int t = 0; t < deg; ++t) {
                const int slot = O - deg + t;
                const double r = rc[at + t];
                const double a = dl[O - 1 - deg + t] * r;
                b0[slot - 1] = fma(1.0 - a, b0[slot], b0[slot - 1]);
                b0[slot] *= a;
                const double g = (double)deg * r;
                b1[slot - 1] = fma(-g, b1[slot], b1[slot - 1]);
                b1[slot] *= g;
            }
        } else if (deg < nValue) {
#pragma unroll
            for (int t = 0; t < deg; ++t) {
                const int slot = O - deg + t;
                const double a = dl[O - 1 - deg + t] * rc[at + t];
                b0[slot - 1] = fma(1.0 - a, b0[slot], b0[slot - 1]);
                b0[slot] *= a;
            }
        } else {
#pragma unroll
            for (int t = 0; t < deg; ++t) {
                const int slot = O - deg + t;
                const double g = (double)deg * rc[at + t];
                b0[slot - 1] = fma(-g, b0[slot], b0[slot - 1]);
                b0[slot] *= g;
            }
        }
        at += deg;
    }
}

// from the knot window kw[j] = knots[ix - (O-1) + j], j = 0 .. 2(O-1)-1: knots[ix-deg+t] is kw[O-1-deg+t] and
// knots[ix+t] is kw[O-1+t]
template <int O, bool DER>
__device__ __forceinline__ void basis_regs(const double (&kw)[2 * (O - 1) > 0 ? 2 * (O - 1) : 1], double u, int d,
                                           double (&b0)[O], double (&b1)[O])
{
    double dl[O > 1 ? O - 1 : 1], rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
    for (int j = 0; j < O - 1; ++j) dl[j] = u - kw[j];
    int at = 0;
#pragma unroll
    for (int deg = 1; deg < O; ++deg) {
#pragma unroll
        for (int t = 0; t < deg; ++t) rc[at + t] = 1.0 / (kw[O - 1 + t] - kw[O - 1 - deg + t]);
        at += deg;
    }
    basis_core<O, DER>(dl, rc, d, b0, b1);
}

// from the per-span record rec = { left knots | reciprocal gaps } (SpanRec<O>::stride doubles, 16-byte aligned)
template <int O, bool DER>
__device__ __forceinline__ void basis_from_span_record(const double *__restrict__ rec, double u, int d, double (&b0)[O],
                                                       double (&b1)[O])
{
    using R = SpanRec<O>;
    double r[R::stride > 0 ? R::stride : 1];
    if constexpr (R::stride > 0) {
        const double2 *rp = reinterpret_cast<const double2 *>(rec);
#pragma unroll
        for (int j = 0; j < R::stride / 2; ++j) {
            const double2 x = __ldg(rp + j);
            r[2 * j] = x.x;
            r[2 * j + 1] = x.y;
        }
    }
    double dl[O > 1 ? O - 1 : 1], rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
    for (int j = 0; j < O - 1; ++j) dl[j] = u - r[j];
#pragma unroll
    for (int j = 0; j < O * (O - 1) / 2; ++j) rc[j] = r[O - 1 + j];
    basis_core<O, DER>(dl, rc, d, b0, b1);
}

template <int O>
__device__ __forceinline__ void load_knot_window(const double *__restrict__ knots, int ix, double (&kw)[2 * (O - 1) > 0 ? 2 * (O - 1) : 1])
{
#pragma unroll
    for (int j = 0; j < 2 * (O - 1); ++j) kw[j] = __ldg(knots + ix - (O - 1) + j);
}

__device__ __forceinline__ double fetch_param(const PointsDev &in, long long p, int iv, long long &rem)
{
    // grid mode: decode the multi-index from the flat index, last variable fastest; callers walk
    // iv from nInd-1 down to 0 and thread `rem` through.
    if (in.grid) {
        const long long n = in.nAxis[iv];
        const long long idx = rem % n;
        rem /= n;
        return __ldg(in.axes[iv] + idx);
    }
    return __ldg(in.uvw + p * in.pointStride + iv * in.varStride);
}

__device__ __forceinline__ bool gate_closed(const PointsDev &in) { return in.gate != nullptr && *in.gate != in.gateWant; }

// first point outside the domain: keep the smallest index
__device__ __forceinline__ void report_outside(int64_t *flag, int64_t p)
{
    // flag holds a negative value when nothing was reported yet; compare as unsigned so that
    // "negative" (huge unsigned) loses against every real index
    atomicMin(reinterpret_cast<unsigned long long *>(flag), (unsigned long long)p);
}

}  // namespace bspy
