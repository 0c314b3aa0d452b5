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
    many_kernel<O, NDEP, DER><<<(unsigned)blocks, threads, smem, stream>>>(P);
    count_launch();
    return check_launch("bspy_cuda_eval_many");
}

template <int O>
static int launch_many(const ManyParams &P, cudaStream_t stream)
{
    const bool der = P.deriv1 != nullptr;
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
