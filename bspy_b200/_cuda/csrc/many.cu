// Many independent curves of one shape: bspy_cuda_eval_many.
//
// Replaces a Python loop of `spline(uArray)` over a list of splines (reference ufunc-style
// evaluate, bspy/spline.py:940-947, one interpreter pass per point) for S curves that share
// order / nCoef / nDep but have their own knots, coefficients and parameters.
// One warp per curve: the curve's knots and coefficients (a few hundred bytes) are staged once in
// the warp's slice of shared memory with coalesced loads, each lane then evaluates every 32nd
// parameter (span bisection over the staged knots, register recurrence, window dot products) and
// stores are coalesced along the point index.  No block-level synchronisation.
#include "common.cuh"

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

template <int O, bool DER>
__global__ void __launch_bounds__(256) many_kernel(const ManyParams P)
{
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    double *sk = sm + (long long)warp * P.slice;   // knots
    const int nKnots = O + P.nCoef;
    double *sc = sk + ((nKnots + 1) & ~1);          // coefs (nDep, nCoef)
    const int nC = P.nDep * P.nCoef;
    for (long long s = (long long)blockIdx.x * warpsPerBlock + warp; s < P.nSplines; s += (long long)gridDim.x * warpsPerBlock) {
        const double *gk = P.knots + s * P.knotStride;
        const double *gc = P.coefs + s * P.coefStride;
        __syncwarp();
        for (int i = lane; i < nKnots; i += 32) sk[i] = __ldcs(gk + i);
        for (int i = lane; i < nC; i += 32) sc[i] = __ldcs(gc + i);
        __syncwarp();
        const double lo = sk[O - 1], hi = sk[P.nCoef];
        const double *gu = P.u + s * P.nPts;
        for (int p = lane; p < P.nPts; p += 32) {
            const double u = __ldcs(gu + p);
            if (((u < lo) | (u > hi)) && P.firstOutside) report_outside((int64_t *)P.firstOutside, s * P.nPts + p);
            // span: knots <= u among sk[O .. nCoef)
            int ix = O, n = P.nCoef - O;
            if (u != u) { ix = P.nCoef; n = 0; }
            while (n > 0) {
                const int half = n >> 1;
                const bool le = sk[ix + half] <= u;
                ix = le ? ix + half + 1 : ix;
                n = le ? n - half - 1 : half;
            }
            double kw[2 * (O - 1) > 0 ? 2 * (O - 1) : 1];
#pragma unroll
            for (int j = 0; j < 2 * (O - 1); ++j) kw[j] = sk[ix - (O - 1) + j];
            double b0[O], b1[O];
            basis_regs<O, DER>(kw, u, 0, b0, b1);
            for (int d = 0; d < P.nDep; ++d) {
                const double *row = sc + d * P.nCoef + ix - O;
                double v = 0.0, g = 0.0;
#pragma unroll
                for (int j = 0; j < O; ++j) {
                    const double x = row[j];
                    v = fma(x, b0[j], v);
                    if (DER) g = fma(x, b1[j], g);
                }
                __stcs(P.values + (s * P.nDep + d) * P.nPts + p, v);
                if (DER) __stcs(P.deriv1 + (s * P.nDep + d) * P.nPts + p, g);
            }
        }
    }
}

template <int O>
static int launch_many(const ManyParams &P, cudaStream_t stream)
{
    const int threads = 256, warps = threads / 32;
    const size_t smem = (size_t)P.slice * warps * sizeof(double);
    if (smem > 200 * 1024) { set_error("bspy_cuda_eval_many: curve too large for shared memory"); return BSPY_E_UNSUPPORTED; }
    long long blocks = (P.nSplines + warps - 1) / warps;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    cudaError_t e = cudaSuccess;
    if (P.deriv1) {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(many_kernel<O, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) many_kernel<O, true><<<(unsigned)blocks, threads, smem, stream>>>(P);
    } else {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(many_kernel<O, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) many_kernel<O, false><<<(unsigned)blocks, threads, smem, stream>>>(P);
    }
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    count_launch();
    return check_launch("bspy_cuda_eval_many");
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
    ManyParams P{};
    P.nCoef = nCoef; P.nDep = nDep; P.nPts = nPts; P.nSplines = nSplines;
    P.knots = knots; P.coefs = coefs; P.u = u;
    P.knotStride = knotStride; P.coefStride = coefStride;
    P.values = values; P.deriv1 = deriv1;
    P.firstOutside = (long long *)firstOutside;
    P.slice = ((order + nCoef + 1) & ~1) + ((nDep * nCoef + 1) & ~1);
    cudaStream_t st = (cudaStream_t)stream;
    switch (order) {
        case 1: return launch_many<1>(P, st);
        case 2: return launch_many<2>(P, st);
        case 3: return launch_many<3>(P, st);
        case 4: return launch_many<4>(P, st);
        case 5: return launch_many<5>(P, st);
        case 6: return launch_many<6>(P, st);
        case 7: return launch_many<7>(P, st);
        case 8: return launch_many<8>(P, st);
        default:
            set_error("bspy_cuda_eval_many: order %d not supported (1..8)", order);
            return BSPY_E_UNSUPPORTED;
    }
}
