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
};

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
}

template <int O, int NDEP, bool DER>
static int launch_curve3(const CurveParams &P, size_t smem, cudaStream_t stream)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(eval_curve_kernel<O, NDEP, DER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const int threads = 256;
    const char *e = getenv("BSPY_CURVE_PPT");
    const int ppt = e ? atoi(e) : 8;
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

}  // namespace bspy
