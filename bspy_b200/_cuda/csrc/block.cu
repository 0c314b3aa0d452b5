// Building blocks of SURVEY 8(f) row 1: Spline.contract and SplineBlock evaluation.
//
//   bspy_cuda_contract_axis        one variable of a coefficient array contracted against `order` basis values
//                                  (reference bspy/_spline_operations.py:184-223: coefs[..., ix-o:ix, ...] @ bValues)
//   bspy_cuda_block_accumulate     dst[rows[r], :] += src[r, :]: the row sums of a block of splines
//                                  (reference bspy/spline_block.py:37-44, 231-245)
//   bspy_cuda_normal_from_jacobian cofactor normals of N jacobians (reference bspy/_spline_evaluation.py:215-246 with a
//                                  SplineBlock as `self`, bspy/spline_block.py:282)
// and of row 3:
//   bspy_cuda_collocation          rows of the collocation matrix of least_squares / contour (reference
//                                  bspy/_spline_fitting.py:736-750, 190-219): A[r, ix-order:ix] = bspline_values(...)
#include "common.cuh"

namespace bspy {

// out[a, b] = sum_j coefs[a, first + j, b] * basis[j]; one thread per output element, consecutive threads along b
__global__ void __launch_bounds__(256) contract_axis_kernel(const double *__restrict__ coefs, const long long outer, const long long n,
                                                            const long long inner, const int first, const int order,
                                                            const double *__restrict__ basis, double *__restrict__ out)
{
    const long long total = outer * inner;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long a = t / inner, b = t - a * inner;
        const double *src = coefs + (a * n + first) * inner + b;
        double acc = 0.0;
        for (int j = 0; j < order; ++j) acc = fma(__ldg(src + j * inner), __ldg(basis + j), acc);
        out[t] = acc;
    }
}

// one thread per row: bit-exact basis (basis_strict, the reference's operation order) of derivative order deriv[r]
// scattered into the zeroed row r of the dense matrix A (N x nCoef, leading dimension ldA)
struct GlobalColumn {
    double *base;
    __device__ __forceinline__ double &operator()(int j) const { return base[j]; }
};

__global__ void __launch_bounds__(128) collocation_kernel(const double *__restrict__ knots, const int nKnots, const int order,
                                                          const double *__restrict__ u, const int32_t *__restrict__ deriv,
                                                          const long long N, int32_t *__restrict__ spansOut, double *__restrict__ A,
                                                          const long long ldA)
{
    const int nCoef = nKnots - order;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) {
        const double x = u[r];
        const int ix = span_search(knots, nKnots, order, x);
        if (spansOut) spansOut[r] = ix;
        double *row = A + r * ldA;
        for (int c = 0; c < nCoef; ++c) row[c] = 0.0;
        basis_strict(knots, order, ix, x, deriv ? deriv[r] : 0, false, GlobalColumn{row + ix - order});
    }
}

struct RowMap {
    int rows[64];
};

__global__ void __launch_bounds__(256) block_accumulate_kernel(double *__restrict__ dst, const long long dstLd,
                                                               const double *__restrict__ src, const long long srcLd, const int nRows,
                                                               const RowMap map, const long long N)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x)
        for (int r = 0; r < nRows; ++r) {
            double *d = dst + map.rows[r] * dstLd + p;
            *d = *d + __ldcs(src + r * srcLd + p);
        }
}

// LU determinant with partial pivoting of an n x n matrix held in local memory (n <= BSPY_MAX_IND)
__device__ static double det_lu_local(double *a, int n)
{
    double det = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = fabs(a[c * n + c]);
        for (int r = c + 1; r < n; ++r) {
            const double v = fabs(a[r * n + c]);
            if (v > best) { best = v; piv = r; }
        }
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

constexpr int NJ_MAX = 9;   // D = max(nInd, nDep) <= 9: minors up to 8 x 8

__global__ void __launch_bounds__(128) normal_from_jacobian_kernel(const double *__restrict__ jac, const int nDep, const int nInd,
                                                                   const long long N, const int sign, const unsigned normalize,
                                                                   const unsigned mask, double *__restrict__ normal)
{
    const int D = nInd > nDep ? nInd : nDep, M = D - 1;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
        double T[NJ_MAX * (NJ_MAX - 1)];          // tangent space, D rows x M columns (transposed when nInd > nDep)
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < M; ++c)
                T[r * M + c] = nInd > nDep ? jac[((long long)c * nInd + r) * N + p] : jac[((long long)r * nInd + c) * N + p];
        double n[NJ_MAX];
        double sq = 0.0;
        for (int i = 0; i < D; ++i) {
            double minor[(NJ_MAX - 1) * (NJ_MAX - 1)];
            int rr = 0;
            for (int r = 0; r < D; ++r) {
                if (r == i) continue;
                for (int c = 0; c < M; ++c) minor[rr * M + c] = T[r * M + c];
                ++rr;
            }
            double det;
            if (M == 0) det = 1.0;
            else if (M == 1) det = minor[0];
            else if (M == 2) det = minor[0] * minor[3] - minor[1] * minor[2];
            else det = det_lu_local(minor, M);
            n[i] = ((i & 1) ? -det : det) * (double)sign;
            if (mask & (1u << i)) sq += n[i] * n[i];
        }
        const double len = sqrt(sq);
        for (int i = 0; i < D; ++i) normal[(long long)i * N + p] = normalize ? n[i] / len : n[i];
    }
}

}  // namespace bspy

using namespace bspy;

extern "C" int bspy_cuda_contract_axis(const double *coefs, int64_t outer, int64_t n, int64_t inner, int32_t first, int32_t order,
                                       const double *basis, double *out, void *stream)
{
    if (!coefs || !basis || !out || outer < 0 || inner < 0 || n < 1 || order < 1 || first < 0 || first + order > n) {
        set_error("bspy_cuda_contract_axis: bad argument");
        return BSPY_E_ARG;
    }
    const long long total = outer * inner;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    contract_axis_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(coefs, outer, n, inner, first, order, basis, out);
    count_launch();
    return check_launch("bspy_cuda_contract_axis");
}

extern "C" int bspy_cuda_block_accumulate(double *dst, int64_t dstLd, const double *src, int64_t srcLd, int32_t nRows,
                                          const int32_t *dstRows_host, int64_t N, void *stream)
{
    if (!dst || !src || !dstRows_host || nRows < 0 || N < 0) {
        set_error("bspy_cuda_block_accumulate: bad argument");
        return BSPY_E_ARG;
    }
    if (N == 0) return 0;
    long long blocks = (N + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    for (int r0 = 0; r0 < nRows; r0 += 64) {
        RowMap map{};
        const int m = nRows - r0 < 64 ? nRows - r0 : 64;
        for (int r = 0; r < m; ++r) {
            if (dstRows_host[r0 + r] < 0) { set_error("bspy_cuda_block_accumulate: negative row"); return BSPY_E_ARG; }
            map.rows[r] = dstRows_host[r0 + r];
        }
        block_accumulate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, dstLd, src + (long long)r0 * srcLd, srcLd, m, map, N);
        count_launch();
        const int rc = check_launch("bspy_cuda_block_accumulate");
        if (rc) return rc;
    }
    return 0;
}

extern "C" int bspy_cuda_normal_from_jacobian(const double *jacobian, int32_t nDep, int32_t nInd, int64_t N, int32_t normalSign,
                                              uint32_t flags, uint32_t normalMask, double *normal, void *stream)
{
    if (!jacobian || !normal || N < 0 || nDep < 0 || nInd < 0) {
        set_error("bspy_cuda_normal_from_jacobian: bad argument");
        return BSPY_E_ARG;
    }
    if (nInd - nDep != 1 && nDep - nInd != 1) {
        set_error("The number of independent variables must be one different than the number of dependent variables.");
        return BSPY_E_NORMAL_DIMS;
    }
    const int D = nInd > nDep ? nInd : nDep;
    if (D > NJ_MAX) { set_error("bspy_cuda_normal_from_jacobian: max(nInd, nDep) = %d > %d", D, NJ_MAX); return BSPY_E_UNSUPPORTED; }
    if (N == 0) return 0;
    if (normalMask == 0) normalMask = 0xffffffffu;
    long long blocks = (N + 127) / 128;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    normal_from_jacobian_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(jacobian, nDep, nInd, N, normalSign < 0 ? -1 : 1,
                                                                                  (flags & BSPY_NORMALIZE) ? 1u : 0u, normalMask, normal);
    count_launch();
    return check_launch("bspy_cuda_normal_from_jacobian");
}

extern "C" int bspy_cuda_collocation(const double *knots, int32_t nKnots, int32_t order, const double *u, const int32_t *derivOrders,
                                     int64_t N, int32_t *spansOut, double *A, int64_t ldA, void *stream)
{
    if (!knots || !u || !A || N < 0 || order < 1 || nKnots < 2 * order || ldA < nKnots - order) {
        set_error("bspy_cuda_collocation: bad argument");
        return BSPY_E_ARG;
    }
    if (N == 0) return 0;
    long long blocks = (N + 127) / 128;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    collocation_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(knots, nKnots, order, u, derivOrders, N, spansOut, A, ldA);
    count_launch();
    return check_launch("bspy_cuda_collocation");
}
