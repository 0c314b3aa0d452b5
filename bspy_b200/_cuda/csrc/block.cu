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

int make_spline_dev(const bspy_spline *sp, SplineDev &s, const char *who);

// ---- curvature in one pass -------------------------------------------------------------------------------
// bspy/_spline_evaluation.py:80-107 for N points of a curve (nInd == 1) or surface (nInd == 2): the reference makes two
// (curves) or five derivative calls plus a normal per point; here one thread walks the point's coefficient window ONCE with
// the basis values, first and second derivatives of every variable (relaxed recurrence on the knots, like eval_generic)
// and finishes with the curvature formula -- no intermediate derivative arrays.  Orders up to 8, nDep up to 3 (nDep == 1:
// the graph of the function, x(u) = u, like the reference's self.graph()).
constexpr int CURV_MAXO = 8;

__device__ __forceinline__ void basis_local(const double *__restrict__ knots, const int order, const int ix, const double u,
                                            const int deriv, double (&b)[CURV_MAXO])
{
#pragma unroll
    for (int j = 0; j < CURV_MAXO; ++j) b[j] = 0.0;
    if (deriv >= order) return;
    b[order - 1] = 1.0;
    const int nValue = order - deriv;
    for (int deg = 1; deg < order; ++deg) {
        int slot = order - deg;
        for (int i = ix - deg; i < ix; ++i, ++slot) {
            const double ki = __ldg(knots + i);
            const double gap = __ldg(knots + i + deg) - ki;
            // static indexing only (the array lives in registers): walk the slots with a compile-time loop
#pragma unroll
            for (int sl = 1; sl < CURV_MAXO; ++sl) {
                if (sl == slot) {
                    if (deg < nValue) {
                        const double a = (u - ki) / gap;
                        b[sl - 1] = fma(1.0 - a, b[sl], b[sl - 1]);
                        b[sl] *= a;
                    } else {
                        const double a = (double)deg / gap;
                        b[sl - 1] = fma(-a, b[sl], b[sl - 1]);
                        b[sl] *= a;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) curvature_points_kernel(const SplineDev s, const double *__restrict__ uvw, const long long pointStride,
                                                               const long long varStride, const long long N, double *__restrict__ out,
                                                               long long *firstOutside)
{
    const int nDep = s.nDep;
    const bool graph = nDep == 1;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
        double B[2][3][CURV_MAXO];
        int ix[2] = {0, 0};
        bool outside = false;
        for (int iv = 0; iv < s.nInd; ++iv) {
            const double *k = s.knots[iv];
            const int o = s.order[iv];
            const double u = __ldg(uvw + p * pointStride + iv * varStride);
            outside |= (u < __ldg(k + o - 1)) | (u > __ldg(k + s.nCoef[iv]));
            ix[iv] = span_search(k, o + s.nCoef[iv], o, u);
            basis_local(k, o, ix[iv], u, 0, B[iv][0]);
            basis_local(k, o, ix[iv], u, 1, B[iv][1]);
            basis_local(k, o, ix[iv], u, 2, B[iv][2]);
        }
        if (outside && firstOutside) report_outside((int64_t *)firstOutside, p);
        if (s.nInd == 1) {
            double fp[3] = {0, 0, 0}, fpp[3] = {0, 0, 0};
            const int o = s.order[0];
            for (int d = 0; d < nDep; ++d) {
                const double *c = s.coefs + d * s.depStride + (ix[0] - o);
                double a1 = 0.0, a2 = 0.0;
#pragma unroll
                for (int j = 0; j < CURV_MAXO; ++j)
                    if (j < o) {
                        const double x = __ldg(c + j);
                        a1 = fma(x, B[0][1][j], a1);
                        a2 = fma(x, B[0][2][j], a2);
                    }
                fp[d] = a1;
                fpp[d] = a2;
            }
            double pp = graph ? 1.0 : 0.0, pq = 0.0, qq = 0.0;     // f'.f', f'.f'', f''.f''
            for (int d = 0; d < nDep; ++d) {
                pp = fma(fp[d], fp[d], pp);
                pq = fma(fp[d], fpp[d], pq);
                qq = fma(fpp[d], fpp[d], qq);
            }
            double num;
            if (graph) num = fpp[0];                                // (1, f') x (0, f'')
            else if (nDep == 2) num = fp[0] * fpp[1] - fp[1] * fpp[0];
            else num = sqrt(qq * pp - pq * pq);
            out[p] = num / (pp * sqrt(pp));
        } else {
            double su[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, suu[3] = {0, 0, 0}, suv[3] = {0, 0, 0}, svv[3] = {0, 0, 0};
            const int ou = s.order[0], ov = s.order[1];
            for (int d = 0; d < nDep; ++d) {
                const double *c = s.coefs + d * s.depStride + (long long)(ix[0] - ou) * s.stride[0] + (ix[1] - ov);
                double a_u = 0.0, a_v = 0.0, a_uu = 0.0, a_uv = 0.0, a_vv = 0.0;
#pragma unroll
                for (int i = 0; i < CURV_MAXO; ++i)
                    if (i < ou) {
                        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
                        for (int j = 0; j < CURV_MAXO; ++j)
                            if (j < ov) {
                                const double x = __ldg(c + i * s.stride[0] + j);
                                t0 = fma(x, B[1][0][j], t0);
                                t1 = fma(x, B[1][1][j], t1);
                                t2 = fma(x, B[1][2][j], t2);
                            }
                        a_u = fma(B[0][1][i], t0, a_u);
                        a_v = fma(B[0][0][i], t1, a_v);
                        a_uu = fma(B[0][2][i], t0, a_uu);
                        a_uv = fma(B[0][1][i], t1, a_uv);
                        a_vv = fma(B[0][0][i], t2, a_vv);
                    }
                const int dd = graph ? 2 : d;
                su[dd] = a_u; sv[dd] = a_v; suu[dd] = a_uu; suv[dd] = a_uv; svv[dd] = a_vv;
            }
            if (graph) { su[0] = 1.0; sv[1] = 1.0; }
            double n[3] = {su[1] * sv[2] - su[2] * sv[1], su[2] * sv[0] - su[0] * sv[2], su[0] * sv[1] - su[1] * sv[0]};
            const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            double E = 0, F = 0, G = 0, L = 0, M = 0, Nn = 0;
            for (int d = 0; d < 3; ++d) {
                const double nd = n[d] / len;
                E = fma(su[d], su[d], E); F = fma(su[d], sv[d], F); G = fma(sv[d], sv[d], G);
                L = fma(suu[d], nd, L); M = fma(suv[d], nd, M); Nn = fma(svv[d], nd, Nn);
            }
            out[p] = (L * Nn - M * M) / (E * G - F * F);
        }
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

extern "C" int bspy_cuda_curvature_points(const bspy_spline *spline, const double *uvw, int64_t pointStride, int64_t varStride,
                                          int64_t N, double *out, int64_t *firstOutside, void *stream)
{
    SplineDev s;
    int rc = make_spline_dev(spline, s, "bspy_cuda_curvature_points");
    if (rc) return rc;
    if (!uvw || !out || N < 0) { set_error("bspy_cuda_curvature_points: bad argument"); return BSPY_E_ARG; }
    const bool shape = (s.nInd == 1 && s.nDep >= 1 && s.nDep <= 3) || (s.nInd == 2 && (s.nDep == 3 || s.nDep == 1));
    if (!shape || s.order[0] > CURV_MAXO || (s.nInd == 2 && s.order[1] > CURV_MAXO)) {
        set_error("bspy_cuda_curvature_points: curves of nDep 1..3 and surfaces of nDep 3 / 1 with orders <= %d", CURV_MAXO);
        return BSPY_E_UNSUPPORTED;
    }
    if (N == 0) return 0;
    long long blocks = (N + 127) / 128;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    curvature_points_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(s, uvw, pointStride, varStride, N, out,
                                                                                (long long *)firstOutside);
    count_launch();
    return check_launch("bspy_cuda_curvature_points");
}
