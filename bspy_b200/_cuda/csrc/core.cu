// libbspy_cuda.so: error handling, launch accounting, spans and (bit-exact) basis kernels.
#include <stdarg.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace bspy {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- experiment switches --------------------------------------------------------------------------
static const char *const kOptionNames[OPT_COUNT] = {
    "CURVE_REPL", "CURVE_PPT", "BIN_MODE", "BIN_OVERLAP", "STAGED", "DEP_TILE", "SPAN_RECORDS", "BIN_CHUNK",
    "BIN_REC_CHUNK_LOG2", "GRID_CHUNK", "GRID_ROWS", "GRID_GROUP", "CELL_KERNEL", "CURVE_TMA", "MANY_MODE",
    "GRID3_ROWS", "GRID3_CHUNK", "EXP_A", "EXP_B", "IMAGE", "STAGED_PAIR", "STAGED_WAVES", "CURVE_POLY", "CELL_POLY", "BIN_PERM",
};
constexpr long long kOptionUnset = INT64_MIN;
static std::atomic<long long> g_options[OPT_COUNT];
static std::once_flag g_optionsOnce;

static void options_from_environment()
{
    for (int i = 0; i < OPT_COUNT; ++i) {
        char name[64];
        snprintf(name, sizeof name, "BSPY_%s", kOptionNames[i]);
        const char *e = getenv(name);
        g_options[i].store(e && *e ? atoll(e) : kOptionUnset, std::memory_order_relaxed);
    }
}

long long option(Option o, long long unset)
{
    std::call_once(g_optionsOnce, options_from_environment);
    const long long v = g_options[o].load(std::memory_order_relaxed);
    return v == kOptionUnset ? unset : v;
}

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// ---- spans ------------------------------------------------------------------------------------
// One thread per parameter, grid-stride; knots are staged in shared memory when they fit so the
// log2(nKnots) probes of every thread stay on-chip.
__global__ void __launch_bounds__(256) spans_kernel(const double *__restrict__ knots, int nKnots, int order,
                                                    const double *__restrict__ u, int64_t N, int32_t *__restrict__ out,
                                                    int useSmem)
{
    extern __shared__ double sk[];
    const double *k = knots;
    if (useSmem) {
        for (int i = threadIdx.x; i < nKnots; i += blockDim.x) sk[i] = knots[i];
        __syncthreads();
        k = sk;
    }
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        const double x = u[p];
        int lo = 0, n = nKnots;
        if (x != x) {
            lo = nKnots;
        } else {
            while (n > 0) {
                const int half = n >> 1;
                const bool le = k[lo + half] <= x;
                lo = le ? lo + half + 1 : lo;
                n = le ? n - half - 1 : half;
            }
        }
        out[p] = min(max(lo, order), nKnots - order);
    }
}

// ---- basis (strict) ---------------------------------------------------------------------------
// Thread per parameter.  The `order` running values live in shared memory, one column per thread
// ([slot][thread], conflict-free), because `order` is a runtime value up to BSPY_MAX_ORDER.
struct SmemColumn {
    double *base;
    int stride;
    __device__ __forceinline__ double &operator()(int j) const { return base[j * stride]; }
};

__global__ void __launch_bounds__(128) basis_kernel(const double *__restrict__ knots, int nKnots, int order,
                                                    const double *__restrict__ u, const int32_t *__restrict__ spansIn,
                                                    int64_t N, int deriv, int taylor, int32_t *__restrict__ spansOut,
                                                    double *__restrict__ basis)
{
    extern __shared__ double scratch[];
    SmemColumn col{scratch + threadIdx.x, (int)blockDim.x};
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        const double x = u[p];
        const int ix = spansIn ? spansIn[p] : span_search(knots, nKnots, order, x);
        if (spansOut) spansOut[p] = ix;
        // a caller-supplied span whose knot window knots[ix-order+1 .. ix+order-2] leaves the array is not evaluated
        // (the host wrapper raises IndexError for it, like numpy indexing would): NaN row, no out-of-bounds read
        if (ix < order - 1 || ix > nKnots - order + 1) {
            for (int j = 0; j < order; ++j) basis[p * order + j] = __longlong_as_double(0x7ff8000000000000LL);
            continue;
        }
        basis_strict(knots, order, ix, x, deriv, taylor != 0, col);
        for (int j = 0; j < order; ++j) basis[p * order + j] = col(j);
    }
}

// ---- curvature ------------------------------------------------------------------------------
// bspy/_spline_evaluation.py:80-107 from batched derivatives.  Curves: kappa = numerator / (f'.f')^1.5 with the
// signed 2-D cross product for planar curves and sqrt(|f''|^2 |f'|^2 - (f'.f'')^2) otherwise.  Surfaces: Gaussian
// curvature (L N - M^2) / (E G - F^2) from the first and second fundamental forms.  graph != 0: the spline is a
// scalar function and the curve / surface is its graph (x(u) = u: first derivative 1, second 0).
__global__ void __launch_bounds__(256) curvature_kernel(int nInd, int nDep, int graph, long long N, const double *__restrict__ d1,
                                                        const double *__restrict__ d2, const double *__restrict__ normal,
                                                        double *__restrict__ out)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
        if (nInd == 1) {
            double pp = graph ? 1.0 : 0.0, pq = 0.0, qq = 0.0;     // f'.f', f'.f'', f''.f''
            for (int d = 0; d < nDep; ++d) {
                const double a = d1[d * N + p], b = d2[d * N + p];
                pp = fma(a, a, pp);
                pq = fma(a, b, pq);
                qq = fma(b, b, qq);
            }
            double num;
            if (graph)                       // (1, f') x (0, f'')
                num = d2[p];
            else if (nDep == 2)
                num = d1[p] * d2[N + p] - d1[N + p] * d2[p];
            else
                num = sqrt(qq * pp - pq * pq);
            out[p] = num / (pp * sqrt(pp));
        } else {
            double su[3], sv[3], suu[3], suv[3], svv[3], n[3];
            if (graph) {
                const double fu = d1[p], fv = d1[N + p];
                su[0] = 1.0; su[1] = 0.0; su[2] = fu;
                sv[0] = 0.0; sv[1] = 1.0; sv[2] = fv;
                suu[0] = suu[1] = suv[0] = suv[1] = svv[0] = svv[1] = 0.0;
                suu[2] = d2[p]; suv[2] = d2[N + p]; svv[2] = d2[2 * N + p];
                const double len = sqrt(fu * fu + fv * fv + 1.0);
                n[0] = -fu / len; n[1] = -fv / len; n[2] = 1.0 / len;
            } else {
                for (int d = 0; d < 3; ++d) {
                    su[d] = d1[(d * 2 + 0) * N + p];
                    sv[d] = d1[(d * 2 + 1) * N + p];
                    suu[d] = d2[(0 * 3 + d) * N + p];
                    suv[d] = d2[(1 * 3 + d) * N + p];
                    svv[d] = d2[(2 * 3 + d) * N + p];
                    n[d] = normal[d * N + p];
                }
            }
            double E = 0, F = 0, G = 0, L = 0, M = 0, Nn = 0;
            for (int d = 0; d < 3; ++d) {
                E = fma(su[d], su[d], E); F = fma(su[d], sv[d], F); G = fma(sv[d], sv[d], G);
                L = fma(suu[d], n[d], L); M = fma(suv[d], n[d], M); Nn = fma(svv[d], n[d], Nn);
            }
            out[p] = (L * Nn - M * M) / (E * G - F * F);
        }
    }
}

}  // namespace bspy

using namespace bspy;

extern "C" int bspy_cuda_curvature(int32_t nInd, int32_t nDep, int32_t graph, int64_t N, const double *d1, const double *d2,
                                   const double *normal, double *out, void *stream)
{
    if (!d1 || !d2 || !out || N < 0 || nDep < 1 || (nInd != 1 && nInd != 2)) {
        set_error("bspy_cuda_curvature: bad argument");
        return BSPY_E_ARG;
    }
    if (graph && nDep != 1) { set_error("bspy_cuda_curvature: graph mode needs nDep == 1"); return BSPY_E_ARG; }
    if (nInd == 2 && !graph && (nDep != 3 || !normal)) {
        set_error("bspy_cuda_curvature: surfaces need nDep == 3 (or a scalar function) and their unit normals");
        return BSPY_E_UNSUPPORTED;
    }
    if (N == 0) return 0;
    long long blocks = (N + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    curvature_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(nInd, nDep, graph, N, d1, d2, normal, out);
    count_launch();
    return check_launch("bspy_cuda_curvature");
}

extern "C" {

int bspy_cuda_abi_version(void) { return BSPY_ABI_VERSION; }

const char *bspy_cuda_last_error_string(void) { return g_err; }

int64_t bspy_cuda_launch_count(void) { return (int64_t)g_launches.load(); }

int bspy_cuda_copy_2d(void *dst, int64_t dstPitchBytes, const void *src, int64_t srcPitchBytes, int64_t widthBytes,
                      int64_t rows, void *stream)
{
    if (!dst || !src || widthBytes < 0 || rows < 0 || dstPitchBytes < widthBytes || srcPitchBytes < widthBytes) {
        set_error("bspy_cuda_copy_2d: bad argument");
        return BSPY_E_ARG;
    }
    if (widthBytes == 0 || rows == 0) return 0;
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dstPitchBytes, src, (size_t)srcPitchBytes, (size_t)widthBytes, (size_t)rows,
                                      cudaMemcpyDefault, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("cudaMemcpy2DAsync: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int bspy_cuda_set_option(const char *name, int64_t value, int32_t isSet)
{
    if (!name) { set_error("bspy_cuda_set_option: name is NULL"); return BSPY_E_ARG; }
    option(OPT_COUNT == 0 ? OPT_CURVE_REPL : OPT_CURVE_REPL, 0);   // environment defaults are read first
    for (int i = 0; i < OPT_COUNT; ++i) {
        if (strcmp(name, kOptionNames[i]) == 0) {
            g_options[i].store(isSet ? (long long)value : kOptionUnset, std::memory_order_relaxed);
            return 0;
        }
    }
    set_error("bspy_cuda_set_option: unknown option %s", name);
    return BSPY_E_ARG;
}

int bspy_cuda_spans(const double *knots, int32_t nKnots, int32_t order, const double *u, int64_t N,
                    int32_t *spans, void *stream)
{
    if (!knots || !u || !spans || N < 0 || order < 1 || nKnots < 2 * order) {
        set_error("bspy_cuda_spans: bad argument");
        return BSPY_E_ARG;
    }
    if (N == 0) return 0;
    const int threads = 256;
    int64_t blocks = (N + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)nKnots * sizeof(double);
    const int useSmem = smem <= 48 * 1024;
    spans_kernel<<<(unsigned)blocks, threads, useSmem ? smem : 0, (cudaStream_t)stream>>>(knots, nKnots, order, u, N,
                                                                                          spans, useSmem);
    count_launch();
    return check_launch("bspy_cuda_spans");
}

int bspy_cuda_basis(const double *knots, int32_t nKnots, int32_t order, const double *u, const int32_t *spansIn,
                    int64_t N, int32_t derivOrder, int32_t taylorCoefs, int32_t *spansOut, double *basis, void *stream)
{
    if (!knots || !u || !basis || N < 0 || order < 1 || nKnots < 2 * order || derivOrder < 0) {
        set_error("bspy_cuda_basis: bad argument");
        return BSPY_E_ARG;
    }
    if (order > BSPY_MAX_ORDER) {
        set_error("bspy_cuda_basis: order %d > BSPY_MAX_ORDER", order);
        return BSPY_E_UNSUPPORTED;
    }
    if (N == 0) return 0;
    const int threads = 128;
    int64_t blocks = (N + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)order * threads * sizeof(double);
    basis_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(knots, nKnots, order, u, spansIn, N, derivOrder,
                                                                          taylorCoefs, spansOut, basis);
    count_launch();
    return check_launch("bspy_cuda_basis");
}

}  // extern "C"
