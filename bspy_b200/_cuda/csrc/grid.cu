// Regular-grid evaluation: bspy_cuda_eval_grid / bspy_cuda_eval_grid_batch.
//
// A tensor grid of parameters makes the evaluation a true dense contraction
//     out[d, a, b] = sum_ij  Bu[a][i] * C[d][su(a)-ou+i][sv(b)-ov+j] * Bv[b][j]
// (reference semantics: spline(*np.meshgrid(u, v, indexing="ij")), bspy/spline.py:940-947, plus
// jacobian / normal of bspy/_spline_evaluation.py:205-246 at every grid point).  For surfaces
// (nInd == 2) it runs on the FP64 tensor pipe: per warp an 8x8 output tile is
//     D(8x8) = A(8x4) * B(4x8),  A = T[a][j] = sum_i Bu[a][i] C[..i..][..j..]   (rows of the strip)
//                                B = Bv[b][j]^T                                   (columns of the tile)
// with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); K = 4 is exactly a cubic's window, higher orders take
// two K steps, lower orders are zero-padded.  value, d/du and d/dv are three accumulators that share
// A or B, the normal is formed in registers from the thread's own accumulator fragment, and every
// thread stores 16-byte pairs that are contiguous along the fastest (last-variable) axis.
// The kernel is bound by the HBM write stream (96 B per point for value+jacobian+normal of a 3-D
// surface against ~90 flops), so the design goal is store efficiency, not tensor utilisation.
//
// Everything else (nInd != 2, orders above 8, nDep above 4) goes through the scattered kernels in
// grid mode: parameters are decoded from the flat index, nothing is materialised.
#include "common.cuh"

namespace bspy {

// from scattered.cu
struct SplineDev {
    int nInd, nDep;
    int order[BSPY_MAX_IND];
    int nCoef[BSPY_MAX_IND];
    const double *knots[BSPY_MAX_IND];
    const double *coefs;
    long long stride[BSPY_MAX_IND];
    long long depStride;
    int normalSign;
};
struct PointsDev {
    const double *uvw;
    long long pointStride, varStride;
    const double *axes[BSPY_MAX_IND];
    long long nAxis[BSPY_MAX_IND];
    int grid;
};
int make_spline_dev(const bspy_spline *sp, SplineDev &s, const char *who);
int eval_common(const bspy_spline *spline, const PointsDev &in, long long N, const int32_t *wrt, uint32_t flags,
                uint32_t normalMask, double *values, double *deriv, double *jacobian, double *normal, int32_t *spans,
                int64_t *firstOutside, void *stream, const char *who);

constexpr int GRID_WARPS = 4;             // warps per CTA, one 8-row strip each
constexpr int GRID_ROWS = 8 * GRID_WARPS; // rows of a CTA tile
constexpr int GRID_COLS = 256;            // columns of a CTA tile
constexpr int GRID_COLS_PAD = GRID_COLS + 8;  // table row stride: == 8 (mod 16) doubles -> 2-wavefront B loads
constexpr int GRID_MAX_ORDER = 8;

struct Grid2Params {
    // spline batch: element s uses knots0 + s*knotStride0, knots1 + s*knotStride1, coefs + s*coefStride
    const double *knots0, *knots1, *coefs;
    long long knotStride0, knotStride1, coefStride;
    int ou, ov, nCu, nCv;      // orders and coefficient counts of the two variables
    long long depStride;       // nCu * nCv
    const double *axisU, *axisV;
    long long nU, nV;
    double *values, *jacobian, *normal;   // (S, nDep, nU, nV), (S, nDep, 2, nU, nV), (S, D, nU, nV)
    long long *firstOutside;
    int normalSign;
    unsigned normalize, normalMask;
    int vec2;                  // 16-byte stores allowed (nV even and bases 16-byte aligned)
    int colChunks, rowBlocks;
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// runtime-order recurrence, values and first derivatives, into strided table columns
__device__ __forceinline__ void axis_basis(const double *__restrict__ knots, int order, int ix, double u,
                                           double *__restrict__ val, double *__restrict__ der, int stride, int padTo)
{
    double b0[GRID_MAX_ORDER], b1[GRID_MAX_ORDER];
#pragma unroll
    for (int j = 0; j < GRID_MAX_ORDER; ++j) { b0[j] = 0.0; b1[j] = 0.0; }
    // slots are addressed from the top so that indices stay compile-time: slot s <-> b[GRID_MAX_ORDER-order+s]
    // (runtime `order` only shifts which stages execute)
    b0[GRID_MAX_ORDER - 1] = 1.0;
#pragma unroll
    for (int deg = 1; deg < GRID_MAX_ORDER; ++deg) {
        if (deg < order) {
            const bool lastStage = deg == order - 1;
            if (lastStage) {
#pragma unroll
                for (int j = 0; j < GRID_MAX_ORDER; ++j) b1[j] = b0[j];
            }
#pragma unroll
            for (int t = 0; t < deg; ++t) {
                const int slot = GRID_MAX_ORDER - deg + t;
                const double kl = __ldg(knots + ix - deg + t);
                const double r = 1.0 / (__ldg(knots + ix + t) - kl);
                const double a = (u - kl) * r;
                b0[slot - 1] += (1.0 - a) * b0[slot];
                b0[slot] *= a;
                if (lastStage) {
                    const double g = (double)deg * r;
                    b1[slot - 1] -= g * b1[slot];
                    b1[slot] *= g;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < GRID_MAX_ORDER; ++j) {
        const int s = j - (GRID_MAX_ORDER - order);   // slot within the order-long row
        if (s >= 0) {
            val[s * stride] = b0[j];
            der[s * stride] = order > 1 ? b1[j] : 0.0;
        }
    }
    for (int s = order; s < padTo; ++s) { val[s * stride] = 0.0; der[s * stride] = 0.0; }
}

template <int NDEP>
__global__ void __launch_bounds__(GRID_WARPS * 32) grid2_dmma_kernel(const Grid2Params P)
{
    constexpr int D = NDEP > 2 ? NDEP : 2;
    const int ksteps = (P.ov + 3) >> 2;
    const int kpad = ksteps * 4;
    // shared tables: V-axis basis [kind][j][col], U-axis basis [kind][i][row], spans
    extern __shared__ double sm[];
    double *tabV = sm;                                        // 2 * kpad * GRID_COLS_PAD
    double *tabU = tabV + 2 * kpad * GRID_COLS_PAD;           // 2 * GRID_MAX_ORDER * GRID_ROWS
    int *spanV = reinterpret_cast<int *>(tabU + 2 * GRID_MAX_ORDER * GRID_ROWS);  // GRID_COLS
    int *spanU = spanV + GRID_COLS;                           // GRID_ROWS

    const long long tile = blockIdx.x;
    const int cc = (int)(tile % P.colChunks);
    const int rb = (int)((tile / P.colChunks) % P.rowBlocks);
    const long long sIdx = tile / ((long long)P.colChunks * P.rowBlocks);
    const double *ku = P.knots0 + sIdx * P.knotStride0;
    const double *kv = P.knots1 + sIdx * P.knotStride1;
    const double *coefs = P.coefs + sIdx * P.coefStride;
    const long long row0 = (long long)rb * GRID_ROWS;
    const long long col0 = (long long)cc * GRID_COLS;
    const long long plane = P.nU * P.nV;

    // ---- per-tile axis tables ----
    for (int c = threadIdx.x; c < GRID_COLS; c += blockDim.x) {
        const long long b = col0 + c;
        int ix = P.ov;
        double v = 0.0;
        if (b < P.nV) {
            v = __ldg(P.axisV + b);
            ix = span_search_inner(kv, P.ov + P.nCv, P.ov, v);
            if ((v < __ldg(kv + P.ov - 1)) | (v > __ldg(kv + P.nCv)))
                if (P.firstOutside) report_outside((int64_t *)P.firstOutside, sIdx * plane + b);
        } else {
            v = __ldg(kv + P.ov - 1);
        }
        spanV[c] = ix;
        axis_basis(kv, P.ov, ix, v, tabV + c, tabV + kpad * GRID_COLS_PAD + c, GRID_COLS_PAD, kpad);
    }
    for (int r = threadIdx.x; r < GRID_ROWS; r += blockDim.x) {
        const long long a = row0 + r;
        int ix = P.ou;
        double u = 0.0;
        if (a < P.nU) {
            u = __ldg(P.axisU + a);
            ix = span_search_inner(ku, P.ou + P.nCu, P.ou, u);
            if ((u < __ldg(ku + P.ou - 1)) | (u > __ldg(ku + P.nCu)))
                if (P.firstOutside) report_outside((int64_t *)P.firstOutside, sIdx * plane + a * P.nV);
        } else {
            u = __ldg(ku + P.ou - 1);
        }
        spanU[r] = ix;
        axis_basis(ku, P.ou, ix, u, tabU + r, tabU + GRID_MAX_ORDER * GRID_ROWS + r, GRID_ROWS, P.ou);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane >> 2;        // row of the A fragment / column of the B fragment
    const int r4 = lane & 3;        // k index of the fragments
    const int myRow = warp * 8 + q;
    const long long a = row0 + myRow;
    const int su = spanU[myRow];
    // this thread's u-basis row (values and derivatives)
    double bu[GRID_MAX_ORDER], dbu[GRID_MAX_ORDER];
#pragma unroll
    for (int i = 0; i < GRID_MAX_ORDER; ++i) {
        bu[i] = i < P.ou ? tabU[i * GRID_ROWS + myRow] : 0.0;
        dbu[i] = i < P.ou ? tabU[(GRID_MAX_ORDER + i) * GRID_ROWS + myRow] : 0.0;
    }
    // A fragments for the cached v-span: T[d][kind][ks]
    double tv[NDEP][2], td[NDEP][2];
    int cached = -1;

    const long long nColsHere = min((long long)GRID_COLS, P.nV - col0);
    for (int c0 = 0; c0 < nColsHere; c0 += 8) {
        const int myCol = c0 + q;
        const int sv = spanV[myCol];
        double acc[NDEP][3][2];
#pragma unroll
        for (int d = 0; d < NDEP; ++d)
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[d][k][0] = acc[d][k][1] = 0.0;
        unsigned todo = 0xffffffffu;
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int cur = __shfl_sync(0xffffffffu, sv, leader);
            const bool mine = sv == cur;
            if (cur != cached) {
                // T[a][j] = sum_i Bu[a][i] * C[d][su-ou+i][cur-ov+j],  j = r4 + 4 ks
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const int j = r4 + 4 * ks;
#pragma unroll
                    for (int d = 0; d < NDEP; ++d) {
                        double s0 = 0.0, s1 = 0.0;
                        if (ks < ksteps && j < P.ov) {
                            const double *cp = coefs + d * P.depStride + (long long)(su - P.ou) * P.nCv + (cur - P.ov + j);
#pragma unroll
                            for (int i = 0; i < GRID_MAX_ORDER; ++i)
                                if (i < P.ou) {
                                    const double x = __ldg(cp + (long long)i * P.nCv);
                                    s0 = fma(x, bu[i], s0);
                                    s1 = fma(x, dbu[i], s1);
                                }
                        }
                        tv[d][ks] = s0;
                        td[d][ks] = s1;
                    }
                }
                cached = cur;
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                if (ks < ksteps) {
                    const int j = r4 + 4 * ks;
                    const double bv = mine ? tabV[j * GRID_COLS_PAD + myCol] : 0.0;
                    const double dbv = mine ? tabV[(kpad + j) * GRID_COLS_PAD + myCol] : 0.0;
#pragma unroll
                    for (int d = 0; d < NDEP; ++d) {
                        dmma884(acc[d][0][0], acc[d][0][1], tv[d][ks], bv);
                        dmma884(acc[d][1][0], acc[d][1][1], td[d][ks], bv);
                        dmma884(acc[d][2][0], acc[d][2][1], tv[d][ks], dbv);
                    }
                }
            }
            todo &= ~__ballot_sync(0xffffffffu, mine);
        }
        // ---- epilogue: this thread owns row `a`, columns col0 + c0 + 2*r4 + {0,1} ----
        const long long b = col0 + c0 + 2 * r4;
        if (a < P.nU && b < P.nV) {
            const bool two = b + 1 < P.nV;
            const long long at = a * P.nV + b;
            auto put = [&](double *base, double x0, double x1) {
                if (P.vec2 && two) {
                    __stcs(reinterpret_cast<double2 *>(base + at), make_double2(x0, x1));
                } else {
                    __stcs(base + at, x0);
                    if (two) __stcs(base + at + 1, x1);
                }
            };
            if (P.values) {
#pragma unroll
                for (int d = 0; d < NDEP; ++d) put(P.values + (sIdx * NDEP + d) * plane, acc[d][0][0], acc[d][0][1]);
            }
            if (P.jacobian) {
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    put(P.jacobian + ((sIdx * NDEP + d) * 2 + 0) * plane, acc[d][1][0], acc[d][1][1]);
                    put(P.jacobian + ((sIdx * NDEP + d) * 2 + 1) * plane, acc[d][2][0], acc[d][2][1]);
                }
            }
            if constexpr (NDEP == 3 || NDEP == 1) {
                if (P.normal) {
                    double n[2][D];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if constexpr (NDEP == 3) {
                            const double ux = acc[0][1][e], uy = acc[1][1][e], uz = acc[2][1][e];
                            const double vx = acc[0][2][e], vy = acc[1][2][e], vz = acc[2][2][e];
                            n[e][0] = (uy * vz - uz * vy) * P.normalSign;
                            n[e][1] = -(ux * vz - uz * vx) * P.normalSign;
                            n[e][2] = (ux * vy - uy * vx) * P.normalSign;
                        } else {
                            // nInd 2 > nDep 1: T = J^T is 2x1, n = (dv, -du) * sign
                            n[e][0] = acc[0][2][e] * P.normalSign;
                            n[e][1] = -acc[0][1][e] * P.normalSign;
                        }
                        if (P.normalize) {
                            double sq = 0.0;
#pragma unroll
                            for (int i = 0; i < D; ++i)
                                if (P.normalMask & (1u << i)) sq += n[e][i] * n[e][i];
                            const double len = sqrt(sq);
#pragma unroll
                            for (int i = 0; i < D; ++i) n[e][i] = n[e][i] / len;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < D; ++i) put(P.normal + (sIdx * D + i) * plane, n[0][i], n[1][i]);
                }
            }
        }
    }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int NDEP>
static int launch_grid2(const Grid2Params &P, long long nSplines, cudaStream_t stream)
{
    const int kpad = ((P.ov + 3) / 4) * 4;
    const size_t smem = sizeof(double) * (2 * kpad * GRID_COLS_PAD + 2 * GRID_MAX_ORDER * GRID_ROWS) +
                        sizeof(int) * (GRID_COLS + GRID_ROWS);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(grid2_dmma_kernel<NDEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const long long tiles = nSplines * P.colChunks * P.rowBlocks;
    if (tiles > 0x7fffffffLL) { set_error("grid too large for one launch"); return BSPY_E_UNSUPPORTED; }
    grid2_dmma_kernel<NDEP><<<(unsigned)tiles, GRID_WARPS * 32, smem, stream>>>(P);
    count_launch();
    return check_launch("bspy_cuda_eval_grid");
}

// surfaces with orders <= 8 and nDep <= 4 take the tensor path
static bool grid2_supported(int nInd, int nDep, const int32_t *order)
{
    return nInd == 2 && nDep >= 1 && nDep <= 4 && order[0] <= GRID_MAX_ORDER && order[1] <= GRID_MAX_ORDER;
}

static int grid2_run(const bspy_spline *sp, long long nSplines, long long knotStride0, long long knotStride1,
                     long long coefStride, const double *const *axes, const int64_t *nAxis, uint32_t flags,
                     uint32_t normalMask, double *values, double *jacobian, double *normal, int64_t *firstOutside,
                     cudaStream_t stream)
{
    Grid2Params P{};
    P.knots0 = sp->knots[0]; P.knots1 = sp->knots[1]; P.coefs = sp->coefs;
    P.knotStride0 = knotStride0; P.knotStride1 = knotStride1; P.coefStride = coefStride;
    P.ou = sp->order[0]; P.ov = sp->order[1]; P.nCu = sp->nCoef[0]; P.nCv = sp->nCoef[1];
    P.depStride = (long long)P.nCu * P.nCv;
    P.axisU = axes[0]; P.axisV = axes[1]; P.nU = nAxis[0]; P.nV = nAxis[1];
    P.values = values; P.jacobian = jacobian; P.normal = normal;
    P.firstOutside = (long long *)firstOutside;
    P.normalSign = sp->normalSign < 0 ? -1 : 1;
    P.normalize = (flags & BSPY_NORMALIZE) ? 1u : 0u;
    P.normalMask = normalMask ? normalMask : 0xffffffffu;
    P.vec2 = (P.nV % 2 == 0) && aligned16(values) && aligned16(jacobian) && aligned16(normal);
    P.colChunks = (int)((P.nV + GRID_COLS - 1) / GRID_COLS);
    P.rowBlocks = (int)((P.nU + GRID_ROWS - 1) / GRID_ROWS);
    if (P.nU == 0 || P.nV == 0 || nSplines == 0) return 0;
    switch (sp->nDep) {
        case 1: return launch_grid2<1>(P, nSplines, stream);
        case 2: return launch_grid2<2>(P, nSplines, stream);
        case 3: return launch_grid2<3>(P, nSplines, stream);
        default: return launch_grid2<4>(P, nSplines, stream);
    }
}

}  // namespace bspy

using namespace bspy;

extern "C" int bspy_cuda_eval_grid(const bspy_spline *spline, const double *const *axes, const int64_t *nAxis,
                                   uint32_t flags, uint32_t normalMask, double *values, double *jacobian, double *normal,
                                   int64_t *firstOutside, void *stream)
{
    if (!spline || (spline->nInd > 0 && (!axes || !nAxis))) {
        set_error("bspy_cuda_eval_grid: NULL argument");
        return BSPY_E_ARG;
    }
    if (spline->nInd > BSPY_MAX_IND) { set_error("bspy_cuda_eval_grid: nInd too large"); return BSPY_E_UNSUPPORTED; }
    long long N = 1;
    for (int i = 0; i < spline->nInd; ++i) {
        if (nAxis[i] < 0 || (nAxis[i] > 0 && !axes[i])) { set_error("bspy_cuda_eval_grid: bad axis %d", i); return BSPY_E_ARG; }
        N *= nAxis[i];
    }
    if (normal && (spline->nInd - spline->nDep != 1 && spline->nDep - spline->nInd != 1)) {
        set_error("The number of independent variables must be one different than the number of dependent variables.");
        return BSPY_E_NORMAL_DIMS;
    }
    if (grid2_supported(spline->nInd, spline->nDep, spline->order) && (values || jacobian || normal)) {
        SplineDev chk;
        int rc = make_spline_dev(spline, chk, "bspy_cuda_eval_grid");
        if (rc) return rc;
        return grid2_run(spline, 1, 0, 0, 0, axes, nAxis, flags, normalMask, values, jacobian, normal, firstOutside,
                         (cudaStream_t)stream);
    }
    PointsDev in{};
    in.grid = 1;
    for (int i = 0; i < spline->nInd; ++i) { in.axes[i] = axes[i]; in.nAxis[i] = nAxis[i]; }
    return eval_common(spline, in, N, nullptr, flags, normalMask, values, nullptr, jacobian, normal, nullptr, firstOutside,
                       stream, "bspy_cuda_eval_grid");
}

extern "C" int bspy_cuda_eval_grid_batch(const bspy_spline *first, int64_t nSplines, const int64_t *knotStride,
                                         int64_t coefStride, const double *const *axes, const int64_t *nAxis,
                                         uint32_t flags, uint32_t normalMask, double *values, double *jacobian,
                                         double *normal, int64_t *firstOutside, void *stream)
{
    if (!first || !axes || !nAxis || !knotStride || nSplines < 0) {
        set_error("bspy_cuda_eval_grid_batch: NULL argument");
        return BSPY_E_ARG;
    }
    if (!grid2_supported(first->nInd, first->nDep, first->order)) {
        set_error("bspy_cuda_eval_grid_batch: only surfaces (nInd == 2, nDep <= 4, orders <= 8) are batched");
        return BSPY_E_UNSUPPORTED;
    }
    if (normal && first->nDep != 3 && first->nDep != 1) {
        set_error("The number of independent variables must be one different than the number of dependent variables.");
        return BSPY_E_NORMAL_DIMS;
    }
    SplineDev chk;
    int rc = make_spline_dev(first, chk, "bspy_cuda_eval_grid_batch");
    if (rc) return rc;
    return grid2_run(first, nSplines, knotStride[0], knotStride[1], coefStride, axes, nAxis, flags, normalMask, values,
                     jacobian, normal, firstOutside, (cudaStream_t)stream);
}
