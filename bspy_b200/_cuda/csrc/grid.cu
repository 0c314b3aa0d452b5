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
#include <stdlib.h>

#include "common.cuh"

namespace bspy {

// from scattered.cu
int make_spline_dev(const bspy_spline *sp, SplineDev &s, const char *who);
int eval_common(const bspy_spline *spline, const PointsDev &in, long long N, const int32_t *wrt, uint32_t flags,
                uint32_t normalMask, double *values, double *deriv, double *jacobian, double *normal, int32_t *spans,
                int64_t *firstOutside, void *stream, const char *who);

constexpr int GRID_WARPS = 8;              // warps per CTA, side by side along the row: 16 columns each per step
constexpr int GRID_ROWS = 8;               // rows of a work unit (one MMA row strip)
constexpr int GRID_STEP = 16 * GRID_WARPS; // columns a CTA covers per step: 128 (1 KB contiguous per row and plane)
constexpr int GRID_TILE_ROWS = 64;         // rows of a CTA tile (8 strips)
constexpr int GRID_GROUP = 4;              // splines of a batch that share one CTA's axis tables (shared knots only)
constexpr int GRID_MAX_ORDER = 8;

struct Grid2Params {
    // spline batch: element s uses knots0 + s*knotStride0, knots1 + s*knotStride1, coefs + s*coefStride
    const double *knots0, *knots1, *coefs;
    long long knotStride0, knotStride1, coefStride;
    long long nSplines;
    int ou, ov, nCu, nCv;      // orders and coefficient counts of the two variables
    long long depStride;       // nCu * nCv
    const double *axisU, *axisV;
    long long nU, nV;
    double *values, *jacobian, *normal;   // (S, nDep, nU, nV), (S, nDep, 2, nU, nV), (S, D, nU, nV)
    long long *firstOutside;
    int normalSign;
    unsigned normalize, normalMask;
    int vec;                   // widest store allowed: 4 (32 B), 2 (16 B) or 1 doubles
    int chunkCols;             // columns of a CTA's chunk (multiple of GRID_STEP); table row stride = chunkCols + 2
    int colChunks;             // ceil(nV / chunkCols)
    int tileRows;              // rows of a CTA tile (multiple of 8, <= GRID_TILE_ROWS)
    long long rowBlocks;       // ceil(nU / tileRows)
    int group;                 // splines per CTA
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void st_cs_v4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// Runtime order <= MAXO: values (b0) and first derivatives (b1) in registers.  Rows are addressed from the
// top so that every index is a compile-time constant: slot s of the order-long row is b[MAXO - order + s].
template <int MAXO>
__device__ __forceinline__ void axis_basis(const double *__restrict__ knots, int order, int ix, double u,
                                           double (&b0)[MAXO], double (&b1)[MAXO])
{
#pragma unroll
    for (int j = 0; j < MAXO; ++j) { b0[j] = 0.0; b1[j] = 0.0; }
    b0[MAXO - 1] = 1.0;
#pragma unroll
    for (int deg = 1; deg < MAXO; ++deg) {
        if (deg < order) {
            const bool lastStage = deg == order - 1;
            if (lastStage) {
#pragma unroll
                for (int j = 0; j < MAXO; ++j) b1[j] = b0[j];
            }
#pragma unroll
            for (int t = 0; t < deg; ++t) {
                const int slot = MAXO - deg + t;
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
}

// CTA = (column chunk of chunkCols columns) x (block of GRID_TILE_ROWS rows) x (group of splines that share the
// axis tables).  The 8 warps work SIDE BY SIDE on one 8-row strip (16 columns each per step, so the CTA writes
// 1 KB contiguous per row and output plane per step), finish the strip's chunk, then move to the next strip.
// Measured with a pure store-pattern probe (tools/pattern_probe*.py): HBM sustains the most when the set of
// rows being written at any moment is small and each row receives long contiguous runs; one warp per strip
// sweeping along the row (the previous layout) topped out at 5.0 TB/s of the 6.5 TB/s a linear fill reaches.
// F32: the outputs are float arrays of the same shapes (computed in float64, rounded to nearest on the store): the
// tessellation path of the reference's viewer, which hands float32 buffers to OpenGL (bspy/splineOpenGLFrame.py:1461-1513);
// half the bytes on a kernel that is bound by its store stream.
template <int NDEP, int MAXO, bool F32 = false>
__global__ void __launch_bounds__(GRID_WARPS * 32, 2) grid2_dmma_kernel(const Grid2Params P)
{
    constexpr int D = NDEP > 2 ? NDEP : 2;
    constexpr int KS = MAXO / 4;
    extern __shared__ double sm[];
    // table row stride == 2 (mod 16) doubles: the B-fragment loads of a warp take the minimum 2 wavefronts
    const int VS = P.chunkCols + 2;
    double *tabV = sm;                                          // [kind][j][col]: 2 * MAXO * VS
    double *tabU = tabV + 2 * MAXO * VS;                        // [kind][i][row]: 2 * MAXO * GRID_TILE_ROWS
    int *spanV = reinterpret_cast<int *>(tabU + 2 * MAXO * GRID_TILE_ROWS);  // chunkCols
    int *spanU = spanV + P.chunkCols;                           // GRID_TILE_ROWS

    const long long tile = blockIdx.x;
    const int cc = (int)(tile % P.colChunks);
    const long long rb = (tile / P.colChunks) % P.rowBlocks;
    const long long s0 = tile / ((long long)P.colChunks * P.rowBlocks) * P.group;
    const long long s1 = min(s0 + P.group, P.nSplines);
    const double *ku = P.knots0 + s0 * P.knotStride0;
    const double *kv = P.knots1 + s0 * P.knotStride1;
    const long long row0 = rb * P.tileRows;
    const long long col0 = (long long)cc * P.chunkCols;
    const long long plane = P.nU * P.nV;

    // ---- axis tables of this tile (shared by the `group` splines of this CTA: their knots are shared) ----
    for (int c = threadIdx.x; c < P.chunkCols; c += blockDim.x) {
        const long long b = col0 + c;
        int ix = P.ov;
        double v = __ldg(kv + P.ov - 1);
        if (b < P.nV) {
            v = __ldg(P.axisV + b);
            ix = span_search_inner(kv, P.ov + P.nCv, P.ov, v);
            if ((v < __ldg(kv + P.ov - 1)) | (v > __ldg(kv + P.nCv)))
                if (P.firstOutside) report_outside((int64_t *)P.firstOutside, s0 * plane + b);
        }
        spanV[c] = ix;
        double b0[MAXO], b1[MAXO];
        axis_basis<MAXO>(kv, P.ov, ix, v, b0, b1);
#pragma unroll
        for (int j = 0; j < MAXO; ++j) {
            const int sl = j - (MAXO - P.ov);
            const int row = sl >= 0 ? sl : P.ov + j;     // slots ov .. MAXO-1 are the zero padding of K
            tabV[row * VS + c] = sl >= 0 ? b0[j] : 0.0;
            tabV[(MAXO + row) * VS + c] = sl >= 0 ? b1[j] : 0.0;
        }
    }
    for (int r = threadIdx.x; r < P.tileRows; r += blockDim.x) {
        const long long a = row0 + r;
        int ix = P.ou;
        double u = __ldg(ku + P.ou - 1);
        if (a < P.nU) {
            u = __ldg(P.axisU + a);
            ix = span_search_inner(ku, P.ou + P.nCu, P.ou, u);
            if (((u < __ldg(ku + P.ou - 1)) | (u > __ldg(ku + P.nCu))) && cc == 0)
                if (P.firstOutside) report_outside((int64_t *)P.firstOutside, s0 * plane + a * P.nV);
        }
        spanU[r] = ix;
        double b0[MAXO], b1[MAXO];
        axis_basis<MAXO>(ku, P.ou, ix, u, b0, b1);
#pragma unroll
        for (int j = 0; j < MAXO; ++j) {
            const int sl = j - (MAXO - P.ou);
            const int row = sl >= 0 ? sl : P.ou + j;
            tabU[row * GRID_TILE_ROWS + r] = sl >= 0 ? b0[j] : 0.0;
            tabU[(MAXO + row) * GRID_TILE_ROWS + r] = sl >= 0 ? b1[j] : 0.0;
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane >> 2;        // row of the A fragment / column of the B fragment
    const int r4 = lane & 3;        // k index of the fragments
    // B-fragment column n = q of MMA tile t is grid column c0 + 4*(q>>1) + 2*t + (q&1): the four accumulator
    // entries a thread ends up with (two per tile) are then the four CONSECUTIVE columns c0 + 4*r4 .. +3, so
    // every thread stores 32 contiguous bytes and a quad covers a full 128-byte line of the row.
    const int bcol = 4 * (q >> 1) + (q & 1);
    const long long nColsHere = min((long long)P.chunkCols, P.nV - col0);
    const int nStrips = (int)min((long long)(P.tileRows / 8), (P.nU - row0 + 7) / 8);

    for (long long sIdx = s0; sIdx < s1; ++sIdx)
    for (int strip = 0; strip < nStrips; ++strip) {
        const double *coefs = P.coefs + sIdx * P.coefStride;
        const int myRow = strip * 8 + q;
        const long long a = row0 + myRow;
        const int su = spanU[myRow];
        double tv[NDEP][KS], td[NDEP][KS];   // A fragments for the cached v-span
        int cached = -1;

        for (int c0 = warp * 16; c0 < nColsHere; c0 += GRID_STEP) {
            const int colA = c0 + bcol, colB = colA + 2;
            const int svA = spanV[colA], svB = spanV[colB];
            double acc[NDEP][3][4];
#pragma unroll
            for (int d = 0; d < NDEP; ++d)
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[d][k][e] = 0.0;
            unsigned todoA = 0xffffffffu, todoB = 0xffffffffu;
            while (todoA | todoB) {
                const int cur = todoA ? __shfl_sync(0xffffffffu, svA, __ffs(todoA) - 1)
                                      : __shfl_sync(0xffffffffu, svB, __ffs(todoB) - 1);
                const bool mineA = svA == cur, mineB = svB == cur;
                if (cur != cached) {
                    // T[a][j] = sum_i Bu[a][i] * C[d][su-ou+i][cur-ov+j],  j = r4 + 4 ks
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const int j = r4 + 4 * ks;
#pragma unroll
                        for (int d = 0; d < NDEP; ++d) {
                            double t0 = 0.0, t1 = 0.0;
                            if (j < P.ov) {
                                const double *cp = coefs + d * P.depStride + (long long)(su - P.ou) * P.nCv + (cur - P.ov + j);
#pragma unroll
                                for (int i = 0; i < MAXO; ++i)
                                    if (i < P.ou) {
                                        const double x = __ldg(cp + (long long)i * P.nCv);
                                        t0 = fma(x, tabU[i * GRID_TILE_ROWS + myRow], t0);
                                        t1 = fma(x, tabU[(MAXO + i) * GRID_TILE_ROWS + myRow], t1);
                                    }
                            }
                            tv[d][ks] = t0;
                            td[d][ks] = t1;
                        }
                    }
                    cached = cur;
                }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int j = r4 + 4 * ks;
                    const double bvA = mineA ? tabV[j * VS + colA] : 0.0;
                    const double dvA = mineA ? tabV[(MAXO + j) * VS + colA] : 0.0;
                    const double bvB = mineB ? tabV[j * VS + colB] : 0.0;
                    const double dvB = mineB ? tabV[(MAXO + j) * VS + colB] : 0.0;
                    if (P.values) {
#pragma unroll
                        for (int d = 0; d < NDEP; ++d) {
                            dmma884(acc[d][0][0], acc[d][0][1], tv[d][ks], bvA);
                            dmma884(acc[d][0][2], acc[d][0][3], tv[d][ks], bvB);
                        }
                    }
                    if (P.jacobian || P.normal) {
#pragma unroll
                        for (int d = 0; d < NDEP; ++d) {
                            dmma884(acc[d][1][0], acc[d][1][1], td[d][ks], bvA);
                            dmma884(acc[d][1][2], acc[d][1][3], td[d][ks], bvB);
                            dmma884(acc[d][2][0], acc[d][2][1], tv[d][ks], dvA);
                            dmma884(acc[d][2][2], acc[d][2][3], tv[d][ks], dvB);
                        }
                    }
                }
                todoA &= ~__ballot_sync(0xffffffffu, mineA);
                todoB &= ~__ballot_sync(0xffffffffu, mineB);
            }
            // ---- epilogue: this thread owns row `a`, columns col0 + c0 + 4*r4 + {0,1,2,3} ----
            const long long b = col0 + c0 + 4 * r4;
            if (a < P.nU && b < P.nV) {
                const int left = (int)min((long long)4, P.nV - b);
                const long long at = a * P.nV + b;
                auto put = [&](double *base, const long long planeIdx, const double (&x)[4]) {
                    if constexpr (F32) {
                        float *p = reinterpret_cast<float *>(base) + planeIdx * plane + at;
                        if (P.vec == 4 && left == 4) {
                            __stcs(reinterpret_cast<float4 *>(p), make_float4(__double2float_rn(x[0]), __double2float_rn(x[1]),
                                                                              __double2float_rn(x[2]), __double2float_rn(x[3])));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (e < left) __stcs(p + e, __double2float_rn(x[e]));
                        }
                    } else {
                        double *p = base + planeIdx * plane + at;
                        if (P.vec == 4 && left == 4) {
                            st_cs_v4(p, x[0], x[1], x[2], x[3]);
                        } else if (P.vec >= 2 && left == 4) {
                            __stcs(reinterpret_cast<double2 *>(p), make_double2(x[0], x[1]));
                            __stcs(reinterpret_cast<double2 *>(p + 2), make_double2(x[2], x[3]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (e < left) __stcs(p + e, x[e]);
                        }
                    }
                };
                if (P.values) {
#pragma unroll
                    for (int d = 0; d < NDEP; ++d) put(P.values, sIdx * NDEP + d, acc[d][0]);
                }
                if (P.jacobian) {
#pragma unroll
                    for (int d = 0; d < NDEP; ++d) {
                        put(P.jacobian, (sIdx * NDEP + d) * 2 + 0, acc[d][1]);
                        put(P.jacobian, (sIdx * NDEP + d) * 2 + 1, acc[d][2]);
                    }
                }
                if constexpr (NDEP == 3 || NDEP == 1) {
                    if (P.normal) {
                        double n[D][4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if constexpr (NDEP == 3) {
                                const double ux = acc[0][1][e], uy = acc[1][1][e], uz = acc[2][1][e];
                                const double vx = acc[0][2][e], vy = acc[1][2][e], vz = acc[2][2][e];
                                n[0][e] = (uy * vz - uz * vy) * P.normalSign;
                                n[1][e] = -(ux * vz - uz * vx) * P.normalSign;
                                n[2][e] = (ux * vy - uy * vx) * P.normalSign;
                            } else {
                                // nInd 2 > nDep 1: T = J^T is 2x1, n = (dv, -du) * sign
                                n[0][e] = acc[0][2][e] * P.normalSign;
                                n[1][e] = -acc[0][1][e] * P.normalSign;
                            }
                            if (P.normalize) {
                                double sq = 0.0;
#pragma unroll
                                for (int i = 0; i < D; ++i)
                                    if (P.normalMask & (1u << i)) sq = fma(n[i][e], n[i][e], sq);
                                // n / |n| as n * (1/sqrt(sq)): zero normal -> 0 * inf = NaN like the reference's 0/0
                                const double inv = 1.0 / sqrt(sq);
#pragma unroll
                                for (int i = 0; i < D; ++i) n[i][e] *= inv;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < D; ++i) put(P.normal, sIdx * D + i, n[i]);
                    }
                }
            }
        }
    }
}

// ---- volumes (nInd == 3): rows of the MMA are (a, b) pairs, columns run along the third axis ------------------
//     out[d,a,b,c] = sum_k ( sum_ij Bu[a][i] Bv[b][j] C[d][..i][..j][sw(c)-ow+k] ) * Bw[c][k]
// A = T[(a,b)][k] for three kinds (value, d/du, d/dv), B = Bw^T and dBw^T; value, the three first partials and
// nothing else (normals of a volume need nDep == 2 or 4 and go through the scattered kernels).  Same store
// pattern as the surface kernel; dependent variables are processed one after the other to keep 16 accumulators.
constexpr int GRID3_TCAP = 32;        // coefficient columns a chunk may touch in the banded path
constexpr int GRID3_TS = GRID3_TCAP + 4;   // row stride of the staged T: == 4 (mod 32) doubles, A-fragment loads take 2 wavefronts
constexpr int GRID3_BAND_ROWS = 16;   // rows of a tile in the banded path
constexpr int GRID3_SJ = 8;           // v-coefficient rows the rows of a tile may touch together in the separable first stage

struct Grid3Params {
    const double *knots[3], *coefs;
    int o[3], nC[3];
    long long depStride;
    const double *axis[3];
    long long n[3];
    double *values, *jacobian;            // (nDep, nU, nV, nW), (nDep, 3, nU, nV, nW)
    long long *firstOutside;
    int vec, chunkCols, colChunks, tileRows;
    long long rowBlocks;
};

template <int NDEP, int MAXO>
__global__ void __launch_bounds__(GRID_WARPS * 32, 2) grid3_dmma_kernel(const Grid3Params P)
{
    constexpr int KS = MAXO / 4;
    extern __shared__ double sm[];
    const int VS = P.chunkCols + 2;
    double *tabW = sm;                                          // [kind][k][col]: 2 * MAXO * VS
    double *tabU = tabW + 2 * MAXO * VS;                        // [kind][i][row]: 2 * MAXO * GRID_TILE_ROWS
    double *tabV = tabU + 2 * MAXO * GRID_TILE_ROWS;            // [kind][j][row]
    int *spanW = reinterpret_cast<int *>(tabV + 2 * MAXO * GRID_TILE_ROWS);  // chunkCols
    int *spanU = spanW + P.chunkCols;                           // GRID_TILE_ROWS
    int *spanVr = spanU + GRID_TILE_ROWS;                       // GRID_TILE_ROWS
    int *krange = spanVr + GRID_TILE_ROWS;                      // [0] = min first coefficient, [1] = max span of the chunk
    // first-stage contraction of the tile, T[kind*NDEP+d][row][k'] (k' = coefficient index - krange[0]), 8-byte aligned
    double *Tsm = reinterpret_cast<double *>(krange + 2);       // 3 * NDEP * GRID3_BAND_ROWS * GRID3_TS
    double *Ssm = Tsm + 3 * NDEP * GRID3_BAND_ROWS * GRID3_TS;  // u-contracted coefficients: 2 * NDEP * GRID3_SJ * GRID3_TS

    const long long tile = blockIdx.x;
    const int cc = (int)(tile % P.colChunks);
    const long long rb = tile / P.colChunks;
    const long long nRows = P.n[0] * P.n[1];
    if (threadIdx.x == 0) { krange[0] = 0x7fffffff; krange[1] = 0; }
    __syncthreads();
    const long long row0 = rb * P.tileRows;
    const long long col0 = (long long)cc * P.chunkCols;
    const long long nW = P.n[2];
    const long long plane = nRows * nW;

    for (int c = threadIdx.x; c < P.chunkCols; c += blockDim.x) {
        const long long b = col0 + c;
        const double *kw = P.knots[2];
        int ix = P.o[2];
        double w = __ldg(kw + P.o[2] - 1);
        if (b < nW) {
            w = __ldg(P.axis[2] + b);
            ix = span_search_inner(kw, P.o[2] + P.nC[2], P.o[2], w);
            if (((w < __ldg(kw + P.o[2] - 1)) | (w > __ldg(kw + P.nC[2]))) && rb == 0)
                if (P.firstOutside) report_outside((int64_t *)P.firstOutside, b);
            atomicMin(krange, ix - P.o[2]);
            atomicMax(krange + 1, ix);
        }
        spanW[c] = ix;
        double b0[MAXO], b1[MAXO];
        axis_basis<MAXO>(kw, P.o[2], ix, w, b0, b1);
#pragma unroll
        for (int j = 0; j < MAXO; ++j) {
            const int sl = j - (MAXO - P.o[2]);
            const int row = sl >= 0 ? sl : P.o[2] + j;
            tabW[row * VS + c] = sl >= 0 ? b0[j] : 0.0;
            tabW[(MAXO + row) * VS + c] = sl >= 0 ? b1[j] : 0.0;
        }
    }
    for (int t = threadIdx.x; t < 2 * P.tileRows; t += blockDim.x) {
        const int which = t / P.tileRows, r = t - which * P.tileRows;      // 0: u of the row, 1: v of the row
        const long long row = row0 + r;
        const long long idx = which == 0 ? row / P.n[1] : row % P.n[1];
        const double *kk = P.knots[which];
        const int o = P.o[which], nC = P.nC[which];
        int ix = o;
        double x = __ldg(kk + o - 1);
        if (row < nRows) {
            x = __ldg(P.axis[which] + idx);
            ix = span_search_inner(kk, o + nC, o, x);
            if (((x < __ldg(kk + o - 1)) | (x > __ldg(kk + nC))) && cc == 0)
                if (P.firstOutside) report_outside((int64_t *)P.firstOutside, (which == 0 ? idx * P.n[1] : idx) * nW);
        }
        (which == 0 ? spanU : spanVr)[r] = ix;
        double b0[MAXO], b1[MAXO];
        axis_basis<MAXO>(kk, o, ix, x, b0, b1);
        double *tab = which == 0 ? tabU : tabV;
#pragma unroll
        for (int j = 0; j < MAXO; ++j) {
            const int sl = j - (MAXO - o);
            const int rowj = sl >= 0 ? sl : o + j;
            tab[rowj * GRID_TILE_ROWS + r] = sl >= 0 ? b0[j] : 0.0;
            tab[(MAXO + rowj) * GRID_TILE_ROWS + r] = sl >= 0 ? b1[j] : 0.0;
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane >> 2, r4 = lane & 3;
    const int bcol = 4 * (q >> 1) + (q & 1);
    const long long nColsHere = min((long long)P.chunkCols, nW - col0);
    const int nStrips = (int)min((long long)(P.tileRows / 8), (nRows - row0 + 7) / 8);
    const long long s0 = (long long)P.nC[1] * P.nC[2], s1 = P.nC[2];

    // ---- banded path: the coefficient columns this chunk touches are few (a fine grid relative to the knots) ----
    // Stage 1, once per CTA: T[kind][d][row][k'] = sum_ij {Bu,dBu,Bu}[i] {Bv,Bv,dBv}[j] C[d][..i][..j][kmin+k'] into
    // shared memory (coalesced along k').  Stage 2: every 16-column step multiplies T by the banded collocation
    // matrix of its columns on the tensor pipe, 4 coefficient columns per K step, skipping K steps whose B fragment
    // is structurally zero.  Without this the first-stage sums were redone per step and per knot span straight from
    // L2 (measured: 11.5 Gpts/s at ~6.6 TB/s of L2 traffic on a 512^3 grid of a 32^3-coefficient volume).
    const int kmin = krange[0], range = krange[1] - krange[0];
    if (range > 0 && range <= GRID3_TCAP && P.tileRows <= GRID3_BAND_ROWS) {
        // Separable first stage.  The rows of a tile are consecutive (a, b) grid nodes: when they share a (tileRows
        // divides n[1], the usual case) the u contraction S[kind][d][jj][k'] = sum_i {Bu,dBu}[i] C[d][su-ou+i][jlo+jj][kmin+k']
        // is common to all of them -- 4 loads per entry once per tile instead of 16 per entry and row -- and every row
        // finishes with its own v basis out of shared memory.
        const long long lastRow = (row0 + P.tileRows < nRows ? row0 + P.tileRows : nRows) - 1;
        int jlo = 0x7fffffff, jhi = 0;
        for (int r = 0; r <= (int)(lastRow - row0); ++r) {
            jlo = min(jlo, spanVr[r] - P.o[1]);
            jhi = max(jhi, spanVr[r]);
        }
        const bool separable = (row0 / P.n[1]) == (lastRow / P.n[1]) && jhi - jlo <= GRID3_SJ;
        if (separable) {
            const int jr = jhi - jlo, su = spanU[0];
            for (int item = threadIdx.x; item < NDEP * jr * GRID3_TS; item += blockDim.x) {
                const int kp = item % GRID3_TS, rest = item / GRID3_TS;
                const int jj = rest % jr, dd = rest / jr;
                double a0 = 0.0, a1 = 0.0;
                if (kp < range) {
                    const double *cp = P.coefs + dd * P.depStride + (long long)(su - P.o[0]) * s0 + (long long)(jlo + jj) * s1 + (kmin + kp);
#pragma unroll
                    for (int i = 0; i < MAXO; ++i)
                        if (i < P.o[0]) {
                            const double x = __ldg(cp + i * s0);
                            a0 = fma(x, tabU[i * GRID_TILE_ROWS], a0);
                            a1 = fma(x, tabU[(MAXO + i) * GRID_TILE_ROWS], a1);
                        }
                }
                Ssm[((0 * NDEP + dd) * GRID3_SJ + jj) * GRID3_TS + kp] = a0;
                Ssm[((1 * NDEP + dd) * GRID3_SJ + jj) * GRID3_TS + kp] = a1;
            }
            __syncthreads();
            for (int item = threadIdx.x; item < P.tileRows * GRID3_TS; item += blockDim.x) {
                const int r = item / GRID3_TS, kp = item - r * GRID3_TS;
                const bool liveItem = kp < range && row0 + r < nRows;
                const int j0 = liveItem ? spanVr[r] - P.o[1] - jlo : 0;
#pragma unroll
                for (int dd = 0; dd < NDEP; ++dd) {
                    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                    if (liveItem) {
#pragma unroll
                        for (int j = 0; j < MAXO; ++j)
                            if (j < P.o[1]) {
                                const double x0 = Ssm[((0 * NDEP + dd) * GRID3_SJ + j0 + j) * GRID3_TS + kp];
                                const double x1 = Ssm[((1 * NDEP + dd) * GRID3_SJ + j0 + j) * GRID3_TS + kp];
                                const double bv = tabV[j * GRID_TILE_ROWS + r];
                                t0 = fma(x0, bv, t0);
                                t1 = fma(x1, bv, t1);
                                t2 = fma(x0, tabV[(MAXO + j) * GRID_TILE_ROWS + r], t2);
                            }
                    }
                    Tsm[((0 * NDEP + dd) * GRID3_BAND_ROWS + r) * GRID3_TS + kp] = t0;
                    Tsm[((1 * NDEP + dd) * GRID3_BAND_ROWS + r) * GRID3_TS + kp] = t1;
                    Tsm[((2 * NDEP + dd) * GRID3_BAND_ROWS + r) * GRID3_TS + kp] = t2;
                }
            }
        } else
        for (int item = threadIdx.x; item < P.tileRows * GRID3_TS; item += blockDim.x) {
            const int r = item / GRID3_TS, kp = item - r * GRID3_TS;
            const bool liveItem = kp < range && row0 + r < nRows;
            const int su = spanU[r], sv = spanVr[r];
#pragma unroll
            for (int dd = 0; dd < NDEP; ++dd) {
                double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                if (liveItem) {
                    const double *cp = P.coefs + dd * P.depStride + (long long)(su - P.o[0]) * s0 + (long long)(sv - P.o[1]) * s1 +
                                       (kmin + kp);
#pragma unroll
                    for (int i = 0; i < MAXO; ++i)
                        if (i < P.o[0]) {
                            double r0 = 0.0, r1 = 0.0;
#pragma unroll
                            for (int j = 0; j < MAXO; ++j)
                                if (j < P.o[1]) {
                                    const double x = __ldg(cp + i * s0 + j * s1);
                                    r0 = fma(x, tabV[j * GRID_TILE_ROWS + r], r0);
                                    r1 = fma(x, tabV[(MAXO + j) * GRID_TILE_ROWS + r], r1);
                                }
                            const double bu = tabU[i * GRID_TILE_ROWS + r];
                            t0 = fma(r0, bu, t0);
                            t1 = fma(r0, tabU[(MAXO + i) * GRID_TILE_ROWS + r], t1);
                            t2 = fma(r1, bu, t2);
                        }
                }
                Tsm[((0 * NDEP + dd) * GRID3_BAND_ROWS + r) * GRID3_TS + kp] = t0;
                Tsm[((1 * NDEP + dd) * GRID3_BAND_ROWS + r) * GRID3_TS + kp] = t1;
                Tsm[((2 * NDEP + dd) * GRID3_BAND_ROWS + r) * GRID3_TS + kp] = t2;
            }
        }
        __syncthreads();
        const int ksteps = (range + 3) >> 2;
        for (int strip = 0; strip < nStrips; ++strip) {
            const int myRow = strip * 8 + q;
            const long long row = row0 + myRow;
            for (int c0 = warp * 16; c0 < nColsHere; c0 += GRID_STEP) {
                const int colA = c0 + bcol, colB = colA + 2;
                const int firstA = spanW[colA] - P.o[2] - kmin, firstB = spanW[colB] - P.o[2] - kmin;   // k' of slot 0
                const long long b = col0 + c0 + 4 * r4;
                const bool live = row < nRows && b < nW;
                const int left = (int)min((long long)4, nW - b);
                const long long at = row * nW + b;
                auto put = [&](double *base, const double (&x)[4]) {
                    double *p = base + at;
                    if (P.vec == 4 && left == 4) {
                        st_cs_v4(p, x[0], x[1], x[2], x[3]);
                    } else if (P.vec >= 2 && left == 4) {
                        __stcs(reinterpret_cast<double2 *>(p), make_double2(x[0], x[1]));
                        __stcs(reinterpret_cast<double2 *>(p + 2), make_double2(x[2], x[3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (e < left) __stcs(p + e, x[e]);
                    }
                };
#pragma unroll
                for (int d = 0; d < NDEP; ++d) {
                    double acc[4][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[k][e] = 0.0;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const int kp = 4 * ks + r4;
                        const int slA = kp - firstA, slB = kp - firstB;          // slot of this K row in the column's window
                        const bool inA = slA >= 0 && slA < P.o[2], inB = slB >= 0 && slB < P.o[2];
                        if (!__any_sync(0xffffffffu, inA || inB)) continue;
                        const double bwA = inA ? tabW[slA * VS + colA] : 0.0;
                        const double dwA = inA ? tabW[(MAXO + slA) * VS + colA] : 0.0;
                        const double bwB = inB ? tabW[slB * VS + colB] : 0.0;
                        const double dwB = inB ? tabW[(MAXO + slB) * VS + colB] : 0.0;
                        const double a0 = Tsm[((0 * NDEP + d) * GRID3_BAND_ROWS + myRow) * GRID3_TS + kp];
                        if (P.values) {
                            dmma884(acc[0][0], acc[0][1], a0, bwA);
                            dmma884(acc[0][2], acc[0][3], a0, bwB);
                        }
                        if (P.jacobian) {
                            const double a1 = Tsm[((1 * NDEP + d) * GRID3_BAND_ROWS + myRow) * GRID3_TS + kp];
                            const double a2 = Tsm[((2 * NDEP + d) * GRID3_BAND_ROWS + myRow) * GRID3_TS + kp];
                            dmma884(acc[1][0], acc[1][1], a1, bwA);
                            dmma884(acc[1][2], acc[1][3], a1, bwB);
                            dmma884(acc[2][0], acc[2][1], a2, bwA);
                            dmma884(acc[2][2], acc[2][3], a2, bwB);
                            dmma884(acc[3][0], acc[3][1], a0, dwA);
                            dmma884(acc[3][2], acc[3][3], a0, dwB);
                        }
                    }
                    if (live) {
                        if (P.values) put(P.values + d * plane, acc[0]);
                        if (P.jacobian) {
                            put(P.jacobian + (d * 3 + 0) * plane, acc[1]);
                            put(P.jacobian + (d * 3 + 1) * plane, acc[2]);
                            put(P.jacobian + (d * 3 + 2) * plane, acc[3]);
                        }
                    }
                }
            }
        }
        return;
    }

    for (int strip = 0; strip < nStrips; ++strip) {
        const int myRow = strip * 8 + q;
        const long long row = row0 + myRow;
        const int su = spanU[myRow], sv = spanVr[myRow];
        double tv[NDEP][KS], tdu[NDEP][KS], tdv[NDEP][KS];
        int cached = -1;
        for (int c0 = warp * 16; c0 < nColsHere; c0 += GRID_STEP) {
            const int colA = c0 + bcol, colB = colA + 2;
            const int swA = spanW[colA], swB = spanW[colB];
            const long long b = col0 + c0 + 4 * r4;
            const bool live = row < nRows && b < nW;
            const int left = (int)min((long long)4, nW - b);
            const long long at = row * nW + b;
            auto put = [&](double *base, const double (&x)[4]) {
                double *p = base + at;
                if (P.vec == 4 && left == 4) {
                    st_cs_v4(p, x[0], x[1], x[2], x[3]);
                } else if (P.vec >= 2 && left == 4) {
                    __stcs(reinterpret_cast<double2 *>(p), make_double2(x[0], x[1]));
                    __stcs(reinterpret_cast<double2 *>(p + 2), make_double2(x[2], x[3]));
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (e < left) __stcs(p + e, x[e]);
                }
            };
#pragma unroll
            for (int d = 0; d < NDEP; ++d) {
                double acc[4][4];                      // value, du, dv, dw
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[k][e] = 0.0;
                unsigned todoA = 0xffffffffu, todoB = 0xffffffffu;
                while (todoA | todoB) {
                    const int cur = todoA ? __shfl_sync(0xffffffffu, swA, __ffs(todoA) - 1)
                                          : __shfl_sync(0xffffffffu, swB, __ffs(todoB) - 1);
                    const bool mineA = swA == cur, mineB = swB == cur;
                    if (cur != cached) {
                        // T[(a,b)][k] = sum_ij Bu[a][i] Bv[b][j] C[dd][su-ou+i][sv-ov+j][cur-ow+k], all dd at once
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const int k = r4 + 4 * ks;
#pragma unroll
                            for (int dd = 0; dd < NDEP; ++dd) {
                                double t0 = 0.0, t1 = 0.0, t2 = 0.0;
                                if (k < P.o[2]) {
                                    const double *cp = P.coefs + dd * P.depStride + (long long)(su - P.o[0]) * s0 +
                                                       (long long)(sv - P.o[1]) * s1 + (cur - P.o[2] + k);
#pragma unroll
                                    for (int i = 0; i < MAXO; ++i)
                                        if (i < P.o[0]) {
                                            double r0 = 0.0, r1 = 0.0;   // sum_j Bv, dBv
#pragma unroll
                                            for (int j = 0; j < MAXO; ++j)
                                                if (j < P.o[1]) {
                                                    const double x = __ldg(cp + i * s0 + j * s1);
                                                    r0 = fma(x, tabV[j * GRID_TILE_ROWS + myRow], r0);
                                                    r1 = fma(x, tabV[(MAXO + j) * GRID_TILE_ROWS + myRow], r1);
                                                }
                                            const double bu = tabU[i * GRID_TILE_ROWS + myRow];
                                            t0 = fma(r0, bu, t0);
                                            t1 = fma(r0, tabU[(MAXO + i) * GRID_TILE_ROWS + myRow], t1);
                                            t2 = fma(r1, bu, t2);
                                        }
                                }
                                tv[dd][ks] = t0;
                                tdu[dd][ks] = t1;
                                tdv[dd][ks] = t2;
                            }
                        }
                        cached = cur;
                    }
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const int k = r4 + 4 * ks;
                        const double bwA = mineA ? tabW[k * VS + colA] : 0.0;
                        const double dwA = mineA ? tabW[(MAXO + k) * VS + colA] : 0.0;
                        const double bwB = mineB ? tabW[k * VS + colB] : 0.0;
                        const double dwB = mineB ? tabW[(MAXO + k) * VS + colB] : 0.0;
                        if (P.values) {
                            dmma884(acc[0][0], acc[0][1], tv[d][ks], bwA);
                            dmma884(acc[0][2], acc[0][3], tv[d][ks], bwB);
                        }
                        if (P.jacobian) {
                            dmma884(acc[1][0], acc[1][1], tdu[d][ks], bwA);
                            dmma884(acc[1][2], acc[1][3], tdu[d][ks], bwB);
                            dmma884(acc[2][0], acc[2][1], tdv[d][ks], bwA);
                            dmma884(acc[2][2], acc[2][3], tdv[d][ks], bwB);
                            dmma884(acc[3][0], acc[3][1], tv[d][ks], dwA);
                            dmma884(acc[3][2], acc[3][3], tv[d][ks], dwB);
                        }
                    }
                    todoA &= ~__ballot_sync(0xffffffffu, mineA);
                    todoB &= ~__ballot_sync(0xffffffffu, mineB);
                }
                if (live) {
                    if (P.values) put(P.values + d * plane, acc[0]);
                    if (P.jacobian) {
                        put(P.jacobian + (d * 3 + 0) * plane, acc[1]);
                        put(P.jacobian + (d * 3 + 1) * plane, acc[2]);
                        put(P.jacobian + (d * 3 + 2) * plane, acc[3]);
                    }
                }
            }
        }
    }
}

static bool aligned_to(const void *p, unsigned n) { return (reinterpret_cast<uintptr_t>(p) & (n - 1)) == 0; }

template <int NDEP, int MAXO, bool F32 = false>
static int launch_grid2(const Grid2Params &P, cudaStream_t stream)
{
    const size_t smem = sizeof(double) * (2 * MAXO * (P.chunkCols + 2) + 2 * MAXO * GRID_TILE_ROWS) +
                        sizeof(int) * (P.chunkCols + GRID_TILE_ROWS);
    if (int rc = allow_dynamic_smem(grid2_dmma_kernel<NDEP, MAXO, F32>, smem)) return rc;
    const long long groups = (P.nSplines + P.group - 1) / P.group;
    const long long tiles = groups * P.colChunks * P.rowBlocks;
    if (tiles > 0x7fffffffLL) { set_error("grid too large for one launch"); return BSPY_E_UNSUPPORTED; }
    grid2_dmma_kernel<NDEP, MAXO, F32><<<(unsigned)tiles, GRID_WARPS * 32, smem, stream>>>(P);
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
    P.nSplines = nSplines;
    P.ou = sp->order[0]; P.ov = sp->order[1]; P.nCu = sp->nCoef[0]; P.nCv = sp->nCoef[1];
    P.depStride = (long long)P.nCu * P.nCv;
    P.axisU = axes[0]; P.axisV = axes[1]; P.nU = nAxis[0]; P.nV = nAxis[1];
    P.values = values; P.jacobian = jacobian; P.normal = normal;
    P.firstOutside = (long long *)firstOutside;
    P.normalSign = sp->normalSign < 0 ? -1 : 1;
    P.normalize = (flags & BSPY_NORMALIZE) ? 1u : 0u;
    P.normalMask = normalMask ? normalMask : 0xffffffffu;
    const bool f32 = (flags & BSPY_OUT_F32) != 0;
    P.vec = 1;
    if (f32) {
        if (P.nV % 4 == 0 && aligned_to(values, 16) && aligned_to(jacobian, 16) && aligned_to(normal, 16)) P.vec = 4;
    } else {
        if (P.nV % 2 == 0 && aligned_to(values, 16) && aligned_to(jacobian, 16) && aligned_to(normal, 16)) P.vec = 2;
        if (P.nV % 4 == 0 && aligned_to(values, 32) && aligned_to(jacobian, 32) && aligned_to(normal, 32)) P.vec = 4;
    }
    {
        // Tile shape.  Write-heavy requests (>= 6 doubles per point, e.g. value + jacobian + normal = 12) are bound
        // by the HBM store stream, which is fastest when CTAs are short-lived and close together: 16 rows x 256
        // columns, one spline per CTA (measured 6.76 TB/s against 5.0 TB/s for 64 x 256 x 4 splines).  Light
        // requests are bound by the per-CTA table set-up instead and take the large tile.
        const int D = sp->nDep > 2 ? sp->nDep : 2;
        const int doubles = (values ? sp->nDep : 0) + (jacobian ? 2 * sp->nDep : 0) + (normal ? D : 0);
        const bool heavy = doubles >= 6;
        // float32 outputs halve the store stream: the kernel is then bound by the per-CTA tables and the arithmetic, and
        // larger tiles win (measured, 32 patches x 2048^2, 12 floats per point: 16 x 256 x 1 -> 81.6 Gpts/s,
        // 32 x 512 x 4 -> 103.8 Gpts/s = 5.0 TB/s written)
        const long long cap = option(OPT_GRID_CHUNK, f32 ? 512 : 256);
        const long long chunks = (P.nV + cap - 1) / cap;
        long long per = (P.nV + chunks - 1) / chunks;
        per = (per + GRID_STEP - 1) / GRID_STEP * GRID_STEP;
        P.chunkCols = (int)per;
        P.colChunks = (int)((P.nV + per - 1) / per);
        P.tileRows = (int)option(OPT_GRID_ROWS, f32 ? 32 : (heavy ? 16 : GRID_TILE_ROWS));
        if (P.tileRows < 8 || P.tileRows > GRID_TILE_ROWS || P.tileRows % 8) P.tileRows = GRID_TILE_ROWS;
        P.rowBlocks = (P.nU + P.tileRows - 1) / P.tileRows;
        const int g = (int)option(OPT_GRID_GROUP, (heavy && !f32) ? 1 : GRID_GROUP);
        P.group = (knotStride0 == 0 && knotStride1 == 0) ? (int)(nSplines < g ? nSplines : g) : 1;
    }
    if (P.nU == 0 || P.nV == 0 || nSplines == 0) return 0;
    const bool small = P.ou <= 4 && P.ov <= 4;
    if (f32) {
        switch (sp->nDep) {
            case 1: return small ? launch_grid2<1, 4, true>(P, stream) : launch_grid2<1, 8, true>(P, stream);
            case 2: return small ? launch_grid2<2, 4, true>(P, stream) : launch_grid2<2, 8, true>(P, stream);
            case 3: return small ? launch_grid2<3, 4, true>(P, stream) : launch_grid2<3, 8, true>(P, stream);
            default: return small ? launch_grid2<4, 4, true>(P, stream) : launch_grid2<4, 8, true>(P, stream);
        }
    }
    switch (sp->nDep) {
        case 1: return small ? launch_grid2<1, 4>(P, stream) : launch_grid2<1, 8>(P, stream);
        case 2: return small ? launch_grid2<2, 4>(P, stream) : launch_grid2<2, 8>(P, stream);
        case 3: return small ? launch_grid2<3, 4>(P, stream) : launch_grid2<3, 8>(P, stream);
        default: return small ? launch_grid2<4, 4>(P, stream) : launch_grid2<4, 8>(P, stream);
    }
}

template <int NDEP, int MAXO>
static int launch_grid3(const Grid3Params &P, cudaStream_t stream)
{
    const size_t smem = sizeof(double) * (2 * MAXO * (P.chunkCols + 2) + 4 * MAXO * GRID_TILE_ROWS) +
                        sizeof(int) * (P.chunkCols + 2 * GRID_TILE_ROWS + 2) +
                        sizeof(double) * (3 * NDEP * GRID3_BAND_ROWS * GRID3_TS + 2 * NDEP * GRID3_SJ * GRID3_TS);
    if (int rc = allow_dynamic_smem(grid3_dmma_kernel<NDEP, MAXO>, smem)) return rc;
    const long long tiles = P.colChunks * P.rowBlocks;
    if (tiles > 0x7fffffffLL) { set_error("grid too large for one launch"); return BSPY_E_UNSUPPORTED; }
    grid3_dmma_kernel<NDEP, MAXO><<<(unsigned)tiles, GRID_WARPS * 32, smem, stream>>>(P);
    count_launch();
    return check_launch("bspy_cuda_eval_grid");
}

static bool grid3_supported(const bspy_spline *sp, const double *normal)
{
    if (sp->nInd != 3 || sp->nDep < 1 || sp->nDep > 4 || normal) return false;
    for (int i = 0; i < 3; ++i)
        if (sp->order[i] > GRID_MAX_ORDER) return false;
    return true;
}

static int grid3_run(const bspy_spline *sp, const double *const *axes, const int64_t *nAxis, double *values, double *jacobian,
                     int64_t *firstOutside, cudaStream_t stream)
{
    Grid3Params P{};
    for (int i = 0; i < 3; ++i) {
        P.knots[i] = sp->knots[i]; P.o[i] = sp->order[i]; P.nC[i] = sp->nCoef[i];
        P.axis[i] = axes[i]; P.n[i] = nAxis[i];
    }
    P.coefs = sp->coefs;
    P.depStride = (long long)sp->nCoef[0] * sp->nCoef[1] * sp->nCoef[2];
    P.values = values; P.jacobian = jacobian;
    P.firstOutside = (long long *)firstOutside;
    if (P.n[0] == 0 || P.n[1] == 0 || P.n[2] == 0) return 0;
    P.vec = 1;
    if (P.n[2] % 2 == 0 && aligned_to(values, 16) && aligned_to(jacobian, 16)) P.vec = 2;
    if (P.n[2] % 4 == 0 && aligned_to(values, 32) && aligned_to(jacobian, 32)) P.vec = 4;
    // measured on the 512^3 grid of the config-4 volume (value + jacobian): 128 / 256 / 512 columns -> 46.1 / 53.2 / 62.2 Gpts/s
    const long long cap = option(OPT_GRID3_CHUNK, 512);
    const long long chunks = (P.n[2] + cap - 1) / cap;
    long long per = (P.n[2] + chunks - 1) / chunks;
    per = (per + GRID_STEP - 1) / GRID_STEP * GRID_STEP;
    P.chunkCols = (int)per;
    P.colChunks = (int)((P.n[2] + per - 1) / per);
    P.tileRows = (int)option(OPT_GRID3_ROWS, GRID3_BAND_ROWS);
    if (P.tileRows < 8 || P.tileRows > GRID3_BAND_ROWS || P.tileRows % 8) P.tileRows = GRID3_BAND_ROWS;
    P.rowBlocks = (P.n[0] * P.n[1] + P.tileRows - 1) / P.tileRows;
    const bool small = P.o[0] <= 4 && P.o[1] <= 4 && P.o[2] <= 4;
    switch (sp->nDep) {
        case 1: return small ? launch_grid3<1, 4>(P, stream) : launch_grid3<1, 8>(P, stream);
        case 2: return small ? launch_grid3<2, 4>(P, stream) : launch_grid3<2, 8>(P, stream);
        case 3: return small ? launch_grid3<3, 4>(P, stream) : launch_grid3<3, 8>(P, stream);
        default: return small ? launch_grid3<4, 4>(P, stream) : launch_grid3<4, 8>(P, stream);
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
    if (flags & BSPY_OUT_F32) {
        set_error("bspy_cuda_eval_grid: float32 outputs are available for surfaces (nInd == 2, nDep <= 4, orders <= 8) only");
        return BSPY_E_UNSUPPORTED;
    }
    if (grid3_supported(spline, normal) && (values || jacobian)) {
        SplineDev chk;
        int rc = make_spline_dev(spline, chk, "bspy_cuda_eval_grid");
        if (rc) return rc;
        return grid3_run(spline, axes, nAxis, values, jacobian, firstOutside, (cudaStream_t)stream);
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
