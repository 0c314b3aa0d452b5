/* bspy_cuda.h -- C ABI of libbspy_cuda.so: batched float64 B-spline evaluation on sm_100a.
 *
 * This is the drop-in boundary for ONE path of ericbrec/BSpy: the evaluation functions of
 * bspy/_spline_evaluation.py.  The reference has no FFI for that path (it is plain Python
 * calling numpy); every entry point below names the reference function it replaces so a
 * maintainer can bind it from a new `bspy/_cuda` module with ctypes.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - Every function returns int: 0 = ok, negative = BSPY_E_* (bad argument / unsupported),
 *    positive = cudaError_t of a failed CUDA call.  bspy_cuda_last_error_string() gives text
 *    for the last non-zero return on the calling thread.
 *  - All data pointers are DEVICE pointers owned by the caller unless the name ends in
 *    `_host`.  The library never allocates or frees caller-visible memory, never
 *    synchronises, and launches everything on the stream passed in (`void*` = cudaStream_t;
 *    NULL = legacy default stream).
 *  - `bspy_spline` / `axes` / `nAxis` / `wrt` / `knotStride` arguments are HOST structures (they
 *    hold device pointers); they are read during the call and need not outlive it.
 *  - float64 throughout; spans are int32; coefficients are `(nDep, nCoef[0], ..,
 *    nCoef[nInd-1])` C-contiguous, exactly `Spline.coefs` (bspy/spline.py:70-75) made
 *    contiguous; knots[i] has order[i] + nCoef[i] entries (bspy/spline.py:57-59).
 *  - Batched outputs are struct-of-arrays so that consecutive points are consecutive in
 *    memory: values (nDep, N), deriv (nDep, N), jacobian (nDep, nInd, N),
 *    normal (max(nInd,nDep), N), spans (nInd, N).  Any output pointer may be NULL (= not
 *    requested).
 */
#ifndef BSPY_CUDA_H
#define BSPY_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSPY_MAX_IND 8     /* independent variables supported by the kernels            */
#define BSPY_MAX_ORDER 32  /* polynomial order per variable supported by the kernels    */
#define BSPY_ABI_VERSION 2

/* error codes (negative returns) */
#define BSPY_E_ARG (-1)          /* NULL / negative / inconsistent argument              */
#define BSPY_E_UNSUPPORTED (-2)  /* nInd > BSPY_MAX_IND, order > BSPY_MAX_ORDER, ...     */
#define BSPY_E_NORMAL_DIMS (-3)  /* normal requested with |nInd - nDep| != 1 (reference:
                                    ValueError, bspy/_spline_evaluation.py:219)          */

/* The attributes of `Spline` the evaluation path reads (bspy/spline.py:46-76). */
typedef struct bspy_spline {
    int32_t nInd;
    int32_t nDep;
    int32_t order[BSPY_MAX_IND];
    int32_t nCoef[BSPY_MAX_IND];
    const double *knots[BSPY_MAX_IND]; /* device; knots[i] has order[i] + nCoef[i] entries */
    const double *coefs;               /* device; (nDep, nCoef[0..nInd-1]) C-contiguous   */
    int32_t normalSign;                /* -1 iff metadata["negateNormal"], else +1         */
    int32_t reserved;
    /* curves (nInd == 1) only, optional: device image of the span tables of this curve, built ONCE per spline
     * by bspy_cuda_curve_table_build into caller-owned memory (bspy_cuda_curve_table_bytes bytes); big batches
     * then fetch it by TMA instead of rebuilding it in every thread block.  NULL / 0: tables are rebuilt per call. */
    const void *curveTable;
    int64_t curveTableBytes;
} bspy_spline;

/* what bspy_cuda_eval_* computes */
#define BSPY_NORMALIZE 1u /* divide the normal by its 2-norm over the selected components */
#define BSPY_OUT_F32 2u   /* grid entry points, surfaces only: values / jacobian / normal point to float arrays */
#define BSPY_WANT_JACOBIAN 4u /* bspy_cuda_eval_points_aos: records carry the jacobian after the values       */
#define BSPY_WANT_NORMAL 8u   /* bspy_cuda_eval_points_aos: ... and the normal after the jacobian             */

/* ---- library ------------------------------------------------------------------------- */
int bspy_cuda_abi_version(void);
const char *bspy_cuda_last_error_string(void);
/* number of kernels launched by this library in this process so far (bench.py reports it) */
int64_t bspy_cuda_launch_count(void);
/* Experiment switches (kernel variants, tile shapes, chunk sizes; names listed in DESIGN.md section 7).  Defaults are
 * the measured best; the environment variable BSPY_<NAME> is read ONCE when the library is first used and this call
 * overrides it (isSet == 0 returns the switch to "unset" = built-in default).  Process-wide; not an evaluation entry. */
int bspy_cuda_set_option(const char *name, int64_t value, int32_t isSet);

/* Host-pipeline helper: rows x widthBytes 2-D copy between device and PINNED host memory (either direction) on the
 * stream, one call per output array and chunk instead of one per row (cudaMemcpy2DAsync, cudaMemcpyDefault).        */
int bspy_cuda_copy_2d(void *dst, int64_t dstPitchBytes, const void *src, int64_t srcPitchBytes,
                      int64_t widthBytes, int64_t rows, void *stream);

/* ---- knot spans: replaces the search in bspline_values
 *      (bspy/_spline_evaluation.py:7-8: np.searchsorted(knots, u, 'right') clamped to
 *      [order, len(knots)-order]).  Bit-exact.  NaN selects the last span like numpy.    */
int bspy_cuda_spans(const double *knots, int32_t nKnots, int32_t order, const double *u, int64_t N,
                    int32_t *spans, void *stream);

/* ---- basis values: replaces bspline_values (bspy/_spline_evaluation.py:4-27; facade
 *      Spline.bspline_values, bspy/spline.py:207-252) for N parameters at once.
 *      spansIn == NULL searches (the reference's `knot is None`); spansOut may be NULL.
 *      basis is (N, order) row-major: row p is the array the reference returns for u[p].
 *      Arithmetic is IEEE with separate multiply/add and true divisions, in the reference's
 *      order: the result is bit-identical to the reference.  derivOrder >= order gives zeros. */
int bspy_cuda_basis(const double *knots, int32_t nKnots, int32_t order, const double *u,
                    const int32_t *spansIn, int64_t N, int32_t derivOrder, int32_t taylorCoefs,
                    int32_t *spansOut, double *basis, void *stream);

/* ---- scattered points: replaces evaluate / derivative / jacobian / normal
 *      (bspy/_spline_evaluation.py:140-164, 109-133, 205-213, 215-246) for N points.
 *      Point p has parameter i at uvw[p*pointStride + i*varStride] (elements): (N,nInd)
 *      row-major is (nInd,1); (nInd,N) is (1,N).
 *      values  : spline value                               (nDep, N)      or NULL
 *      wrt/deriv: mixed partial of order wrt[i] in variable i (nDep, N)    or both NULL
 *                 (wrt is a HOST array of nInd ints)
 *      jacobian: first partials, [d][i][p]                  (nDep, nInd, N) or NULL
 *      normal  : sign * (-1)^i * det(J without row i)       (max(nInd,nDep), N) or NULL;
 *                with BSPY_NORMALIZE divided by the norm over the components in normalMask
 *                (bit i = component i; 0 = all), as `indices` does in the reference
 *      spans   : rightmost-knot index per variable          (nInd, N) int32 or NULL
 *      firstOutside: device int64, must hold -1 (or any negative) on entry; receives the
 *                smallest point index with a parameter outside the closed domain
 *                [knots[o-1], knots[nCoef]] (reference raises ValueError there,
 *                bspy/_spline_evaluation.py:149-152; NaN is inside).  May be NULL.        */
int bspy_cuda_eval_points(const bspy_spline *spline, const double *uvw, int64_t pointStride,
                          int64_t varStride, int64_t N, const int32_t *wrt, uint32_t flags,
                          uint32_t normalMask, double *values, double *deriv, double *jacobian,
                          double *normal, int32_t *spans, int64_t *firstOutside, void *stream);

/* ---- span tables of one curve (see bspy_spline.curveTable): bytes needed (0: shape without tables -- nInd != 1,
 *      order > 6, nDep > 4, or tables larger than shared memory) and the build (one small kernel on `stream`; the
 *      image must be rebuilt when knots or coefficients change, like the device copies themselves).              */
int64_t bspy_cuda_curve_table_bytes(const bspy_spline *spline);
int bspy_cuda_curve_table_build(const bspy_spline *spline, void *table, int64_t tableBytes, void *stream);

/* ---- the same with cell binning for big scattered batches: the points of each 512 Ki chunk are
 *      counting-sorted by knot-span cell (in `workspace`, caller-owned device memory) so that the
 *      lanes of a warp share their coefficient window, evaluated in that order and written to their
 *      original positions.  Results are bit-identical to bspy_cuda_eval_points.
 *      bspy_cuda_binned_workspace_bytes returns the workspace size for (spline, N), or 0 when
 *      binning does not apply (curves, small windows, splines that fit in L1, N < 65536, shapes
 *      without a compiled kernel); with a NULL / too small workspace the call falls back to
 *      bspy_cuda_eval_points.                                                               */
int64_t bspy_cuda_binned_workspace_bytes(const bspy_spline *spline, int64_t N);
int bspy_cuda_eval_points_binned(const bspy_spline *spline, const double *uvw, int64_t pointStride,
                                 int64_t varStride, int64_t N, const int32_t *wrt, uint32_t flags,
                                 uint32_t normalMask, double *values, double *deriv, double *jacobian,
                                 double *normal, int32_t *spans, int64_t *firstOutside,
                                 void *workspace, int64_t workspaceBytes, void *stream);

/* ---- array-of-structs results: one record per point, records[p * recordStride ..] =
 *      [ values (nDep) | jacobian (nDep x nInd, [d][i]; with BSPY_WANT_JACOBIAN or BSPY_WANT_NORMAL) |
 *        normal (max(nInd,nDep); with BSPY_WANT_NORMAL) ], which is the shape the reference returns per point
 *      (evaluate -> (nDep,), jacobian -> (nDep, nInd), normal -> (D,): bspy/_spline_evaluation.py:140-164, 205-246).
 *      recordStride (doubles) must be even and >= the record length, records 16-byte aligned; a stride that is a
 *      multiple of 4 makes every record whole 32-byte sectors: the cell-sorted path for big batches then writes each
 *      result straight to its point's original position and needs no un-permute pass over the outputs (padding doubles
 *      of a record may be written as zeros).  workspace: bspy_cuda_aos_workspace_bytes(spline, N) bytes of device
 *      memory (0 = the direct kernels are used; NULL / too small: the direct kernels are used as well).             */
int64_t bspy_cuda_aos_workspace_bytes(const bspy_spline *spline, int64_t N);
int bspy_cuda_eval_points_aos(const bspy_spline *spline, const double *uvw, int64_t pointStride,
                              int64_t varStride, int64_t N, uint32_t flags, uint32_t normalMask,
                              double *records, int64_t recordStride, int32_t *spans, int64_t *firstOutside,
                              void *workspace, int64_t workspaceBytes, void *stream);

/* ---- regular grid: the tensor product of one parameter axis per variable; axis i has
 *      nAxis[i] device doubles.  Outputs as above with N = prod(nAxis) laid out C-order with
 *      the LAST variable fastest: values (nDep, nAxis[0], .., nAxis[nInd-1]) etc.  This is
 *      `spline(*np.meshgrid(*axes, indexing="ij"))` in the reference's ufunc style
 *      (bspy/spline.py:940-947), plus jacobian / normal which the reference can only do one
 *      point at a time.  Surfaces (nInd == 2) run the B_u * C * B_v^T contraction on the
 *      FP64 tensor pipe.                                                                  */
int bspy_cuda_eval_grid(const bspy_spline *spline, const double *const *axes, const int64_t *nAxis,
                        uint32_t flags, uint32_t normalMask, double *values, double *jacobian,
                        double *normal, int64_t *firstOutside, void *stream);

/* ---- the same for a batch of nSplines surfaces of identical shape (nInd == 2, nDep <= 4,
 *      orders <= 8) on ONE shared pair of axes, in one launch: element s uses
 *      first->knots[i] + s*knotStride[i] (0 = shared knots) and first->coefs + s*coefStride.
 *      Outputs gain a leading batch dimension: values (S, nDep, nU, nV), jacobian
 *      (S, nDep, 2, nU, nV), normal (S, D, nU, nV).  This is the bulk container for a list of
 *      patches such as examples/teapot.py's 32 bicubic patches.                            */
int bspy_cuda_eval_grid_batch(const bspy_spline *first, int64_t nSplines, const int64_t *knotStride,
                              int64_t coefStride, const double *const *axes, const int64_t *nAxis,
                              uint32_t flags, uint32_t normalMask, double *values, double *jacobian,
                              double *normal, int64_t *firstOutside, void *stream);

/* ---- many independent curves/splines of identical shape (nInd == 1): spline s has knots
 *      knots[s*knotStride ..][order+nCoef], coefficients coefs[s*coefStride ..] as
 *      (nDep, nCoef), and nPts parameters u[s*nPts ..].  values is (nSplines, nDep, nPts),
 *      deriv1 (first derivative, same shape) may be NULL.  firstOutside receives the
 *      smallest flat index s*nPts + p outside its curve's domain.                         */
int bspy_cuda_eval_many(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines,
                        const double *knots, int64_t knotStride, const double *coefs, int64_t coefStride,
                        const double *u, int32_t nPts, double *values, double *deriv1,
                        int64_t *firstOutside, void *stream);

/*      Cached per-curve images for a resident batch of curves (values; deriv1 != NULL: + first derivatives) (nothing in the reference: what
 *      bspline_values, bspy/_spline_evaluation.py:4-27, recomputes per point -- span search, knot gaps, the recurrence --
 *      precomputed once per curve): per curve { bucket scale, flag | knots | bucket table of the span search | one row per span
 *      with the span's polynomial in powers of (u - mid-span) }, rows validated against the Cox-de Boor recurrence when they are
 *      built (a curve that fails keeps the recurrence, decided per curve inside the kernel).  The caller owns the buffer.
 *      bspy_cuda_many_table_bytes: bytes for nSplines curves (0: shape without tables -- order outside 2..6, nDep > 3, image
 *      of a curve above 12 KB); bspy_cuda_many_table_build fills it (16-byte aligned); bspy_cuda_eval_many_tab evaluates
 *      values (and deriv1, may be NULL), both (nSplines, nDep, nPts), from it -- tableBytes must be exactly what _bytes returned for this batch.            */
int64_t bspy_cuda_many_table_bytes(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines);
int bspy_cuda_many_table_build(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines,
                               const double *knots, int64_t knotStride, const double *coefs, int64_t coefStride,
                               void *table, int64_t tableBytes, void *stream);
int bspy_cuda_eval_many_tab(int32_t order, int32_t nCoef, int32_t nDep, int64_t nSplines,
                            const double *knots, int64_t knotStride, const double *coefs, int64_t coefStride,
                            const void *table, int64_t tableBytes, const double *u, int32_t nPts,
                            double *values, double *deriv1, int64_t *firstOutside, void *stream);

/* ---- curvature: replaces curvature, bspy/_spline_evaluation.py:80-107, for N points
 *      from batched derivatives produced by bspy_cuda_eval_points.
 *      nInd == 1: d1 (nDep, N) first, d2 (nDep, N) second derivative; signed for planar curves.
 *      nInd == 2: Gaussian curvature; d1 = jacobian (3, 2, N), d2 = (3 kinds uu|uv|vv, 3, N), normal (3, N) unit.
 *      graph != 0: nDep == 1, the curve / surface is the graph of the scalar function (reference: self.graph());
 *                  d1 (nInd, N), d2 (1 or 3 kinds, N), normal unused.                                  */
int bspy_cuda_curvature(int32_t nInd, int32_t nDep, int32_t graph, int64_t N, const double *d1,
                        const double *d2, const double *normal, double *out, void *stream);

/*      bspy_cuda_curvature_points: the same from the spline itself, one fused pass per point (basis values, first and
 *      second derivatives of every variable, one walk over the coefficient window, curvature formula): curves of nDep
 *      1..3 and surfaces of nDep 3 or 1 (nDep == 1: graph of the function) with orders <= 8; BSPY_E_UNSUPPORTED otherwise
 *      (callers then compose bspy_cuda_eval_points passes + bspy_cuda_curvature).  firstOutside as in eval_points.   */
int bspy_cuda_curvature_points(const bspy_spline *spline, const double *uvw, int64_t pointStride,
                               int64_t varStride, int64_t N, double *out, int64_t *firstOutside, void *stream);

/* ---- Spline.contract and SplineBlock -------------------------------------------------
 *      bspy_cuda_contract_axis: out[a, b] = sum_j coefs[a, first + j, b] * basis[j], a < outer, b < inner, j < order:
 *      one variable of a C-contiguous coefficient array (outer, n, inner) contracted against the `order` non-zero
 *      basis values of one parameter (basis: device pointer, e.g. the output of bspy_cuda_basis); replaces
 *      `coefs[..., ix-order:ix, ...] @ bValues` of bspy/_spline_operations.py:184-223 (Spline.contract).
 *      bspy_cuda_block_accumulate: dst[dstRows[r] * dstLd + p] += src[r * srcLd + p], r < nRows, p < N: the row sums
 *      of SplineBlock._block_evaluation / SplineBlock.jacobian (bspy/spline_block.py:37-44, 231-245); dstRows is a
 *      HOST array.
 *      bspy_cuda_normal_from_jacobian: cofactor normals (D = max(nInd, nDep) components, SoA (D, N)) of N jacobians
 *      stored SoA as (nDep, nInd, N): bspy/_spline_evaluation.py:215-246 applied to a block's jacobian
 *      (bspy/spline_block.py:282).  flags: BSPY_NORMALIZE; normalMask as in bspy_cuda_eval_points.          */
int bspy_cuda_contract_axis(const double *coefs, int64_t outer, int64_t n, int64_t inner, int32_t first,
                            int32_t order, const double *basis, double *out, void *stream);
int bspy_cuda_block_accumulate(double *dst, int64_t dstLd, const double *src, int64_t srcLd, int32_t nRows,
                               const int32_t *dstRows_host, int64_t N, void *stream);
int bspy_cuda_normal_from_jacobian(const double *jacobian, int32_t nDep, int32_t nInd, int64_t N,
                                   int32_t normalSign, uint32_t flags, uint32_t normalMask, double *normal,
                                   void *stream);

/* ---- collocation rows ------------------------------------------------------------------
 *      A[r, ix_r - order : ix_r] = bspline_values(None, knots, order, u[r], derivOrders[r]) (bit-identical to the
 *      reference), every other entry of row r zero: the matrix assembled row by row in Spline.least_squares
 *      (bspy/_spline_fitting.py:736-750; consecutive equal parameters raise the derivative order) and in contour
 *      (bspy/_spline_fitting.py:190-219).  A: dense (N, nKnots - order), row-major with leading dimension ldA;
 *      derivOrders NULL = values; spansOut optional.                                                          */
int bspy_cuda_collocation(const double *knots, int32_t nKnots, int32_t order, const double *u,
                          const int32_t *derivOrders, int64_t N, int32_t *spansOut, double *A, int64_t ldA,
                          void *stream);

/* ---- measurement helpers used by bench.py (not part of the evaluation path) --------------
 *      bspy_cuda_probe_fp64: runs `iters` dependent-free FP64 FMA (kind 0) or DMMA m8n8k4
 *      (kind 1) chains on every SM and returns the flop count; time it with events on `stream`.
 *      bspy_cuda_probe_hbm: kind 0 = copy src->dst (bytes read + bytes written),
 *      kind 1 = write-only fill of dst; returns bytes moved.                               */
int bspy_cuda_probe_fp64(int32_t kind, int32_t iters, double *sink, double *flopsOut_host, void *stream);
int bspy_cuda_probe_hbm(int32_t kind, const double *src, double *dst, int64_t nDoubles,
                        double *bytesOut_host, void *stream);
/*      bspy_cuda_probe_tiles: writes `planes` arrays of nU x nV doubles tile by tile with the grid
 *      kernel's store pattern and no arithmetic (the attainable rate of the pattern itself).      */
int bspy_cuda_probe_tiles(double *dst, int32_t planes, int64_t nU, int64_t nV, int32_t tileRows,
                          int32_t tileCols, double *bytesOut_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BSPY_CUDA_H */
