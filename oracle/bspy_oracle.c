/* CPU oracle for the BSpy evaluation path, plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * A second restatement of the reference algorithm (the first is oracle/bspy_oracle.py),
 * used where the numpy tier is too slow or too memory-hungry: full-size parity samples
 * and the "native port" line of bench.py's CPU baseline.  Nothing in bspy_b200/ links,
 * loads or calls this file.  Parity status: pinned -- tests/test_oracle_c.py checks it
 * against the golden vectors generated from the unmodified reference (spans and basis
 * values bit-for-bit: build with -ffp-contract=off so no FMA is formed).
 *
 * Reference lines followed (bspy/_spline_evaluation.py):
 *   span      :7-8      upper bound of u in knots, clamped to [order, nKnots-order]
 *   basis     :9-26     triangular recurrence, value stages then derivative stages
 *   evaluate  :140-164  window slice, contraction from the last variable to the first
 *   derivative:109-133
 *   jacobian  :205-213  one derivative call per independent variable
 *   normal    :215-246  signed cofactors (LU with partial pivoting, like LAPACK getrf
 *                       behind np.linalg.det), optional division by the 2-norm
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_IND 8
#define ORACLE_MAX_ORDER 64

/* number of knots <= u, clamped; NaN compares false everywhere and numpy sorts it last */
static int span_of(const double *knots, int nKnots, int order, double u)
{
    int lo = 0, hi = nKnots;
    if (u != u)
        lo = nKnots;
    else
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (knots[mid] <= u) lo = mid + 1; else hi = mid;
        }
    if (lo < order) lo = order;
    if (lo > nKnots - order) lo = nKnots - order;
    return lo;
}

static void basis_of(const double *knots, int order, int ix, double u, int deriv, int taylor, double *b)
{
    int deg, i, slot;
    for (i = 0; i < order; ++i) b[i] = 0.0;
    if (deriv >= order) return;
    b[order - 1] = 1.0;
    for (deg = 1; deg < order; ++deg) {
        slot = order - deg;
        if (deg < order - deriv) {
            for (i = ix - deg; i < ix; ++i, ++slot) {
                double a = (u - knots[i]) / (knots[i + deg] - knots[i]);
                double t = (1.0 - a) * b[slot];
                b[slot - 1] = b[slot - 1] + t;
                b[slot] = b[slot] * a;
            }
        } else {
            double scale = (double)deg / (taylor ? (double)(order - deg) : 1.0);
            for (i = ix - deg; i < ix; ++i, ++slot) {
                double a = scale / (knots[i + deg] - knots[i]);
                double t = -a * b[slot];
                b[slot - 1] = b[slot - 1] + t;
                b[slot] = b[slot] * a;
            }
        }
    }
}

int bspy_oracle_spans(const double *knots, int nKnots, int order, const double *u, int64_t N, int32_t *out)
{
    int64_t p;
#pragma omp parallel for schedule(static)
    for (p = 0; p < N; ++p) out[p] = span_of(knots, nKnots, order, u[p]);
    return 0;
}

/* basis[N][order]; ixIn may be NULL (search) */
int bspy_oracle_basis(const double *knots, int nKnots, int order, const double *u, const int32_t *ixIn,
                      int64_t N, int deriv, int taylor, int32_t *ixOut, double *basis)
{
    int64_t p;
    if (order > ORACLE_MAX_ORDER) return -1;
#pragma omp parallel for schedule(static)
    for (p = 0; p < N; ++p) {
        int ix = ixIn ? ixIn[p] : span_of(knots, nKnots, order, u[p]);
        if (ixOut) ixOut[p] = ix;
        basis_of(knots, order, ix, u[p], deriv, taylor, basis + p * order);
    }
    return 0;
}

typedef struct {
    int nInd, nDep;
    const int32_t *order, *nCoef;
    const double *const *knots;
    const double *coefs;
} spline_t;

/* one mixed partial at one point -> out[nDep]; scratch holds nDep * prod(order) doubles */
static void derivative_at(const spline_t *s, const int32_t *wrt, const double *uvw, double *out,
                          double *scratch, int32_t *spansOut)
{
    double rows[ORACLE_MAX_IND][ORACLE_MAX_ORDER];
    int ix[ORACLE_MAX_IND];
    int64_t stride[ORACLE_MAX_IND];
    int64_t win = 1, depStride = 1, n, j;
    int iv, d, k;
    for (iv = s->nInd - 1; iv >= 0; --iv) { stride[iv] = depStride; depStride *= s->nCoef[iv]; }
    for (iv = 0; iv < s->nInd; ++iv) {
        ix[iv] = span_of(s->knots[iv], s->order[iv] + s->nCoef[iv], s->order[iv], uvw[iv]);
        if (spansOut) spansOut[iv] = ix[iv];
        basis_of(s->knots[iv], s->order[iv], ix[iv], uvw[iv], wrt ? wrt[iv] : 0, 0, rows[iv]);
        win *= s->order[iv];
    }
    /* gather the window, C order (dep, i0, i1, ...) */
    for (d = 0; d < s->nDep; ++d)
        for (j = 0; j < win; ++j) {
            int64_t rem = j, off = 0;
            for (iv = s->nInd - 1; iv >= 0; --iv) {
                int o = s->order[iv];
                off += (ix[iv] - o + (rem % o)) * stride[iv];
                rem /= o;
            }
            scratch[d * win + j] = s->coefs[d * depStride + off];
        }
    n = s->nDep * win;
    for (iv = s->nInd - 1; iv >= 0; --iv) {
        int o = s->order[iv];
        n /= o;
        for (j = 0; j < n; ++j) {
            double acc = 0.0;
            for (k = 0; k < o; ++k) acc += scratch[j * o + k] * rows[iv][k];
            scratch[j] = acc;
        }
    }
    for (d = 0; d < s->nDep; ++d) out[d] = scratch[d];
}

/* determinant by LU with partial pivoting (row swaps flip the sign) */
static double det_lu(double *a, int n)
{
    double det = 1.0;
    int c, r, k;
    for (c = 0; c < n; ++c) {
        int piv = c;
        double best = fabs(a[c * n + c]);
        for (r = c + 1; r < n; ++r)
            if (fabs(a[r * n + c]) > best) { best = fabs(a[r * n + c]); piv = r; }
        if (piv != c) {
            for (k = 0; k < n; ++k) { double t = a[c * n + k]; a[c * n + k] = a[piv * n + k]; a[piv * n + k] = t; }
            det = -det;
        }
        if (a[c * n + c] == 0.0) return 0.0;
        for (r = c + 1; r < n; ++r) {
            double l = a[r * n + c] / a[c * n + c];
            for (k = c + 1; k < n; ++k) a[r * n + k] -= l * a[c * n + k];
        }
        det *= a[c * n + c];
    }
    return det;
}

enum { WANT_VALUES = 1, WANT_DERIV = 2, WANT_JACOBIAN = 4, WANT_NORMAL = 8, WANT_NORMALIZE = 16, WANT_SPANS = 32 };

/* uvw[N][nInd]; values[N][nDep]; deriv[N][nDep]; jac[N][nDep][nInd]; normal[N][D]; spans[N][nInd].
 * normalMask selects the components that enter the norm (bit i = component i); all D are written.
 * firstOob receives the index of the first point outside the closed domain, or -1. */
int bspy_oracle_eval(int nInd, int nDep, const int32_t *order, const int32_t *nCoef, const double *const *knots,
                     const double *coefs, const double *uvw, int64_t N, const int32_t *wrt, int flags,
                     int normalSign, uint32_t normalMask, double *values, double *deriv, double *jac,
                     double *normal, int32_t *spans, int64_t *firstOob)
{
    spline_t s = { nInd, nDep, order, nCoef, knots, coefs };
    int64_t win = 1, p, oob = -1;
    int iv, D = nInd > nDep ? nInd : nDep;
    if (nInd > ORACLE_MAX_IND) return -1;
    for (iv = 0; iv < nInd; ++iv) { if (order[iv] > ORACLE_MAX_ORDER) return -1; win *= order[iv]; }
    if ((flags & WANT_NORMAL) && abs(nInd - nDep) != 1) return -2;
    for (p = 0; p < N && oob < 0; ++p)
        for (iv = 0; iv < nInd; ++iv) {
            double u = uvw[p * nInd + iv];
            if (u < knots[iv][order[iv] - 1] || u > knots[iv][nCoef[iv]]) { oob = p; break; }
        }
    if (firstOob) *firstOob = oob;
#pragma omp parallel
    {
        double *scratch = (double *)malloc(sizeof(double) * (size_t)(nDep * win + 1));
        double *J = (double *)malloc(sizeof(double) * (size_t)(nDep * nInd + 1));
        double *col = (double *)malloc(sizeof(double) * (size_t)(nDep + 1));
        double *minor = (double *)malloc(sizeof(double) * (size_t)(D * D + 1));
        int32_t e[ORACLE_MAX_IND];
        int64_t q;
#pragma omp for schedule(static)
        for (q = 0; q < N; ++q) {
            const double *pt = uvw + q * nInd;
            int i, d, r, c;
            if (flags & WANT_VALUES) derivative_at(&s, NULL, pt, values + q * nDep, scratch, (flags & WANT_SPANS) ? spans + q * nInd : NULL);
            else if (flags & WANT_SPANS) for (i = 0; i < nInd; ++i) spans[q * nInd + i] = span_of(knots[i], order[i] + nCoef[i], order[i], pt[i]);
            if (flags & WANT_DERIV) derivative_at(&s, wrt, pt, deriv + q * nDep, scratch, NULL);
            if (flags & (WANT_JACOBIAN | WANT_NORMAL)) {
                for (i = 0; i < nInd; ++i) {
                    memset(e, 0, sizeof e);
                    e[i] = 1;
                    derivative_at(&s, e, pt, col, scratch, NULL);
                    for (d = 0; d < nDep; ++d) J[d * nInd + i] = col[d];
                }
                if (flags & WANT_JACOBIAN) memcpy(jac + q * nDep * nInd, J, sizeof(double) * (size_t)(nDep * nInd));
            }
            if (flags & WANT_NORMAL) {
                double *nrm = normal + q * D, sq = 0.0;
                int m = D - 1;
                for (i = 0; i < D; ++i) {
                    int rr = 0;
                    for (r = 0; r < D; ++r) {
                        if (r == i) continue;
                        for (c = 0; c < m; ++c) /* T = J (nDep>nInd) or J^T */
                            minor[rr * m + c] = (nInd > nDep) ? J[c * nInd + r] : J[r * nInd + c];
                        ++rr;
                    }
                    nrm[i] = normalSign * ((i & 1) ? -1.0 : 1.0) * det_lu(minor, m);
                }
                if (flags & WANT_NORMALIZE) {
                    for (i = 0; i < D; ++i) if (normalMask & (1u << i)) sq += nrm[i] * nrm[i];
                    sq = sqrt(sq);
                    for (i = 0; i < D; ++i) nrm[i] /= sq;
                }
            }
        }
        free(scratch); free(J); free(col); free(minor);
    }
    return 0;
}
