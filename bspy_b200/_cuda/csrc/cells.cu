// Cell-sorted evaluation of big scattered batches: bspy_cuda_eval_points_binned (struct-of-arrays outputs) and
// bspy_cuda_eval_points_aos (one result record per point).
//
// Replaces evaluate / derivative / jacobian / normal of bspy/_spline_evaluation.py:109-246 for N scattered points on a
// spline whose coefficients do not fit in L1: the points of a chunk are counting-sorted by knot-span cell so that the
// points a warp works on share their coefficient window.
#include <stdlib.h>

#include <mutex>

#include "scattered.cuh"

namespace bspy {

// ---- cell binning --------------------------------------------------------------------------------
// Scattered points on a spline whose coefficients do not fit in L1 gather a window of prod(order)*nDep doubles
// per point from L2 (1.5 KB for a tricubic volume, 3.9 KB for the 4-variate manifold): measured, the
// thread-per-point kernel is then bound by L2->SM traffic (~7 TB/s) at 10% of the FP64 roofline.  Binning makes
// the lanes of a warp share their window: the points of a chunk (small enough that its outputs stay in L2) are
// counting-sorted by knot-span cell, evaluated in cell order (window loads become L1 broadcasts) and written
// straight back to their original positions.  Same arithmetic per point, so results are bit-identical to the
// unbinned kernel.
// per-span records of one variable (runtime order): rec[s] = { knots[ix-(o-1)..ix-1] | 1/(knots[ix+t]-knots[ix-deg+t]) }
// with ix = o + s, the layout basis_from_span_record<O> reads; one thread per span
__global__ void __launch_bounds__(128) span_records_kernel(const double *__restrict__ kn, const int o, const int nCoef,
                                                           double *__restrict__ rec, const int stride)
{
    const int sp = blockIdx.x * blockDim.x + threadIdx.x;
    if (sp > nCoef - o) return;
    const int ix = o + sp;
    double *r = rec + (long long)sp * stride;
    for (int j = 0; j < o - 1; ++j) r[j] = kn[ix - (o - 1) + j];
    int at = o - 1;
    for (int deg = 1; deg < o; ++deg)
        for (int t = 0; t < deg; ++t) r[at++] = 1.0 / (kn[ix + t] - kn[ix - deg + t]);
    for (; at < stride; ++at) r[at] = 0.0;
}

static int span_rec_stride(int o) { return ((o - 1 + o * (o - 1) / 2) + 1) & ~1; }

__global__ void __launch_bounds__(256) bin_keys_kernel(const SplineDev s, const PointsDev in, const long long base, const int n,
                                                       int *__restrict__ keys, int *__restrict__ hist, int *__restrict__ rank,
                                                       const OutDev out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned active = __ballot_sync(0xffffffffu, t < n);
    if (t >= n) return;
    const long long p = base + t;
    int key = 0;
    bool outside = false;
    for (int iv = 0; iv < s.nInd; ++iv) {
        const double *k = s.knots[iv];
        const int o = s.order[iv];
        const double u = __ldg(in.uvw + p * in.pointStride + iv * in.varStride);
        outside |= (u < __ldg(k + o - 1)) | (u > __ldg(k + s.nCoef[iv]));
        const int ix = span_search_inner(k, o + s.nCoef[iv], o, u);
        if (out.spans) __stcs(out.spans + iv * out.ld + p, ix);
        key = key * (s.nCoef[iv] - o + 1) + (ix - o);
    }
    if (outside && out.firstOutside) report_outside((int64_t *)out.firstOutside, p);
    keys[t] = key;
    // one atomic per distinct cell in the warp (coherent inputs would otherwise serialise on one counter)
    const unsigned peers = __match_any_sync(active, key);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int first = 0;
    if (lane == leader) first = atomicAdd(hist + key, __popc(peers));
    if (rank) {   // position inside the cell: the scatter pass then needs no second round of atomics
        first = __shfl_sync(peers, first, leader);
        rank[t] = first + __popc(peers & ((1u << lane) - 1));
    }
}

// The same pass with P points per thread in flight (coordinates of all P points requested first, the P bisections
// advance together, P atomics outstanding), plus the rank of every point inside its cell; used by the sorted-record
// pipeline.  The pass is latency-bound (halving its occupancy doubles its time), and a quarter of the threads with
// four points each keeps more in flight than one point per thread.
template <int P>
__global__ void __launch_bounds__(128) bin_keys_batched_kernel(const SplineDev s, const PointsDev in, const long long base,
                                                               const int n, int *__restrict__ keys, int *__restrict__ hist,
                                                               int *__restrict__ rank, const OutDev out)
{
    const int t0 = blockIdx.x * (blockDim.x * P) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int key[P];
    bool outside[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { key[j] = 0; outside[j] = false; }
    for (int iv = 0; iv < s.nInd; ++iv) {
        const double *k = s.knots[iv];
        const int o = s.order[iv], nKnots = o + s.nCoef[iv];
        const double lo = __ldg(k + o - 1), hi = __ldg(k + s.nCoef[iv]);
        double u[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int t = t0 + j * blockDim.x;
            u[j] = t < n ? __ldcs(in.uvw + (base + t) * in.pointStride + iv * in.varStride) : lo;
        }
        int at[P], cnt[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            outside[j] |= (u[j] < lo) | (u[j] > hi);
            at[j] = o;
            cnt[j] = (u[j] != u[j]) ? 0 : nKnots - 2 * o;
        }
        bool more = true;
        while (more) {                                    // P upper-bound bisections side by side (span_search_inner)
            more = false;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                if (cnt[j] > 0) {
                    const int half = cnt[j] >> 1, mid = at[j] + half;
                    const bool le = __ldg(k + mid) <= u[j];
                    at[j] = le ? mid + 1 : at[j];
                    cnt[j] = le ? cnt[j] - half - 1 : half;
                    more |= cnt[j] > 0;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int ix = (u[j] != u[j]) ? nKnots - o : at[j];
            const int t = t0 + j * blockDim.x;
            if (out.spans && t < n) __stcs(out.spans + iv * out.ld + base + t, ix);
            key[j] = key[j] * (s.nCoef[iv] - o + 1) + (ix - o);
        }
    }
    int first[P];
    unsigned peers[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        const unsigned active = __ballot_sync(0xffffffffu, t < n);
        first[j] = 0;
        peers[j] = 0;
        if (t < n) {
            if (outside[j] && out.firstOutside) report_outside((int64_t *)out.firstOutside, base + t);
            keys[t] = key[j];
            peers[j] = __match_any_sync(active, key[j]);
            if (lane == __ffs(peers[j]) - 1) first[j] = atomicAdd(hist + key[j], __popc(peers[j]));
        }
    }
    if (rank) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int t = t0 + j * blockDim.x;
            if (t < n) {
                const int f = __shfl_sync(peers[j], first[j], __ffs(peers[j]) - 1);
                rank[t] = f + __popc(peers[j] & ((1u << lane) - 1));
            }
        }
    }
}

// ---- sorted-record variant (large chunks; no L2-residency assumption) --------------------------------------
// scatter: 32-byte point records in cell order (a full sector per point, so the scattered write needs no
// read-modify-write) at offset[cell] + rank (no atomics) + the inverse permutation (coalesced); P points per thread
template <int P>
__global__ void __launch_bounds__(128) bin_scatter_records_batched_kernel(const SplineDev s, const PointsDev in, const long long base,
                                                                          const int n, const int *__restrict__ keys,
                                                                          const int *__restrict__ offset, double *__restrict__ records,
                                                                          int2 *__restrict__ recKI, int *__restrict__ inv,
                                                                          const int writeInv)
{
    const int t0 = blockIdx.x * (blockDim.x * P) + threadIdx.x;
    int key[P], pos[P];
    double r[P][4];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        key[j] = t < n ? keys[t] : 0;
        pos[j] = t < n ? inv[t] : 0;
#pragma unroll
        for (int iv = 0; iv < 4; ++iv)
            r[j][iv] = (t < n && iv < s.nInd) ? __ldcs(in.uvw + (base + t) * in.pointStride + iv * in.varStride) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < P; ++j) pos[j] += __ldg(offset + key[j]) & 0x7fffffff;   // sign bit: parity flag of an even-padded sort
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        if (t < n) {
            // cell key in the low, index inside the chunk in the high 32 bits
            if (s.nInd <= 3) r[j][3] = __longlong_as_double(((long long)t << 32) | (unsigned)key[j]);
            else recKI[pos[j]] = make_int2(key[j], t);
            double2 *q = reinterpret_cast<double2 *>(records + 4LL * pos[j]);
            q[0] = make_double2(r[j][0], r[j][1]);
            q[1] = make_double2(r[j][2], r[j][3]);
            if (writeInv) inv[t] = pos[j];   // inverse permutation for the un-permute pass (rank before, position after)
        }
    }
}

// the same without the records: only the (cell key, index in the chunk) pair of every point goes to its sorted slot (8 bytes into
// an array that stays in L2) -- the evaluation kernel gathers the parameters of its points itself, from the chunk the keys
// pass has just read.  Saves the second read of the points and the 32-byte record write of every point.
template <int P>
__global__ void __launch_bounds__(128) bin_scatter_pairs_batched_kernel(const int n, const int *__restrict__ keys,
                                                                        const int *__restrict__ offset, int2 *__restrict__ recKI,
                                                                        int *__restrict__ inv, const int writeInv)
{
    const int t0 = blockIdx.x * (blockDim.x * P) + threadIdx.x;
    int key[P], pos[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        key[j] = t < n ? keys[t] : 0;
        pos[j] = t < n ? inv[t] : 0;
    }
#pragma unroll
    for (int j = 0; j < P; ++j) pos[j] += __ldg(offset + key[j]) & 0x7fffffff;
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int t = t0 + j * blockDim.x;
        if (t < n) {
            recKI[pos[j]] = make_int2(key[j], t);
            if (writeInv) inv[t] = pos[j];
        }
    }
}

// warp-aggregated slot claim: the lanes of a warp that share a cell take consecutive slots with one atomic
__device__ __forceinline__ int claim_slot(int *cursor, int key, unsigned active)
{
    const unsigned peers = __match_any_sync(active, key);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int first = 0;
    if (lane == leader) first = atomicAdd(cursor + key, __popc(peers));
    first = __shfl_sync(peers, first, leader);
    return first + __popc(peers & ((1u << lane) - 1));
}

// exclusive scan of hist[0..cells) in place, one CTA: 4096 counters per round (coalesced 16-byte loads, warp
// shuffles, one shared-memory hop between the warps), running total carried from round to round
// evenPad != 0: every cell's segment is rounded up to an even number of slots so that the aligned pairs (2i, 2i+1) of the
// sorted sequence never straddle a cell (two points per thread share their window loads); the odd cells' spare slot
// receives a dummy record (NaN parameters, index -1: evaluated, never stored).  hist[cells] receives the total.
__global__ void __launch_bounds__(1024) bin_scan_kernel(int *__restrict__ hist, const int cells, const int evenPad = 0,
                                                        double *__restrict__ records = nullptr, int2 *__restrict__ recKI = nullptr,
                                                        const int nInd = 0)
{
    __shared__ int warpSum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < cells; base += 4096) {
        const int i = base + 4 * threadIdx.x;
        int4 c = make_int4(0, 0, 0, 0);
        if (i + 3 < cells) c = *reinterpret_cast<const int4 *>(hist + i);
        else {
            if (i < cells) c.x = hist[i];
            if (i + 1 < cells) c.y = hist[i + 1];
            if (i + 2 < cells) c.z = hist[i + 2];
        }
        const int4 raw = c;
        if (evenPad) { c.x = (c.x + 1) & ~1; c.y = (c.y + 1) & ~1; c.z = (c.z + 1) & ~1; c.w = (c.w + 1) & ~1; }
        const int mine = c.x + c.y + c.z + c.w;
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = warpSum[lane];
            int wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, wi, off);
                if (lane >= off) wi += v;
            }
            warpSum[lane] = wi - w;                       // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int start = carry + warpSum[warp] + incl - mine;
        const int4 o = make_int4(start, start + c.x, start + c.x + c.y, start + c.x + c.y + c.z);
        if (i + 3 < cells) *reinterpret_cast<int4 *>(hist + i) = o;
        else {
            if (i < cells) hist[i] = o.x;
            if (i + 1 < cells) hist[i + 1] = o.y;
            if (i + 2 < cells) hist[i + 2] = o.z;
        }
        if (evenPad) {
            // odd cells: the dummy record of the spare slot is written by bin_pad_kernel (all SMs; 19 000 scattered stores from
            // this one CTA cost 36 us per chunk); the cell's parity travels in the sign bit of its offset
            const int cnt[4] = {raw.x, raw.y, raw.z, raw.w};
            int flagged[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i + j < cells && (cnt[j] & 1)) flagged[j] |= (int)0x80000000;
            if (i + 3 < cells) *reinterpret_cast<int4 *>(hist + i) = make_int4(flagged[0], flagged[1], flagged[2], flagged[3]);
            else {
                if (i < cells) hist[i] = flagged[0];
                if (i + 1 < cells) hist[i + 1] = flagged[1];
                if (i + 2 < cells) hist[i + 2] = flagged[2];
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = start + mine;
        __syncthreads();
    }
    if (threadIdx.x == 0) hist[cells] = carry;
}

// dummy records of the odd cells of an even-padded sort (offsets flagged by bin_scan_kernel): NaN parameters, index -1
__global__ void __launch_bounds__(256) bin_pad_kernel(const int *__restrict__ hist, const int cells, double *__restrict__ records,
                                                      int2 *__restrict__ recKI, const int nInd)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells || hist[c] >= 0) return;
    const long long pos = (long long)(hist[c + 1] & 0x7fffffff) - 1;        // last slot of the cell's (even) segment; hist[cells] = total
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double2 *q = reinterpret_cast<double2 *>(records + 4 * pos);
    const long long ki = (-1LL << 32) | (unsigned)c;
    q[0] = make_double2(nan, nan);
    q[1] = make_double2(nan, nInd <= 3 ? __longlong_as_double(ki) : nan);
    if (nInd > 3) recKI[pos] = make_int2(c, -1);
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const int *__restrict__ keys, int *__restrict__ cursor, const int n,
                                                          int *__restrict__ perm, int *__restrict__ sortedKey)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned active = __ballot_sync(0xffffffffu, t < n);
    if (t >= n) return;
    const int key = keys[t];
    const int pos = claim_slot(cursor, key, active);
    perm[pos] = t;
    sortedKey[pos] = key;
}

// un-permute: a warp takes 32 consecutive points and pulls their 32 result records (each a run of whole sectors
// somewhere in the sorted array) into shared memory with 16-byte cp.async copies -- lanes run along the records,
// every sector is requested once, and all of a warp's copies are in flight together -- then every lane reads its
// own record and the warp writes the struct-of-arrays outputs coalesced (8 warps side by side: 2 KB per plane).
constexpr int UNPERM_WARPS = 8;
template <int S_>
__global__ void __launch_bounds__(UNPERM_WARPS * 32) bin_unpermute_kernel(const double *__restrict__ aos, const int aosStride,
                                                                          const int *__restrict__ inv, const long long base,
                                                                          const int n, const int nDep, const int nJ,
                                                                          const int nNormal, const OutDev out)
{
    extern __shared__ __align__(16) double tile[];         // per warp: 32 rows of (stride + 2) doubles
    const int S = S_ ? S_ : aosStride;
    const int pitch = S + 2, S2 = S >> 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *tw = tile + (long long)warp * 32 * pitch;
    const int first = (blockIdx.x * UNPERM_WARPS + warp) * 32;
    if (first >= n) return;
    const int t = first + lane;
    const int myRec = t < n ? __ldg(inv + t) : -1;
    const unsigned twAddr = (unsigned)__cvta_generic_to_shared(tw);
    const int total = 32 * S2;
#pragma unroll 8
    for (int idx = lane; idx < total; idx += 32) {
        const int r = idx / S2, c = idx - r * S2;
        const int rec = __shfl_sync(0xffffffffu, myRec, r);
        if (rec >= 0) {
            const double *src = aos + (long long)rec * S + 2 * c;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(twAddr + (unsigned)(r * pitch + 2 * c) * 8u), "l"(src) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (t >= n) return;
    const long long p = base + t;
    const double *mine = tw + lane * pitch;
    if (out.values) {
        for (int k = 0; k < nDep; ++k) __stcs(out.values + k * out.ld + p, mine[k]);
    }
    if (out.jacobian) {
        for (int k = 0; k < nJ; ++k) __stcs(out.jacobian + k * out.ld + p, mine[nDep + k]);
    }
    if (out.normal) {
        for (int k = 0; k < nNormal; ++k) __stcs(out.normal + k * out.ld + p, mine[nDep + nJ + k]);
    }
}

template <int S_>
static int launch_unpermute(const double *aos, int stride, const int *inv, long long base, int n, int nDep, int nJ, int nN,
                            const OutDev &out, cudaStream_t st)
{
    const size_t smem = sizeof(double) * UNPERM_WARPS * 32 * (stride + 2);
    if (int rc = allow_dynamic_smem(bin_unpermute_kernel<S_>, smem)) return rc;
    bin_unpermute_kernel<S_><<<(n + UNPERM_WARPS * 32 - 1) / (UNPERM_WARPS * 32), UNPERM_WARPS * 32, smem, st>>>(
        aos, stride, inv, base, n, nDep, nJ, nN, out);
    return 0;
}

// points per chunk in sorted-record mode (BSPY_BIN_REC_CHUNK_LOG2 overrides for experiments)
static long long bin_rec_chunk()
{
    const long long lg = option(OPT_BIN_REC_CHUNK_LOG2, 22);
    return 1LL << (lg < 16 ? 16 : (lg > 26 ? 26 : lg));
}
#define BIN_REC_CHUNK bin_rec_chunk()
constexpr long long BIN_CHUNK_MAX = 1 << 20;  // workspace is sized for this many points per chunk

// Points per chunk: the outputs of a chunk are scattered back to their original positions 8 bytes at a time, which
// is only cheap while the chunk's output region stays in L2 (measured optimum: ~56 MB of outputs per chunk; larger
// chunks give more points per cell but turn the scatter into DRAM read-modify-writes and are 2x slower).
static long long bin_chunk(long long outBytesPerPoint)
{
    long long v = option(OPT_BIN_CHUNK, (56LL << 20) / (outBytesPerPoint > 0 ? outBytesPerPoint : 8));
    v = v / 1024 * 1024;
    if (v < 65536) v = 65536;
    if (v > BIN_CHUNK_MAX) v = BIN_CHUNK_MAX;
    return v;
}
constexpr long long BIN_MAX_CELLS = 1 << 18;  // histogram / scan size limit

static long long binned_cells(const SplineDev &s)
{
    long long cells = 1;
    for (int i = 0; i < s.nInd; ++i) {
        cells *= (s.nCoef[i] - s.order[i] + 1);
        if (cells > BIN_MAX_CELLS) return 0;
    }
    return cells;
}

static int bin_mode(long long N)
{
    // 0: scatter results 8 bytes at a time within L2-sized chunks; 1: sorted 32-byte point records, array-of-structs
    // results and an un-permute pass over 4 Mi-point chunks (every scattered access is a whole sector)
    return option(OPT_BIN_MODE, N >= (1 << 21) ? 1 : 0) ? 1 : 0;
}

static int aos_stride(const SplineDev &s)
{
    const int D = (s.nInd - s.nDep == 1 || s.nDep - s.nInd == 1) ? (s.nInd > s.nDep ? s.nInd : s.nDep) : 0;
    return (s.nDep + s.nDep * s.nInd + D + 3) & ~3;
}

static long long pad64(long long n) { return (n + 63) / 64 * 64; }

static long long span_records_bytes(const SplineDev &s)
{
    long long doubles = 0;
    for (int i = 0; i < s.nInd; ++i) doubles += (long long)(s.nCoef[i] - s.order[i] + 1) * span_rec_stride(s.order[i]);
    return 8 * pad64(doubles);
}

// bytes of workspace for the binned path, 0 when binning does not apply to this spline; aosOut: results go to
// caller-visible records (always the sorted-record pipeline, no private result records)
static long long records_half_bytes(const SplineDev &s, long long chunk, bool ownAos);

static bool binning_applies(const SplineDev &s, long long N)
{
    if (s.nInd < 2 || s.nInd > 4 || N < 65536) return false;
    long long window = s.nDep;
    for (int i = 0; i < s.nInd; ++i) window *= s.order[i];
    if (window * 8 < 512) return false;                         // small windows: the gather is cheap anyway
    if (s.depStride * s.nDep * 8 < 128 * 1024) return false;    // the whole spline fits in L1
    if (!binned_cells(s)) return false;
    return find_fixed(s, 0) != nullptr;
}

struct ImageEntry;
static const ImageEntry *find_image(const SplineDev &s, int jac, int code);

// run-time shape version of the layout (the builder is not templated)
struct ImageLayout {
    int n, nDep, OL, Q, window, recs, size;
    int recOffset[BSPY_MAX_IND], recStride[BSPY_MAX_IND];
};

static ImageLayout image_layout(const SplineDev &s)
{
    ImageLayout L{};
    L.n = s.nInd; L.nDep = s.nDep; L.OL = s.order[s.nInd - 1];
    L.Q = 1;
    for (int i = 0; i < s.nInd - 1; ++i) L.Q *= s.order[i];
    L.window = s.nDep * L.Q * 4;
    int at = 0;
    for (int i = 0; i < s.nInd; ++i) {
        L.recOffset[i] = at;
        L.recStride[i] = span_rec_stride(s.order[i]);
        at += L.recStride[i];
    }
    L.recs = (at + 3) & ~3;
    L.size = L.window + L.recs;
    return L;
}


// cell images are built once per call: worth it when the batch holds a few points per cell at least
static bool images_apply(const SplineDev &s, long long N)
{
    const long long code = option(OPT_IMAGE, 0);
    if (code < 0 || !option(OPT_SPAN_RECORDS, 1)) return false;
    if (!find_image(s, 1, (int)code) && !find_image(s, 0, (int)code)) return false;
    return N >= 8 * binned_cells(s);
}

static long long cell_images_bytes(const SplineDev &s, long long N)
{
    if (!images_apply(s, N)) return 0;
    return 8 * pad64(binned_cells(s) * (long long)image_layout(s).size);
}

struct PolyEntry;
static const PolyEntry *find_poly(const SplineDev &s, int code);
static bool poly_entry_pair(const PolyEntry *e);

static int poly_matrix_stride(int o) { return (o * o + 2 + 1) & ~1; }

struct PolyLayout {
    int n, nDep, o[BSPY_MAX_IND], cs[BSPY_MAX_IND], pm[BSPY_MAX_IND];   // orders, compact strides, matrix strides
    int perDep, perDepPad, E, slot;                                       // slot = doubles per cell image
    int padded;                                                           // rows of the last variable padded to 4 doubles ([d][q][4], ImageShape)
};

static PolyLayout poly_layout(const SplineDev &s, bool padded = false)
{
    PolyLayout L{};
    L.padded = padded ? 1 : 0;
    L.n = s.nInd; L.nDep = s.nDep;
    L.perDep = 1;
    for (int i = s.nInd - 1; i >= 0; --i) {
        L.o[i] = s.order[i];
        L.cs[i] = L.perDep;
        L.perDep *= s.order[i];
        L.pm[i] = poly_matrix_stride(s.order[i]);
    }
    L.perDepPad = padded ? L.perDep / s.order[s.nInd - 1] * 4 : (L.perDep + 1) & ~1;
    L.E = L.nDep * L.perDep;
    L.slot = L.nDep * L.perDepPad + 4;
    return L;
}


// cell polynomial images (built once per call): worth it when the batch holds a few points per cell at least
static bool poly_applies(const SplineDev &s, long long N)
{
    const long long code = option(OPT_CELL_POLY, 1);
    if (code <= 0 || s.nInd < 2 || s.nInd > 4) return false;
    for (int i = 0; i < s.nInd; ++i)
        if (s.order[i] > 4) return false;
    const PolyEntry *e = find_poly(s, (int)code);
    if (!e) return false;
    // one image per cell in the caller's workspace: not for splines whose images would pass 1 GiB
    if (binned_cells(s) * (long long)poly_layout(s, poly_entry_pair(e)).slot * 8 > (1LL << 30)) return false;
    return N >= 8 * binned_cells(s);
}

static long long cell_poly_bytes(const SplineDev &s, long long N)
{
    if (!poly_applies(s, N)) return 0;
    long long doubles = 8;                                            // flag (+ padding)
    for (int i = 0; i < s.nInd; ++i) doubles += (long long)(s.nCoef[i] - s.order[i] + 1) * poly_matrix_stride(s.order[i]);
    doubles = pad64(doubles) + pad64(binned_cells(s) * (long long)poly_layout(s, poly_entry_pair(find_poly(s, (int)option(OPT_CELL_POLY, 1)))).slot);
    return 8 * doubles;
}

long long binned_workspace(const SplineDev &s, long long N, bool aosOut)
{
    if (!binning_applies(s, N)) return 0;
    const long long cells = binned_cells(s);
    if (aosOut || bin_mode(N) == 1) {
        const long long chunk = N < BIN_REC_CHUNK ? N : BIN_REC_CHUNK;
        return 2 * records_half_bytes(s, chunk, !aosOut) + span_records_bytes(s) + cell_images_bytes(s, N) + cell_poly_bytes(s, N);
    }
    const long long chunk = N < BIN_CHUNK_MAX ? N : BIN_CHUNK_MAX;
    return 3 * 4 * pad64(chunk) + 4 * pad64(cells + 1);
}

long long binned_workspace(const SplineDev &s, long long N) { return binned_workspace(s, N, false); }

// Two internal helper streams (+ the events that fork from / join to the caller's stream) for the sorted-record
// pipeline: the sort / un-permute passes are memory-bound, the evaluation FP64-bound, so chunk c+1 is sorted
// (high-priority stream) while chunk c is evaluated (low-priority stream).  Capturable in a CUDA graph.
// Sets are pooled per device behind a mutex: a call owns its set exclusively while it ENQUEUES (event record / wait
// pairs of two host threads can therefore never interleave) and returns it when it is done enqueueing; work already
// enqueued keeps the dependencies it captured, so the next owner may reuse streams and events right away.
struct BinStreams {
    cudaStream_t sort = nullptr, eval = nullptr;
    cudaEvent_t fork = nullptr, sorted[2] = {nullptr, nullptr}, evaluated[2] = {nullptr, nullptr}, joinSort = nullptr,
                joinEval = nullptr;
    int device = -1;
    BinStreams *next = nullptr;
};

static std::mutex g_binStreamsMutex;
static BinStreams *g_binStreamsFree[64] = {};

static BinStreams *acquire_bin_streams()
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    {
        std::lock_guard<std::mutex> lock(g_binStreamsMutex);
        if (BinStreams *b = g_binStreamsFree[dev]) {
            g_binStreamsFree[dev] = b->next;
            b->next = nullptr;
            return b;
        }
    }
    BinStreams *b = new BinStreams();
    b->device = dev;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = lowest priority (numerically largest)
    bool ok = cudaStreamCreateWithPriority(&b->sort, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&b->eval, cudaStreamNonBlocking, lo) == cudaSuccess;
    cudaEvent_t *evs[] = {&b->fork, &b->sorted[0], &b->sorted[1], &b->evaluated[0], &b->evaluated[1], &b->joinSort, &b->joinEval};
    for (cudaEvent_t *e : evs) ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {   // leave the partially built set to the process; the caller runs without overlap
        cudaGetLastError();
        delete b;
        return nullptr;
    }
    return b;
}

static void release_bin_streams(BinStreams *b)
{
    std::lock_guard<std::mutex> lock(g_binStreamsMutex);
    b->next = g_binStreamsFree[b->device];
    g_binStreamsFree[b->device] = b;
}

// ---- cell evaluation with the first contraction stage on the FP64 tensor pipe ---------------------------------------
// In cell order the points of an 8-row group share their coefficient window, so the contraction of the LAST variable
// (contiguous in the spline, order OL <= 4 = K of DMMA.8x8x4) is a small dense product per dependent variable d:
//     T[p][q] = sum_k Bw[p][k] * C[d][q][k],   Tw[p][q] = sum_k dBw[p][k] * C[d][q][k]
// with p = 8 points (rows of A), q = the prod(outer orders) window positions of the other variables (columns, 8 per
// tile) -- (cubic volume: 2 tiles x 2 kinds per group and dependent variable).  The D fragments go to a per-warp tile
// in shared memory, T[point][kind][q], and every lane then finishes ITS point with the register-resident
// sum-factorised contraction over the outer variables (ContractT, same recursion as Contract<>): value, derivatives with
// respect to the outer variables from T, derivative with respect to the last variable from Tw.
//   * warps are persistent and walk contiguous runs of sorted 32-point tiles; the window image of a cell ([d][q][k], k
//     padded to 4) is staged once per cell with cp.async into one of two slots per warp; a tile that straddles cells
//     runs the MMAs of the straddling group once per cell with the other rows' A operand zeroed;
//   * B fragments are one conflict-free LDS.64 per tile (lane l reads element 32 t + l of the image);
//   * ~450 warp instructions per tile instead of ~1060 for the all-DFMA staged kernel, the same FP64-pipe work.
// Summation order differs from Contract<> in the last variable only (the MMA's k order): results agree with the
// other kernels to rounding, not bit for bit.
template <class Ord, int NDEP>
struct CellShape {
    static constexpr int n = Ord::n;
    static constexpr int OL = Ord::at(n - 1);
    __host__ __device__ static constexpr int qstride(int iv)          // stride of outer variable iv in q (iv < n-1)
    {
        int st = 1;
        for (int m = n - 2; m > iv; --m) st *= Ord::at(m);
        return st;
    }
    static constexpr int Q = n > 1 ? qstride(0) * Ord::at(0) : 1;     // window positions of the outer variables
    static constexpr int NT = (Q + 7) / 8;                            // MMA column tiles per dependent variable
    static constexpr int QP = NT * 8;
    static constexpr int winDoubles = NDEP * QP * 4;                  // image [d][QP][4]
    static constexpr int TS = 2 * QP + 2;                             // T row [kind][QP] + 2: odd number of 16-byte units
    static constexpr int ES = 12;                                     // A-operand exchange row: Bw[4] | dBw[4] | pad
    static constexpr int perWarp = 2 * winDoubles + 32 * TS + 32 * ES;
    static_assert(OL <= 4, "the last variable's order is the K of DMMA.8x8x4");
};

template <class Ord, int NDEP>
__device__ __forceinline__ void stage_window_k4(const SplineDev &s, int key, double *dst, const int lane)
{
    using CS = CellShape<Ord, NDEP>;
    long long base = 0;
#pragma unroll
    for (int iv = Ord::n - 1; iv >= 0; --iv) {
        const int m = s.nCoef[iv] - Ord::at(iv) + 1;
        base += (long long)(key % m) * s.stride[iv];
        key /= m;
    }
    const unsigned dstAddr = (unsigned)__cvta_generic_to_shared(dst);
    constexpr int perDep = CS::Q * CS::OL, total = perDep * NDEP;
#pragma unroll
    for (int e0 = 0; e0 < total; e0 += 32) {
        const int e = e0 + lane;
        if (total % 32 == 0 || e < total) {
            const int d = e / perDep;
            const int rem = e - d * perDep;
            const int q = rem / CS::OL, k = rem - q * CS::OL;
            long long src = base + (long long)d * s.depStride + k;
#pragma unroll
            for (int iv = 0; iv < Ord::n - 1; ++iv) src += (long long)((q / CS::qstride(iv)) % Ord::at(iv)) * s.stride[iv];
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dstAddr + (unsigned)((d * CS::QP + q) * 4 + k) * 8u), "l"(s.coefs + src) : "memory");
        }
    }
}

// Contract the outer variables L .. n-2 of one dependent variable's T row (Tv: value kind, Tw: last-variable-derivative
// kind) starting at window position off: v value, gw derivative w.r.t. the last variable, g[m] w.r.t. outer variable m.
template <int L, class Ord>
struct ContractT {
    static constexpr int n = Ord::n;
    __device__ __forceinline__ static void run(const double *__restrict__ Tv, const double *__restrict__ Tw, const int off,
                                               const FixedCtx<Ord, 1, true> &c, double &v, double &gw, double (&g)[n])
    {
        constexpr int O = Ord::at(L);
        v = 0.0;
        gw = 0.0;
#pragma unroll
        for (int m = L; m < n - 1; ++m) g[m] = 0.0;
        if constexpr (L == n - 2) {
            double x[O], y[O];
            load_run<O>(Tv, off, x);
            load_run<O>(Tw, off, y);
#pragma unroll
            for (int i = 0; i < O; ++i) {
                v = fma(x[i], c.B[L][i], v);
                g[L] = fma(x[i], c.dB[L][i], g[L]);
                gw = fma(y[i], c.B[L][i], gw);
            }
        } else {
            using CS = CellShape<Ord, 1>;
#pragma unroll
            for (int i = 0; i < O; ++i) {
                double cv, cgw, cg[n];
                ContractT<L + 1, Ord>::run(Tv, Tw, off + i * CS::qstride(L), c, cv, cgw, cg);
                v = fma(cv, c.B[L][i], v);
                g[L] = fma(cv, c.dB[L][i], g[L]);
                gw = fma(cgw, c.B[L][i], gw);
#pragma unroll
                for (int m = L + 1; m < n - 1; ++m) g[m] = fma(cg[m], c.B[L][i], g[m]);
            }
        }
    }
};

__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

constexpr int CELL_WARPS = 2;   // warps per CTA (small CTAs: the per-warp shared-memory slice decides the occupancy)

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, int MINB>
__global__ void __launch_bounds__(CELL_WARPS * 32, MINB) eval_cell_mma_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                               const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using CS = CellShape<Ord, NDEP>;
    static_assert(NIND >= 2, "curves have no outer variables");
    extern __shared__ __align__(16) double cellSmem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = lane >> 2, c2 = lane & 3;
    double *win = cellSmem + warp * CS::perWarp;                      // two window slots
    double *T = win + 2 * CS::winDoubles;
    double *exch = T + 32 * CS::TS;
    for (int i = lane; i < 2 * CS::winDoubles; i += 32) win[i] = 0.0;  // K / column padding stays zero for good
    __syncwarp();
    int slotKey0 = -1, slotKey1 = -1;
    const long long tiles = (N + 31) >> 5, nWarps = gridDim.x * (long long)CELL_WARPS;
    const long long per = (tiles + nWarps - 1) / nWarps;
    const long long firstTile = (blockIdx.x * (long long)CELL_WARPS + warp) * per;
    const long long endTile = firstTile + per < tiles ? firstTile + per : tiles;
    double2 r0 = make_double2(0.0, 0.0), r1 = r0;
    long long k4 = -1;
    auto fetch = [&](long long tile) {
        const long long t = tile * 32 + lane;
        if (tile < endTile && t < N) {
            const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
            r0 = __ldcs(rp);
            r1 = __ldcs(rp + 1);
            if constexpr (NIND > 3) k4 = __ldcs(reinterpret_cast<const long long *>(in.recKI) + t);
        }
    };
    fetch(firstTile);
    for (long long tile = firstTile; tile < endTile; ++tile) {
        const long long t = tile * 32 + lane;
        const bool live = t < N;
        FixedCtx<Ord, 1, true> c;
        int ix[NIND];
        double u[NIND];
        u[0] = r0.x;
        if constexpr (NIND > 1) u[1] = r0.y;
        if constexpr (NIND > 2) u[2] = r1.x;
        if constexpr (NIND > 3) u[3] = r1.y;
        const long long ki = NIND > 3 ? k4 : __double_as_longlong(r1.y);
        const int key = live ? (int)ki : -1;
        const long long dest = out.aosScatter ? out.aosBase + (ki >> 32) : t;
        fetch(tile + 1);                                            // next tile's records arrive under this tile's arithmetic
        {
            int k = key < 0 ? 0 : key;
#pragma unroll
            for (int iv = NIND - 1; iv >= 0; --iv) {
                const int m = s.nCoef[iv] - Ord::at(iv) + 1;
                ix[iv] = Ord::at(iv) + k % m;
                k /= m;
            }
        }
        bool outside = false;
        setup_variable<0, Ord, 1, true>(s, u[0], 0, c, ix, outside, true, in.spanRec[0]);
        if constexpr (NIND > 1) setup_variable<1, Ord, 1, true>(s, u[1], 0, c, ix, outside, true, in.spanRec[1]);
        if constexpr (NIND > 2) setup_variable<2, Ord, 1, true>(s, u[2], 0, c, ix, outside, true, in.spanRec[2]);
        if constexpr (NIND > 3) setup_variable<3, Ord, 1, true>(s, u[3], 0, c, ix, outside, true, in.spanRec[3]);
        __syncwarp();                                               // the previous tile is done with exch / T
        {
            // A operand of this lane's point: basis values and first derivatives of the last variable, K padded to 4
            double2 *e = reinterpret_cast<double2 *>(exch + lane * CS::ES);
            double a[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                a[k] = (k < CS::OL && live) ? c.B[NIND - 1][k < CS::OL ? k : 0] : 0.0;
                a[4 + k] = (k < CS::OL && live) ? c.dB[NIND - 1][k < CS::OL ? k : 0] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] = make_double2(a[2 * k], a[2 * k + 1]);
        }
        double v[NDEP];
        double g[NIND][NDEP];
        bool done = !live;
        while (true) {
            const unsigned pending = __ballot_sync(0xffffffffu, !done);
            if (!pending) break;
            // up to two distinct cells per round, each in the slot that already holds it or freshly staged
            const int k0 = __shfl_sync(0xffffffffu, key, __ffs(pending) - 1);
            const bool in0 = !done && key == k0;
            const unsigned rest = __ballot_sync(0xffffffffu, !done && !in0);
            const int k1 = rest ? __shfl_sync(0xffffffffu, key, __ffs(rest) - 1) : -1;
            const bool in1 = !done && !in0 && key == k1;
            int s0, s1 = -1;
            bool staged = false;
            if (k0 == slotKey0) s0 = 0;
            else if (k0 == slotKey1) s0 = 1;
            else {
                s0 = (k1 >= 0 && k1 == slotKey0) ? 1 : 0;
                stage_window_k4<Ord, NDEP>(s, k0, win + s0 * CS::winDoubles, lane);
                if (s0) slotKey1 = k0; else slotKey0 = k0;
                staged = true;
            }
            if (k1 >= 0) {
                s1 = 1 - s0;
                if ((s1 ? slotKey1 : slotKey0) != k1) {
                    stage_window_k4<Ord, NDEP>(s, k1, win + s1 * CS::winDoubles, lane);
                    if (s1) slotKey1 = k1; else slotKey0 = k1;
                    staged = true;
                }
            }
            if (staged) asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            __syncwarp();                                            // windows and the A exchange are visible
            const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
            const bool member = in0 || in1;
#pragma unroll
            for (int d = 0; d < NDEP; ++d) {
                // stage 1: DMMA per cell of the round, group of 8 rows and column tile
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const unsigned mask = cc ? m1 : m0;
                    if (!mask) continue;
                    const double *w = win + (cc ? s1 : s0) * CS::winDoubles + d * CS::QP * 4;
                    double bf[CS::NT];
#pragma unroll
                    for (int tt = 0; tt < CS::NT; ++tt) bf[tt] = w[32 * tt + lane];
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) {
                        const unsigned gm = (mask >> (8 * gq)) & 0xffu;
                        if (!gm) continue;
                        const bool rowIn = (gm >> row) & 1u;
                        const double *e = exch + (8 * gq + row) * CS::ES;
                        const double av = rowIn ? e[c2] : 0.0, ad = rowIn ? e[4 + c2] : 0.0;
                        double *trow = T + (8 * gq + row) * CS::TS + 2 * c2;
#pragma unroll
                        for (int tt = 0; tt < CS::NT; ++tt) {
                            double x0, x1, y0, y1;
                            dmma_8x8x4(x0, x1, av, bf[tt]);
                            dmma_8x8x4(y0, y1, ad, bf[tt]);
                            if (rowIn) {
                                *reinterpret_cast<double2 *>(trow + 8 * tt) = make_double2(x0, x1);
                                *reinterpret_cast<double2 *>(trow + CS::QP + 8 * tt) = make_double2(y0, y1);
                            }
                        }
                    }
                }
                __syncwarp();
                // stage 2: this lane's point, outer variables
                if (member) {
                    double vd, gw, gd[NIND];
                    ContractT<0, Ord>::run(T + lane * CS::TS, T + lane * CS::TS + CS::QP, 0, c, vd, gw, gd);
                    v[d] = vd;
                    g[NIND - 1][d] = gw;
#pragma unroll
                    for (int m = 0; m < NIND - 1; ++m) g[m][d] = gd[m];
                }
                __syncwarp();                                        // T is rewritten by the next dependent variable
            }
            if (member) {
                store_result_record<NIND, NDEP, true>(s, out, out.aos + dest * out.aosStride, v, g);
                done = true;
            }
        }
    }
}

// ---- cell polynomials: the window of every cell converted to powers of (u - mid-cell) ------------------------------------
// In cell order every point of a warp evaluates the SAME polynomial piece.  The Cox-de Boor kernels spend a tenth of their
// FP64 instructions on the basis recurrence and contract the window with value / derivative dot products (2 O FMAs per row
// and variable); a piece written in powers of t_v = u_v - (middle of the span of variable v),
//     S(u) = sum_k P[k_0 .. k_{n-1}] t_0^k_0 .. t_{n-1}^k_{n-1},    P = (M_0 x .. x M_{n-1}) C,   M_v[k][j] = B_j^(k)(m_v) / k!
// needs no basis at all and a nested Horner scheme with derivatives: 2 O - 3 FMAs per row instead of 2 O -- 369 instead of
// 678 FP64 instructions per point for the tricubic nDep-3 volume with its jacobian, 936 instead of 1854 for the 4-variate
// nDep-6 manifold -- with the same loads.  Centred at mid-cell, every |t_v| <= h_v / 2 and every knot gap of the derivative
// stages contains the span itself, so the terms stay of the order of the coefficients: measured against long-double
// arithmetic the values and jacobians are as close to the exact result as the recurrence's (worst 0.08 vs 0.09 of the
// parity bar on the config-4 shape).  A pre-pass per call builds
//   * per variable and span the conversion matrix M_v (derivative stages of the recurrence at the mid-span), m_v, h_v / 2;
//   * per cell the image { P in the compact layout of WindowShape | m_0 .. m_{n-1} } and checks its conditioning
//     (sum |P| prod (h_v/2)^k_v <= 64 max|C| of the cell's window per dependent variable, everything finite; a cell no
//     parameter can reach -- an interior span of zero width -- is skipped): one failing cell clears a device flag and the
//     evaluation of the whole call is done by the recurrence kernels instead (both kernels are enqueued, each looks at
//     the flag first: PointsDev::gate).
// The evaluation kernel is eval_staged2_kernel's structure: persistent warps over contiguous runs of sorted tiles, the
// cell's image staged once per cell (contiguous 16-byte cp.async copies) into one of two slots per warp.

template <int O>
__device__ __forceinline__ void poly_matrix_span(const double *__restrict__ kn, const int sp, double *__restrict__ out)
{
    const int ix = O + sp;
    double left[O > 1 ? O - 1 : 1], rc[O > 1 ? O * (O - 1) / 2 : 1], dl[O > 1 ? O - 1 : 1];
#pragma unroll
    for (int j = 0; j < O - 1; ++j) left[j] = kn[ix - (O - 1) + j];
    int at = 0;
#pragma unroll
    for (int deg = 1; deg < O; ++deg)
#pragma unroll
        for (int t = 0; t < deg; ++t) rc[at++] = 1.0 / (kn[ix + t] - kn[ix - deg + t]);
    const double k0 = kn[ix - 1], k1 = kn[ix], m = 0.5 * (k0 + k1);
#pragma unroll
    for (int j = 0; j < O - 1; ++j) dl[j] = m - left[j];
    double invFact = 1.0;
#pragma unroll
    for (int k = 0; k < O; ++k) {
        double bk[O], unused[O];
        basis_core<O, false>(dl, rc, k, bk, unused);
        if (k > 1) invFact /= (double)k;
#pragma unroll
        for (int j = 0; j < O; ++j) out[k * O + j] = bk[j] * invFact;
    }
    out[O * O] = m;
    out[O * O + 1] = 0.5 * (k1 - k0);
}

__global__ void __launch_bounds__(128) poly_matrix_kernel(const double *__restrict__ kn, const int o, const int nCoef,
                                                          double *__restrict__ out, const int stride)
{
    const int sp = blockIdx.x * blockDim.x + threadIdx.x;
    if (sp > nCoef - o) return;
    double *r = out + (long long)sp * stride;
    switch (o) {
        case 1: poly_matrix_span<1>(kn, sp, r); break;
        case 2: poly_matrix_span<2>(kn, sp, r); break;
        case 3: poly_matrix_span<3>(kn, sp, r); break;
        default: poly_matrix_span<4>(kn, sp, r); break;
    }
}

constexpr int POLY_BUILD_WARPS = 4, POLY_BUILD_MAX_E = 1024;

struct PolyMatrices { const double *m[BSPY_MAX_IND]; };

// one warp per cell: window -> tensor transform, axis by axis, in the warp's two shared-memory buffers -> image
__global__ void __launch_bounds__(POLY_BUILD_WARPS * 32) build_cell_poly_kernel(const SplineDev s, const PolyLayout L, const long long cells,
                                                                                const PolyMatrices PM, double *__restrict__ images,
                                                                                int *__restrict__ flag)
{
    extern __shared__ __align__(16) double polyBuild[];
    // per window position r (all variables, last fastest): its digits k_v (8 bits each) and its offset in the coefficient
    // array -- computed once per CTA, so that the per-cell loops below are free of integer divisions
    __shared__ unsigned digTab[256];
    __shared__ long long offTab[256];
    for (int r = threadIdx.x; r < L.perDep; r += blockDim.x) {
        unsigned dig = 0;
        long long off = 0;
        int rr = r;
        for (int iv = L.n - 1; iv >= 0; --iv) {
            const int k = rr % L.o[iv];
            rr /= L.o[iv];
            dig |= (unsigned)k << (8 * iv);
            off += (long long)k * s.stride[iv];
        }
        digTab[r] = dig;
        offTab[r] = off;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *bufA = polyBuild + (size_t)warp * (2 * L.E + 32), *bufB = bufA + L.E, *cm = bufB + L.E;   // cm: max |C| per dependent variable
    for (long long cell = blockIdx.x * (long long)POLY_BUILD_WARPS + warp; cell < cells; cell += gridDim.x * (long long)POLY_BUILD_WARPS) {
        int span[BSPY_MAX_IND];
        const double *M[BSPY_MAX_IND];
        double hp[BSPY_MAX_IND];
        long long base = 0;
        bool unreachable = false, bad = false;
        {
            long long key = cell;
            for (int iv = L.n - 1; iv >= 0; --iv) {
                const int m = s.nCoef[iv] - L.o[iv] + 1;
                span[iv] = (int)(key % m);
                key /= m;
                base += (long long)span[iv] * s.stride[iv];
                M[iv] = PM.m[iv] + (long long)span[iv] * L.pm[iv];
                const double hh = M[iv][L.o[iv] * L.o[iv] + 1];
                hp[iv] = hh;
                if (!(hh > 0.0)) {
                    if (hh == 0.0 && span[iv] > 0 && span[iv] < m - 1) unreachable = true;   // interior span of zero width
                    else bad = true;
                }
            }
        }
        __syncwarp();
        if (!unreachable && !bad) {
            for (int d = 0; d < L.nDep; ++d) {
                const double *gsrc = s.coefs + base + (long long)d * s.depStride;
                double mx = 0.0;
                for (int r = lane; r < L.perDep; r += 32) {
                    const double c = __ldg(gsrc + offTab[r]);
                    bufA[d * L.perDep + r] = c;
                    mx = fmax(mx, fabs(c));
                }
#pragma unroll
                for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
                if (lane == 0 && d < 32) cm[d] = mx;
            }
            __syncwarp();
            double *src = bufA, *dst = bufB;
            for (int v = 0; v < L.n; ++v) {
                const int cs = L.cs[v], ov = L.o[v];
                const double *Mv = M[v];
                for (int d = 0; d < L.nDep; ++d) {
                    const double *sd = src + d * L.perDep;
                    for (int r = lane; r < L.perDep; r += 32) {
                        const int k = (digTab[r] >> (8 * v)) & 0xff;
                        const double *col = sd + r - k * cs;
                        double acc = 0.0;
                        for (int j = 0; j < ov; ++j) acc = fma(Mv[k * ov + j], col[j * cs], acc);
                        dst[d * L.perDep + r] = acc;
                    }
                }
                __syncwarp();
                double *t = src; src = dst; dst = t;
            }
            // conditioning: the Horner terms against the largest coefficient of the window, per dependent variable
            for (int d = 0; d < L.nDep; ++d) {
                double terms = 0.0;
                for (int r = lane; r < L.perDep; r += 32) {
                    double w = fabs(src[d * L.perDep + r]);
                    const unsigned dig = digTab[r];
                    for (int iv = 0; iv < L.n; ++iv)
                        for (int k = (dig >> (8 * iv)) & 0xff; k > 0; --k) w *= hp[iv];
                    terms += w;
                }
#pragma unroll
                for (int off = 16; off; off >>= 1) terms += __shfl_xor_sync(0xffffffffu, terms, off);
                if (!(terms <= 64.0 * cm[d < 32 ? d : 31])) bad = true;     // NaN / inf fail too
            }
            double *img = images + cell * (long long)L.slot;
            if (L.padded) {
                const int ol = L.o[L.n - 1];
                for (int d = 0; d < L.nDep; ++d)
                    for (int r = lane; r < L.perDepPad; r += 32) {
                        const int q = r >> 2, k = r & 3;
                        img[d * L.perDepPad + r] = k < ol ? src[d * L.perDep + q * ol + k] : 0.0;
                    }
            } else {
                for (int d = 0; d < L.nDep; ++d)
                    for (int r = lane; r < L.perDep; r += 32) img[d * L.perDepPad + r] = src[d * L.perDep + r];
                if (L.perDepPad > L.perDep && lane < L.nDep) img[lane * L.perDepPad + L.perDep] = 0.0;
            }
            if (lane < 4) img[L.nDep * L.perDepPad + lane] = lane < L.n ? M[lane][L.o[lane] * L.o[lane]] : 0.0;
        }
        if (bad && lane == 0) atomicExch(flag, 0);
    }
}

// nested Horner with derivatives over variables L .. n-1 of the compact image (NDT dependent variables starting at w);
// q = offset of the fixed indices of variables < L.  v value, g[m] derivative with respect to variable m >= L.
template <int L, class Ord, int NDEP, int NDT>
struct HornerS {
    using WS = WindowShape<Ord, NDEP>;
    static constexpr int n = Ord::n;
    __device__ __forceinline__ static void run(const double *__restrict__ w, const int q, const double (&t)[n], double (&v)[NDT],
                                               double (&g)[n][NDT])
    {
        constexpr int O = Ord::at(L);
        if constexpr (L == n - 1) {
#pragma unroll
            for (int d = 0; d < NDT; ++d) {
                double x[O];
                load_run<O>(w, q + d * WS::perDepPad, x);
                double val = x[O - 1], der = 0.0;
#pragma unroll
                for (int k = O - 2; k >= 0; --k) {
                    der = (k == O - 2) ? val : fma(der, t[L], val);
                    val = fma(val, t[L], x[k]);
                }
                v[d] = val;
                g[L][d] = der;
            }
        } else {
#pragma unroll
            for (int i = O - 1; i >= 0; --i) {
                double cv[NDT];
                double cg[n][NDT];
                HornerS<L + 1, Ord, NDEP, NDT>::run(w, q + i * WS::stride(L), t, cv, cg);
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    if (i == O - 1) {
                        v[d] = cv[d];
                        g[L][d] = 0.0;
#pragma unroll
                        for (int m = L + 1; m < n; ++m) g[m][d] = cg[m][d];
                    } else {
                        g[L][d] = (i == O - 2) ? v[d] : fma(g[L][d], t[L], v[d]);
                        v[d] = fma(v[d], t[L], cv[d]);
#pragma unroll
                        for (int m = L + 1; m < n; ++m) g[m][d] = fma(g[m][d], t[L], cg[m][d]);
                    }
                }
            }
        }
    }
};

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_poly_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                               const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using WS = WindowShape<Ord, NDEP>;
    constexpr int SLOT = WS::size + 4;                              // image: compact polynomial | mid-cell of every variable
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    if (gate_closed(in)) return;
    extern __shared__ __align__(16) double polySlots[];             // per warp: two slots
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *w0 = polySlots + warp * 2 * SLOT;
    int slotKey0 = -1, slotKey1 = -1;                               // cells held by the two slots (warp-uniform)
    bool prefetched = false;                                        // an image requested ahead of its first use is still in flight
    int cells = 1;
#pragma unroll
    for (int iv = 0; iv < NIND; ++iv) cells *= s.nCoef[iv] - Ord::at(iv) + 1;
    const long long tiles = (N + 31) >> 5, nWarps = gridDim.x * 4LL;
    const long long per = (tiles + nWarps - 1) / nWarps;
    const long long firstTile = (blockIdx.x * 4LL + warp) * per;
    const long long endTile = firstTile + per < tiles ? firstTile + per : tiles;
    double2 r0 = make_double2(0.0, 0.0), r1 = r0;
    long long k4 = -1, k4next = -1;
    // records == nullptr: the sorted sequence holds (cell key, index) pairs only and the lane gathers its point's parameters from
    // the caller's array (one 32-byte sector of a chunk that is still in L2 from the keys pass); the pair of tile + 2 and the
    // point of tile + 1 are requested under tile's arithmetic
    const bool gather = in.records == nullptr;
    auto fetchKey = [&](long long tile) {
        const long long t = tile * 32 + lane;
        k4next = (tile < endTile && t < N) ? __ldcs(reinterpret_cast<const long long *>(in.recKI) + t) : -1;
    };
    auto fetch = [&](long long tile) {
        const long long t = tile * 32 + lane;
        if (gather) {
            k4 = k4next;                                            // requested a tile ago
            fetchKey(tile + 1);
            if (tile < endTile && t < N) {
                const double *up = in.uvw + (in.base + (k4 >> 32)) * in.pointStride;
                r0.x = __ldg(up);
                if constexpr (NIND > 1) r0.y = __ldg(up + in.varStride);
                if constexpr (NIND > 2) r1.x = __ldg(up + 2 * in.varStride);
                if constexpr (NIND > 3) r1.y = __ldg(up + 3 * in.varStride);
            }
            return;
        }
        if (tile < endTile && t < N) {
            const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
            r0 = __ldcs(rp);
            r1 = __ldcs(rp + 1);
            if constexpr (NIND > 3) k4 = __ldcs(reinterpret_cast<const long long *>(in.recKI) + t);
        }
    };
    auto stage = [&](int key, double *dst) {                        // the warp copies the image of cell `key` into a slot
        const double *src = in.images + (long long)key * SLOT;
        const unsigned dstAddr = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
        for (int c0 = 0; c0 < SLOT / 2; c0 += 32) {
            const int c = c0 + lane;
            if ((SLOT / 2) % 32 == 0 || c < SLOT / 2)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstAddr + 16u * c), "l"(src + 2 * c) : "memory");
        }
    };
    if (gather) fetchKey(firstTile);
    fetch(firstTile);
    for (long long tile = firstTile; tile < endTile; ++tile) {
        const long long t = tile * 32 + lane;
        const bool live = t < N;
        double u[NIND];
        u[0] = r0.x;
        if constexpr (NIND > 1) u[1] = r0.y;
        if constexpr (NIND > 2) u[2] = r1.x;
        if constexpr (NIND > 3) u[3] = r1.y;
        const long long ki = (NIND > 3 || gather) ? k4 : __double_as_longlong(r1.y);
        const int key = live ? (int)ki : -1;
        const long long dest = out.aosScatter ? out.aosBase + (ki >> 32) : t;
        fetch(tile + 1);                                            // next tile's records arrive under this tile's arithmetic
        if (prefetched) {                                           // requested a tile ago: long since there
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            prefetched = false;
        }
        bool done = !live;
        while (true) {
            const unsigned pending = __ballot_sync(0xffffffffu, !done);
            if (!pending) break;
            // up to two distinct cells per pass, each in the slot that already holds it or freshly staged
            const int k0 = __shfl_sync(0xffffffffu, key, __ffs(pending) - 1);
            const bool in0 = !done && key == k0;
            const unsigned rest = __ballot_sync(0xffffffffu, !done && !in0);
            const int k1 = rest ? __shfl_sync(0xffffffffu, key, __ffs(rest) - 1) : -1;
            const bool in1 = !done && !in0 && key == k1;
            int s0, s1 = -1;
            bool staged = false;
            if (k0 == slotKey0) s0 = 0;
            else if (k0 == slotKey1) s0 = 1;
            else {
                s0 = (k1 >= 0 && k1 == slotKey0) ? 1 : 0;
                stage(k0, w0 + s0 * SLOT);
                if (s0) slotKey1 = k0; else slotKey0 = k0;
                staged = true;
            }
            if (k1 >= 0) {
                s1 = 1 - s0;
                if ((s1 ? slotKey1 : slotKey0) != k1) {
                    stage(k1, w0 + s1 * SLOT);
                    if (s1) slotKey1 = k1; else slotKey0 = k1;
                    staged = true;
                }
            }
            if (staged) {
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            if (k1 < 0 && k0 + 1 < cells) {
                // the last pass of this tile needs one slot: the image of the NEXT cell of the sorted sequence (cells are dense:
                // its points follow within a tile or a few) travels to the other slot under this tile's arithmetic
                const int other = 1 - s0;
                if ((other ? slotKey1 : slotKey0) != k0 + 1) {
                    stage(k0 + 1, w0 + other * SLOT);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    if (other) slotKey1 = k0 + 1; else slotKey0 = k0 + 1;
                    prefetched = true;
                }
            }
            if (in0 || in1) {
                const double *w = w0 + (in1 ? s1 : s0) * SLOT;
                double tt[NIND];
                {
                    const double2 m01 = *reinterpret_cast<const double2 *>(w + WS::size);
                    tt[0] = u[0] - m01.x;
                    if constexpr (NIND > 1) tt[1] = u[1] - m01.y;
                    if constexpr (NIND > 2) {
                        const double2 m23 = *reinterpret_cast<const double2 *>(w + WS::size + 2);
                        tt[2] = u[2] - m23.x;
                        if constexpr (NIND > 3) tt[3] = u[3] - m23.y;
                    }
                }
                double *rec = out.aos + dest * out.aosStride;
                if constexpr (NDT == NDEP) {
                    double v[NDEP];
                    double g[NIND][NDEP];
                    HornerS<0, Ord, NDEP, NDEP>::run(w, 0, tt, v, g);
                    store_result_record<NIND, NDEP, true>(s, out, rec, v, g);
                } else {
#pragma unroll 1
                    for (int d0 = 0; d0 < NDEP; d0 += NDT) {
                        double vt[NDT];
                        double gt[NIND][NDT];
                        HornerS<0, Ord, NDEP, NDT>::run(w + d0 * WS::perDepPad, 0, tt, vt, gt);
                        store_result_tile<NIND, NDEP, NDT, true>(rec, d0, vt, gt);
                    }
                }
                done = true;
            }
            __syncwarp();                                            // slots may be overwritten by the next pass / tile
        }
    }
}

// ---- cell images: one compact, padded copy of every cell's window (+ its span records) in global memory ------------
// The thread-per-point kernel in cell order walks its window through L1 with run-time strides: 486 scalar loads and
// ~700 address instructions per point for the 4-variate nDep-6 manifold (ncu: 4.0 long-scoreboard + 1.9 LG-throttle stalls
// per issue, FP64 pipe 40 %), and the warp-staged shared-memory variant of that shape spills.  Here a pre-pass (once per
// call) writes, for every cell, the image  [d][q][4]  (q = window position of all variables but the last, rows of the last
// variable padded to 4 doubles = one 32-byte sector) followed by the cell's span records; the evaluation kernel is then
// straight-line code per point: one 256-bit load per window row at an immediate offset from the cell's base pointer
// (all lanes of a warp in the same cell: one broadcast sector from L1), the span records by broadcast loads, no
// per-lane key decode, no address arithmetic.  Arithmetic and summation order are those of Contract<>: bit-identical.
template <class Ord, int NDEP>
struct ImageShape {
    static constexpr int n = Ord::n;
    static constexpr int OL = Ord::at(n - 1);
    __host__ __device__ static constexpr int qstride(int iv)
    {
        int st = 1;
        for (int m = n - 2; m > iv; --m) st *= Ord::at(m);
        return st;
    }
    static constexpr int Q = n > 1 ? qstride(0) * Ord::at(0) : 1;
    static constexpr int perDep = Q * 4;
    static constexpr int window = NDEP * perDep;
    static constexpr int recs = (CellRecords<Ord>::size + 3) & ~3;
    static constexpr int size = window + recs;                        // doubles per cell: a multiple of 4
    static_assert(OL <= 4, "rows of the last variable are padded to 4 doubles");
};

// one warp per cell
__global__ void __launch_bounds__(256) build_cell_images_kernel(const SplineDev s, const ImageLayout L, const long long cells,
                                                                const double *const *spanRecUnused, const PointsDev recs,
                                                                double *__restrict__ images)
{
    const int lane = threadIdx.x & 31;
    for (long long cell = blockIdx.x * 8LL + (threadIdx.x >> 5); cell < cells; cell += gridDim.x * 8LL) {
        int span[BSPY_MAX_IND];
        long long base = 0;
        {
            long long k = cell;
            for (int iv = L.n - 1; iv >= 0; --iv) {
                const int m = s.nCoef[iv] - s.order[iv] + 1;
                span[iv] = (int)(k % m);
                k /= m;
                base += (long long)span[iv] * s.stride[iv];
            }
        }
        double *img = images + cell * L.size;
        for (int e = lane; e < L.window; e += 32) {
            const int k = e & 3, row = e >> 2;
            const int d = row / L.Q;
            int q = row - d * L.Q;
            double x = 0.0;
            if (k < L.OL) {
                long long src = base + (long long)d * s.depStride + k;
                for (int iv = L.n - 2; iv >= 0; --iv) {
                    const int o = s.order[iv];
                    src += (long long)(q % o) * s.stride[iv];
                    q /= o;
                }
                x = __ldg(s.coefs + src);
            }
            img[e] = x;
        }
        for (int iv = 0; iv < L.n; ++iv)
            for (int j = lane; j < L.recStride[iv]; j += 32)
                img[L.window + L.recOffset[iv] + j] = __ldg(recs.spanRec[iv] + (long long)span[iv] * L.recStride[iv] + j);
        for (int j = L.recOffset[L.n - 1] + L.recStride[L.n - 1] + lane; j < L.recs; j += 32) img[L.window + j] = 0.0;
    }
}

__device__ __forceinline__ void ld_row256(const double *__restrict__ p, double (&x)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
}

// Contract variables L .. n-1 of the image (dependent variables d0 .. d0+NDT-1) at window position q of the outer
// variables; same recursion and summation order as Contract<>.
template <int L, class Ord, int NDEP, int NDT, bool JAC>
struct ContractI {
    using IS = ImageShape<Ord, NDEP>;
    __device__ __forceinline__ static void run(const double *__restrict__ img, const int q, const FixedCtx<Ord, NDT, JAC> &c,
                                               double (&v)[NDT], double (&g)[Ord::n][NDT])
    {
        constexpr int O = Ord::at(L);
#pragma unroll
        for (int d = 0; d < NDT; ++d) v[d] = 0.0;
        if constexpr (JAC) {
#pragma unroll
            for (int m = L; m < Ord::n; ++m)
#pragma unroll
                for (int d = 0; d < NDT; ++d) g[m][d] = 0.0;
        }
        if constexpr (L == Ord::n - 1) {
#pragma unroll
            for (int d = 0; d < NDT; ++d) {
                double x[4];
                ld_row256(img + (d * IS::Q + q) * 4, x);
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    v[d] = fma(x[i], c.B[L][i], v[d]);
                    if constexpr (JAC) g[L][d] = fma(x[i], c.dB[L][i], g[L][d]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < O; ++i) {
                double cv[NDT];
                double cg[Ord::n][NDT];
                ContractI<L + 1, Ord, NDEP, NDT, JAC>::run(img, q + i * IS::qstride(L), c, cv, cg);
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    v[d] = fma(cv[d], c.B[L][i], v[d]);
                    if constexpr (JAC) {
                        g[L][d] = fma(cv[d], c.dB[L][i], g[L][d]);
#pragma unroll
                        for (int m = L + 1; m < Ord::n; ++m) g[m][d] = fma(cg[m][d], c.B[L][i], g[m][d]);
                    }
                }
            }
        }
    }
};

template <int IV, class Ord, int NDT, bool JAC>
__device__ __forceinline__ void setup_variable_image(const double *__restrict__ rec, double u, int d, FixedCtx<Ord, NDT, JAC> &c)
{
    constexpr int O = Ord::at(IV);
    double b0[O], b1[O];
    basis_from_span_record<O, JAC>(rec + CellRecords<Ord>::offset(IV), u, d, b0, b1);
#pragma unroll
    for (int j = 0; j < O; ++j) {
        c.B[IV][j] = b0[j];
        if constexpr (JAC) c.dB[IV][j] = b1[j];
    }
}

// STAGE: the passes over the dependent variables (rolled loop) leave their results in a [slot][thread] tile of shared
// memory and every thread then writes its whole record in one go (whole sectors at the record's final position)
template <int NIND, int O0, int O1, int O2, int O3, int NDEP, bool JAC, int NDT, int MINB, bool STAGE>
__global__ void __launch_bounds__(128, MINB) eval_image_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using IS = ImageShape<Ord, NDEP>;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    constexpr int R = JAC ? NDEP * (1 + NIND) : NDEP;
    extern __shared__ __align__(16) double recTile[];                // STAGE: R slots x 128 threads
    const long long t = blockIdx.x * 128LL + threadIdx.x;
    if (t >= N) return;
    const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
    const double2 r0 = __ldcs(rp), r1 = __ldcs(rp + 1);
    double u[NIND];
    u[0] = r0.x;
    if constexpr (NIND > 1) u[1] = r0.y;
    if constexpr (NIND > 2) u[2] = r1.x;
    if constexpr (NIND > 3) u[3] = r1.y;
    const long long ki = NIND > 3 ? __ldcs(reinterpret_cast<const long long *>(in.recKI) + t) : __double_as_longlong(r1.y);
    const long long dest = out.aosScatter ? out.aosBase + (ki >> 32) : t;
    const double *img = in.images + (long long)(int)ki * IS::size;
    FixedCtx<Ord, NDT, JAC> c;
    setup_variable_image<0, Ord, NDT, JAC>(img + IS::window, u[0], wrt.d[0], c);
    if constexpr (NIND > 1) setup_variable_image<1, Ord, NDT, JAC>(img + IS::window, u[1], wrt.d[1], c);
    if constexpr (NIND > 2) setup_variable_image<2, Ord, NDT, JAC>(img + IS::window, u[2], wrt.d[2], c);
    if constexpr (NIND > 3) setup_variable_image<3, Ord, NDT, JAC>(img + IS::window, u[3], wrt.d[3], c);
    double *rec = out.aos + dest * out.aosStride;
    if constexpr (NDT == NDEP) {
        double v[NDEP];
        double g[NIND][NDEP];
        ContractI<0, Ord, NDEP, NDEP, JAC>::run(img, 0, c, v, g);
        store_result_record<NIND, NDEP, JAC>(s, out, rec, v, g);
    } else {
        double *mine = recTile + threadIdx.x;
#pragma unroll 1
        for (int d0 = 0; d0 < NDEP; d0 += NDT) {
            double vt[NDT];
            double gt[NIND][NDT];
            ContractI<0, Ord, NDEP, NDT, JAC>::run(img + d0 * IS::perDep, 0, c, vt, gt);
            if constexpr (STAGE) {
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    mine[(d0 + d) * 128] = vt[d];
                    if constexpr (JAC) {
#pragma unroll
                        for (int iv = 0; iv < NIND; ++iv) mine[(NDEP + (d0 + d) * NIND + iv) * 128] = gt[iv][d];
                    }
                }
            } else {
                store_result_tile<NIND, NDEP, NDT, JAC>(rec, d0, vt, gt);
            }
        }
        if constexpr (STAGE) {
            constexpr int RP = (R + 3) & ~3;
            double2 *q2 = reinterpret_cast<double2 *>(rec);
#pragma unroll
            for (int j = 0; j < RP / 2; ++j) {
                const double a = 2 * j < R ? mine[(2 * j) * 128] : 0.0, b = 2 * j + 1 < R ? mine[(2 * j + 1) * 128] : 0.0;
                if (2 * j < out.aosStride) __stcs(q2 + j, make_double2(a, b));
            }
        }
    }
}

// ---- two points per thread --------------------------------------------------------------------------------------------
// The innermost contraction stage needs every window coefficient in a register of every lane: 1.5 KB (tricubic, nDep 3) /
// 3.9 KB (4-variate, nDep 6) per point through the 128 B/clk load-return path of the SM, which is as many cycles as the
// FP64 pipe needs for the point's arithmetic (ncu: l1tex 68-77 % busy on every variant of the one-point kernels, which is
// why they all run at the same speed whatever their occupancy or instruction count).  Two points of the SAME cell per
// thread use every loaded row twice: half the load-return traffic per point, twice the independent FMA chains.  The sort
// rounds every cell's segment up to an even length (bin_scan_kernel, evenPad), so the aligned pairs of the sorted sequence
// never straddle a cell; spare slots hold dummy records that are evaluated and not stored.
template <int L, class Ord, int NDEP, int NDT>
struct ContractI2 {
    using IS = ImageShape<Ord, NDEP>;
    using Ctx = FixedCtx<Ord, NDT, true>;
    __device__ __forceinline__ static void run(const double *__restrict__ img, const int q, const Ctx &c0, const Ctx &c1,
                                               double (&v0)[NDT], double (&g0)[Ord::n][NDT], double (&v1)[NDT], double (&g1)[Ord::n][NDT])
    {
        constexpr int O = Ord::at(L);
#pragma unroll
        for (int d = 0; d < NDT; ++d) { v0[d] = 0.0; v1[d] = 0.0; }
#pragma unroll
        for (int m = L; m < Ord::n; ++m)
#pragma unroll
            for (int d = 0; d < NDT; ++d) { g0[m][d] = 0.0; g1[m][d] = 0.0; }
        if constexpr (L == Ord::n - 1) {
#pragma unroll
            for (int d = 0; d < NDT; ++d) {
                double x[4];
                ld_row256(img + (d * IS::Q + q) * 4, x);
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    v0[d] = fma(x[i], c0.B[L][i], v0[d]);
                    v1[d] = fma(x[i], c1.B[L][i], v1[d]);
                    g0[L][d] = fma(x[i], c0.dB[L][i], g0[L][d]);
                    g1[L][d] = fma(x[i], c1.dB[L][i], g1[L][d]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < O; ++i) {
                double cv0[NDT], cv1[NDT];
                double cg0[Ord::n][NDT], cg1[Ord::n][NDT];
                ContractI2<L + 1, Ord, NDEP, NDT>::run(img, q + i * IS::qstride(L), c0, c1, cv0, cg0, cv1, cg1);
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    v0[d] = fma(cv0[d], c0.B[L][i], v0[d]);
                    v1[d] = fma(cv1[d], c1.B[L][i], v1[d]);
                    g0[L][d] = fma(cv0[d], c0.dB[L][i], g0[L][d]);
                    g1[L][d] = fma(cv1[d], c1.dB[L][i], g1[L][d]);
#pragma unroll
                    for (int m = L + 1; m < Ord::n; ++m) {
                        g0[m][d] = fma(cg0[m][d], c0.B[L][i], g0[m][d]);
                        g1[m][d] = fma(cg1[m][d], c1.B[L][i], g1[m][d]);
                    }
                }
            }
        }
    }
};

// basis values and first derivatives of TWO parameters of the same span from one read of the span record
template <int IV, class Ord, int NDT>
__device__ __forceinline__ void setup_variable_pair(const double *__restrict__ rec, double ua, double ub, FixedCtx<Ord, NDT, true> &ca,
                                                    FixedCtx<Ord, NDT, true> &cb)
{
    constexpr int O = Ord::at(IV);
    using R = SpanRec<O>;
    const double *rp0 = rec + CellRecords<Ord>::offset(IV);
    double r[R::stride > 0 ? R::stride : 1];
    if constexpr (R::stride > 0) {
        const double2 *rp = reinterpret_cast<const double2 *>(rp0);
#pragma unroll
        for (int j = 0; j < R::stride / 2; ++j) {
            const double2 x = __ldg(rp + j);
            r[2 * j] = x.x;
            r[2 * j + 1] = x.y;
        }
    }
    double rc[O > 1 ? O * (O - 1) / 2 : 1];
#pragma unroll
    for (int j = 0; j < O * (O - 1) / 2; ++j) rc[j] = r[O - 1 + j];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const double u = p ? ub : ua;
        double dl[O > 1 ? O - 1 : 1], b0[O], b1[O];
#pragma unroll
        for (int j = 0; j < O - 1; ++j) dl[j] = u - r[j];
        basis_core<O, true>(dl, rc, 0, b0, b1);
        FixedCtx<Ord, NDT, true> &c = p ? cb : ca;
#pragma unroll
        for (int j = 0; j < O; ++j) { c.B[IV][j] = b0[j]; c.dB[IV][j] = b1[j]; }
    }
}

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_image2_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                 const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using IS = ImageShape<Ord, NDEP>;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    constexpr int R = NDEP * (1 + NIND), RP = (R + 3) & ~3;
    extern __shared__ __align__(16) double recTile[];                // [point 0 | point 1][R slots][128 threads]
    if (gate_closed(in)) return;
    // one pair per thread; a launch with fewer CTAs than pairs / 128 (the fallback behind a gate: cheap to skip) strides over them
    const long long total = (long long)__ldg(in.sortedTotal);
    for (long long P = blockIdx.x * 128LL + threadIdx.x; 2 * P < total; P += gridDim.x * 128LL) {   // pair index: sorted slots 2P and 2P + 1
    double ra[4], rb[4];
    ld_row256(in.records + 8 * P, ra);
    ld_row256(in.records + 8 * P + 4, rb);
    long long kia, kib;
    if constexpr (NIND > 3) {
        const longlong2 kk = __ldcs(reinterpret_cast<const longlong2 *>(in.recKI) + P);
        kia = kk.x;
        kib = kk.y;
    } else {
        kia = __double_as_longlong(ra[3]);
        kib = __double_as_longlong(rb[3]);
    }
    const double *img = in.images + (long long)(int)kia * IS::size;
    if (in.prefetchImages) {
        // the first warp to touch a cell would otherwise take its misses one dependent row at a time: request every
        // 128-byte line of the image at once (the lanes of a warp ask for the same lines: one request per line)
        constexpr int lines = (IS::size * 8 + 127) / 128;
#pragma unroll
        for (int l = 0; l < lines; ++l) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(img) + 128 * l));
    }
    FixedCtx<Ord, NDT, true> c0, c1;
    setup_variable_pair<0, Ord, NDT>(img + IS::window, ra[0], rb[0], c0, c1);
    if constexpr (NIND > 1) setup_variable_pair<1, Ord, NDT>(img + IS::window, ra[1], rb[1], c0, c1);
    if constexpr (NIND > 2) setup_variable_pair<2, Ord, NDT>(img + IS::window, ra[2], rb[2], c0, c1);
    if constexpr (NIND > 3) setup_variable_pair<3, Ord, NDT>(img + IS::window, ra[3], rb[3], c0, c1);
    double *mine = recTile + threadIdx.x;
#pragma unroll 1
    for (int d0 = 0; d0 < NDEP; d0 += NDT) {
        double v0[NDT], v1[NDT];
        double g0[NIND][NDT], g1[NIND][NDT];
        ContractI2<0, Ord, NDEP, NDT>::run(img + d0 * IS::perDep, 0, c0, c1, v0, g0, v1, g1);
#pragma unroll
        for (int d = 0; d < NDT; ++d) {
            mine[(d0 + d) * 128] = v0[d];
            mine[(R + d0 + d) * 128] = v1[d];
#pragma unroll
            for (int iv = 0; iv < NIND; ++iv) {
                mine[(NDEP + (d0 + d) * NIND + iv) * 128] = g0[iv][d];
                mine[(R + NDEP + (d0 + d) * NIND + iv) * 128] = g1[iv][d];
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const long long ki = p ? kib : kia;
        const long long idx = ki >> 32;
        if (idx < 0) continue;                                        // dummy slot of an odd cell
        const long long dest = out.aosScatter ? out.aosBase + idx : 2 * P + p;
        double *rec = out.aos + dest * out.aosStride;
        const double *src = mine + p * R * 128;
#pragma unroll
        for (int j = 0; j < RP / 4; ++j) {
            if (4 * j < out.aosStride) {
                double x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = 4 * j + e < R ? src[(4 * j + e) * 128] : 0.0;
                if (out.aosWide)
                    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(rec + 4 * j), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
                else {
                    __stcs(reinterpret_cast<double2 *>(rec + 4 * j), make_double2(x[0], x[1]));
                    __stcs(reinterpret_cast<double2 *>(rec + 4 * j) + 1, make_double2(x[2], x[3]));
                }
            }
        }
    }
    }
}

// ---- cell polynomials, two points of the same cell per thread -------------------------------------------------------------
// eval_image2_kernel's structure (images read from global memory through L1 by 256-bit broadcast loads at immediate offsets,
// even-padded cell segments, results through a shared-memory tile) on the padded polynomial images: no span records, no basis,
// nested Horner for both points from every loaded row.  Half the FP64 instructions and 14 * NDT instead of (14 + 6) * NDT + 24
// live doubles per point against the recurrence version.
template <int L, class Ord, int NDEP, int NDT>
struct HornerI2 {
    using IS = ImageShape<Ord, NDEP>;
    static constexpr int n = Ord::n;
    __device__ __forceinline__ static void run(const double *__restrict__ img, const int q, const double (&t0)[n], const double (&t1)[n],
                                               double (&v0)[NDT], double (&g0)[n][NDT], double (&v1)[NDT], double (&g1)[n][NDT])
    {
        constexpr int O = Ord::at(L);
        if constexpr (L == n - 1) {
#pragma unroll
            for (int d = 0; d < NDT; ++d) {
                double x[4];
                ld_row256(img + (d * IS::Q + q) * 4, x);
                double a = x[O - 1], b = x[O - 1], da = 0.0, db = 0.0;
#pragma unroll
                for (int k = O - 2; k >= 0; --k) {
                    da = (k == O - 2) ? a : fma(da, t0[L], a);
                    db = (k == O - 2) ? b : fma(db, t1[L], b);
                    a = fma(a, t0[L], x[k]);
                    b = fma(b, t1[L], x[k]);
                }
                v0[d] = a; v1[d] = b;
                g0[L][d] = da; g1[L][d] = db;
            }
        } else {
#pragma unroll
            for (int i = O - 1; i >= 0; --i) {
                double cv0[NDT], cv1[NDT];
                double cg0[n][NDT], cg1[n][NDT];
                HornerI2<L + 1, Ord, NDEP, NDT>::run(img, q + i * IS::qstride(L), t0, t1, cv0, cg0, cv1, cg1);
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    if (i == O - 1) {
                        v0[d] = cv0[d]; v1[d] = cv1[d];
                        g0[L][d] = 0.0; g1[L][d] = 0.0;
#pragma unroll
                        for (int m = L + 1; m < n; ++m) { g0[m][d] = cg0[m][d]; g1[m][d] = cg1[m][d]; }
                    } else {
                        g0[L][d] = (i == O - 2) ? v0[d] : fma(g0[L][d], t0[L], v0[d]);
                        g1[L][d] = (i == O - 2) ? v1[d] : fma(g1[L][d], t1[L], v1[d]);
                        v0[d] = fma(v0[d], t0[L], cv0[d]);
                        v1[d] = fma(v1[d], t1[L], cv1[d]);
#pragma unroll
                        for (int m = L + 1; m < n; ++m) {
                            g0[m][d] = fma(g0[m][d], t0[L], cg0[m][d]);
                            g1[m][d] = fma(g1[m][d], t1[L], cg1[m][d]);
                        }
                    }
                }
            }
        }
    }
};

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_poly2_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using IS = ImageShape<Ord, NDEP>;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    constexpr int SLOT = IS::window + 4;                             // padded polynomial | mid-cell of every variable
    constexpr int R = NDEP * (1 + NIND), RP = (R + 3) & ~3;
    if (gate_closed(in)) return;
    extern __shared__ __align__(16) double recTile[];                // [point 0 | point 1][R slots][128 threads]
    const long long P = blockIdx.x * 128LL + threadIdx.x;            // pair index: sorted slots 2P and 2P + 1
    if (2 * P >= (long long)__ldg(in.sortedTotal)) return;
    double ra[4], rb[4];
    ld_row256(in.records + 8 * P, ra);
    ld_row256(in.records + 8 * P + 4, rb);
    long long kia, kib;
    if constexpr (NIND > 3) {
        const longlong2 kk = __ldcs(reinterpret_cast<const longlong2 *>(in.recKI) + P);
        kia = kk.x;
        kib = kk.y;
    } else {
        kia = __double_as_longlong(ra[3]);
        kib = __double_as_longlong(rb[3]);
    }
    const double *img = in.images + (long long)(int)kia * SLOT;
    if (in.prefetchImages) {
        constexpr int lines = (SLOT * 8 + 127) / 128;
#pragma unroll
        for (int l = 0; l < lines; ++l) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(img) + 128 * l));
    }
    double t0[NIND], t1[NIND];
    {
        double mid[4];
        ld_row256(img + IS::window, mid);
#pragma unroll
        for (int iv = 0; iv < NIND; ++iv) { t0[iv] = ra[iv] - mid[iv]; t1[iv] = rb[iv] - mid[iv]; }
    }
    double *mine = recTile + threadIdx.x;
#pragma unroll 1
    for (int d0 = 0; d0 < NDEP; d0 += NDT) {
        double v0[NDT], v1[NDT];
        double g0[NIND][NDT], g1[NIND][NDT];
        HornerI2<0, Ord, NDEP, NDT>::run(img + d0 * IS::perDep, 0, t0, t1, v0, g0, v1, g1);
#pragma unroll
        for (int d = 0; d < NDT; ++d) {
            mine[(d0 + d) * 128] = v0[d];
            mine[(R + d0 + d) * 128] = v1[d];
#pragma unroll
            for (int iv = 0; iv < NIND; ++iv) {
                mine[(NDEP + (d0 + d) * NIND + iv) * 128] = g0[iv][d];
                mine[(R + NDEP + (d0 + d) * NIND + iv) * 128] = g1[iv][d];
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const long long ki = p ? kib : kia;
        const long long idx = ki >> 32;
        if (idx < 0) continue;                                        // dummy slot of an odd cell
        const long long dest = out.aosScatter ? out.aosBase + idx : 2 * P + p;
        double *rec = out.aos + dest * out.aosStride;
        const double *src = mine + p * R * 128;
#pragma unroll
        for (int j = 0; j < RP / 4; ++j) {
            if (4 * j < out.aosStride) {
                double x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = 4 * j + e < R ? src[(4 * j + e) * 128] : 0.0;
                if (out.aosWide)
                    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(rec + 4 * j), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
                else {
                    __stcs(reinterpret_cast<double2 *>(rec + 4 * j), make_double2(x[0], x[1]));
                    __stcs(reinterpret_cast<double2 *>(rec + 4 * j) + 1, make_double2(x[2], x[3]));
                }
            }
        }
    }
}

struct ImageEntry {
    int nInd, o[4], nDep, jac, code;      // code = 1000 * PAIR + 100 * STAGE + 10 * (dependent variables per pass) + CTAs per SM
    FixedFn fn;
    int recDoubles;                       // shared memory per thread (STAGE / PAIR), doubles
    int optIn;
    int pair;                             // two points per thread: needs even-padded cell segments
};
#define BSPY_IMAGE(NI, A, B, C, D_, ND, J, NDT, MB, ST, OPT)                                                            \
    {NI, {A, B, C, D_}, ND, J, 100 * ST + 10 * NDT + MB, eval_image_kernel<NI, A, B, C, D_, ND, J != 0, NDT, MB, ST != 0>, \
     ST ? (J ? ND * (1 + NI) : ND) : 0, OPT, 0}
#define BSPY_IMAGE2(NI, A, B, C, D_, ND, NDT, MB, OPT)                                                                  \
    {NI, {A, B, C, D_}, ND, 1, 1000 + 10 * NDT + MB, eval_image2_kernel<NI, A, B, C, D_, ND, NDT, MB>, 2 * ND * (1 + NI), OPT, 1}
static const ImageEntry kImage[] = {
    // two points per thread (value + jacobian requests)
    // measured on config 5 (Gpts/s, whole step): 1022 -> 3.60, 1012 -> 3.50, 1013 -> 3.35; one point per thread: 114 -> 2.71,
    // the L1-gather kernel 2.61
    BSPY_IMAGE2(4, 3, 3, 3, 3, 6, 2, 2, 0), BSPY_IMAGE2(4, 3, 3, 3, 3, 6, 1, 2, 1), BSPY_IMAGE2(4, 3, 3, 3, 3, 6, 1, 3, 1),
    BSPY_IMAGE2(3, 4, 4, 4, 0, 3, 1, 3, 1), BSPY_IMAGE2(3, 4, 4, 4, 0, 3, 1, 4, 1), BSPY_IMAGE2(3, 4, 4, 4, 0, 3, 3, 2, 1),
    // the 4-variate nDep-6 manifold of config 5 (value + jacobian); variants for measurement are opt-in (IMAGE=<code>)
    BSPY_IMAGE(4, 3, 3, 3, 3, 6, 1, 1, 4, 1, 0),
    BSPY_IMAGE(4, 3, 3, 3, 3, 6, 1, 1, 4, 0, 1), BSPY_IMAGE(4, 3, 3, 3, 3, 6, 1, 2, 3, 1, 1), BSPY_IMAGE(4, 3, 3, 3, 3, 6, 1, 3, 2, 1, 1),
    BSPY_IMAGE(4, 3, 3, 3, 3, 6, 1, 1, 3, 1, 1), BSPY_IMAGE(4, 3, 3, 3, 3, 6, 1, 6, 2, 0, 1), BSPY_IMAGE(4, 3, 3, 3, 3, 6, 1, 1, 5, 1, 1),
    BSPY_IMAGE(4, 3, 3, 3, 3, 6, 0, 6, 4, 0, 0),
    // tricubic volume (config 4): measured against the warp-staged kernel (10.6 Gpts/s): 34 -> 7.1, pairs 1013 -> 8.1; opt-in
    BSPY_IMAGE(3, 4, 4, 4, 0, 3, 1, 3, 4, 0, 1), BSPY_IMAGE(3, 4, 4, 4, 0, 3, 1, 3, 3, 0, 1),
};

static const ImageEntry *find_image(const SplineDev &s, int jac, int code)
{
    for (const ImageEntry &e : kImage) {
        if (e.nInd != s.nInd || e.nDep != s.nDep || e.jac != jac) continue;
        bool same = true;
        for (int i = 0; i < s.nInd; ++i) same &= e.o[i] == s.order[i];
        if (same && (code == 0 ? !e.optIn : code == e.code)) return &e;
    }
    return nullptr;
}

// ---- cell polynomials, staged images, two points of the same cell per lane --------------------------------------------------
// The global-image pair kernel above is latency-bound (ncu: 47 % of the warp time waits for image rows on their way from L2 /
// HBM, FP64 pipe 29 %, 77 % of its instructions are FMAs): every CTA starts cold, and the 61 KB result tile of a CTA leaves L1
// too small to keep prefetched images.  Here warps are persistent over contiguous runs of 64-slot tiles of the (even-padded)
// sorted sequence, lane l takes slots 2l and 2l + 1; the cell's compact image is staged in shared memory once per cell (two
// slots per warp) and the image of the NEXT cell of the sorted sequence is requested a tile ahead, so that the point loop
// never waits for memory; results of the passes over the dependent variables are collected in the warp's [2][R][32] tile and
// leave as whole 32-byte sectors.
template <int L, class Ord, int NDEP, int NDT>
struct HornerS2 {
    using WS = WindowShape<Ord, NDEP>;
    static constexpr int n = Ord::n;
    __device__ __forceinline__ static void run(const double *__restrict__ w, const int q, const double (&t0)[n], const double (&t1)[n],
                                               double (&v0)[NDT], double (&g0)[n][NDT], double (&v1)[NDT], double (&g1)[n][NDT])
    {
        constexpr int O = Ord::at(L);
        if constexpr (L == n - 1) {
#pragma unroll
            for (int d = 0; d < NDT; ++d) {
                double x[O];
                load_run<O>(w, q + d * WS::perDepPad, x);
                double a = x[O - 1], b = x[O - 1], da = 0.0, db = 0.0;
#pragma unroll
                for (int k = O - 2; k >= 0; --k) {
                    da = (k == O - 2) ? a : fma(da, t0[L], a);
                    db = (k == O - 2) ? b : fma(db, t1[L], b);
                    a = fma(a, t0[L], x[k]);
                    b = fma(b, t1[L], x[k]);
                }
                v0[d] = a; v1[d] = b;
                g0[L][d] = da; g1[L][d] = db;
            }
        } else {
#pragma unroll
            for (int i = O - 1; i >= 0; --i) {
                double cv0[NDT], cv1[NDT];
                double cg0[n][NDT], cg1[n][NDT];
                HornerS2<L + 1, Ord, NDEP, NDT>::run(w, q + i * WS::stride(L), t0, t1, cv0, cg0, cv1, cg1);
#pragma unroll
                for (int d = 0; d < NDT; ++d) {
                    if (i == O - 1) {
                        v0[d] = cv0[d]; v1[d] = cv1[d];
                        g0[L][d] = 0.0; g1[L][d] = 0.0;
#pragma unroll
                        for (int m = L + 1; m < n; ++m) { g0[m][d] = cg0[m][d]; g1[m][d] = cg1[m][d]; }
                    } else {
                        g0[L][d] = (i == O - 2) ? v0[d] : fma(g0[L][d], t0[L], v0[d]);
                        g1[L][d] = (i == O - 2) ? v1[d] : fma(g1[L][d], t1[L], v1[d]);
                        v0[d] = fma(v0[d], t0[L], cv0[d]);
                        v1[d] = fma(v1[d], t1[L], cv1[d]);
#pragma unroll
                        for (int m = L + 1; m < n; ++m) {
                            g0[m][d] = fma(g0[m][d], t0[L], cg0[m][d]);
                            g1[m][d] = fma(g1[m][d], t1[L], cg1[m][d]);
                        }
                    }
                }
            }
        }
    }
};

template <int NIND, int O0, int O1, int O2, int O3, int NDEP, int NDT, int MINB>
__global__ void __launch_bounds__(128, MINB) eval_poly_s2_kernel(const SplineDev s, const PointsDev in, const long long N,
                                                                  const WrtDev wrt, const OutDev out)
{
    using Ord = Orders<NIND, O0, O1, O2, O3>;
    using WS = WindowShape<Ord, NDEP>;
    constexpr int SLOT = WS::size + 4;
    constexpr int R = NDEP * (1 + NIND), RP = (R + 3) & ~3;
    static_assert(NDEP % NDT == 0, "dependent-variable tile must divide nDep");
    if (gate_closed(in)) return;
    extern __shared__ __align__(16) double polyPair[];              // per warp: two slots | result tile [2][R][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *w0 = polyPair + warp * (2 * SLOT + 2 * R * 32);
    double *mine = w0 + 2 * SLOT + lane;                            // results of this lane: mine[(p * R + slot) * 32]
    int slotKey0 = -1, slotKey1 = -1;
    bool prefetched = false;
    int cells = 1;
#pragma unroll
    for (int iv = 0; iv < NIND; ++iv) cells *= s.nCoef[iv] - Ord::at(iv) + 1;
    const long long total = __ldg(in.sortedTotal);                  // slots of the sorted sequence (segments padded to even)
    const long long tiles = (total + 63) >> 6, nWarps = gridDim.x * 4LL;
    const long long per = (tiles + nWarps - 1) / nWarps;
    const long long firstTile = (blockIdx.x * 4LL + warp) * per;
    const long long endTile = firstTile + per < tiles ? firstTile + per : tiles;
    double2 ra0 = make_double2(0.0, 0.0), ra1 = ra0, rb0 = ra0, rb1 = ra0;
    longlong2 kk = make_longlong2(-1, -1);
    auto fetch = [&](long long tile) {
        const long long t = tile * 64 + 2 * lane;
        if (tile < endTile && t < total) {
            const double2 *rp = reinterpret_cast<const double2 *>(in.records + 4 * t);
            ra0 = __ldcs(rp); ra1 = __ldcs(rp + 1); rb0 = __ldcs(rp + 2); rb1 = __ldcs(rp + 3);
            if constexpr (NIND > 3) kk = __ldcs(reinterpret_cast<const longlong2 *>(in.recKI + t));
        }
    };
    auto stage = [&](int key, double *dst) {
        const double *src = in.images + (long long)key * SLOT;
        const unsigned dstAddr = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
        for (int c0 = 0; c0 < SLOT / 2; c0 += 32) {
            const int c = c0 + lane;
            if ((SLOT / 2) % 32 == 0 || c < SLOT / 2)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstAddr + 16u * c), "l"(src + 2 * c) : "memory");
        }
    };
    fetch(firstTile);
    for (long long tile = firstTile; tile < endTile; ++tile) {
        const long long t = tile * 64 + 2 * lane;
        const bool live = t < total;
        double ua[NIND], ub[NIND];
        ua[0] = ra0.x; ub[0] = rb0.x;
        if constexpr (NIND > 1) { ua[1] = ra0.y; ub[1] = rb0.y; }
        if constexpr (NIND > 2) { ua[2] = ra1.x; ub[2] = rb1.x; }
        if constexpr (NIND > 3) { ua[3] = ra1.y; ub[3] = rb1.y; }
        const long long kia = NIND > 3 ? kk.x : __double_as_longlong(ra1.y), kib = NIND > 3 ? kk.y : __double_as_longlong(rb1.y);
        const int key = live ? (int)kia : -1;
        fetch(tile + 1);
        if (prefetched) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            prefetched = false;
        }
        bool done = !live;
        while (true) {
            const unsigned pending = __ballot_sync(0xffffffffu, !done);
            if (!pending) break;
            const int k0 = __shfl_sync(0xffffffffu, key, __ffs(pending) - 1);
            const bool in0 = !done && key == k0;
            const unsigned rest = __ballot_sync(0xffffffffu, !done && !in0);
            const int k1 = rest ? __shfl_sync(0xffffffffu, key, __ffs(rest) - 1) : -1;
            const bool in1 = !done && !in0 && key == k1;
            int s0, s1 = -1;
            bool staged = false;
            if (k0 == slotKey0) s0 = 0;
            else if (k0 == slotKey1) s0 = 1;
            else {
                s0 = (k1 >= 0 && k1 == slotKey0) ? 1 : 0;
                stage(k0, w0 + s0 * SLOT);
                if (s0) slotKey1 = k0; else slotKey0 = k0;
                staged = true;
            }
            if (k1 >= 0) {
                s1 = 1 - s0;
                if ((s1 ? slotKey1 : slotKey0) != k1) {
                    stage(k1, w0 + s1 * SLOT);
                    if (s1) slotKey1 = k1; else slotKey0 = k1;
                    staged = true;
                }
            }
            if (staged) {
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            }
            if (k1 < 0 && k0 + 1 < cells) {
                const int other = 1 - s0;
                if ((other ? slotKey1 : slotKey0) != k0 + 1) {
                    stage(k0 + 1, w0 + other * SLOT);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    if (other) slotKey1 = k0 + 1; else slotKey0 = k0 + 1;
                    prefetched = true;
                }
            }
            if (in0 || in1) {
                const double *w = w0 + (in1 ? s1 : s0) * SLOT;
                double t0[NIND], t1[NIND];
                {
                    const double2 m01 = *reinterpret_cast<const double2 *>(w + WS::size);
                    t0[0] = ua[0] - m01.x; t1[0] = ub[0] - m01.x;
                    if constexpr (NIND > 1) { t0[1] = ua[1] - m01.y; t1[1] = ub[1] - m01.y; }
                    if constexpr (NIND > 2) {
                        const double2 m23 = *reinterpret_cast<const double2 *>(w + WS::size + 2);
                        t0[2] = ua[2] - m23.x; t1[2] = ub[2] - m23.x;
                        if constexpr (NIND > 3) { t0[3] = ua[3] - m23.y; t1[3] = ub[3] - m23.y; }
                    }
                }
                if constexpr (NDT == NDEP) {
                    // one pass: both records leave from registers as whole sectors, no result tile
                    double v0[NDEP], v1[NDEP];
                    double g0[NIND][NDEP], g1[NIND][NDEP];
                    HornerS2<0, Ord, NDEP, NDEP>::run(w, 0, t0, t1, v0, g0, v1, g1);
                    if ((kia >> 32) >= 0)
                        store_result_record<NIND, NDEP, true>(s, out, out.aos + (out.aosScatter ? out.aosBase + (kia >> 32) : t) * out.aosStride, v0, g0);
                    if ((kib >> 32) >= 0)
                        store_result_record<NIND, NDEP, true>(s, out, out.aos + (out.aosScatter ? out.aosBase + (kib >> 32) : t + 1) * out.aosStride, v1, g1);
                    done = true;
                    __syncwarp();
                    continue;
                }
#pragma unroll 1
                for (int d0 = 0; d0 < NDEP; d0 += NDT) {
                    double v0[NDT], v1[NDT];
                    double g0[NIND][NDT], g1[NIND][NDT];
                    HornerS2<0, Ord, NDEP, NDT>::run(w + d0 * WS::perDepPad, 0, t0, t1, v0, g0, v1, g1);
#pragma unroll
                    for (int d = 0; d < NDT; ++d) {
                        mine[(d0 + d) * 32] = v0[d];
                        mine[(R + d0 + d) * 32] = v1[d];
#pragma unroll
                        for (int iv = 0; iv < NIND; ++iv) {
                            mine[(NDEP + (d0 + d) * NIND + iv) * 32] = g0[iv][d];
                            mine[(R + NDEP + (d0 + d) * NIND + iv) * 32] = g1[iv][d];
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const long long ki = p ? kib : kia;
                    const long long idx = ki >> 32;
                    if (idx < 0) continue;                          // dummy slot of an odd cell
                    const long long dest = out.aosScatter ? out.aosBase + idx : t + p;
                    double *rec = out.aos + dest * out.aosStride;
                    const double *src = mine + p * R * 32;
#pragma unroll
                    for (int j = 0; j < RP / 4; ++j) {
                        if (4 * j < out.aosStride) {
                            double x[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) x[e] = 4 * j + e < R ? src[(4 * j + e) * 32] : 0.0;
                            if (out.aosWide)
                                asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(rec + 4 * j), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
                            else {
                                __stcs(reinterpret_cast<double2 *>(rec + 4 * j), make_double2(x[0], x[1]));
                                __stcs(reinterpret_cast<double2 *>(rec + 4 * j) + 1, make_double2(x[2], x[3]));
                            }
                        }
                    }
                }
                done = true;
            }
            __syncwarp();
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// cell-polynomial kernels: code = 10 * (dependent variables per pass) + CTAs per SM
struct PolyEntry {
    int nInd, o[4], nDep, code;
    FixedFn fn;
    int slotDoubles, optIn;
    int pair;                  // 1: two points per thread on padded images read from global memory (code 1000 + ...);
                               // 2: two points per lane on staged compact images, persistent warps (code 2000 + ...)
    int recDoubles;            // pair 1: shared-memory result tile per thread; pair 2: doubles per warp (two slots + result tile)
};
#define BSPY_POLY(NI, A, B, C, D_, ND, NDT, MB, OPT)                                                   \
    {NI, {A, B, C, D_}, ND, 10 * NDT + MB, eval_poly_kernel<NI, A, B, C, D_, ND, NDT, MB>,              \
     WindowShape<Orders<NI, A, B, C, D_>, ND>::size + 4, OPT, 0, 0}
#define BSPY_POLY2(NI, A, B, C, D_, ND, NDT, MB, OPT)                                                  \
    {NI, {A, B, C, D_}, ND, 1000 + 10 * NDT + MB, eval_poly2_kernel<NI, A, B, C, D_, ND, NDT, MB>,      \
     ImageShape<Orders<NI, A, B, C, D_>, ND>::window + 4, OPT, 1, 2 * ND * (1 + NI)}
#define BSPY_POLYS2(NI, A, B, C, D_, ND, NDT, MB, OPT)                                                 \
    {NI, {A, B, C, D_}, ND, 2000 + 10 * NDT + MB, eval_poly_s2_kernel<NI, A, B, C, D_, ND, NDT, MB>,    \
     WindowShape<Orders<NI, A, B, C, D_>, ND>::size + 4, OPT, 2,                                       \
     2 * (WindowShape<Orders<NI, A, B, C, D_>, ND>::size + 4) + 2 * ND * (1 + NI) * 32}
static const PolyEntry kPoly[] = {
    // staged images, two points per lane: measured on config 5 (whole step) 2022 -> 2.92, 2032 -> 2.92, 2062 -> 2.68 Gpts/s against 3.83
    // for the global-image pair kernel -- 23 KB of shared memory per warp (two slots + the result tile) leave 8 warps per SM, too
    // few to cover the dependent FMA chains and shared-memory latencies (ncu: issue 24 %, FP64 34 %, no memory stalls left);
    // opt-in, kept with its parity test
    BSPY_POLYS2(4, 3, 3, 3, 3, 6, 2, 2, 1), BSPY_POLYS2(4, 3, 3, 3, 3, 6, 3, 2, 1),
    // (config 4, records stored from registers without the result tile: 2033 -> 11.56 Gpts/s, 242 us per chunk against 213 for one point
    //  per lane -- 168 registers, 12 warps per SM; with the tile and IMAGE=1013 for the even-padded sort: 2033 -> 11.26, 2032 -> 11.00, 2014 -> 11.23 Gpts/s against 12.1 for one point per lane;
    //  a bucket table instead of the bisection in the keys pass: 87 us either way -- the pass is bound by its atomics, not by the search)
    // the 4-variate nDep-6 manifold (config 5): two points per thread
    BSPY_POLY2(4, 3, 3, 3, 3, 6, 2, 3, 0), BSPY_POLY2(4, 3, 3, 3, 3, 6, 2, 4, 1), BSPY_POLY2(4, 3, 3, 3, 3, 6, 3, 3, 1), BSPY_POLY2(4, 3, 3, 3, 3, 6, 1, 4, 1),
    BSPY_POLY2(4, 3, 3, 3, 3, 6, 2, 2, 1),
    // tricubic nDep-3 volume (config 4), whole step: 35 -> 12.00, 34 -> 11.98, 36 -> 11.73, 16 -> 5.85 Gpts/s; recurrence 10.53
    BSPY_POLY(3, 4, 4, 4, 0, 3, 3, 5, 0), BSPY_POLY(3, 4, 4, 4, 0, 3, 3, 4, 1), BSPY_POLY(3, 4, 4, 4, 0, 3, 3, 6, 1),
    // the 4-variate nDep-6 manifold: one point per lane moves 3.9 KB per point through the load-return path -- 2.75 Gpts/s against
    // 3.61 for two points per thread on the recurrence images (eval_image2_kernel); opt-in until a pair version exists
    BSPY_POLY(4, 3, 3, 3, 3, 6, 2, 5, 1), BSPY_POLY(4, 3, 3, 3, 3, 6, 3, 4, 1), BSPY_POLY(4, 3, 3, 3, 3, 6, 2, 4, 1), BSPY_POLY(4, 3, 3, 3, 3, 6, 1, 6, 1),
    BSPY_POLY(3, 4, 4, 4, 0, 1, 1, 6, 0), BSPY_POLY(3, 4, 4, 4, 0, 4, 2, 5, 0), BSPY_POLY(3, 3, 3, 3, 0, 3, 3, 6, 0), BSPY_POLY(3, 4, 4, 4, 0, 2, 2, 6, 0),
    BSPY_POLY(3, 3, 3, 3, 0, 2, 2, 6, 0),
};

static const PolyEntry *find_poly(const SplineDev &s, int code)
{
    for (const PolyEntry &e : kPoly) {
        if (e.nInd != s.nInd || e.nDep != s.nDep) continue;
        bool same = true;
        for (int i = 0; i < s.nInd; ++i) same &= e.o[i] == s.order[i];
        if (same && (code <= 1 ? !e.optIn : code == e.code)) return &e;
    }
    return nullptr;
}

static bool poly_entry_pair(const PolyEntry *e) { return e && e->pair == 1; }   // padded image layout

struct CellEntry {
    int nInd, o[4], nDep;
    FixedFn fn;
    int warpDoubles, minBlocks;
};
#define BSPY_CELL(NI, A, B, C, D_, ND, MB)                                                     \
    {NI, {A, B, C, D_}, ND, eval_cell_mma_kernel<NI, A, B, C, D_, ND, MB>, CellShape<Orders<NI, A, B, C, D_>, ND>::perWarp, MB}
static const CellEntry kCell[] = {
    BSPY_CELL(3, 4, 4, 4, 0, 3, 7),      // config 4: 14.8 KB per warp -> 7 CTAs of 2 warps
    BSPY_CELL(3, 4, 4, 4, 0, 1, 8),
    BSPY_CELL(3, 3, 3, 3, 0, 3, 7),
    BSPY_CELL(4, 3, 3, 3, 3, 6, 3),      // config 5: 32 KB per warp -> 3 CTAs of 2 warps
};

static const CellEntry *find_cell(const SplineDev &s)
{
    for (const CellEntry &e : kCell) {
        if (e.nInd != s.nInd || e.nDep != s.nDep) continue;
        bool same = true;
        for (int i = 0; i < s.nInd; ++i) same &= e.o[i] == s.order[i];
        if (same) return &e;
    }
    return nullptr;
}

// ---- sorted-record pipeline ------------------------------------------------------------------------------------------
// per workspace half: keys | rank -> inverse permutation | (key, index) pairs (nInd == 4) | histogram | point records |
// result records (struct-of-arrays outputs only: with caller-visible records the evaluation writes them in place)
static long long records_half_bytes(const SplineDev &s, long long chunk, bool ownAos)
{
    const long long cells = binned_cells(s);
    const long long slots = pad64(chunk) + pad64(cells);          // sorted slots: cell segments may be padded to even lengths
    return 4 * (2 * pad64(chunk) + pad64(cells + 1)) + 8 * slots + 8 * (4 * slots + (ownAos ? (long long)aos_stride(s) * slots : 0));
}

static int eval_binned_records(const SplineDev &s, PointsDev in, long long N, const WrtDev &wrt, OutDev out, int jac,
                               void *workspace, cudaStream_t stream)
{
    const bool userAos = out.aos != nullptr;                     // caller-visible records: written in place, no un-permute
    const long long cells = binned_cells(s);
    const long long chunk = N < BIN_REC_CHUNK ? N : BIN_REC_CHUNK;
    const long long cpad = pad64(chunk);
    const long long half = records_half_bytes(s, chunk, !userAos);
    const int D = (s.nInd - s.nDep == 1 || s.nDep - s.nInd == 1) ? (s.nInd > s.nDep ? s.nInd : s.nDep) : 0;
    const int nJ = jac ? s.nDep * s.nInd : 0;
    const int nN = (jac && (out.normal || out.aosNormal)) ? D : 0;
    const int stride = userAos ? out.aosStride : (s.nDep + nJ + nN + 3) & ~3;
    const bool plainWrt = wrt.d[0] == 0 && wrt.d[1] == 0 && wrt.d[2] == 0 && wrt.d[3] == 0;
    FixedFn fn = find_fixed(s, jac);
    {
        const int ndt = (int)option(OPT_DEP_TILE, 14);
        // one dependent variable per pass (smaller accumulator set, more resident warps); no normals there
        FixedFn tiled = (ndt > 0 && !nN) ? find_fixed_tiled(s, jac, ndt) : nullptr;
        if (tiled) fn = tiled;
    }
    // evaluation kernel for dense chunks: warp-staged windows where compiled, else the L1-gather kernel.  CELL_KERNEL=1
    // selects the tensor-pipe cell kernel instead (measured SLOWER: 493 vs 326 us per 4 Mi points on config 4, 3400 vs
    // 1070 us on config 5 -- the T round trip through shared memory saturates the LSU data pipe at 92 % with the FP64
    // pipe 15 % active, profiles/r02_cfg4_eval_cell_mma_ncu_full.txt; kept as a tested experiment, off by default)
    const CellEntry *cell = (jac && plainWrt && option(OPT_CELL_KERNEL, 0) == 1) ? find_cell(s) : nullptr;
    const size_t cellSmem = cell ? sizeof(double) * CELL_WARPS * cell->warpDoubles : 0;
    if (cell)
        if (int rc = allow_dynamic_smem(cell->fn, cellSmem)) return rc;
    // cell polynomials (value + jacobian requests): the default where compiled; CELL_POLY=0 turns them off, CELL_POLY=<10 * deps
    // per pass + CTAs per SM> picks a variant
    const PolyEntry *poly = (jac && plainWrt && !nN && !cell && poly_applies(s, N)) ? find_poly(s, (int)option(OPT_CELL_POLY, 1)) : nullptr;
    const size_t polySmem = !poly ? 0 : poly->pair == 1 ? sizeof(double) * 128 * poly->recDoubles : poly->pair == 2 ? sizeof(double) * 4 * poly->recDoubles : sizeof(double) * 4 * 2 * poly->slotDoubles;
    if (poly)
        if (int rc = allow_dynamic_smem(poly->fn, polySmem)) return rc;
    const StagedEntry *staged = nullptr;
    {
        const int code = (int)option(OPT_STAGED, 0);
        if (code >= 0 && !nN) staged = find_staged(s, jac, code);
    }
    // two points per lane on staged windows (value + jacobian requests).  STAGED_PAIR=<10 * deps per pass + CTAs per SM> selects
    // it; measured on config 4 (whole step, Gpts/s): 13 -> 9.15, 12 -> 9.49, 32 -> 9.61, 14 -> 8.13 against 10.5 for one point
    // per lane (eval_staged2_kernel) -- unlike the 4-variate manifold, where two points per thread gave +38 %, the tricubic
    // kernel loses more to the lower occupancy (168-246 registers) than it gains from halving the window loads.  Off by default.
    const StagedEntry *stagedPair = nullptr;
    if (jac && plainWrt && !nN && option(OPT_SPAN_RECORDS, 1) && option(OPT_STAGED_PAIR, -1) >= 0 && !(cell != nullptr) && !poly)
        stagedPair = find_staged_pair(s, (int)option(OPT_STAGED_PAIR, -1));
    const size_t stagedPairSmem = stagedPair ? sizeof(double) * 4 * stagedPair->windowDoubles : 0;
    if (stagedPair)
        if (int rc = allow_dynamic_smem(stagedPair->fn, stagedPairSmem)) return rc;
    // second-generation staged kernel (span records staged with the window) unless EXP_A=1 or the records are off
    const bool staged2 = staged && option(OPT_SPAN_RECORDS, 1) && !option(OPT_EXP_A, 0);
    const FixedFn stagedFn = staged ? (staged2 ? staged->fn2 : staged->fn) : nullptr;
    const size_t stagedSmem = staged ? sizeof(double) * 4 * 2 * (staged2 ? staged->slotDoubles : staged->windowDoubles) : 0;
    if (staged)
        if (int rc = allow_dynamic_smem(stagedFn, stagedSmem)) return rc;
    // per-span records (left knots | reciprocal gaps) for every variable: no divisions in the evaluation kernel
    const double *spanRec[BSPY_MAX_IND] = {};
    {
        if (option(OPT_SPAN_RECORDS, 1)) {
            double *at = (double *)((char *)workspace + 2 * half);
            for (int i = 0; i < s.nInd; ++i) {
                const int spans = s.nCoef[i] - s.order[i] + 1, st = span_rec_stride(s.order[i]);
                span_records_kernel<<<(spans + 127) / 128, 128, 0, stream>>>(s.knots[i], s.order[i], s.nCoef[i], at, st);
                spanRec[i] = at;
                at += (long long)spans * st;
            }
            count_launch(s.nInd);
        }
    }
    // cell images (padded window + span records per cell, built once per call) for the shapes compiled for them
    const ImageEntry *image = nullptr;
    const double *images = nullptr;
    if (images_apply(s, N) && !cell && !stagedPair && !(poly && poly->pair != 1)) {   // the global-image pair polynomial kernel keeps the recurrence images as its fallback
        image = find_image(s, jac, (int)option(OPT_IMAGE, 0));
        if (image && nN && (image->pair || (image->code % 100) / 10 != s.nDep)) image = nullptr;   // normals need the whole jacobian in one pass
        if (image && image->pair && !plainWrt) image = nullptr;
        if (image && !spanRec[0]) image = nullptr;
    }
    if (image) {
        const ImageLayout L = image_layout(s);
        double *dst = (double *)((char *)workspace + 2 * half + span_records_bytes(s));
        PointsDev rp{};
        for (int i = 0; i < s.nInd; ++i) rp.spanRec[i] = spanRec[i];
        long long blocks = (cells + 7) / 8;
        const long long cap = (long long)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        build_cell_images_kernel<<<(unsigned)blocks, 256, 0, stream>>>(s, L, cells, nullptr, rp, dst);
        count_launch(1);
        images = dst;
        if (int rc = allow_dynamic_smem(image->fn, sizeof(double) * 128 * image->recDoubles)) return rc;
    }
    // cell polynomial images + the flag their validation leaves behind
    const double *polyImages = nullptr;
    const int *polyFlag = nullptr;
    if (poly && poly->pair == 1 && !(image && image->pair)) poly = nullptr;
    if (poly) {
        const PolyLayout L = poly_layout(s, poly->pair == 1);
        if (L.E > POLY_BUILD_MAX_E) { set_error("cell polynomial build: window of %d doubles", L.E); return BSPY_E_UNSUPPORTED; }
        double *at = (double *)((char *)workspace + 2 * half + span_records_bytes(s) + cell_images_bytes(s, N));
        int *flag = (int *)at;
        cudaError_t e = cudaMemsetAsync(flag, 0xff, sizeof(int), stream);              // != 0: valid until a cell says otherwise
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        at += 8;
        PolyMatrices PM{};
        long long used = 8;
        for (int i = 0; i < s.nInd; ++i) {
            const int spans = s.nCoef[i] - s.order[i] + 1, st = poly_matrix_stride(s.order[i]);
            poly_matrix_kernel<<<(spans + 127) / 128, 128, 0, stream>>>(s.knots[i], s.order[i], s.nCoef[i], at, st);
            PM.m[i] = at;
            at += (long long)spans * st;
            used += (long long)spans * st;
        }
        double *dst = (double *)((char *)workspace + 2 * half + span_records_bytes(s) + cell_images_bytes(s, N)) + pad64(used);
        const size_t bsm = sizeof(double) * POLY_BUILD_WARPS * (2 * L.E + 32);
        if (int rc = allow_dynamic_smem(build_cell_poly_kernel, bsm)) return rc;
        long long blocks = (cells + POLY_BUILD_WARPS - 1) / POLY_BUILD_WARPS;
        const long long cap = (long long)num_sms() * 8;
        if (blocks > cap) blocks = cap;
        build_cell_poly_kernel<<<(unsigned)blocks, POLY_BUILD_WARPS * 32, bsm, stream>>>(s, L, cells, PM, dst, flag);
        count_launch(s.nInd + 1);
        polyImages = dst;
        polyFlag = flag;
    }
    // cell segments are padded to even lengths for the chunks a two-points-per-thread kernel evaluates (the staged pair
    // kernel only takes dense chunks: a sparse tail chunk is sorted without padding and goes to the one-point kernels)
    // BIN_PERM=1: (cell key, index) pairs instead of point records for the staged polynomial kernel -- no record scatter (64 -> 34 us
    // per chunk), but the kernel's own gather of the points costs its load path more than that (213 -> 276 us): 11.54 against
    // 12.17 Gpts/s on config 4.  Off by default, kept with its bit-for-bit test.
    const bool usePairs = poly && !poly->pair && option(OPT_BIN_PERM, 0) != 0;
    auto pair_pad = [&](int n) { return (image && image->pair) || (poly && poly->pair == 2) || (stagedPair != nullptr && n >= 48 * cells); };
    // Sort (and un-permute) of the neighbouring chunks on a second stream under the evaluation of this one: the sort
    // passes are memory / latency bound, the evaluation FP64 bound.  BIN_OVERLAP=0/1 overrides.
    const bool wantOverlap = option(OPT_BIN_OVERLAP, (staged != nullptr || cell != nullptr || stagedPair != nullptr || (poly && !poly->pair)) ? 1 : 0) != 0;
    BinStreams *bs = wantOverlap ? acquire_bin_streams() : nullptr;
    struct Release { BinStreams *b; ~Release() { if (b) release_bin_streams(b); } } releaseOnExit{bs};
    const long long nChunks = (N + chunk - 1) / chunk;
    const bool overlap = bs != nullptr && nChunks > 1;
    cudaStream_t sSort = overlap ? bs->sort : stream, sEval = overlap ? bs->eval : stream;
    if (overlap) {
        cudaEventRecord(bs->fork, stream);
        cudaStreamWaitEvent(sSort, bs->fork, 0);
        cudaStreamWaitEvent(sEval, bs->fork, 0);
    }
    struct Buf { int *keys, *inv, *hist; int2 *recKI; double *records, *aos; } buf[2];
    for (int h = 0; h < 2; ++h) {
        char *base = (char *)workspace + h * half;
        buf[h].keys = (int *)base; buf[h].inv = buf[h].keys + cpad;
        buf[h].hist = buf[h].inv + cpad;
        buf[h].recKI = (int2 *)(buf[h].hist + pad64(cells + 1));
        buf[h].records = (double *)(buf[h].recKI + cpad + pad64(cells));
        buf[h].aos = buf[h].records + 4 * (cpad + pad64(cells));
    }
    auto sortChunk = [&](long long c) -> int {
        const Buf &B = buf[c & 1];
        const long long base = c * chunk;
        const int n = (int)(N - base < chunk ? N - base : chunk);
        // the evaluation of chunk c-2 (same half) must be done with the records before they are overwritten
        if (overlap && c >= 2) cudaStreamWaitEvent(sSort, bs->evaluated[c & 1], 0);
        cudaError_t e = cudaMemsetAsync(B.hist, 0, sizeof(int) * (cells + 1), sSort);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        OutDev o1{};
        o1.ld = out.ld; o1.spans = out.spans; o1.firstOutside = out.firstOutside;
        // four points per thread in flight (measured against one point per thread at full occupancy: keys 106 -> 95 us,
        // scatter 88 -> 71 us per 4 Mi points)
        bin_keys_batched_kernel<4><<<(n + 511) / 512, 128, 0, sSort>>>(s, in, base, n, B.keys, B.hist, B.inv, o1);
        if (pair_pad(n)) {
            bin_scan_kernel<<<1, 1024, 0, sSort>>>(B.hist, (int)cells, 1, B.records, B.recKI, s.nInd);
            bin_pad_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, sSort>>>(B.hist, (int)cells, B.records, B.recKI, s.nInd);
            count_launch(1);
        }
        else bin_scan_kernel<<<1, 1024, 0, sSort>>>(B.hist, (int)cells);
        if (usePairs)
            bin_scatter_pairs_batched_kernel<4><<<(n + 511) / 512, 128, 0, sSort>>>(n, B.keys, B.hist, B.recKI, B.inv, userAos ? 0 : 1);
        else
            bin_scatter_records_batched_kernel<4><<<(n + 511) / 512, 128, 0, sSort>>>(s, in, base, n, B.keys, B.hist, B.records, B.recKI,
                                                                                     B.inv, userAos ? 0 : 1);
        if (overlap) cudaEventRecord(bs->sorted[c & 1], sSort);
        count_launch(3);
        return check_launch("bspy_cuda_eval_points_binned(sort)");
    };
    auto evalChunk = [&](long long c) -> int {
        const Buf &B = buf[c & 1];
        const long long base = c * chunk;
        const int n = (int)(N - base < chunk ? N - base : chunk);
        if (overlap) cudaStreamWaitEvent(sEval, bs->sorted[c & 1], 0);
        PointsDev pin{};
        pin.records = usePairs ? nullptr : B.records; pin.recKI = B.recKI;
        if (usePairs) { pin.uvw = in.uvw; pin.pointStride = in.pointStride; pin.varStride = in.varStride; pin.base = base; }
        for (int i = 0; i < s.nInd; ++i) pin.spanRec[i] = spanRec[i];
        OutDev o2 = out;
        o2.spans = nullptr; o2.firstOutside = nullptr;
        o2.aosStride = stride;
        if (userAos) { o2.aosScatter = 1; o2.aosBase = base; }
        else { o2.aos = B.aos; o2.aosScatter = 0; o2.aosBase = 0; }
        o2.aosWide = (stride % 4 == 0 && (reinterpret_cast<uintptr_t>(o2.aos) & 31) == 0) ? 1 : 0;
        // the window-sharing kernels live on reuse: they need cells that hold a few tiles' worth of points (a sparse
        // tail chunk makes every tile straddle several cells); below that the L1-gather kernel is the faster one
        pin.images = images;
        pin.sortedTotal = B.hist + cells;
        pin.prefetchImages = (int)option(OPT_EXP_B, 1);
        if (poly && poly->pair) {
            PointsDev pp = pin;
            pp.images = polyImages; pp.gate = polyFlag; pp.gateWant = -1;
            const long long pairs = (n + cells + 1) / 2;          // upper bound; the kernel reads the exact slot count
            long long blocks = (pairs + 127) / 128;
            if (poly->pair == 2 && blocks > (long long)num_sms() * (poly->code % 10)) blocks = (long long)num_sms() * (poly->code % 10);
            poly->fn<<<(unsigned)blocks, 128, polySmem, sEval>>>(s, pp, n, wrt, o2);
            count_launch(1);
            pin.gate = polyFlag; pin.gateWant = 0;
        } else if (poly && n >= 48 * cells) {
            // persistent warps over contiguous runs of tiles; the recurrence kernel of the same chunk follows and runs only
            // if the validation of the images cleared the flag
            PointsDev pp = pin;
            pp.images = polyImages; pp.gate = polyFlag; pp.gateWant = -1;
            long long blocks = (long long)num_sms() * (poly->code % 10);
            if (blocks > (n + 127) / 128) blocks = (n + 127) / 128;
            poly->fn<<<(unsigned)blocks, 128, polySmem, sEval>>>(s, pp, n, wrt, o2);
            count_launch(1);
            pin.gate = polyFlag; pin.gateWant = 0;
        }
        if (image && image->pair) {
            const long long pairs = (n + cells + 1) / 2;          // upper bound; the kernel reads the exact slot count
            long long blocks = (pairs + 127) / 128;
            if (pin.gate && blocks > (long long)num_sms() * 2) blocks = (long long)num_sms() * 2;   // behind a gate: normally skipped
            image->fn<<<(unsigned)blocks, 128, sizeof(double) * 128 * image->recDoubles, sEval>>>(s, pin, n, wrt, o2);
        } else if (image) {
            image->fn<<<(unsigned)((n + 127) / 128), 128, sizeof(double) * 128 * image->recDoubles, sEval>>>(s, pin, n, wrt, o2);
        } else if (stagedPair && n >= 48 * cells) {
            long long blocks = (long long)num_sms() * (stagedPair->code % 10);
            if (blocks > (n + 255) / 256) blocks = (n + 255) / 256;
            stagedPair->fn<<<(unsigned)blocks, 128, stagedPairSmem, sEval>>>(s, pin, n, wrt, o2);
        } else if (cell && n >= 48 * cells) {
            long long blocks = (long long)num_sms() * cell->minBlocks;
            const long long most = (n + 32 * CELL_WARPS - 1) / (32 * CELL_WARPS);
            if (blocks > most) blocks = most;
            cell->fn<<<(unsigned)blocks, CELL_WARPS * 32, cellSmem, sEval>>>(s, pin, n, wrt, o2);
        } else if (staged && n >= 48 * cells && !(poly && poly->pair == 2) && !usePairs) {
            // persistent warps over contiguous runs of tiles (window reuse between consecutive tiles)
            // STAGED_WAVES > 1: that many times more, shorter CTAs (runs of tiles stay long enough for the window reuse), so
            // that CTAs retire all along the kernel and the high-priority sort stream finds room before the tail
            long long blocks = (long long)num_sms() * (staged->code % 10) * option(OPT_STAGED_WAVES, 1);
            if (blocks > (n + 127) / 128) blocks = (n + 127) / 128;
            stagedFn<<<(unsigned)blocks, 128, stagedSmem, sEval>>>(s, pin, n, wrt, o2);
        }
        else {
            // even-padded records (staged pair polynomial kernels): up to one spare slot per cell, the kernel reads the exact count
            const long long upTo = (poly && poly->pair == 2) ? (long long)n + cells : n;
            long long blocks = (upTo + 127) / 128;
            if (pin.gate && blocks > (long long)num_sms() * 4) blocks = (long long)num_sms() * 4;   // behind a gate: normally skipped (grid-stride kernel)
            fn<<<(unsigned)blocks, 128, 0, sEval>>>(s, pin, upTo, wrt, o2);
        }
        if (overlap) cudaEventRecord(bs->evaluated[c & 1], sEval);
        count_launch(1);
        return check_launch("bspy_cuda_eval_points_binned(eval)");
    };
    auto unpermChunk = [&](long long c) -> int {
        if (userAos) return 0;
        const Buf &B = buf[c & 1];
        const long long base = c * chunk;
        const int n = (int)(N - base < chunk ? N - base : chunk);
        if (overlap) cudaStreamWaitEvent(sSort, bs->evaluated[c & 1], 0);
        int urc;
        switch (stride) {
            case 4: urc = launch_unpermute<4>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 8: urc = launch_unpermute<8>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 12: urc = launch_unpermute<12>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 16: urc = launch_unpermute<16>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            case 32: urc = launch_unpermute<32>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
            default: urc = launch_unpermute<0>(B.aos, stride, B.inv, base, n, s.nDep, nJ, nN, out, sSort); break;
        }
        if (urc) return urc;
        count_launch(1);
        return check_launch("bspy_cuda_eval_points_binned(unpermute)");
    };
    // software pipeline: sort(c+1) is enqueued before unpermute(c) so that it runs under eval(c)
    int rc = sortChunk(0);
    for (long long c = 0; c < nChunks && !rc; ++c) {
        rc = evalChunk(c);
        if (!rc && c + 1 < nChunks) rc = sortChunk(c + 1);
        if (!rc) rc = unpermChunk(c);
    }
    if (overlap) {
        cudaEventRecord(bs->joinSort, sSort);
        cudaEventRecord(bs->joinEval, sEval);
        cudaStreamWaitEvent(stream, bs->joinSort, 0);
        cudaStreamWaitEvent(stream, bs->joinEval, 0);
    }
    return rc;
}

int eval_binned(const SplineDev &s, PointsDev in, long long N, const WrtDev &wrt, OutDev out, int jac, void *workspace,
                cudaStream_t stream)
{
    if (out.aos || (bin_mode(N) == 1 && s.nInd <= 4)) return eval_binned_records(s, in, N, wrt, out, jac, workspace, stream);
    const long long cells = binned_cells(s);
    const int D = s.nInd > s.nDep ? s.nInd : s.nDep;
    const long long outBytes = 8LL * ((out.values ? s.nDep : 0) + (out.jacobian ? s.nDep * s.nInd : 0) + (out.normal ? D : 0));
    const long long cap = N < BIN_CHUNK_MAX ? N : BIN_CHUNK_MAX;      // what the workspace was sized for
    const long long chunk = bin_chunk(outBytes) < cap ? bin_chunk(outBytes) : cap;
    const long long cpad = pad64(cap);
    int *keys = (int *)workspace, *perm = keys + cpad, *skey = perm + cpad, *hist = skey + cpad;
    FixedFn fn = find_fixed(s, jac);
    int32_t *spans = out.spans;
    long long *flag = out.firstOutside;
    for (long long base = 0; base < N; base += chunk) {
        const int n = (int)(N - base < chunk ? N - base : chunk);
        cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(int) * (cells + 1), stream);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
        OutDev o1{};
        o1.ld = out.ld; o1.spans = spans; o1.firstOutside = flag;
        bin_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(s, in, base, n, keys, hist, nullptr, o1);
        bin_scan_kernel<<<1, 1024, 0, stream>>>(hist, (int)cells);
        bin_scatter_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, hist, n, perm, skey);
        PointsDev pin = in;
        pin.perm = perm; pin.cellKey = skey; pin.base = base;
        OutDev o2 = out;
        o2.spans = nullptr; o2.firstOutside = nullptr;
        long long blocks = (n + 127) / 128;
        fn<<<(unsigned)blocks, 128, 0, stream>>>(s, pin, n, wrt, o2);
        count_launch(4);
        int rc = check_launch("bspy_cuda_eval_points_binned");
        if (rc) return rc;
    }
    return 0;
}

}  // namespace bspy

using namespace bspy;

extern "C" int64_t bspy_cuda_binned_workspace_bytes(const bspy_spline *spline, int64_t N)
{
    SplineDev s;
    if (make_spline_dev(spline, s, "bspy_cuda_binned_workspace_bytes")) return 0;
    return binned_workspace(s, N);
}

extern "C" int bspy_cuda_eval_points_binned(const bspy_spline *spline, const double *uvw, int64_t pointStride, int64_t varStride,
                                            int64_t N, const int32_t *wrt, uint32_t flags, uint32_t normalMask, double *values,
                                            double *deriv, double *jacobian, double *normal, int32_t *spans,
                                            int64_t *firstOutside, void *workspace, int64_t workspaceBytes, void *stream)
{
    const char *who = "bspy_cuda_eval_points_binned";
    SplineDev s;
    int rc = make_spline_dev(spline, s, who);
    if (rc) return rc;
    const long long need = binned_workspace(s, N);
    const bool onePass = !(deriv && (values || jacobian || normal));
    if (!need || !workspace || workspaceBytes < need || !onePass || !uvw)
        return bspy_cuda_eval_points(spline, uvw, pointStride, varStride, N, wrt, flags, normalMask, values, deriv, jacobian,
                                     normal, spans, firstOutside, stream);
    if ((deriv != nullptr) != (wrt != nullptr)) { set_error("%s: wrt and deriv must be given together", who); return BSPY_E_ARG; }
    const int D = s.nInd > s.nDep ? s.nInd : s.nDep;
    if (normal && (s.nInd - s.nDep != 1 && s.nDep - s.nInd != 1)) {
        set_error("The number of independent variables must be one different than the number of dependent variables.");
        return BSPY_E_NORMAL_DIMS;
    }
    if (normalMask == 0 || D >= 32) normalMask = 0xffffffffu;
    PointsDev in{};
    in.uvw = uvw; in.pointStride = pointStride; in.varStride = varStride;
    OutDev out{};
    out.ld = N;
    out.firstOutside = (long long *)firstOutside;
    out.spans = spans;
    out.normalize = (flags & BSPY_NORMALIZE) ? 1u : 0u;
    out.normalMask = normalMask;
    WrtDev w{};
    int jac = 0;
    if (jacobian || normal) {
        jac = 1;
        out.values = values; out.jacobian = jacobian; out.normal = normal;
    } else if (deriv) {
        for (int i = 0; i < s.nInd; ++i) {
            if (wrt[i] < 0) { set_error("%s: negative derivative order", who); return BSPY_E_ARG; }
            w.d[i] = wrt[i];
        }
        out.values = deriv;
    } else {
        out.values = values;
    }
    return eval_binned(s, in, N, w, out, jac, workspace, (cudaStream_t)stream);
}

extern "C" int64_t bspy_cuda_aos_workspace_bytes(const bspy_spline *spline, int64_t N)
{
    SplineDev s;
    if (make_spline_dev(spline, s, "bspy_cuda_aos_workspace_bytes")) return 0;
    return binned_workspace(s, N, true);
}

extern "C" int bspy_cuda_eval_points_aos(const bspy_spline *spline, const double *uvw, int64_t pointStride, int64_t varStride,
                                         int64_t N, uint32_t flags, uint32_t normalMask, double *records, int64_t recordStride,
                                         int32_t *spans, int64_t *firstOutside, void *workspace, int64_t workspaceBytes,
                                         void *stream)
{
    const char *who = "bspy_cuda_eval_points_aos";
    SplineDev s;
    int rc = make_spline_dev(spline, s, who);
    if (rc) return rc;
    if (N < 0 || (!uvw && N > 0 && s.nInd > 0) || (!records && N > 0)) { set_error("%s: bad argument", who); return BSPY_E_ARG; }
    const bool wantNormal = (flags & BSPY_WANT_NORMAL) != 0;
    const int jac = (flags & (BSPY_WANT_JACOBIAN | BSPY_WANT_NORMAL)) ? 1 : 0;
    const int D = s.nInd > s.nDep ? s.nInd : s.nDep;
    if (wantNormal && (s.nInd - s.nDep != 1 && s.nDep - s.nInd != 1)) {
        set_error("The number of independent variables must be one different than the number of dependent variables.");
        return BSPY_E_NORMAL_DIMS;
    }
    const long long length = s.nDep + (jac ? (long long)s.nDep * s.nInd : 0) + (wantNormal ? D : 0);
    if (recordStride < length || recordStride > 0x7fffffff || (recordStride & 1) || (reinterpret_cast<uintptr_t>(records) & 15)) {
        set_error("%s: recordStride must be even and >= %lld doubles, records 16-byte aligned", who, length);
        return BSPY_E_ARG;
    }
    if (N == 0) return 0;
    if (normalMask == 0 || D >= 32) normalMask = 0xffffffffu;
    PointsDev in{};
    in.uvw = uvw; in.pointStride = pointStride; in.varStride = varStride;
    OutDev out{};
    out.ld = N;
    out.firstOutside = (long long *)firstOutside;
    out.spans = spans;
    out.normalize = (flags & BSPY_NORMALIZE) ? 1u : 0u;
    out.normalMask = normalMask;
    out.aos = records;
    out.aosStride = (int)recordStride;
    out.aosNormal = wantNormal ? 1 : 0;
    WrtDev w{};
    out.aosWide = (recordStride % 4 == 0 && (reinterpret_cast<uintptr_t>(records) & 31) == 0) ? 1 : 0;
    const long long need = binned_workspace(s, N, true);
    if (need && workspace && workspaceBytes >= need) return eval_binned(s, in, N, w, out, jac, workspace, (cudaStream_t)stream);
    return launch_eval(s, in, N, w, out, jac, (cudaStream_t)stream);
}
