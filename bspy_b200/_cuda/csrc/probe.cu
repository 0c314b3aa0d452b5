// Measurement probes for bench.py: achievable FP64 FMA / DMMA rate and HBM copy / write-only
// bandwidth on the device the library runs on.  Not part of the evaluation path; they give the
// roofline denominators that MEASURED_PEAKS.json does not carry (FP64, write-only stream).
#include "common.cuh"

namespace bspy {

__global__ void __launch_bounds__(256) probe_dfma_kernel(int iters, double *sink)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) sink[0] = r;
}

__global__ void __launch_bounds__(256) probe_dmma_kernel(int iters, double *sink)
{
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1])
                         : "d"(a), "d"(b));
    }
    const double r = c[0][0] + c[1][1] + c[2][0] + c[3][1];
    if (r == 123.456) sink[0] = r;
}

// Are DFMA and DMMA one pipe or two?  kind 2: every warp interleaves 4 DMMA with 32 DFMA per iteration (equal pipe time if one pipe); kind 3: even
// warps run the DMMA loop, odd warps the DFMA loop.  If the sum of the two rates exceeds either alone, the pipes overlap.
__global__ void __launch_bounds__(256) probe_mixed_kernel(int iters, double *sink, int split)
{
    double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, cc = 1e-9;
    const bool doMma = !split || ((threadIdx.x >> 5) & 1) == 0, doFma = !split || ((threadIdx.x >> 5) & 1) == 1;
    for (int i = 0; i < iters; ++i) {
        if (doMma) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[k][0]), "+d"(c[k][1])
                             : "d"(a), "d"(b));
        }
        if (doFma) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                a0 = fma(a0, m, cc); a1 = fma(a1, m, cc); a2 = fma(a2, m, cc); a3 = fma(a3, m, cc);
                a4 = fma(a4, m, cc); a5 = fma(a5, m, cc); a6 = fma(a6, m, cc); a7 = fma(a7, m, cc);
            }
        }
    }
    const double r = c[0][0] + c[1][1] + c[2][0] + c[3][1] + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) sink[0] = r;
}

// DFMA as the contraction kernels issue it: a register-tiled outer product acc[i][j] += x[i] * b[j] (three distinct,
// varying register operands per instruction, no constant operand in the reuse cache)
__global__ void __launch_bounds__(256) probe_dfma_tile_kernel(int iters, double *sink)
{
    double x[4], b[4], acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-7;
        b[i] = 1.0 - threadIdx.x * 1e-9 - i * 1e-7;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = i + j;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(x[i], b[j], acc[i][j]);
        // keep x / b varying without extra FP64 work: integer twiddle of the low mantissa bits
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[i] = __longlong_as_double(__double_as_longlong(x[i]) ^ (long long)(it & 1));
            b[i] = __longlong_as_double(__double_as_longlong(b[i]) ^ (long long)(it & 2));
        }
    }
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) r += acc[i][j];
    if (r == 123.456) sink[0] = r;
}

__global__ void __launch_bounds__(256) probe_copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, long long n2)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
        __stcs(dst + i, __ldcs(src + i));
}

__global__ void __launch_bounds__(256) probe_fill_kernel(double2 *__restrict__ dst, long long n2)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
        __stcs(dst + i, make_double2((double)i, 1.0));
}

// Write-pattern probe: `planes` arrays of nU x nV doubles, written tile by tile exactly like the grid
// kernel does (one warp = 8 rows x 16 columns per step, 32-byte stores, a CTA = (tileRows x tileCols)),
// with no arithmetic: the attainable HBM rate of the store pattern itself.
__global__ void __launch_bounds__(256) probe_tiles_kernel(double *dst, int planes, long long nU, long long nV,
                                                          int tileRows, int tileCols, int colChunks, int rowBlocks)
{
    const long long tile = blockIdx.x;
    const int cc = (int)(tile % colChunks);
    const int rb = (int)(tile / colChunks);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane >> 2, r4 = lane & 3;
    const int stripsPerCol = tileRows / 8;            // warps stacked along rows
    const int colGroups = 8 / stripsPerCol;           // the remaining warps split the columns
    const int strip = warp % stripsPerCol, cg = warp / stripsPerCol;
    const long long a = (long long)rb * tileRows + strip * 8 + q;
    const long long plane = nU * nV;
    // tileRows == 8 with a NEGATIVE tileCols request (passed as interleave flag) is handled by the caller:
    // interleave = 1 makes the 8 warps write side by side (16 columns each, 1 KB per row per step).
    const bool interleave = stripsPerCol == 1 && (tileCols & 1);
    const int tc = tileCols & ~1;
    const int first = interleave ? warp * 16 : cg * (tc / colGroups);
    const int last = interleave ? tc : (cg + 1) * (tc / colGroups);
    const int step = interleave ? 128 : 16;
    for (int c0 = first; c0 < last; c0 += step) {
        const long long b = (long long)cc * tc + c0 + 4 * r4;
        if (a < nU && b + 3 < nV) {
            for (int p = 0; p < planes; ++p) {
                double *ptr = dst + p * plane + a * nV + b;
                asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "d"(1.0), "d"((double)c0), "d"(3.0), "d"(4.0) : "memory");
            }
        }
    }
}

}  // namespace bspy

using namespace bspy;

extern "C" int bspy_cuda_probe_tiles(double *dst, int32_t planes, int64_t nU, int64_t nV, int32_t tileRows, int32_t tileCols,
                                     double *bytesOut_host, void *stream)
{
    if (!dst || planes <= 0 || tileRows % 8 || 64 % tileRows || (tileCols & ~1) % (16 * (64 / tileRows)) || nV % 4) {
        set_error("bspy_cuda_probe_tiles: bad argument");
        return BSPY_E_ARG;
    }
    const int tcols = tileCols & ~1;   // bit 0 of tileCols = "interleave the warps along the row"
    const int colChunks = (int)((nV + tcols - 1) / tcols), rowBlocks = (int)((nU + tileRows - 1) / tileRows);
    probe_tiles_kernel<<<(unsigned)(colChunks * rowBlocks), 256, 0, (cudaStream_t)stream>>>(dst, planes, nU, nV, tileRows, tileCols,
                                                                                          colChunks, rowBlocks);
    if (bytesOut_host) *bytesOut_host = 8.0 * planes * (double)nU * (double)nV;
    count_launch();
    return check_launch("bspy_cuda_probe_tiles");
}

extern "C" int bspy_cuda_probe_fp64(int32_t kind, int32_t iters, double *sink, double *flopsOut_host, void *stream)
{
    if (!sink || iters <= 0) { set_error("bspy_cuda_probe_fp64: bad argument"); return BSPY_E_ARG; }
    const int blocks = num_sms() * 8, threads = 256;
    if (kind == 0) {
        probe_dfma_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
        if (flopsOut_host) *flopsOut_host = (double)blocks * threads * (double)iters * 8.0 * 2.0;
    } else if (kind == 1) {
        probe_dmma_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink);
        if (flopsOut_host) *flopsOut_host = (double)blocks * (threads / 32) * (double)iters * 4.0 * 512.0;
    } else if (kind >= 4) {
        // kind 4: outer-product DFMA at 8 CTAs of 256 threads per SM; kind 5: the same at 2 CTAs per SM (4 warps per scheduler)
        const int b2 = kind == 5 ? num_sms() * 2 : blocks;
        probe_dfma_tile_kernel<<<b2, threads, 0, (cudaStream_t)stream>>>(iters, sink);
        if (flopsOut_host) *flopsOut_host = (double)b2 * threads * (double)iters * 16.0 * 2.0;
    } else {
        // kind 2: DMMA and DFMA interleaved in every warp; kind 3: alternate warps
        probe_mixed_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink, kind == 3);
        const double warps = (double)blocks * (threads / 32);
        const double perWarp = 4.0 * 512.0 + 32.0 * 32.0 * 2.0;
        if (flopsOut_host) *flopsOut_host = (kind == 3 ? 0.5 : 1.0) * warps * (double)iters * perWarp;
    }
    count_launch();
    return check_launch("bspy_cuda_probe_fp64");
}

extern "C" int bspy_cuda_probe_hbm(int32_t kind, const double *src, double *dst, int64_t nDoubles, double *bytesOut_host,
                                   void *stream)
{
    if (!dst || nDoubles <= 0 || (kind == 0 && !src)) { set_error("bspy_cuda_probe_hbm: bad argument"); return BSPY_E_ARG; }
    const long long n2 = nDoubles / 2;
    const int blocks = num_sms() * 16, threads = 256;
    if (kind == 0) {
        probe_copy_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>((const double2 *)src, (double2 *)dst, n2);
        if (bytesOut_host) *bytesOut_host = 32.0 * (double)n2;
    } else {
        probe_fill_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>((double2 *)dst, n2);
        if (bytesOut_host) *bytesOut_host = 16.0 * (double)n2;
    }
    count_launch();
    return check_launch("bspy_cuda_probe_hbm");
}
