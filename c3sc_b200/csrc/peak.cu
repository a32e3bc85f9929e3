// peak.cu -- in-run measurement of the FP64 FMA-pipe peak (the roofline denominator
// MEASURED_PEAKS.json does not carry): 16 independent DFMA chains per thread, enough
// CTAs to fill every SM, timed with CUDA events on the launching stream.
#include <cuda_runtime.h>
#include "../../include/c3sc_b200.h"

namespace {
constexpr int CHAINS = 16;
__global__ void __launch_bounds__(256) k_dfma_peak(double *sink, int iters, double a, double b)
{
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = (double)(threadIdx.x + c);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) acc[c] = fma(acc[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += acc[c];
    if (s == 123.456) sink[0] = s;     // never true; keeps the chains alive
}
}  // namespace

extern "C" int c3sc_measure_fp64_peak(double *tflops, int iters, int repeats)
{
    if (!tflops) return C3SC_EINVAL;
    if (iters <= 0) iters = 4096;
    if (repeats <= 0) repeats = 5;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return C3SC_ENODEV;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return C3SC_ECUDA;
    const int grid = sms * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dfma_peak<<<grid, block>>>(sink, 64, 0.999999, 1e-9);        // warm-up
    double best = 0.0;
    for (int r = 0; r < repeats; r++) {
        cudaEventRecord(e0);
        k_dfma_peak<<<grid, block>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return C3SC_ECUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * CHAINS * (double)iters * (double)grid * block;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return C3SC_OK;
}

// ---- FP64 tensor path: DMMA m8n8k4 loop, 8 independent accumulator pairs per warp ------------------------------
namespace {
__global__ void __launch_bounds__(256) k_dmma_peak(double *sink, int iters, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}
}  // namespace

/* Best-of-`repeats` throughput of a pure DMMA (mma.sync.m8n8k4.f64) loop, TFLOP/s: the roofline denominator of the
 * kernels bound by the FP64 tensor path (one DMMA = 8*8*4 FMAs = 512 flop per warp). */
extern "C" int c3sc_measure_fp64_tensor_peak(double *tflops, int iters, int repeats)
{
    if (!tflops) return C3SC_EINVAL;
    if (iters <= 0) iters = 4096;
    if (repeats <= 0) repeats = 5;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return C3SC_ENODEV;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return C3SC_ECUDA;
    const int grid = sms * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dmma_peak<<<grid, block>>>(sink, 64, 0.999999, 1e-9);
    double best = 0.0;
    for (int r = 0; r < repeats; r++) {
        cudaEventRecord(e0);
        k_dmma_peak<<<grid, block>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(sink); return C3SC_ECUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 512.0 * 8 * (double)iters * (double)grid * (block / 32);
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return C3SC_OK;
}
