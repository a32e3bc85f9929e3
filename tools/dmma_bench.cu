// dmma_bench.cu -- is the FP64 tensor path (DMMA, mma.sync m8n8k4 f64) worth using on B200?
// Measures sustained DMMA throughput against the plain DFMA loop.   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void __launch_bounds__(256) k_dmma(double *sink, int iters, double a, double b)
{
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}
template <int CH>
__global__ void __launch_bounds__(256) k_dfma(double *sink, int iters, double a, double b)
{
    double c[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i];
    if (s == 123.456) sink[0] = s;
}
// DFMA whose three source operands are all distinct vector registers (what real kernels issue)
template <int CH>
__global__ void __launch_bounds__(256) k_dfma3(double *sink, const double *src, int iters)
{
    double c[CH], x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i] = threadIdx.x + i; x[i] = src[(threadIdx.x + i) & 63]; y[i] = src[(threadIdx.x * 3 + i) & 63]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) c[i] = fma(x[i], y[(i + 1) % CH], c[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i];
    if (s == 123.456) sink[0] = s;
}
// two register operands + accumulator, second multiplicand shared by all chains
template <int CH>
__global__ void __launch_bounds__(256) k_dfma2(double *sink, const double *src, int iters)
{
    double c[CH], x[CH];
    const double y = src[threadIdx.x & 63];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i] = threadIdx.x + i; x[i] = src[(threadIdx.x + i) & 63]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) c[i] = fma(x[i], y, c[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i];
    if (s == 123.456) sink[0] = s;
}
// DMMA and DFMA chains interleaved in the same warp: do the tensor path and the FMA pipe overlap?
template <int CH>
__global__ void __launch_bounds__(256) k_mix(double *sink, int iters, double a, double b)
{
    double c[CH][2], f[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) { dmma(c[i][0], c[i][1], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1] + f[i];
    if (s == 123.456) sink[0] = s;
}
// DMMA with distinct operand registers per instruction (what a real fragment loop issues)
template <int CH>
__global__ void __launch_bounds__(256) k_dmma_ops(double *sink, const double *src, int iters)
{
    double c[CH][2], a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x; c[i][1] = i; a[i] = src[(threadIdx.x + i) & 63]; b[i] = src[(threadIdx.x * 5 + i) & 63]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma(c[i][0], c[i][1], a[i], b[(i + 3) % CH]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}
int main()
{
    double *sink; cudaMalloc(&sink, 8);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int ctas = 1; ctas <= 8; ctas *= 2) {
        float ms;
        k_dmma<8><<<sms * ctas, 256>>>(sink, 100, 1.0, 1e-9);
        cudaEventRecord(e0); k_dmma<8><<<sms * ctas, 256>>>(sink, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 256 * 8.0 * iters * (double)sms * ctas * 8;   // 256 FMA per DMMA, 8 warps, 8 chains
        printf("DMMA m8n8k4  %d CTA/SM x8 chains: %.2f TFLOP/s\n", ctas, fl / ms / 1e9);
        k_dfma<8><<<sms * ctas, 256>>>(sink, 100, 1.0, 1e-9);
        cudaEventRecord(e0); k_dfma<8><<<sms * ctas, 256>>>(sink, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 256 * 8.0 * iters * (double)sms * ctas;
        printf("DFMA         %d CTA/SM x8 chains: %.2f TFLOP/s\n", ctas, fl / ms / 1e9);
    }
    {
        double *src; cudaMalloc(&src, 64 * 8);
        double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i;
        cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            float ms;
            k_dfma3<8><<<sms * ctas, 256>>>(sink, src, 100);
            cudaEventRecord(e0); k_dfma3<8><<<sms * ctas, 256>>>(sink, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 256 * 8.0 * iters * (double)sms * ctas;
            printf("DFMA 3 distinct register operands %d CTA/SM: %.2f TFLOP/s\n", ctas, fl / ms / 1e9);
            k_dfma2<8><<<sms * ctas, 256>>>(sink, src, 100);
            cudaEventRecord(e0); k_dfma2<8><<<sms * ctas, 256>>>(sink, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("DFMA 2 distinct + shared operand   %d CTA/SM: %.2f TFLOP/s\n", ctas, fl / ms / 1e9);
        }
    }
    for (int ctas = 1; ctas <= 4; ctas *= 2) {
        float ms;
        k_mix<8><<<sms * ctas, 256>>>(sink, 100, 1.0, 1e-9);
        cudaEventRecord(e0); k_mix<8><<<sms * ctas, 256>>>(sink, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl_t = 2.0 * 256 * 8.0 * iters * (double)sms * ctas * 8, fl_f = 2.0 * 256 * 8.0 * 4 * iters * (double)sms * ctas;
        printf("MIX 1 DMMA : 4 DFMA  %d CTA/SM: tensor %.2f + fma %.2f = %.2f TFLOP/s\n", ctas, fl_t / ms / 1e9, fl_f / ms / 1e9, (fl_t + fl_f) / ms / 1e9);
    }
    {
        double *src; cudaMalloc(&src, 64 * 8);
        double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i;
        cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            float ms;
            k_dmma_ops<8><<<sms * ctas, 256>>>(sink, src, 100);
            cudaEventRecord(e0); k_dmma_ops<8><<<sms * ctas, 256>>>(sink, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double fl = 2.0 * 256 * 8.0 * iters * (double)sms * ctas * 8;
            printf("DMMA distinct operand registers %d CTA/SM: %.2f TFLOP/s\n", ctas, fl / ms / 1e9);
        }
    }
    // dependent-chain latency: one chain per warp, one warp per SM
    {
        float ms;
        cudaEventRecord(e0); k_dmma<1><<<sms, 32>>>(sink, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA dependent latency ~ %.1f ns/op\n", ms * 1e6 / iters);
        cudaEventRecord(e0); k_dfma<1><<<sms, 32>>>(sink, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA dependent latency ~ %.1f ns/op\n", ms * 1e6 / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
