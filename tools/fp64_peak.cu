// FP64 FMA throughput micro-benchmark for the roofline denominator (SURVEY 8d:
// MEASURED_PEAKS.json has no FP64 entry).  Each thread runs 8 independent FMA
// chains; prints achieved TFLOP/s (FMA = 2 flop) for a burst and for a ~2 s loop.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void fma_chain(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
    for (int w = 0; w < 3; w++) fma_chain<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 10; r++) {
        float ms;
        cudaEventRecord(t0); fma_chain<<<blocks, threads>>>(out, iters, 0.999999, 1e-9); cudaEventRecord(t1);
        cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1);
        if (ms < best) best = ms;
    }
    float total;
    int reps = 0;
    cudaEventRecord(t0);
    do {
        for (int r = 0; r < 20; r++) fma_chain<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        reps += 20;
        cudaEventRecord(t1); cudaEventSynchronize(t1); cudaEventElapsedTime(&total, t0, t1);
    } while (total < 2000.0f);
    printf("{\"fp64_tflops_burst\": %.3f, \"fp64_tflops_sustained\": %.3f, \"sms\": %d, \"clock_khz\": %d}\n",
           flop / (best * 1e-3) / 1e12, flop * reps / (total * 1e-3) / 1e12, prop.multiProcessorCount, prop.clockRate);
    cudaFree(out);
    return 0;
}
