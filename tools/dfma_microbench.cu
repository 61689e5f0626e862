// FP64 latency / throughput probe for B200 (informs the K5 scan design).  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, double a, double b, int iters) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void kf(float* out, float a, float b, int iters) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8 * 2);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    printf("clock %d kHz\n", clk);
    // latency: 1 warp, ILP 1
    float ms = timeit([&] { k<1><<<1, 32>>>(d, 1.0000001, 1e-9, iters); });
    printf("DFMA dependent latency ~ %.1f cycles\n", ms * 1e-3 * clk * 1e3 / iters);
    ms = timeit([&] { kf<1><<<1, 32>>>((float*)d, 1.0000001f, 1e-9f, iters); });
    printf("FFMA dependent latency ~ %.1f cycles\n", ms * 1e-3 * clk * 1e3 / iters);
    for (int warps : {4, 8, 16, 32}) {
        ms = timeit([&] { k<8><<<148, warps * 32>>>(d, 1.0000001, 1e-9, iters); });
        double flops = 2.0 * 148 * warps * 32 * 8.0 * iters;
        printf("DFMA ILP8 warps/SM %2d : %.2f TFLOP/s  (%.1f DFMA lanes/clk/SM)\n", warps, flops / ms / 1e9,
               flops / 2 / (ms * 1e-3) / 148 / (clk * 1e3));
    }
    for (int warps : {4, 16, 32}) {
        ms = timeit([&] { k<2><<<148, warps * 32>>>(d, 1.0000001, 1e-9, iters); });
        double flops = 2.0 * 148 * warps * 32 * 2.0 * iters;
        printf("DFMA ILP2 warps/SM %2d : %.2f TFLOP/s\n", warps, flops / ms / 1e9);
    }
    ms = timeit([&] { kf<8><<<148, 1024>>>((float*)d, 1.0000001f, 1e-9f, iters); });
    printf("FFMA ILP8 32 warps/SM: %.2f TFLOP/s\n", 2.0 * 148 * 1024 * 8.0 * iters / ms / 1e9);
    return 0;
}
