// Issue-rate probes for the K1 redesign (B200, sm_100a): scalar FFMA/FADD vs packed FFMA2/FADD2,
// 64- vs 128-bit shared-memory exchange.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};"
        " fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

template <int ILP, int MODE>  // 0 FFMA, 1 FFMA2, 2 FADD, 3 FADD2, 4 FMUL (scalar x2), 5 FMUL2, 6 FMUL2 with broadcast operand
__global__ void k(float* out, float a, float b, int iters) {
    float2 x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-9f + i, threadIdx.x * 2e-9f + i);
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 1.0001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, A.x, B.x); x[i].y = fmaf(x[i].y, A.y, B.y); }
            if (MODE == 1) x[i] = fma2(x[i], A, B);
            if (MODE == 2) { x[i].x += B.x; x[i].y += B.y; }
            if (MODE == 3) x[i] = add2(x[i], B);
            if (MODE == 4) { x[i].x *= A.x; x[i].y *= A.y; }
            if (MODE == 5) x[i] = mul2(x[i], A);
            if (MODE == 6) x[i] = mul2(x[i], make_float2(A.x, A.x));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory exchange: each thread stores 16 values and loads 16 values (transposed) per round
template <int W>  // W = 8: float2 (64-bit), 16: float4 (128-bit)
__global__ void kx(float* out, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tid = threadIdx.x, nt = blockDim.x;
    float acc = 0.f;
    if (W == 8) {
        float2* s = reinterpret_cast<float2*>(sm);
        float2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = make_float2(tid + i, tid - i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i * (nt + 1) + tid] = v[i];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = s[(tid & 15) * (nt + 1) + (tid >> 4) + i * (nt >> 4)];
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i].x + v[i].y;
    } else {
        float4* s = reinterpret_cast<float4*>(sm);
        float4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = make_float4(tid + i, tid - i, i, tid);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i * (nt + 1) + tid] = v[i];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = s[(tid & 15) * (nt + 1) + (tid >> 4) + i * (nt >> 4)];
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    out[blockIdx.x * nt + tid] = acc;
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
    float* d; cudaMalloc(&d, 148 * 1024 * 4 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    printf("clock %d kHz\n", clk);
    const char* names[7] = {"FFMA (scalar x2)", "FFMA2", "FADD (scalar x2)", "FADD2", "FMUL (scalar x2)", "FMUL2", "FMUL2 (bcast)"};
    for (int warps : {8, 16, 32}) {
        float ms[7];
        ms[0] = timeit([&] { k<8, 0><<<148, warps * 32>>>(d, 1.0000001f, 1e-9f, iters); });
        ms[1] = timeit([&] { k<8, 1><<<148, warps * 32>>>(d, 1.0000001f, 1e-9f, iters); });
        ms[2] = timeit([&] { k<8, 2><<<148, warps * 32>>>(d, 1.0000001f, 1e-9f, iters); });
        ms[3] = timeit([&] { k<8, 3><<<148, warps * 32>>>(d, 1.0000001f, 1e-9f, iters); });
        ms[4] = timeit([&] { k<8, 4><<<148, warps * 32>>>(d, 1.0000001f, 1e-9f, iters); });
        ms[5] = timeit([&] { k<8, 5><<<148, warps * 32>>>(d, 1.0000001f, 1e-9f, iters); });
        ms[6] = timeit([&] { k<8, 6><<<148, warps * 32>>>(d, 1.0000001f, 1e-9f, iters); });
        for (int m = 0; m < 7; ++m) {
            const double lane_ops = 148.0 * warps * 32 * 8 * 2.0 * iters;  // fp32 lane-operations
            printf("%-18s warps/SM %2d: %.3f ms  %.1f fp32 lane-ops/clk/SM\n", names[m], warps, ms[m],
                   lane_ops / (ms[m] * 1e-3) / 148 / (clk * 1e3));
        }
    }
    for (int threads : {256, 512}) {
        const int it2 = 2000;
        cudaFuncSetAttribute(kx<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(kx<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        float m8 = timeit([&] { kx<8><<<148, threads, 16 * (threads + 1) * 8>>>(d, it2); });
        float m16 = timeit([&] { kx<16><<<148, threads, 16 * (threads + 1) * 16>>>(d, it2); });
        printf("smem exchange %d thr: 64-bit %.1f B/clk/SM, 128-bit %.1f B/clk/SM (store+load bytes)\n", threads,
               2.0 * threads * 16 * 8 * it2 / (m8 * 1e-3) / (clk * 1e3), 2.0 * threads * 16 * 16 * it2 / (m16 * 1e-3) / (clk * 1e3));
    }
    return 0;
}
