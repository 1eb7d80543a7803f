// Micro-benchmark: issue rate of packed fp32x2 arithmetic (FADD2/FFMA2) against scalar FADD/FFMA on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o f32x2_rate f32x2_rate.cu ; run without arguments.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float x, float y) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 up(u64 r) { float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r)); return v; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int PACKED>
__global__ void k(float *out, int iters, long long *cycles) {
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 m = make_float2(1.0001f, 0.9999f), a = make_float2(0.001f, 0.002f);
    const long long t0 = clock64();
    if (PACKED) {
        u64 r[8];
        for (int i = 0; i < 8; ++i) r[i] = pk(acc[i].x, acc[i].y);
        const u64 mm = pk(m.x, m.y), aa = pk(a.x, a.y);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = fma2(r[i], mm, aa);
        for (int i = 0; i < 8; ++i) acc[i] = up(r[i]);
    } else {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc[i].x = fma1(acc[i].x, m.x, a.x); acc[i].y = fma1(acc[i].y, m.y, a.y); }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float)); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps = 1; warps <= 32; warps *= 2) {
        for (int packed = 0; packed < 2; ++packed) {
            if (packed) k<1><<<148, warps * 32>>>(out, iters, cyc); else k<0><<<148, warps * 32>>>(out, iters, cyc);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double flops_per_cycle_sm = 2.0 * 16 * iters * warps * 32 / (double)h;   // 16 fp32 FMAs per thread-iteration
            printf("warps/SM=%2d %s: %lld cycles, %.1f fp32 FMA-flops/cycle/SM\n", warps, packed ? "FFMA2 " : "FFMA  ", h, flops_per_cycle_sm);
        }
    }
    return 0;
}
