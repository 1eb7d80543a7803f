// Micro-benchmark (not part of the product): how fast can B200 move [ROWS x SEG-byte] tiles whose rows are
// STRIDE bytes apart (the first-pass access pattern of the four-step transform)?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int LANES>   // LANES threads x 8 bytes = segment
__global__ void tile_copy(const float2 *__restrict__ src, float2 *__restrict__ dst, int rows, long long row_stride_elems,
                          int tiles_per_row, int loads_per_thread) {
    const long long tile = blockIdx.x;
    const long long mat = tile / tiles_per_row, tc = tile % tiles_per_row;
    const float2 *s = src + mat * rows * row_stride_elems + tc * LANES;
    float2 *d = dst + mat * rows * row_stride_elems + tc * LANES;
    const int lane = threadIdx.x % LANES, r0 = threadIdx.x / LANES, rstep = blockDim.x / LANES;
    float2 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < loads_per_thread) v[i] = __ldcs(s + (long long)(r0 + i * rstep) * row_stride_elems + lane);
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < loads_per_thread) __stcs(d + (long long)(r0 + i * rstep) * row_stride_elems + lane, v[i]);
}

template <int LANES> float run(const float2 *src, float2 *dst, int mats, int rows, int cols, int threads) {
    const int tiles_per_row = cols / LANES;
    const int lpt = rows * LANES / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) tile_copy<LANES><<<mats * tiles_per_row, threads>>>(src, dst, rows, cols, tiles_per_row, lpt);
    cudaEventRecord(e0);
    for (int w = 0; w < 5; ++w) tile_copy<LANES><<<mats * tiles_per_row, threads>>>(src, dst, rows, cols, tiles_per_row, lpt);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    const int mats = 128, rows = 1024, cols = 1024;          // 128 x (1024 x 1024) complex64 = 1 GiB
    const size_t bytes = (size_t)mats * rows * cols * sizeof(float2);
    float2 *src, *dst;
    cudaMalloc(&src, bytes); cudaMalloc(&dst, bytes);
    cudaMemset(src, 1, bytes);
    for (int threads : {256, 512, 1024}) {
        float a = run<8>(src, dst, mats, rows, cols, threads);
        float b = run<16>(src, dst, mats, rows, cols, threads);
        float c = run<32>(src, dst, mats, rows, cols, threads);
        printf("threads %4d: seg 64B %.3f ms (%.0f GB/s)  seg 128B %.3f ms (%.0f GB/s)  seg 256B %.3f ms (%.0f GB/s)\n", threads,
               a, 2 * bytes / a / 1e6, b, 2 * bytes / b / 1e6, c, 2 * bytes / c / 1e6);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
