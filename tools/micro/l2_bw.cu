// L2 / DRAM streaming bandwidth seen by plain kernels: read-only, write-only and copy over a working set that either
// fits L2 (32 MiB) or not (1 GiB).  Answers: how much cheaper is an L2 hit than a DRAM access in THROUGHPUT terms?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/l2_bw tools/micro/l2_bw.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_read(const uint4 *__restrict__ p, size_t n, int reps, uint4 *sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    if (acc.x == 0x12345678u) *sink = acc;
}
__global__ void k_write(uint4 *__restrict__ p, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
            p[i] = make_uint4(r, (unsigned)i, 3, 4);
}
__global__ void k_copy(const uint4 *__restrict__ a, uint4 *__restrict__ b, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a + i));
            b[i] = v;
        }
}

int main() {
    const size_t big = (size_t)1 << 30;
    uint4 *a, *b, *sink;
    cudaMalloc(&a, big); cudaMalloc(&b, big); cudaMalloc(&sink, 64);
    cudaMemset(a, 1, big); cudaMemset(b, 2, big);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 512;
    for (size_t bytes : {(size_t)16 << 20, (size_t)32 << 20, (size_t)64 << 20, (size_t)1 << 30}) {
        const size_t n = bytes / 16;
        const int reps = (int)(((size_t)8 << 30) / bytes);
        float ms;
        k_read<<<grid, block>>>(a, n, 2, sink);
        cudaEventRecord(e0); k_read<<<grid, block>>>(a, n, reps, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double rd = (double)bytes * reps / ms / 1e6;
        k_write<<<grid, block>>>(a, n, 2);
        cudaEventRecord(e0); k_write<<<grid, block>>>(a, n, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double wr = (double)bytes * reps / ms / 1e6;
        k_copy<<<grid, block>>>(a, b, n / 2, 2);
        cudaEventRecord(e0); k_copy<<<grid, block>>>(a, b, n / 2, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double cp = (double)bytes * reps / ms / 1e6;     // read + write bytes of a working set of `bytes` in total
        printf("working set %5zu MiB: read %7.0f GB/s   write %7.0f GB/s   copy (read+write) %7.0f GB/s\n", bytes >> 20, rd, wr, cp);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
