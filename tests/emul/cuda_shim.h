// cuda_shim.h -- just enough of the CUDA programming model to compile dsc_b200/csrc/*.cuh
// with g++ and run thread blocks on pthreads.  TEST INFRASTRUCTURE ONLY: it exists so the
// index arithmetic of the kernels (Stockham scatter, padding, pad/crop predicates, real
// un-mixing, four-step geometry) can be checked in the GPU-less build container.  It is
// compiled into tests/emul/libdsc_emul.so by tests/emul/Makefile and loaded by
// tests/test_kernels_emulated.py; the product library never contains it.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <pthread.h>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct double2 { double x, y; };
struct dim3_ { unsigned x = 1, y = 1, z = 1; };

namespace dsc_emul {
struct ThreadCtx {
    dim3_ threadIdx, blockIdx, blockDim, gridDim;
    unsigned char *smem;
    pthread_barrier_t *barrier;
};
inline thread_local ThreadCtx tls;

// Runs `body` once per (block, thread).  A launch owns `block` worker threads; they walk the
// grid block by block, with a barrier between blocks so shared memory can be reused.
template <typename F> void launch(unsigned grid, unsigned block, size_t smem_bytes, F body) {
    std::vector<unsigned char> smem(smem_bytes + 64);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, block);
    auto worker = [&](unsigned tid) {
        tls.barrier = &bar;
        tls.smem = smem.data();
        tls.blockDim.x = block;
        tls.gridDim.x = grid;
        tls.threadIdx.x = tid;
        for (unsigned b = 0; b < grid; ++b) {
            tls.blockIdx.x = b;
            body();
            pthread_barrier_wait(&bar);
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(block);
    for (unsigned t = 0; t < block; ++t) pool.emplace_back(worker, t);
    for (auto &th : pool) th.join();
    pthread_barrier_destroy(&bar);
}
}  // namespace dsc_emul

#define threadIdx (dsc_emul::tls.threadIdx)
#define blockIdx (dsc_emul::tls.blockIdx)
#define blockDim (dsc_emul::tls.blockDim)
#define gridDim (dsc_emul::tls.gridDim)

inline void __syncthreads() { pthread_barrier_wait(dsc_emul::tls.barrier); }
// every call site is in block-uniform control flow, so a block barrier is a valid (stronger) stand-in
inline void __syncwarp() { pthread_barrier_wait(dsc_emul::tls.barrier); }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
#define __shared__ static
template <typename T> inline void __stcs(T *p, const T v) { *p = v; }

inline void sincospi(double x, double *s, double *c) {
    const long double a = 3.14159265358979323846264338327950288L * (long double)x;
    *s = (double)sinl(a);
    *c = (double)cosl(a);
}
