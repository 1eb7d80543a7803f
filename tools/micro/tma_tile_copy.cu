// How fast can one SM's TMA move 64 KiB tiles as 2-D boxes of a given row width, with 3 buffers in flight and no compute?
// Each block loops over tiles: box load (rows x width bytes, row pitch = the matrix row) -> shared memory -> box store to a
// second matrix.  One thread issues everything (the structure of four_step_tma's loader + storer, merged).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I dsc_b200/csrc -I include -o tools/micro/tma_tile_copy tools/micro/tma_tile_copy.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "fft_tma.cuh"

using namespace dscfft;

constexpr int TILE = 64 * 1024, NBUF = 3;

__global__ void __launch_bounds__(64, 1)
tile_copy(const __grid_constant__ CUtensorMap in, const __grid_constant__ CUtensorMap out, int width_elems, int box_rows,
          int boxes_per_tile, int tiles_x, long long tiles, unsigned *ticket) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *buf = smem + ((1024u - (tma::smem_u32(smem) & 1023u)) & 1023u);
    unsigned long long *full = (unsigned long long *)(buf + NBUF * TILE);
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) tma::mbar_init(&full[b], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const unsigned long long pol = tma::policy_evict_first();
    long long mine[NBUF];
    unsigned t = 0;
    auto issue = [&](unsigned tt, long long tile) {
        const int b = tt % NBUF;
        mine[b] = tile;
        tma::mbar_arrive_expect_tx(&full[b], TILE);
        const int tx = (int)(tile % tiles_x), ty = (int)(tile / tiles_x);
        for (int r = 0; r < boxes_per_tile; ++r)
            tma::load_3d(buf + b * TILE + (size_t)r * box_rows * width_elems * 8, &in, tx * width_elems, (ty * boxes_per_tile + r) * box_rows, 0, &full[b], pol);
    };
    long long next = atomicAdd(ticket, 1u);
    // prologue: fill the buffers
    unsigned issued = 0;
    while (issued < NBUF && next < tiles) { issue(issued++, next); next = atomicAdd(ticket, 1u); }
    for (; t < issued; ++t) {
        const int b = t % NBUF;
        tma::mbar_wait(&full[b], (t / NBUF) & 1);
        const long long tile = mine[b];
        const int tx = (int)(tile % tiles_x), ty = (int)(tile / tiles_x);
        for (int r = 0; r < boxes_per_tile; ++r)
            tma::store_3d(&out, tx * width_elems, (ty * boxes_per_tile + r) * box_rows, 0, buf + b * TILE + (size_t)r * box_rows * width_elems * 8, pol);
        tma::store_commit();
        tma::store_wait_read();
        if (next < tiles) { issue(issued++, next); next = atomicAdd(ticket, 1u); }
    }
    tma::store_wait_all();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const size_t bytes = (size_t)1 << 30;
    void *a, *b; unsigned *ticket;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&ticket, 4);
    cudaMemset(a, 1, bytes);
    cudaFuncSetAttribute(tile_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * TILE + 2048);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // matrix: rows of `pitch` bytes; tile = (TILE / width) rows x width bytes
    struct Case { int width_bytes; int pitch_bytes; };
    const Case cases[] = {{64, 8192}, {128, 8192}, {256, 2048}, {256, 8192}, {512, 8192}, {2048, 8192}, {8192, 8192}};
    for (const Case &c : cases) {
        const int width_elems = c.width_bytes / 8;
        const long long pitch_elems = c.pitch_bytes / 8;
        const long long rows = (long long)(bytes / c.pitch_bytes);
        const int tile_rows = TILE / c.width_bytes;
        const int box_rows = tile_rows < 256 ? tile_rows : 256;
        const int boxes_per_tile = tile_rows / box_rows;
        const int tiles_x = (int)(pitch_elems / width_elems);
        const long long tiles = (long long)tiles_x * (rows / tile_rows);
        CUtensorMap mi, mo;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch_elems, (cuuint64_t)rows, 1};
        const cuuint64_t strides[2] = {(cuuint64_t)c.pitch_bytes, (cuuint64_t)bytes};
        const cuuint32_t box[3] = {(cuuint32_t)(width_elems > 256 ? 256 : width_elems), (cuuint32_t)box_rows, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        if (width_elems > 256) { printf("width %d B: skipped (box inner extent > 256 elements)\n", c.width_bytes); continue; }
        enc(&mi, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        enc(&mo, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, b, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaMemset(ticket, 0, 4);
            cudaEventRecord(e0);
            tile_copy<<<148, 64, NBUF * TILE + 2048>>>(mi, mo, width_elems, box_rows, boxes_per_tile, tiles_x, tiles, ticket);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("box rows of %5d B (pitch %5d B), %4d rows per 64 KiB tile: %7.0f GB/s read+write  (%s)\n", c.width_bytes, c.pitch_bytes,
               tile_rows, 2.0 * bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
