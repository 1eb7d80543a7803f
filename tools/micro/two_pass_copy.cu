// What does the MEMORY SYSTEM alone allow a two-pass (four-step) transform?  The data movement of four_step_tma at 2^16
// points per row with no arithmetic at all: every block alternates between
//   A tiles: box load [256 rows][256 B] (row pitch 2 KiB) of x            -> ONE linear 64 KiB store into a ring of work rows
//            (L2 evict_last), and
//   B tiles: box load [256 rows][256 B] of a work row written LAG rows ago -> discard.global.L2 of those lines
//            -> box store [256 rows][256 B] of the result (evict_first),
// with three 64 KiB buffers per SM and one issuing thread (the loader + storer of four_step_tma merged; no row counters:
// the contents do not matter here).  Per point: 8 B DRAM -> L2 -> SM, 8 B SM -> L2, 8 B L2 -> SM, 8 B SM -> L2 -> DRAM --
// twice the L2 trips of a plain copy for the same 16 algorithmic bytes.  Printed: algorithmic GB/s (16 B per point), to be
// read next to tma_tile_copy's 6.5 - 6.8 TB/s for the one-pass copy of the same boxes, and the cycles a tile spends between
// the issue of its load and its landing / in the store's read of the buffer.
// Measured on B200 (profiles/r2b_two_pass_copy.log): 4.7 - 4.8 TB/s when the second pass reads rows written >= 48 rows
// (24 MB) earlier, 3.1 - 3.4 TB/s when its reads chase the first pass's writes by 16 - 24 rows; first pass alone 8.6 - 8.9
// TB/s read+write.  L2 moves 3 x 8 B per point and direction here against 2 x 8 B in a one-pass copy: ~14 TB/s either way.
// usage: two_pass_copy [ring_mb:lag_rows:discard:wpol:rpol ...]
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I dsc_b200/csrc -I include -o tools/micro/two_pass_copy tools/micro/two_pass_copy.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "fft_tma.cuh"

using namespace dscfft;

#ifndef TPC_TILE_KB
#define TPC_TILE_KB 64
#endif
#ifndef TPC_NBUF
#define TPC_NBUF 3
#endif
constexpr int TILE = TPC_TILE_KB * 1024, NBUF = TPC_NBUF;          // -DTPC_TILE_KB=32 -DTPC_NBUF=6: the 16-point variant's tiles
#ifndef TPC_N
#define TPC_N 256
#endif
// rows of TPC_N x TPC_N points (2^16 / 2^18 / 2^20): N2 / L first-pass + N2 / L second-pass tiles per row, box rows of L x 8 bytes
constexpr int N1 = TPC_N, N2 = TPC_N, L = TILE / (N1 * 8);
constexpr int BOX = N1 < 256 ? N1 : 256;                // box extents are at most 256
constexpr int TILES_PER_ROW = N2 / L;

__global__ void __launch_bounds__(64, 1)
two_pass_copy(const __grid_constant__ CUtensorMap in, const __grid_constant__ CUtensorMap wk, const __grid_constant__ CUtensorMap out,
              unsigned char *work, int ring_rows, int lag_rows, int rows, int discard, int skip_a, int skip_b, unsigned *ticket,
              unsigned long long *clk, int wpol, int rpol) {      // L2 hints of the work-row store / load: 0 evict_first, 1 evict_last, 2 evict_normal      // clk[4]: sum of cycles {A load, A store-read, B load, B store-read}, clk[4..5]: tiles A, B
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *buf = smem + ((1024u - (tma::smem_u32(smem) & 1023u)) & 1023u);
    unsigned long long *full = (unsigned long long *)(buf + NBUF * TILE);
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) tma::mbar_init(&full[b], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;                      // lane 0 issues every copy; the warp shares the discards
    const unsigned long long pol_stream = tma::policy_evict_first(), pol_keep = tma::policy_evict_last();
    unsigned long long pol_normal;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_normal));
    const unsigned long long pol_w = wpol == 0 ? pol_stream : wpol == 1 ? pol_keep : pol_normal;
    const unsigned long long pol_r = rpol == 0 ? pol_stream : rpol == 1 ? pol_keep : pol_normal;
    // ticket t: group g = t / 16 is one row's worth: 8 A tiles of row g, then 8 B tiles of row g - lag
    const long long total = (long long)(rows + lag_rows) * 2 * TILES_PER_ROW;
    struct Desc { int role_a, row, r, live; } mine[NBUF];
    long long t_issue[NBUF];
    unsigned long long acc[6] = {0, 0, 0, 0, 0, 0};
    auto decode = [&](long long t) {
        Desc d;
        const int g = (int)(t / (2 * TILES_PER_ROW)), i = (int)(t % (2 * TILES_PER_ROW));
        d.role_a = i < TILES_PER_ROW;
        d.r = i % TILES_PER_ROW;
        d.row = d.role_a ? g : g - lag_rows;
        d.live = d.row >= 0 && d.row < rows && !(d.role_a ? skip_a : skip_b);
        return d;
    };
    auto issue = [&](unsigned tt, const Desc d) {
        const int b = tt % NBUF;
        mine[b] = d;
        t_issue[b] = clock64();
        if (lane != 0) return;
        tma::mbar_arrive_expect_tx(&full[b], TILE);
        for (int r0 = 0; r0 < N1; r0 += BOX) {
            if (d.role_a) tma::load_3d(buf + b * TILE + (size_t)r0 * L * 8, &in, d.r * L, r0, d.row, &full[b], pol_stream);
            else tma::load_3d(buf + b * TILE + (size_t)r0 * L * 8, &wk, d.r * L, r0, d.row % ring_rows, &full[b], pol_r);
        }
    };
    auto next_live = [&](Desc &d) {
        for (;;) {
            unsigned tk = 0;
            if (lane == 0) tk = atomicAdd(ticket, 1u);
            const long long t = __shfl_sync(0xffffffffu, tk, 0);
            if (t >= total) return false;
            d = decode(t);
            if (d.live) return true;
        }
    };
    unsigned issued = 0;
    Desc d;
    bool more = next_live(d);
    while (issued < NBUF && more) { issue(issued++, d); more = next_live(d); }
    for (unsigned t = 0; t < issued; ++t) {
        const int b = t % NBUF;
        tma::mbar_wait(&full[b], (t / NBUF) & 1);
        const Desc m = mine[b];
        const long long t_full = clock64();
        acc[m.role_a ? 0 : 2] += (unsigned long long)(t_full - t_issue[b]);
        acc[m.role_a ? 4 : 5] += 1;
        if (m.role_a) {
            unsigned char *dst = work + ((size_t)(m.row % ring_rows) * TILES_PER_ROW + m.r) * TILE;
            if (lane == 0) tma::store_linear(dst, buf + b * TILE, TILE, pol_w);
        } else {
            if (discard && L * 8 >= 128) {       // narrower rows share their 128-byte lines with the neighbouring tile
                const unsigned char *base = work + (size_t)(m.row % ring_rows) * TILES_PER_ROW * TILE + (size_t)m.r * L * 8;
                constexpr int PER_ROW = L * 8 >= 128 ? L * 8 / 128 : 1;
                for (int i = lane; i < N2 * PER_ROW; i += 32) {
                    const unsigned char *p = base + (size_t)(i / PER_ROW) * N1 * 8 + (i % PER_ROW) * 128;
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
                }
            }
            if (lane == 0)
                for (int r0 = 0; r0 < N2; r0 += BOX) tma::store_3d(&out, m.r * L, r0, m.row, buf + b * TILE + (size_t)r0 * L * 8, pol_stream);
        }
        const long long t_st = clock64();
        if (lane == 0) { tma::store_commit(); tma::store_wait_read(); }
        __syncwarp();
        acc[m.role_a ? 1 : 3] += (unsigned long long)(clock64() - t_st);
        if (more) { issue(issued++, d); more = next_live(d); }
    }
    if (lane == 0) {
        tma::store_wait_all();
        for (int i = 0; i < 6; ++i) atomicAdd(clk + i, acc[i]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const size_t bytes = (size_t)1 << 30, row_bytes = (size_t)N1 * N2 * 8;
    const int rows = (int)(bytes / row_bytes);
    void *a, *b, *w; unsigned *ticket; unsigned long long *clk;
    cudaMalloc(&clk, 6 * sizeof(unsigned long long));
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&w, (size_t)256 << 20); cudaMalloc(&ticket, 4);
    cudaMemset(a, 1, bytes); cudaMemset(w, 0, (size_t)256 << 20);
    cudaFuncSetAttribute(two_pass_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * TILE + 2048);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto encode = [&](CUtensorMap *m, void *base, int nrows) {
        // rows of [256][256] 8-byte elements: box = [256 positions][32 columns]
        const cuuint64_t dims[3] = {(cuuint64_t)N2, (cuuint64_t)N1, (cuuint64_t)nrows};
        const cuuint64_t strides[2] = {(cuuint64_t)N2 * 8, (cuuint64_t)row_bytes};
        const cuuint32_t box[3] = {L, BOX, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    struct Case { int ring_mb, lag_rows, discard, skip_a, skip_b, wpol, rpol; char what[96]; };
    Case cases[64];
    int n_cases = 0;
    // usage: two_pass_copy [ring_mb:lag:discard:wpol:rpol ...]   (no arguments: the default set)
    for (int i = 1; i < argc && n_cases < 64; ++i) {
        Case c{64, 16, 1, 0, 0, 1, 0, ""};
        sscanf(argv[i], "%d:%d:%d:%d:%d", &c.ring_mb, &c.lag_rows, &c.discard, &c.wpol, &c.rpol);
        snprintf(c.what, sizeof(c.what), "ring %d MB lag %d discard %d wpol %d rpol %d", c.ring_mb, c.lag_rows, c.discard, c.wpol, c.rpol);
        cases[n_cases++] = c;
    }
    if (n_cases == 0) {
        const Case dflt[] = {
            {64, 48, 1, 0, 0, 1, 0, "both passes, 64 MB ring, reads 48 rows behind, discard"},
            {64, 48, 0, 0, 0, 1, 0, "both passes, 64 MB ring, reads 48 rows behind, no discard"},
            {128, 64, 1, 0, 0, 1, 0, "both passes, 128 MB ring, reads 64 rows behind, discard"},
            {64, 20, 1, 0, 0, 1, 0, "both passes, reads 20 rows behind (chasing the writes)"},
            {64, 16, 1, 0, 1, 1, 0, "first pass alone (DRAM -> L2 ring)"},
            {64, 16, 1, 1, 0, 1, 0, "second pass alone (stale ring -> DRAM)"},
        };
        for (const Case &c : dflt) cases[n_cases++] = c;
    }
    for (int ci = 0; ci < n_cases; ++ci) {
        const Case &c = cases[ci];
        const int ring_rows = (int)(((size_t)c.ring_mb << 20) / row_bytes);
        CUtensorMap mi, mw, mo;
        if (encode(&mi, a, rows) || encode(&mw, w, ring_rows) || encode(&mo, b, rows)) { printf("encode failed\n"); return 1; }
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaMemset(ticket, 0, 4);
            cudaMemset(clk, 0, 6 * sizeof(unsigned long long));
            cudaEventRecord(e0);
            two_pass_copy<<<148, 64, NBUF * TILE + 2048>>>(mi, mw, mo, (unsigned char *)w, ring_rows, c.lag_rows, rows, c.discard,
                                                          c.skip_a, c.skip_b, ticket, clk, c.wpol, c.rpol);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%-52s: %.3f ms per GiB  %7.0f GB/s %s  (%s)\n", c.what, best, 2.0 * bytes / best / 1e6,
               (c.skip_a || c.skip_b) ? "read+write of that pass" : "algorithmic (16 B per point)", cudaGetErrorString(cudaGetLastError()));
        unsigned long long h[6];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        printf("      cycles per tile: A load %.0f, A store read %.0f | B load %.0f, B store read %.0f\n",
               h[4] ? (double)h[0] / h[4] : 0.0, h[4] ? (double)h[1] / h[4] : 0.0, h[5] ? (double)h[2] / h[5] : 0.0, h[5] ? (double)h[3] / h[5] : 0.0);
    }
    return 0;
}
