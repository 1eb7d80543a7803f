// fft_cluster.cuh -- lines of 2^14 .. 2^17 complex64 points in ONE pass over HBM: a thread-block cluster owns a
// line, split over the shared memory of its C = 2 / 4 / 8 / 16 blocks, and the transpose between the two halves of
// the four-step decomposition goes through distributed shared memory instead of through L2.
//
// Why: a two-pass transform moves every point through L2 four times (x in, work row out, work row in, result out).
// Measured on B200, that caps it near half the copy peak whatever feeds the tiles (register-direct loads 2.5-2.8
// TB/s, TMA-fed tiles 2.55-3.0 TB/s, ncu: butterfly warps waiting for data 72 % of the time, DRAM read exactly the
// algorithmic bytes): the wall is L2 throughput, not HBM.  Keeping the intermediate inside the cluster removes half
// of that traffic.
//
//   n = n1 * n2, line x[i1][q]; block c of the cluster owns the L = n2 / C columns q in [c L, (c+1) L):
//     1. one bulk tensor load (TMA box [n1][L], SASS UTMALDG) into the block's 64 KiB tile, [position][line];
//     2. length-n1 transforms over i1 for its L columns (radix-32 x radix-n1/32, one exchange in shared memory),
//        times W_n^(q k1);
//     3. every thread stores its points straight from registers into the tile of the block that owns row k1
//        (st.shared::cluster; a warp writes 128-256 contiguous bytes), between two cluster barriers: block c' then
//        holds the L' = n1 / C rows k1 in [c' L', (c'+1) L') complete, [line][position];
//     4. length-n2 transforms over q of its L' rows (the threads start as consecutive positions of one row and
//        come back from the exchange as adjacent lanes on adjacent rows);
//     5. one bulk tensor store (box [n2][L'], UTMASTG) of X[k1 + n1 k2].
//   HBM and L2 see one read and one write of the line.
//
// Reference work replaced: /root/reference/dsc/include/dsc_fft.h:57-103 (dsc_fft_pass2), :168-175 (1/N).
#pragma once

#if !defined(DSC_EMUL)

#include "fft_tma.cuh"

namespace dscfft {

namespace cl {
DSC_DEV unsigned cta_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
DSC_DEV unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
DSC_DEV void arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
DSC_DEV void wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// shared::cluster address of `local` (a shared::cta address) in block `rank`
DSC_DEV unsigned map(unsigned local, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
DSC_DEV void store(unsigned addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
DSC_DEV void store(unsigned addr, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
}  // namespace cl

constexpr int CLUSTER_THREADS = 256;

struct ClusterArgs {
    const void *tw_a[DSC_CUDA_MAX_STAGES];     // stage tables of the length-n1 transform
    const void *tw_b[DSC_CUDA_MAX_STAGES];     // stage tables of the length-n2 transform
    const void *tw_lo, *tw_hi;                 // W_n^p split tables
    int four_shift, four_mask;
    double scale;
    int do_scale;
};

template <typename T, int LG_N1> struct ClusterSmem {
    using V = cx<T>;
    static constexpr int E = 1 << tma_lg_e<T>();
    static constexpr int L = tma_tile_points<T>() >> LG_N1;
    alignas(1024) unsigned char buf[TMA_TILE_BYTES];
    V table[L * E];                            // W^(q TT c), [c][line]
    unsigned long long full;
};

template <typename T, int LG_N1, int LG_N2, bool FWD>
__global__ void __launch_bounds__(CLUSTER_THREADS, 2)
fft_cluster(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out, const ClusterArgs a) {
    using V = cx<T>;
    constexpr int LG_E = tma_lg_e<T>(), E = 1 << LG_E;
    constexpr int N1 = 1 << LG_N1, N2 = 1 << LG_N2;
    constexpr int L = tma_tile_points<T>() >> LG_N1;         // columns per block in the first half
    constexpr int C = N2 / L;                                // blocks per cluster
    constexpr int LP = N1 / C;                               // rows per block in the second half
    constexpr int TT1 = N1 >> LG_E, TT2 = N2 >> LG_E;
    static_assert(L * TT1 == CLUSTER_THREADS && LP * TT2 == CLUSTER_THREADS, "one register tile per thread in both halves");
    static_assert(C >= 2 && C <= 16, "cluster of 2..16 blocks");
    using TileA = TmaTile<T, LG_N1, L, FWD, false>;
    using TileB = TmaTile<T, LG_N2, LP, FWD, false>;
    static_assert(TileA::STAGES == 2 && TileB::STAGES == 2, "two stages per half");
    constexpr int SLOTS = 128 / (int)sizeof(V), LG_SLOTS = SLOTS == 16 ? 4 : 3;
    constexpr int LG_TT2 = LG_N2 - LG_E;
    static_assert(TT2 <= SLOTS && LP >= SLOTS, "second-half swizzles assume short rows of threads and wide tiles");
    constexpr int ES = sizeof(T) == 4 ? 1 : 2;

    DSC_DYN_SMEM(smem_raw);
    using Smem = ClusterSmem<T, LG_N1>;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u));
    V *buf = reinterpret_cast<V *>(sm.buf);

    const int tid = threadIdx.x;
    const unsigned rank = cl::cta_rank();
    const unsigned line = cl::cluster_id_x();
    const unsigned long long pol_stream = tma::policy_evict_first();
    if (tid == 0) {
        tma::mbar_init(&sm.full, 1);
        tma::fence_barrier_init();
        tma::mbar_arrive_expect_tx(&sm.full, TMA_TILE_BYTES);
        constexpr int BOX = tma_box_rows(LG_N1);
#pragma unroll
        for (int r0 = 0; r0 < N1; r0 += BOX)
            tma::load_3d(sm.buf + (size_t)r0 * L * sizeof(V), &map_x, (int)rank * L * ES, r0, (int)line, &sm.full, pol_stream);
    }
    // ---- first half: columns q = rank L + l
    const int l = tid % L, j = tid / L;
    const unsigned q = rank * (unsigned)L + (unsigned)l;
    TmaArgs ta{};
    ta.tw_lo = a.tw_lo; ta.tw_hi = a.tw_hi; ta.four_shift = a.four_shift; ta.four_mask = a.four_mask;
    // inter-pass twiddles while the tile travels: W^(q j) per thread, W^(q TT1 c) per (c, column) in shared memory
    const V w0 = tma_twiddle<T>(ta, q * (unsigned)j);
    for (int i = tid; i < L * E; i += CLUSTER_THREADS) {
        const unsigned ll = (unsigned)(i % L), c = (unsigned)(i / L);
        sm.table[i] = tma_twiddle<T>(ta, (rank * (unsigned)L + ll) * (unsigned)TT1 * c);
    }
    __syncthreads();                       // the mbarrier is initialised, the table is complete
    tma::mbar_wait(&sm.full, 0);
    V v[E];
#pragma unroll
    for (int c = 0; c < E; ++c) v[c] = buf[(j + c * TT1) * L + l];
    TileA::stage_first(v, buf, nullptr, ta, l, j, l, j, a.tw_a, 0u, tid, 0);
    // this block has read its tile for the last time: the others may overwrite it
    cl::arrive();
#pragma unroll
    for (int c = 0; c < E; ++c) v[c] = cmul_tw<FWD>(v[c], c == 0 ? w0 : cmul(w0, sm.table[c * L + l]));
    cl::wait();
    // ---- the transpose: point (k1, q) goes to block k1 / LP, row k1 % LP, position q (XOR-swizzled within 128 bytes
    // by the row, so that the second half's first read -- 16 lanes = a few rows x consecutive positions -- is
    // conflict-free)
    {
        const unsigned local = tma::smem_u32(buf);
#pragma unroll
        for (int c = 0; c < E; ++c) {
            const unsigned k1 = (unsigned)(j + c * TT1);
            const unsigned dst = k1 / (unsigned)LP, row = k1 % (unsigned)LP;
            const unsigned pos = q ^ ((row & (unsigned)(SLOTS / TT2 - 1)) * (unsigned)TT2);
            cl::store(cl::map(local + (row * (unsigned)N2 + pos) * (unsigned)sizeof(V), dst), v[c]);
        }
    }
    cl::arrive();
    cl::wait();
    // ---- second half: rows k1 = rank LP + row, transform over q
    {
        const int row = tid / TT2, jj = tid % TT2;           // consecutive positions of one row
        const int sw = (row & (SLOTS / TT2 - 1)) * TT2;
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = buf[row * N2 + ((jj + c * TT2) ^ sw)];
        Dft<E, FWD, T>::run(v);
        __syncthreads();                   // every thread has read the received rows
        // exchange layout [position][row], the row's slot XOR-ed with the writer's position-in-row bits
        constexpr int SH = LG_SLOTS - LG_TT2;
#pragma unroll
        for (int p = 0; p < E; ++p) buf[(jj * E + p) * LP + (row ^ (jj << SH))] = v[p];
        __syncthreads();
        const int l2 = tid % LP, j2 = tid / LP;              // adjacent lanes on adjacent rows
#pragma unroll
        for (int c = 0; c < E; ++c) {
            // position j2 + c TT2: its writer was thread (j2 + c TT2) / E of the row = (c TT2) / E (j2 < TT2 never carries)
            constexpr int unused = 0; (void)unused;
            const int wj = ((c * TT2) >> LG_E);
            v[c] = buf[(j2 + c * TT2) * LP + (l2 ^ (wj << SH))];
        }
        TileB::template stage<1>(v, buf, l2, j2, l2, j2, a.tw_b, 0);
        if (a.do_scale) {
            const T s = (T)a.scale;
#pragma unroll
            for (int c = 0; c < E; ++c) { v[c].x *= s; v[c].y *= s; }
        }
        __syncthreads();                   // every thread has read its last-stage inputs
#pragma unroll
        for (int c = 0; c < E; ++c) buf[(j2 + c * TT2) * LP + l2] = v[c];
    }
    tma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        constexpr int BOX = tma_box_rows(LG_N2);
#pragma unroll
        for (int r0 = 0; r0 < N2; r0 += BOX)
            tma::store_3d(&map_out, (int)rank * LP * ES, r0, (int)line, sm.buf + (size_t)r0 * LP * sizeof(V), pol_stream);
        tma::store_commit();
        tma::store_wait_read();            // the tile must outlive the copy's reads
    }
}

}  // namespace dscfft

#endif  // !DSC_EMUL
