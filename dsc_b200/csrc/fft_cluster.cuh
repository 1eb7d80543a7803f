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


// ------------------------------------------------------------------------------------------------------------------
// The same data path, pipelined: persistent clusters, three tile buffers per block, two butterfly groups, a loader and a
// storer thread (the structure of four_step_tma), and mbarrier-signalled distributed-shared-memory stores instead of
// whole-cluster barriers.
//
// A cluster walks its lines i = 0, 1, ... (line = cluster + i * clusters); line i lives in buffer i % 3 of EVERY block of
// the cluster and is transformed by group i % 2 of every block.  Per block and line:
//   loader : wait empty[b] -> TMA box load of the block's column block -> full[b]
//   group  : wait full[b] -> first half (length n1) -> "my buffer may be overwritten": arrive on free[b] of every block of
//            the cluster (count C) -> wait free[b] -> st.async each point into its owner's buffer, completing bytes on
//            the owner's landed[b] -> wait landed[b] (64 KiB) -> second half (length n2) -> ready[b]
//   storer : wait ready[b] -> TMA box store -> empty[b]
// While one group waits for its peers or for bytes to land, the other group computes and the next line is already loading.
namespace cl {
DSC_DEV void mbar_arrive_remote(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
DSC_DEV void mbar_wait_cluster(void *bar, unsigned parity) {
    const unsigned a = tma::smem_u32(bar);
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
DSC_DEV void store_async(unsigned addr, float2 v, unsigned mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "f"(v.x), "f"(v.y), "r"(mbar) : "memory");
}
DSC_DEV void store_async(unsigned addr, double2 v, unsigned mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "d"(v.x), "d"(v.y), "r"(mbar) : "memory");
}
DSC_DEV void sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
}  // namespace cl

template <typename T, int LG_N1> struct ClusterPipeSmem {
    using V = cx<T>;
    static constexpr int E = 1 << tma_lg_e<T>();
    static constexpr int L = tma_tile_points<T>() >> LG_N1;
    alignas(1024) unsigned char buf[TMA_BUFFERS][TMA_TILE_BYTES];
    V table[L * E];                                // W^(q TT c), [c][line]: the same for every line of the launch
    unsigned long long full[TMA_BUFFERS];          // the block's column block has landed (TMA)
    unsigned long long free_[TMA_BUFFERS];         // every block of the cluster has read its tile: the buffers may be overwritten
    unsigned long long landed[TMA_BUFFERS];        // the block's rows have arrived from all blocks (st.async bytes)
    unsigned long long ready[TMA_BUFFERS];         // the finished rows lie in the buffer
    unsigned long long empty[TMA_BUFFERS];         // the TMA store has read the buffer
};

template <typename T, int LG_N1, int LG_N2, bool FWD>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fft_cluster_pipe(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out, const ClusterArgs a,
                 const unsigned lines, const unsigned clusters) {
    using V = cx<T>;
    constexpr int LG_E = tma_lg_e<T>(), E = 1 << LG_E;
    constexpr int N1 = 1 << LG_N1, N2 = 1 << LG_N2;
    constexpr int L = tma_tile_points<T>() >> LG_N1;
    constexpr int C = N2 / L;
    constexpr int LP = N1 / C;
    constexpr int TT1 = N1 >> LG_E, TT2 = N2 >> LG_E;
    static_assert(L * TT1 == TMA_GROUP_THREADS && LP * TT2 == TMA_GROUP_THREADS, "one register tile per thread in both halves");
    using TileA = TmaTile<T, LG_N1, L, FWD, false>;
    using TileB = TmaTile<T, LG_N2, LP, FWD, false>;
    static_assert(TileA::STAGES == 2 && TileB::STAGES == 2, "two stages per half");
    constexpr int SLOTS = 128 / (int)sizeof(V), LG_SLOTS = SLOTS == 16 ? 4 : 3;
    constexpr int LG_TT2 = LG_N2 - LG_E, SH = LG_SLOTS - LG_TT2;
    static_assert(TT2 <= SLOTS && LP >= SLOTS, "second-half swizzles assume short rows of threads and wide tiles");
    constexpr int ES = sizeof(T) == 4 ? 1 : 2;

    DSC_DYN_SMEM(smem_raw);
    using Smem = ClusterPipeSmem<T, LG_N1>;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u));

    const int tid = threadIdx.x;
    const unsigned rank = cl::cta_rank();
    const unsigned cid = cl::cluster_id_x();
    const unsigned my_lines = cid < lines ? (lines - cid + clusters - 1) / clusters : 0;     // the same in every block of the cluster
    if (tid == 0) {
        for (int b = 0; b < TMA_BUFFERS; ++b) {
            tma::mbar_init(&sm.full[b], 1);
            tma::mbar_init(&sm.free_[b], C);
            tma::mbar_init(&sm.landed[b], 1);
            tma::mbar_init(&sm.ready[b], TMA_GROUP_THREADS);
            tma::mbar_init(&sm.empty[b], 1);
        }
        tma::fence_barrier_init();
    }
    TmaArgs ta{};
    ta.tw_lo = a.tw_lo; ta.tw_hi = a.tw_hi; ta.four_shift = a.four_shift; ta.four_mask = a.four_mask;
    for (int i = tid; i < L * E; i += TMA_THREADS) {
        const unsigned ll = (unsigned)(i % L), c = (unsigned)(i / L);
        sm.table[i] = tma_twiddle<T>(ta, (rank * (unsigned)L + ll) * (unsigned)TT1 * c);
    }
    __syncthreads();
    cl::sync_all();          // every block's barriers exist before anyone arrives on them remotely

    if (tid >= TMA_GROUPS * TMA_GROUP_THREADS) {
        const int warp = (tid - TMA_GROUPS * TMA_GROUP_THREADS) / 32;
        if ((tid & 31) == 0) {
            const unsigned long long pol_stream = tma::policy_evict_first();
            if (warp == 0) {
                // ---- loader
                for (unsigned t = 0; t < my_lines; ++t) {
                    const int b = (int)(t % TMA_BUFFERS);
                    if (t >= TMA_BUFFERS) tma::mbar_wait(&sm.empty[b], (t / TMA_BUFFERS - 1) & 1);
                    const unsigned line = cid + t * clusters;
                    tma::mbar_arrive_expect_tx(&sm.full[b], TMA_TILE_BYTES);
                    constexpr int BOX = tma_box_rows(LG_N1);
#pragma unroll
                    for (int r0 = 0; r0 < N1; r0 += BOX)
                        tma::load_3d(sm.buf[b] + (size_t)r0 * L * sizeof(V), &map_x, (int)rank * L * ES, r0, (int)line, &sm.full[b], pol_stream);
                }
            } else {
                // ---- storer
                for (unsigned t = 0; t < my_lines; ++t) {
                    const int b = (int)(t % TMA_BUFFERS);
                    tma::mbar_wait(&sm.ready[b], (t / TMA_BUFFERS) & 1);
                    const unsigned line = cid + t * clusters;
                    constexpr int BOX = tma_box_rows(LG_N2);
#pragma unroll
                    for (int r0 = 0; r0 < N2; r0 += BOX)
                        tma::store_3d(&map_out, (int)rank * LP * ES, r0, (int)line, sm.buf[b] + (size_t)r0 * LP * sizeof(V), pol_stream);
                    tma::store_commit();
                    tma::store_wait_read();
                    tma::mbar_arrive(&sm.empty[b]);
                }
                tma::store_wait_all();
            }
        }
    } else {
        // ---- butterfly groups
        const int group = tid / TMA_GROUP_THREADS, gtid = tid % TMA_GROUP_THREADS;
        const int bar_id = 1 + group;
        const int l = gtid % L, j = gtid / L;
        const unsigned q = rank * (unsigned)L + (unsigned)l;
        const V w0 = tma_twiddle<T>(ta, q * (unsigned)j);
        const int row = gtid / TT2, jj = gtid % TT2;           // second half, first read: consecutive positions of one row
        const int sw = (row & (SLOTS / TT2 - 1)) * TT2;
        const int l2 = gtid % LP, j2 = gtid / LP;              // second half after its exchange: adjacent lanes, adjacent rows
        for (unsigned t = (unsigned)group; t < my_lines; t += TMA_GROUPS) {
            const int b = (int)(t % TMA_BUFFERS);
            const unsigned par = (t / TMA_BUFFERS) & 1;
            V *buf = reinterpret_cast<V *>(sm.buf[b]);
            tma::mbar_wait(&sm.full[b], par);
            V v[E];
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = buf[(j + c * TT1) * L + l];
            TileA::stage_first(v, buf, nullptr, ta, l, j, l, j, a.tw_a, 0u, gtid, bar_id);
            dsc_group_barrier(bar_id, TMA_GROUP_THREADS);        // the whole group has read its tile for the last time
            if (gtid == 0) {
                // the bytes this block is about to receive, then "my buffer is free" to every block of the cluster
                tma::mbar_arrive_expect_tx(&sm.landed[b], TMA_TILE_BYTES);
                const unsigned fr = tma::smem_u32(&sm.free_[b]);
#pragma unroll
                for (unsigned p = 0; p < (unsigned)C; ++p) cl::mbar_arrive_remote(cl::map(fr, p));
            }
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = cmul_tw<FWD>(v[c], c == 0 ? w0 : cmul(w0, sm.table[c * L + l]));
            cl::mbar_wait_cluster(&sm.free_[b], par);
            {
                const unsigned local = tma::smem_u32(buf), lb = tma::smem_u32(&sm.landed[b]);
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    const unsigned k1 = (unsigned)(j + c * TT1);
                    const unsigned dst = k1 / (unsigned)LP, r = k1 % (unsigned)LP;
                    const unsigned pos = q ^ ((r & (unsigned)(SLOTS / TT2 - 1)) * (unsigned)TT2);
                    cl::store_async(cl::map(local + (r * (unsigned)N2 + pos) * (unsigned)sizeof(V), dst), v[c], cl::map(lb, dst));
                }
            }
            cl::mbar_wait_cluster(&sm.landed[b], par);
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = buf[row * N2 + ((jj + c * TT2) ^ sw)];
            Dft<E, FWD, T>::run(v);
            dsc_group_barrier(bar_id, TMA_GROUP_THREADS);        // every thread has read the received rows
#pragma unroll
            for (int p = 0; p < E; ++p) buf[(jj * E + p) * LP + (row ^ (jj << SH))] = v[p];
            dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = buf[(j2 + c * TT2) * LP + (l2 ^ (((c * TT2) >> LG_E) << SH))];
            TileB::template stage<1>(v, buf, l2, j2, l2, j2, a.tw_b, bar_id);
            if (a.do_scale) {
                const T s = (T)a.scale;
#pragma unroll
                for (int c = 0; c < E; ++c) { v[c].x *= s; v[c].y *= s; }
            }
            dsc_group_barrier(bar_id, TMA_GROUP_THREADS);        // every thread has read its last-stage inputs
#pragma unroll
            for (int c = 0; c < E; ++c) buf[(j2 + c * TT2) * LP + l2] = v[c];
            tma::fence_async_smem();
            tma::mbar_arrive(&sm.ready[b]);
        }
    }
    // nobody leaves while a peer may still arrive on or write into its shared memory
    cl::sync_all();
}

}  // namespace dscfft

#endif  // !DSC_EMUL
