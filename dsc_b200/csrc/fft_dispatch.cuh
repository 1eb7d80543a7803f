// fft_dispatch.cuh -- compile-time tables of fft_lines<> instantiations, indexed by log2(N).
//
// Per precision the register tile is E = 16 (float) / 8 (double) points per thread, so
// butterflies are radix-16 / radix-8.  Two block shapes exist per length:
//   "C" (contiguous lines, last axis): ~256 threads per block, whole lines per warp group;
//   "S" (strided lines, axis != last and both four-step passes): LPB adjacent lines so a
//       warp touches >= 64..128 contiguous bytes per access.
#pragma once

#include <cstdlib>
#include <utility>
#include "fft_kernels.cuh"
#include "fft_tma.cuh"
#include "fft_cluster.cuh"

namespace dscfft {

struct KernelEntry {
    void (*fn)(const FftArgs);
    int lpb;        // lines per block
    int threads;    // block size
    int smem;       // dynamic shared memory bytes
    bool configured;
    // dense lines with one block per SM: persistent blocks that prefetch their next line into L2 (fft_lines_persist)
    void (*fn_persist)(const FftArgs, const long long);
    int grid_persist;
};

template <typename T> struct Tile;
template <> struct Tile<float>  { static constexpr int LG_E = 4; static constexpr int MAX_LG = 14; static constexpr int COAL = 16; };
template <> struct Tile<double> { static constexpr int LG_E = 3; static constexpr int MAX_LG = 13; static constexpr int COAL = 8; };

// points per thread: radix-16 (float) / radix-8 (double) tiles; the longest float lines use radix-32 so
// that three stages (two shared-memory exchanges) still cover them
template <typename T> constexpr int lg_e_for(int lg_n) {
    if (sizeof(T) == 4 && lg_n >= 13) return 5;
    if (sizeof(T) == 8 && lg_n >= 12) return 4;
    // 16- and 32-point lines: four threads per line (direct coalesced accesses, one warp-synchronous exchange)
    // instead of one or two threads per line behind a staged copy -- measured 3.9 -> 5.7 and 2.5 -> 5.7 TB/s
    if (lg_n == 4 || (lg_n == 5 && sizeof(T) == 4)) return lg_n - 2;
    if (lg_n == 6 && sizeof(T) == 4) return 3;              // 64 points as 8 x 8 with eight threads per line: 5.6 -> 6.9 TB/s
    return lg_n < Tile<T>::LG_E ? lg_n : Tile<T>::LG_E;
}

constexpr int imin(int a, int b) { return a < b ? a : b; }
constexpr int imax(int a, int b) { return a > b ? a : b; }

// lines per block, contiguous shape: aim for 256 threads
template <typename T> constexpr int lpb_c(int lg_n) {
    const int tt = 1 << (lg_n - lg_e_for<T>(lg_n));
    return imax(1, 256 / tt);
}
// lines per block, strided shape: enough adjacent lines for coalescing; at most 1024 threads with the
// standard register tile (<= 64 registers per thread), 512 with the double-size tile of the longest lines
template <typename T> constexpr int lpb_s(int lg_n) {
    const int tt = 1 << (lg_n - lg_e_for<T>(lg_n));
    const int cap = lg_e_for<T>(lg_n) > Tile<T>::LG_E ? 512 : 1024;
    return imax(1, imin(imax(Tile<T>::COAL, 256 / tt), cap / tt));
}

template <typename T, bool FWD, int MODE, bool SV, int LG_N>
KernelEntry make_entry() {
    constexpr int LG_E = lg_e_for<T>(LG_N);
    constexpr int LPB = SV ? lpb_s<T>(LG_N) : lpb_c<T>(LG_N);
    using Sc = Sched<LG_N, LG_E>;
    KernelEntry e;
    e.fn = fft_lines<T, LG_N, LG_E, LPB, FWD, MODE>;
    e.lpb = LPB;
    e.threads = LPB * Sc::TT;
    e.smem = LPB * Sc::line_stride(LPB, (int)sizeof(cx<T>), mode_is_dense(MODE) ? Sc::TT : 0) * (int)sizeof(cx<T>);
    e.configured = false;
    e.fn_persist = nullptr;
    e.grid_persist = 0;
#if !defined(DSC_EMUL)
    if constexpr ((MODE == MODE_FAST || MODE == MODE_R2C_FAST || MODE == MODE_C2R_FAST) &&
                  LPB * Sc::line_stride(LPB, (int)sizeof(cx<T>), Sc::TT) * (int)sizeof(cx<T>) > 114 * 1024)
        e.fn_persist = fft_lines_persist<T, LG_N, LG_E, LPB, FWD, MODE>;
#endif
    return e;
}

template <typename T, bool FWD, int MODE, bool SV, int... I>
KernelEntry *build_table(std::integer_sequence<int, I...>) {
    static KernelEntry table[] = {make_entry<T, FWD, MODE, SV, I>()...};
    return table;
}

// table[lg_n] for lg_n in [0, Tile<T>::MAX_LG]; defined by explicit instantiation in inst_*.cu
template <typename T, bool FWD, int MODE, bool SV>
KernelEntry *get_table();


// every table that exists (each defined in exactly one inst_*.cu)
#define DSC_DECLARE_TABLE(T, FWD, MODE, SV) template <> KernelEntry *get_table<T, FWD, MODE, SV>();
DSC_DECLARE_TABLE(float, true, MODE_C2C, false)   DSC_DECLARE_TABLE(float, true, MODE_C2C, true)
DSC_DECLARE_TABLE(float, false, MODE_C2C, false)  DSC_DECLARE_TABLE(float, false, MODE_C2C, true)
DSC_DECLARE_TABLE(double, true, MODE_C2C, false)  DSC_DECLARE_TABLE(double, true, MODE_C2C, true)
DSC_DECLARE_TABLE(double, false, MODE_C2C, false) DSC_DECLARE_TABLE(double, false, MODE_C2C, true)
DSC_DECLARE_TABLE(float, true, MODE_R2C, false)   DSC_DECLARE_TABLE(float, false, MODE_C2R, false)
DSC_DECLARE_TABLE(double, true, MODE_R2C, false)  DSC_DECLARE_TABLE(double, false, MODE_C2R, false)
DSC_DECLARE_TABLE(float, true, MODE_FILTER, false) DSC_DECLARE_TABLE(double, true, MODE_FILTER, false)
DSC_DECLARE_TABLE(float, true, MODE_FAST, false)  DSC_DECLARE_TABLE(float, false, MODE_FAST, false)
DSC_DECLARE_TABLE(double, true, MODE_FAST, false) DSC_DECLARE_TABLE(double, false, MODE_FAST, false)
DSC_DECLARE_TABLE(float, true, MODE_R2C_FAST, false)  DSC_DECLARE_TABLE(float, false, MODE_C2R_FAST, false)
DSC_DECLARE_TABLE(double, true, MODE_R2C_FAST, false) DSC_DECLARE_TABLE(double, false, MODE_C2R_FAST, false)
#undef DSC_DECLARE_TABLE

// ---- fused four-step launches ------------------------------------------------------------------
struct FusedEntry {
    void (*fn)(const FftArgs, const FftArgs, const FourStepSync);
    int lg_n1, lg_n2, threads, lpb_a, lpb_b, smem;
    int grid;            // persistent launch: resident blocks on the whole device (set when first configured)
    bool configured;
    // float factors <= 512: 16 points per thread (half the registers: four 256-thread blocks per SM)
    void (*fn16)(const FftArgs, const FftArgs, const FourStepSync);
    int lpb_a16, lpb_b16, smem16, grid16;
};

// block size of the fused launch: 64 payload registers per thread, 512 threads per SM.  Two 256-thread blocks
// overlap their phases better than one 512-thread block, except for float first passes of 1024 points, where
// 256 threads are only 8 lines = 64-byte segments: there 16 lines (128-byte segments) win (measured +8 %).
template <typename T> constexpr int fused_threads(int lg_n1, int lg_n2) {
    (void)lg_n2;
    return (sizeof(T) == 4 && lg_n1 >= 10) ? 512 : 256;
}

template <typename T, bool FWD, int LG_N1, int LG_N2, int THREADS = fused_threads<T>(LG_N1, LG_N2)>
FusedEntry make_fused() {
    constexpr int LG_E1 = pass_lg_e<T>(LG_N1, LG_N2), LG_E2 = pass_lg_e<T>(LG_N2, LG_N1);
    constexpr int LPB_A = THREADS >> (LG_N1 - LG_E1), LPB_B = THREADS >> (LG_N2 - LG_E2);
    FusedEntry e;
    e.fn = four_step_fused<T, LG_N1, LG_N2, THREADS, FWD>;
    e.lg_n1 = LG_N1; e.lg_n2 = LG_N2; e.threads = THREADS; e.lpb_a = LPB_A; e.lpb_b = LPB_B;
    e.smem = fused_smem_bytes<T, LG_N1, LG_N2, THREADS>();
    e.grid = 0;
    e.configured = false;
    e.fn16 = nullptr; e.lpb_a16 = e.lpb_b16 = e.smem16 = e.grid16 = 0;
    if constexpr (sizeof(T) == 4 && LG_N1 <= 9 && LG_N2 <= 9 && THREADS == 256) {
        e.fn16 = four_step_fused<T, LG_N1, LG_N2, THREADS, FWD, 4>;
        e.lpb_a16 = THREADS >> (LG_N1 - 4); e.lpb_b16 = THREADS >> (LG_N2 - 4);
        e.smem16 = fused_smem_bytes<T, LG_N1, LG_N2, THREADS, 4>();
    }
    return e;
}

// (lg_n1, lg_n2) pairs produced by plan_layout for lengths up to 2^20: lg_n2 = lg_n / 2, lg_n1 = lg_n - lg_n2
#define DSC_FUSED_PAIRS(X) X(7, 7) X(8, 7) X(8, 8) X(9, 8) X(9, 9) X(10, 9) X(10, 10)

// nullptr when the pair has no fused instantiation; defined in inst_*_fused.cu
template <typename T, bool FWD> FusedEntry *fused_entry(int lg_n1, int lg_n2);
template <> FusedEntry *fused_entry<float, true>(int, int);
template <> FusedEntry *fused_entry<float, false>(int, int);
template <> FusedEntry *fused_entry<double, true>(int, int);
template <> FusedEntry *fused_entry<double, false>(int, int);

#define DSC_DEFINE_FUSED(T, FWD)                                                           \
    template <> FusedEntry *fused_entry<T, FWD>(int lg_n1, int lg_n2) {                    \
        static FusedEntry table[] = {DSC_FUSED_PAIRS(DSC_FUSED_MAKE_##FWD##_##T)};         \
        for (auto &e : table) if (e.lg_n1 == lg_n1 && e.lg_n2 == lg_n2) return &e;        \
        return nullptr;                                                                    \
    }
#define DSC_FUSED_MAKE_true_float(A, B) make_fused<float, true, A, B>(),
#define DSC_FUSED_MAKE_false_float(A, B) make_fused<float, false, A, B>(),
#define DSC_FUSED_MAKE_true_double(A, B) make_fused<double, true, A, B>(),
#define DSC_FUSED_MAKE_false_double(A, B) make_fused<double, false, A, B>(),

// ---- TMA-fed four-step launches (fft_tma.cuh) -----------------------------------------------------------------
#if !defined(DSC_EMUL)
struct TmaEntry {
    void (*fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TmaArgs, const FourStepSync);
    // float passes of <= 512 points: 16 points per thread on 32 KiB tiles (half the lines per tile), two blocks per SM
    void (*fn16)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TmaArgs, const FourStepSync);
    int smem16, ctas16;
    // the packed-real bin-pair step fused in: forward launches un-mix in the second pass (runs of at least 64 bytes);
    // inverse launches (double: float bin rows have an odd pitch no tensor map takes) mix in the first pass
    void (*fn_real)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TmaArgs, const FourStepSync);
    // finished tiles stored from the registers, buffers released after the last exchange (four_step_tma DIRECT)
    void (*fn_direct)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TmaArgs, const FourStepSync);
    int lg_n1, lg_n2, l_a, l_b, box_a, box_b, smem;
    int grid;
    bool configured;
};

template <typename T, bool FWD, int LG_N1, int LG_N2>
TmaEntry make_tma() {
    TmaEntry e;
    e.fn = four_step_tma<T, LG_N1, LG_N2, FWD>;
    e.fn16 = nullptr; e.smem16 = 0; e.ctas16 = 0;
    if constexpr (sizeof(T) == 4 && LG_N1 <= 9 && LG_N2 <= 9) {
        e.fn16 = four_step_tma<T, LG_N1, LG_N2, FWD, 4, TMA_TILE_BYTES / 2>;
        e.smem16 = (int)sizeof(TmaSmem<T, TMA_TILE_BYTES / 2>) + 1024;
    }
    e.fn_real = nullptr;
    if constexpr (FWD && (tma_lines<T>(LG_N2) / 2) * (int)sizeof(cx<T>) >= 64)
        e.fn_real = four_step_tma<T, LG_N1, LG_N2, FWD, tma_lg_e<T>(), TMA_TILE_BYTES, 1>;
    if constexpr (!FWD && sizeof(T) == 8 && (tma_lines<T>(LG_N1) / 2) * (int)sizeof(cx<T>) >= 64 && (1 << LG_N1) / 2 <= TMA_GROUP_THREADS)
        e.fn_real = four_step_tma<T, LG_N1, LG_N2, FWD, tma_lg_e<T>(), TMA_TILE_BYTES, 2>;
    e.fn_direct = four_step_tma<T, LG_N1, LG_N2, FWD, tma_lg_e<T>(), TMA_TILE_BYTES, 0, true>;
    e.lg_n1 = LG_N1; e.lg_n2 = LG_N2;
    e.l_a = tma_lines<T>(LG_N1); e.l_b = tma_lines<T>(LG_N2);
    e.box_a = tma_box_rows(LG_N1); e.box_b = tma_box_rows(LG_N2);
    e.smem = (int)sizeof(TmaSmem<T>) + 1024;
    e.grid = 0;
    e.configured = false;
    return e;
}

// float: every fused pair from 2^15 up; double (16-byte elements, 4096-point tiles): passes of at most 512 points
#define DSC_TMA_PAIRS_float(X) X(8, 7) X(8, 8) X(9, 8) X(9, 9) X(10, 9) X(10, 10)
#define DSC_TMA_PAIRS_double(X) X(7, 7) X(8, 7) X(8, 8) X(9, 8) X(9, 9)

template <typename T, bool FWD> TmaEntry *tma_entry(int lg_n1, int lg_n2);
template <> TmaEntry *tma_entry<float, true>(int, int);
template <> TmaEntry *tma_entry<float, false>(int, int);
template <> TmaEntry *tma_entry<double, true>(int, int);
template <> TmaEntry *tma_entry<double, false>(int, int);

#define DSC_DEFINE_TMA(T, FWD)                                                             \
    template <> TmaEntry *tma_entry<T, FWD>(int lg_n1, int lg_n2) {                        \
        static TmaEntry table[] = {DSC_TMA_PAIRS_##T(DSC_TMA_MAKE_##FWD##_##T)};           \
        for (auto &e : table) if (e.lg_n1 == lg_n1 && e.lg_n2 == lg_n2) return &e;        \
        return nullptr;                                                                    \
    }
#define DSC_TMA_MAKE_true_float(A, B) make_tma<float, true, A, B>(),
#define DSC_TMA_MAKE_false_float(A, B) make_tma<float, false, A, B>(),
#define DSC_TMA_MAKE_true_double(A, B) make_tma<double, true, A, B>(),
#define DSC_TMA_MAKE_false_double(A, B) make_tma<double, false, A, B>(),
#endif

// ---- one line per thread-block cluster (fft_cluster.cuh) --------------------------------------------------------
#if !defined(DSC_EMUL)
struct ClusterEntry {
    void (*fn)(const CUtensorMap, const CUtensorMap, const ClusterArgs);
    void (*fn_pipe)(const CUtensorMap, const CUtensorMap, const ClusterArgs, unsigned, unsigned);     // persistent, pipelined
    int lg_n1, lg_n2, l, lp, blocks, box_a, box_b, smem, smem_pipe;
    int state;            // 0 = not configured yet, 1 = usable, -1 = this device cannot co-schedule the cluster
    int state_pipe, clusters_pipe;      // same for the pipelined kernel; resident clusters of its persistent launch
};

template <typename T, bool FWD, int LG_N1, int LG_N2>
ClusterEntry make_cluster() {
    ClusterEntry e;
    e.fn = fft_cluster<T, LG_N1, LG_N2, FWD>;
    e.lg_n1 = LG_N1; e.lg_n2 = LG_N2;
    e.l = tma_tile_points<T>() >> LG_N1;
    e.blocks = (1 << LG_N2) / e.l;
    e.lp = (1 << LG_N1) / e.blocks;
    e.box_a = tma_box_rows(LG_N1); e.box_b = tma_box_rows(LG_N2);
    e.smem = (int)sizeof(ClusterSmem<T, LG_N1>) + 1024;
    e.state = 0;
    e.fn_pipe = fft_cluster_pipe<T, LG_N1, LG_N2, FWD>;
    e.smem_pipe = (int)sizeof(ClusterPipeSmem<T, LG_N1>) + 1024;
    e.state_pipe = 0;
    e.clusters_pipe = 0;
    return e;
}

// complex64 lines of 2^14 (2 blocks), 2^15 (4), 2^16 (8) and 2^17 (16 blocks: non-portable cluster size) points
#define DSC_CLUSTER_PAIRS_float(X) X(7, 7) X(8, 7) X(8, 8) X(9, 8)

template <typename T, bool FWD> ClusterEntry *cluster_entry(int lg_n1, int lg_n2);
template <> ClusterEntry *cluster_entry<float, true>(int, int);
template <> ClusterEntry *cluster_entry<float, false>(int, int);
template <> inline ClusterEntry *cluster_entry<double, true>(int, int) { return nullptr; }
template <> inline ClusterEntry *cluster_entry<double, false>(int, int) { return nullptr; }

#define DSC_DEFINE_CLUSTER(T, FWD)                                                         \
    template <> ClusterEntry *cluster_entry<T, FWD>(int lg_n1, int lg_n2) {                \
        static ClusterEntry table[] = {DSC_CLUSTER_PAIRS_##T(DSC_CLUSTER_MAKE_##FWD##_##T)}; \
        for (auto &e : table) if (e.lg_n1 == lg_n1 && e.lg_n2 == lg_n2) return &e;        \
        return nullptr;                                                                    \
    }
#define DSC_CLUSTER_MAKE_true_float(A, B) make_cluster<float, true, A, B>(),
#define DSC_CLUSTER_MAKE_false_float(A, B) make_cluster<float, false, A, B>(),
#endif

// ---- two-pass transforms along a non-last axis (four_step_columns) --------------------------------------------
struct ColumnsEntry {
    void (*fn)(const FftArgs, const FftArgs, const FourStepSync, const ColumnsGeom);
    int lg_n1, lg_n2, threads, l_a, l_b, smem;
    int grid;
    bool configured;
    // float passes of at most 256 points: 16 points per thread, four 256-thread blocks per SM (like FusedEntry::fn16)
    void (*fn16)(const FftArgs, const FftArgs, const FourStepSync, const ColumnsGeom);
    int l_a16, l_b16, smem16, grid16;
};

template <typename T, bool FWD, int LG_N1, int LG_N2, int THREADS = fused_threads<T>(LG_N1, LG_N2)>
ColumnsEntry make_columns() {
    constexpr int LG_E1 = pass_lg_e<T>(LG_N1, LG_N2), LG_E2 = pass_lg_e<T>(LG_N2, LG_N1);
    ColumnsEntry e;
    e.fn = four_step_columns<T, LG_N1, LG_N2, THREADS, FWD>;
    e.lg_n1 = LG_N1; e.lg_n2 = LG_N2; e.threads = THREADS;
    e.l_a = THREADS >> (LG_N1 - LG_E1); e.l_b = THREADS >> (LG_N2 - LG_E2);
    e.smem = fused_smem_bytes<T, LG_N1, LG_N2, THREADS>();
    e.grid = 0;
    e.configured = false;
    e.fn16 = nullptr; e.l_a16 = e.l_b16 = e.smem16 = e.grid16 = 0;
    if constexpr (sizeof(T) == 4 && LG_N1 <= 8 && LG_N2 <= 8 && LG_N2 >= 7 && THREADS == 256) {
        e.fn16 = four_step_columns<T, LG_N1, LG_N2, THREADS, FWD, 4>;
        e.l_a16 = THREADS >> (LG_N1 - 4); e.l_b16 = THREADS >> (LG_N2 - 4);
        e.smem16 = fused_smem_bytes<T, LG_N1, LG_N2, THREADS, 4>();
    }
    return e;
}

// the fused pairs plus the decompositions of the long single-pass lengths 2^13 and 2^14
#define DSC_COLUMNS_PAIRS(X) X(7, 6) DSC_FUSED_PAIRS(X)

template <typename T, bool FWD> ColumnsEntry *columns_entry(int lg_n1, int lg_n2);
template <> ColumnsEntry *columns_entry<float, true>(int, int);
template <> ColumnsEntry *columns_entry<float, false>(int, int);
template <> ColumnsEntry *columns_entry<double, true>(int, int);
template <> ColumnsEntry *columns_entry<double, false>(int, int);

#define DSC_DEFINE_COLUMNS(T, FWD)                                                         \
    template <> ColumnsEntry *columns_entry<T, FWD>(int lg_n1, int lg_n2) {                \
        static ColumnsEntry table[] = {DSC_COLUMNS_PAIRS(DSC_COLUMNS_MAKE_##FWD##_##T)};   \
        for (auto &e : table) if (e.lg_n1 == lg_n1 && e.lg_n2 == lg_n2) return &e;        \
        return nullptr;                                                                    \
    }
#define DSC_COLUMNS_MAKE_true_float(A, B) make_columns<float, true, A, B>(),
#define DSC_COLUMNS_MAKE_false_float(A, B) make_columns<float, false, A, B>(),
#define DSC_COLUMNS_MAKE_true_double(A, B) make_columns<double, true, A, B>(),
#define DSC_COLUMNS_MAKE_false_double(A, B) make_columns<double, false, A, B>(),

#define DSC_DEFINE_TABLE(T, FWD, MODE, SV)                                                        \
    template <> KernelEntry *get_table<T, FWD, MODE, SV>() {                                      \
        return build_table<T, FWD, MODE, SV>(std::make_integer_sequence<int, Tile<T>::MAX_LG + 1>{}); \
    }

}  // namespace dscfft
