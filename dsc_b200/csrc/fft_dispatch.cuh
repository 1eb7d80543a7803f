// fft_dispatch.cuh -- compile-time tables of fft_lines<> instantiations, indexed by log2(N).
//
// Per precision the register tile is E = 16 (float) / 8 (double) points per thread, so
// butterflies are radix-16 / radix-8.  Two block shapes exist per length:
//   "C" (contiguous lines, last axis): ~256 threads per block, whole lines per warp group;
//   "S" (strided lines, axis != last and both four-step passes): LPB adjacent lines so a
//       warp touches >= 64..128 contiguous bytes per access.
#pragma once

#include <utility>
#include "fft_kernels.cuh"

namespace dscfft {

struct KernelEntry {
    void (*fn)(const FftArgs);
    int lpb;        // lines per block
    int threads;    // block size
    int smem;       // dynamic shared memory bytes
    bool configured;
};

template <typename T> struct Tile;
template <> struct Tile<float>  { static constexpr int LG_E = 4; static constexpr int MAX_LG = 14; static constexpr int COAL = 16; };
template <> struct Tile<double> { static constexpr int LG_E = 3; static constexpr int MAX_LG = 13; static constexpr int COAL = 8; };

template <typename T> constexpr int lg_e_for(int lg_n) { return lg_n < Tile<T>::LG_E ? lg_n : Tile<T>::LG_E; }

constexpr int imin(int a, int b) { return a < b ? a : b; }
constexpr int imax(int a, int b) { return a > b ? a : b; }

// lines per block, contiguous shape: aim for 256 threads
template <typename T> constexpr int lpb_c(int lg_n) {
    const int tt = 1 << (lg_n - lg_e_for<T>(lg_n));
    return imax(1, 256 / tt);
}
// lines per block, strided shape: enough adjacent lines for coalescing, at most 512 threads
template <typename T> constexpr int lpb_s(int lg_n) {
    const int tt = 1 << (lg_n - lg_e_for<T>(lg_n));
    return imax(1, imin(imax(Tile<T>::COAL, 256 / tt), 512 / tt));
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
    e.smem = LPB * Sc::LINE * (int)sizeof(cx<T>);
    e.configured = false;
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
DSC_DECLARE_TABLE(float, true, MODE_FAST, false)  DSC_DECLARE_TABLE(float, false, MODE_FAST, false)
DSC_DECLARE_TABLE(double, true, MODE_FAST, false) DSC_DECLARE_TABLE(double, false, MODE_FAST, false)
#undef DSC_DECLARE_TABLE

#define DSC_DEFINE_TABLE(T, FWD, MODE, SV)                                                        \
    template <> KernelEntry *get_table<T, FWD, MODE, SV>() {                                      \
        return build_table<T, FWD, MODE, SV>(std::make_integer_sequence<int, Tile<T>::MAX_LG + 1>{}); \
    }

}  // namespace dscfft
