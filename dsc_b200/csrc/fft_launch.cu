// fft_launch.cu -- the extern "C" launch layer declared in include/dsc_cuda.h:
// plan construction (radix schedule + twiddle tables in HBM) and the launch logic that
// maps (outer, n, inner) tensors onto fft_lines<> grids, single-pass or four-step.
//
// No cudaMalloc / cudaFree anywhere in this file: plan and work memory belong to the caller.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "dsc_cuda.h"
#include "fft_dispatch.cuh"
#include "pointwise_kernels.cuh"

using namespace dscfft;

namespace {

thread_local char g_err[256] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#if defined(DSC_EMUL)
#define DSC_LAUNCH(fn, grid, block, smem, stream, ...) dsc_emul::launch(grid, block, smem, [=]() { fn(__VA_ARGS__); })
inline int check_launch(const char *) { return 0; }
#else
#define DSC_LAUNCH(fn, grid, block, smem, stream, ...) fn<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__)
inline int check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}
#endif

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int ilog2(long long v) { int l = 0; while ((1LL << l) < v) ++l; return l; }


// ---- plan geometry -------------------------------------------------------------------

struct SubSched {           // stage layout of one shared-memory pass of length 2^lg
    int stages;
    int ns[DSC_CUDA_MAX_STAGES], r[DSC_CUDA_MAX_STAGES];
};

SubSched sub_sched(int lg, int lg_e) {
    SubSched s{};
    s.stages = lg_e == 0 ? 1 : (lg + lg_e - 1) / lg_e;
    for (int i = 0; i < s.stages; ++i) {
        const int rem = lg - i * lg_e;
        s.r[i] = 1 << (rem < lg_e ? rem : lg_e);
        s.ns[i] = 1 << (i * lg_e);
    }
    return s;
}

struct PlanLayout {
    int lg_n, lg_n1, lg_n2, four_shift, real_shift, lg_e1, lg_e2;
    size_t off_tw1[DSC_CUDA_MAX_STAGES], off_tw2[DSC_CUDA_MAX_STAGES];
    size_t off_lo, off_hi, off_real, off_real_lo, off_real_hi, total;
    // column decomposition of long single-pass complex plans (transforms along a non-last axis)
    int col_lg_n1, col_lg_n2, col_shift, col_lg_e1, col_lg_e2;
    size_t off_ctw1[DSC_CUDA_MAX_STAGES], off_ctw2[DSC_CUDA_MAX_STAGES], off_clo, off_chi;
    bool e16;                                   // float fused plans with factors <= 512: tables for 16 points per thread
    size_t off_tw1_e16[DSC_CUDA_MAX_STAGES], off_tw2_e16[DSC_CUDA_MAX_STAGES];
    bool ok;
};

// shortest single-pass length (log2) that also gets a column decomposition
template <typename T> constexpr int columns_min_lg() { return 13; }

template <typename T> PlanLayout plan_layout(int n, int fft_type) {
    PlanLayout L{};
    const size_t es = sizeof(cx<T>);
    L.lg_n = ilog2(n);
    if (L.lg_n <= Tile<T>::MAX_LG) { L.lg_n1 = L.lg_n; L.lg_n2 = 0; }
    else {
        L.lg_n2 = L.lg_n / 2 < 10 ? L.lg_n / 2 : 10;   // rows pass: <= 1024 points
        L.lg_n1 = L.lg_n - L.lg_n2;
    }
    L.ok = L.lg_n1 <= Tile<T>::MAX_LG;
    // tile size of the pass kernels: the fused launch (both factors <= 2^10) has its own choice
    const bool fused = L.lg_n2 != 0 && L.lg_n1 <= 10 && L.lg_n2 <= 10;
    L.lg_e1 = fused ? pass_lg_e<T>(L.lg_n1, L.lg_n2) : lg_e_for<T>(L.lg_n1);
    L.lg_e2 = fused ? pass_lg_e<T>(L.lg_n2, L.lg_n1) : lg_e_for<T>(L.lg_n2);
    size_t off = 0;
    auto take = [&](size_t count) { const size_t o = off; off = align_up(off + count * es, 256); return o; };
    const SubSched s1 = sub_sched(L.lg_n1, L.lg_e1);
    for (int s = 1; s < s1.stages; ++s) L.off_tw1[s] = take((size_t)(s1.r[s] - 1) * s1.ns[s]);
    if (L.lg_n2) {
        const SubSched s2 = sub_sched(L.lg_n2, L.lg_e2);
        for (int s = 1; s < s2.stages; ++s) L.off_tw2[s] = take((size_t)(s2.r[s] - 1) * s2.ns[s]);
        L.four_shift = (L.lg_n + 1) / 2;
        L.off_lo = take((size_t)1 << L.four_shift);
        L.off_hi = take((size_t)1 << (L.lg_n - L.four_shift));
    }
    L.e16 = sizeof(T) == 4 && fused && L.lg_n1 <= 9 && L.lg_n2 <= 9;
    if (L.e16) {
        const SubSched a1 = sub_sched(L.lg_n1, 4), a2 = sub_sched(L.lg_n2, 4);
        for (int s = 1; s < a1.stages; ++s) L.off_tw1_e16[s] = take((size_t)(a1.r[s] - 1) * a1.ns[s]);
        for (int s = 1; s < a2.stages; ++s) L.off_tw2_e16[s] = take((size_t)(a2.r[s] - 1) * a2.ns[s]);
    }
    if (L.lg_n2 == 0 && fft_type == DSC_CUDA_FFT_COMPLEX && L.lg_n >= columns_min_lg<T>()) {
        L.col_lg_n2 = L.lg_n / 2;
        L.col_lg_n1 = L.lg_n - L.col_lg_n2;
        L.col_lg_e1 = pass_lg_e<T>(L.col_lg_n1, L.col_lg_n2);
        L.col_lg_e2 = pass_lg_e<T>(L.col_lg_n2, L.col_lg_n1);
        const SubSched c1 = sub_sched(L.col_lg_n1, L.col_lg_e1), c2 = sub_sched(L.col_lg_n2, L.col_lg_e2);
        for (int s = 1; s < c1.stages; ++s) L.off_ctw1[s] = take((size_t)(c1.r[s] - 1) * c1.ns[s]);
        for (int s = 1; s < c2.stages; ++s) L.off_ctw2[s] = take((size_t)(c2.r[s] - 1) * c2.ns[s]);
        L.col_shift = (L.lg_n + 1) / 2;
        L.off_clo = take((size_t)1 << L.col_shift);
        L.off_chi = take((size_t)1 << (L.lg_n - L.col_shift));
    }
    if (fft_type == DSC_CUDA_FFT_REAL) {
        if (L.lg_n2 == 0) L.off_real = take((size_t)n / 2 + 1);
        else {
            // W_2n^k for k < n/2: index has lg_n - 1 bits
            L.real_shift = L.lg_n / 2;
            L.off_real_lo = take((size_t)1 << L.real_shift);
            L.off_real_hi = take(((size_t)n / 2 >> L.real_shift) + 1);
        }
    }
    L.total = off ? off : 256;
    return L;
}

template <typename T>
int build_tables(dsc_cuda_plan *p, const PlanLayout &L, void *stream) {
    using V = cx<T>;
    char *base = (char *)p->dev_base;
    const int n = p->n;
    auto fill_stage = [&](void *dst, int ns, int r) {
        const int total = (r - 1) * ns;
        DSC_LAUNCH(fill_stage_twiddles<T>, (total + 255) / 256, 256, 0, stream, (V *)dst, ns, r);
    };
    auto fill_pow = [&](void *dst, long long count, long long mult, long long denom) {
        const int blocks = (int)((count + 255) / 256 < 1024 ? (count + 255) / 256 : 1024);
        DSC_LAUNCH(fill_power_twiddles<T>, blocks, 256, 0, stream, (V *)dst, count, mult, denom);
    };
    const SubSched s1 = sub_sched(L.lg_n1, L.lg_e1);
    for (int s = 1; s < s1.stages; ++s) {
        p->tw1[s] = base + L.off_tw1[s];
        fill_stage(p->tw1[s], s1.ns[s], s1.r[s]);
    }
    if (L.lg_n2) {
        const SubSched s2 = sub_sched(L.lg_n2, L.lg_e2);
        for (int s = 1; s < s2.stages; ++s) {
            p->tw2[s] = base + L.off_tw2[s];
            fill_stage(p->tw2[s], s2.ns[s], s2.r[s]);
        }
        p->tw_lo = base + L.off_lo;
        p->tw_hi = base + L.off_hi;
        fill_pow(p->tw_lo, 1LL << L.four_shift, 1, n);
        fill_pow(p->tw_hi, 1LL << (L.lg_n - L.four_shift), 1LL << L.four_shift, n);
    }
    if (L.e16) {
        const SubSched a1 = sub_sched(L.lg_n1, 4), a2 = sub_sched(L.lg_n2, 4);
        for (int s = 1; s < a1.stages; ++s) { p->tw1_e16[s] = base + L.off_tw1_e16[s]; fill_stage(p->tw1_e16[s], a1.ns[s], a1.r[s]); }
        for (int s = 1; s < a2.stages; ++s) { p->tw2_e16[s] = base + L.off_tw2_e16[s]; fill_stage(p->tw2_e16[s], a2.ns[s], a2.r[s]); }
    }
    if (L.lg_n2) {
        // a two-pass plan is its own column decomposition
        p->col_lg_n1 = L.lg_n1; p->col_lg_n2 = L.lg_n2; p->col_shift = L.four_shift;
        for (int s = 0; s < DSC_CUDA_MAX_STAGES; ++s) { p->col_tw1[s] = p->tw1[s]; p->col_tw2[s] = p->tw2[s]; }
        p->col_lo = p->tw_lo; p->col_hi = p->tw_hi;
    } else if (L.col_lg_n2) {
        p->col_lg_n1 = L.col_lg_n1; p->col_lg_n2 = L.col_lg_n2; p->col_shift = L.col_shift;
        const SubSched c1 = sub_sched(L.col_lg_n1, L.col_lg_e1), c2 = sub_sched(L.col_lg_n2, L.col_lg_e2);
        for (int s = 1; s < c1.stages; ++s) { p->col_tw1[s] = base + L.off_ctw1[s]; fill_stage(p->col_tw1[s], c1.ns[s], c1.r[s]); }
        for (int s = 1; s < c2.stages; ++s) { p->col_tw2[s] = base + L.off_ctw2[s]; fill_stage(p->col_tw2[s], c2.ns[s], c2.r[s]); }
        p->col_lo = base + L.off_clo;
        p->col_hi = base + L.off_chi;
        fill_pow(p->col_lo, 1LL << L.col_shift, 1, n);
        fill_pow(p->col_hi, 1LL << (L.lg_n - L.col_shift), 1LL << L.col_shift, n);
    }
    if (p->fft_type == DSC_CUDA_FFT_REAL) {
        if (L.lg_n2 == 0) {
            p->tw_real = base + L.off_real;
            fill_pow(p->tw_real, n / 2 + 1, 1, 2LL * n);
        } else {
            p->tw_real_lo = base + L.off_real_lo;
            p->tw_real_hi = base + L.off_real_hi;
            fill_pow(p->tw_real_lo, 1LL << L.real_shift, 1, 2LL * n);
            fill_pow(p->tw_real_hi, ((long long)n / 2 >> L.real_shift) + 1, 1LL << L.real_shift, 2LL * n);
        }
    }
    return check_launch("twiddle tables");
}

// ---- launches --------------------------------------------------------------------------

inline int pow2_shift(long long v);

int launch_lines(KernelEntry *table, int lg_n, const FftArgs &a_in, void *stream, size_t real_bytes = 0) {
    KernelEntry &e = table[lg_n];
#if !defined(DSC_EMUL)
    if (!e.configured) {
        if (e.smem > 48 * 1024) {
            const cudaError_t err = cudaFuncSetAttribute((const void *)e.fn,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, e.smem);
            if (err != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "smem attribute: %s", cudaGetErrorString(err));
        }
        e.configured = true;
    }
#endif
    if (a_in.lines <= 0) return 0;
    FftArgs a = a_in;
    a.inner_shift = pow2_shift(a.inner);
    // packed real pairs (last-axis rfft / irfft / filter): vector accesses when every line starts aligned
    {
        const size_t vec = 2 * real_bytes;
        a.packed_in = real_bytes && a.gi_pstride == 1 && a.gi.estride == 2 && a.inner == 1 && a.gi.ostride % 2 == 0 &&
                      (uintptr_t)a.x % vec == 0;
        a.packed_out = real_bytes && !a.out_take && a.go_pstride == 1 && a.go.estride == 2 && a.inner == 1 && a.go.ostride % 2 == 0 &&
                       (uintptr_t)a.out % vec == 0;
    }
    if (a.lines % e.lpb != 0) a.no_limit = 0;
    const long long blocks = (a.lines + e.lpb - 1) / e.lpb;
    if (blocks > 0x7fffffffLL) return fail(DSC_CUDA_EINVAL, "too many lines for one grid: %lld", a.lines);
#if !defined(DSC_EMUL)
    // one block per SM (128 KiB lines): persistent blocks with the next line prefetched into L2; DSC_NO_PERSIST=1 disables
    static const bool no_persist = [] { const char *v = getenv("DSC_NO_PERSIST"); return v != nullptr && *v == '1'; }();
    // (measured: float 2^14 4298 -> 5043 GB/s, double 2^13 3987 -> 4668; with two or more blocks per SM the block scheduler
    // already overlaps the lines -- float 2^13 +0 %, double 2^12 -5 % -- so only the one-block-per-SM lengths take it;
    // prefetching two lines ahead instead of one: -2 %)
    if (e.fn_persist != nullptr && !no_persist && a.lines % e.lpb == 0) {
        if (e.grid_persist == 0) {
            cudaError_t err = cudaFuncSetAttribute((const void *)e.fn_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, e.smem);
            int per_sm = 0, dev = 0, sms = 0;
            if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)e.fn_persist, e.threads, e.smem);
            if (err == cudaSuccess) err = cudaGetDevice(&dev);
            if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            e.grid_persist = (err == cudaSuccess && per_sm >= 1) ? per_sm * sms : -1;
            if (err != cudaSuccess) cudaGetLastError();
        }
        if (e.grid_persist > 0 && blocks > e.grid_persist) {
            e.fn_persist<<<(unsigned)e.grid_persist, e.threads, e.smem, (cudaStream_t)stream>>>(a, blocks);
            return check_launch("fft_lines_persist");
        }
    }
#endif
    DSC_LAUNCH(e.fn, (unsigned)blocks, e.threads, e.smem, stream, a);
    return check_launch("fft_lines");
}

template <typename T>
void set_stage_tables(FftArgs &a, void *const tw[DSC_CUDA_MAX_STAGES]) {
    for (int s = 0; s < DSC_CUDA_MAX_STAGES; ++s) a.tw[s] = tw[s];
}

template <typename T, bool FWD>
KernelEntry *c2c_table(bool strided_shape) {
    return strided_shape ? get_table<T, FWD, MODE_C2C, true>() : get_table<T, FWD, MODE_C2C, false>();
}

inline int pow2_shift(long long v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int s = 0;
    while ((1LL << s) < v) ++s;
    return s;
}

// Rows per pass of a multi-kernel pipeline.  Chunks small enough for the intermediate to stay in L2 between
// the kernels (knob DSC_L2_CHUNK_BYTES) were measured SLOWER on B200 than whole-batch launches: the passes are
// not HBM-bound, and short persistent launches lose more in ramp-up and tail than the L2 hits save
// (config 3: 16/32/64 MiB chunks 11.3/7.3/5.8 ms against 5.4 ms unchunked).  Default: no cap.
inline long long l2_chunk_rows(size_t row_bytes) {
    static const size_t budget = [] {
        const char *e = getenv("DSC_L2_CHUNK_BYTES");
        const long long v = e ? atoll(e) : 0;
        return v > 0 ? (size_t)v : (size_t)1 << 40;
    }();
    const long long r = (long long)(budget / (row_bytes ? row_bytes : 1));
    return r > 1 ? r : 1;
}

// shortest complex order (log2) served by the dense packed-real kernels: a line must span at least a warp
template <typename T> constexpr int real_fast_min_lg() { return Tile<T>::LG_E + 5; }

inline size_t in_elem_size(const FftArgs &a, size_t real_size) {
    return (a.in_kind == IN_REAL || a.in_kind == IN_PAIRS) ? real_size : 2 * real_size;
}

// The packed-real bin-pair step fused into a TMA-fed launch (fft_tma.cuh, REAL): forward transforms leave as the bins
// X[0..n] of the real transform (dst rows of pitch >= n + 1) or, with a spectrum, as the packed input z' of the inverse
// transform of the fused filter.
struct RealFuse {
    const void *filt;          // forward: spectrum B[0..n], nullptr = rfft.  Inverse launches (irfft) read rows of n + 1 bins.
};
inline bool real_fuse_disabled() {
    static const bool off = [] { const char *e = getenv("DSC_NO_REAL_FUSE"); return e != nullptr && *e != '\0' && *e != '0'; }();
    return off;
}
// float32 filter: the fused forward launch is parity-green but measures 3 - 5 % SLOWER than transform + bin-pair sweep on
// B200 (1.91 - 1.99 ms against 1.83 - 1.87 ms per GiB at 2^16 .. 2^20 samples): complex64 runs start 8 bytes off the 16-byte
// granule of a bulk tensor store, so half of every tile leaves through the group's own stores, and the pair step (two
// un-mix/mix steps, two spectrum bins per pair) doubles the instructions of a second-pass tile in a launch that is paced by
// its instruction stream.  float64 gains 33 % (rfft) / 65 % (irfft).  Opt-in for float32: DSC_REAL_FUSE_F32=1 (tests do).
inline bool real_fuse_f32() {
    static const bool on = [] { const char *e = getenv("DSC_REAL_FUSE_F32"); return e != nullptr && *e == '1'; }();
    return on;
}


#if !defined(DSC_EMUL)
// ---- TMA-fed four-step (fft_tma.cuh): tensor maps and the launch -------------------------------------------------
// cuTensorMapEncodeTiled comes from the driver through the runtime's entry-point query, so the library keeps
// linking against libcudart only.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// [rows][outer][inner] tensor of elem_bytes-sized elements (8 or 16), described in 8-byte units; box = box_outer x box_inner
bool encode_3d(CUtensorMap *map, const void *base, size_t elem_bytes, unsigned long long inner, unsigned long long outer,
               unsigned long long rows, unsigned long long outer_stride_elems, unsigned long long row_stride_elems,
               unsigned box_inner, unsigned box_outer) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return false;
    const unsigned es = (unsigned)(elem_bytes / 8);
    const cuuint64_t dims[3] = {inner * es, outer, rows};
    const cuuint64_t strides[2] = {outer_stride_elems * elem_bytes, row_stride_elems * elem_bytes};
    const cuuint32_t box[3] = {box_inner * es, box_outer, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool tma_disabled() {
    static const bool off = [] { const char *e = getenv("DSC_NO_TMA"); return e != nullptr && *e != '\0' && *e != '0'; }();
    return off;
}

// Returns 1 when the shape is not covered (the caller continues with four_step_fused), 0 on success, < 0 on error.
template <typename T, bool FWD>
int four_step_tma_launch(const dsc_cuda_plan *p, const FftArgs &first, long long rows, void *work, size_t work_bytes,
                         void *dst, long long dst_row_stride, bool scale, void *stream, bool keep_out,
                         const RealFuse *rf = nullptr) {
    using V = cx<T>;
    const long long n = p->n, n1 = 1LL << p->lg_n1, n2 = 1LL << p->lg_n2;
    TmaEntry *te = tma_entry<T, FWD>(p->lg_n1, p->lg_n2);
    if (te == nullptr || tma_disabled()) return 1;
    if (rf != nullptr && (te->fn_real == nullptr || real_fuse_disabled() || p->tw_real_lo == nullptr ||
                          (sizeof(T) == 4 && !real_fuse_f32()) ||
                          p->real_shift > p->lg_n2 || p->real_shift > p->lg_n1 ||
                          (FWD ? dst_row_stride < n + (rf->filt ? 0 : 1) : (first.in_limit < n + 1 || first.gi.ostride < n + 1))))
        return 1;
    // 16 points per thread on 32 KiB tiles, two blocks per SM (twice the butterfly warps), where the plan carries its tables.
    // With one storer warp per buffer the launch is no longer paced by its copy pipeline and the extra warps pay where the
    // half-size tiles still have 128-byte rows: passes of at most 256 points (2^15: 3338 -> 3586 GB/s, 2^16: 3340 -> 3534).
    // 512-point passes would have 64-byte rows (2^17 / 2^18: 2675 / 2564 against 3386 / 3349) and keep 32 points per thread.
    // DSC_TMA_E16=0 / 1 forces it off / on wherever the tables exist (the GPU parity tests run both).
    static const int e16_env = [] { const char *e = getenv("DSC_TMA_E16"); return e == nullptr || *e == '\0' ? -1 : atoi(e); }();
    const bool want_e16 = e16_env < 0 ? (p->lg_n1 <= 8 && p->lg_n2 <= 8) : e16_env != 0;
    bool e16 = te->fn16 != nullptr && p->tw1_e16[1] != nullptr && p->tw2_e16[1] != nullptr && want_e16 && rf == nullptr;
    // dense complex rows, read in full, 16-byte aligned rows on both sides
    if (first.in_kind != IN_COMPLEX || first.in_limit < n || first.seg_shift != 0 || first.gi.lstride != 1 || first.gi.estride != n2 ||
        first.ring_in != 0 || (uintptr_t)first.x % 16 != 0 || (uintptr_t)dst % 16 != 0 ||
        ((size_t)first.gi.ostride * sizeof(V)) % 16 != 0 || ((size_t)dst_row_stride * sizeof(V)) % 16 != 0 ||
        first.gi.ostride < n || dst_row_stride < n || rows >= 0x7fffffffLL)
        return 1;
    const size_t row_bytes = (size_t)n * sizeof(V);
    const size_t sync_bytes = align_up((size_t)(1 + 2 * rows) * sizeof(unsigned), 256);
    if (work == nullptr || work_bytes < sync_bytes + row_bytes || (uintptr_t)work % 256 != 0) return 1;
    if (!te->configured) {
        cudaError_t err = cudaFuncSetAttribute((const void *)te->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, te->smem);
        if (err == cudaSuccess && te->fn_real != nullptr)
            err = cudaFuncSetAttribute((const void *)te->fn_real, cudaFuncAttributeMaxDynamicSharedMemorySize, te->smem);
        if (err == cudaSuccess && te->fn_direct != nullptr)
            err = cudaFuncSetAttribute((const void *)te->fn_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, te->smem);
        if (err == cudaSuccess && te->fn16 != nullptr) {
            err = cudaFuncSetAttribute((const void *)te->fn16, cudaFuncAttributeMaxDynamicSharedMemorySize, te->smem16);
            int per_sm = 0;
            if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)te->fn16, TMA4_THREADS, te->smem16);
            te->ctas16 = per_sm;
        }
        int dev = 0, sms = 0;
        if (err == cudaSuccess) err = cudaGetDevice(&dev);
        if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) { cudaGetLastError(); return 1; }
        te->grid = sms;                       // one block (two butterfly groups + the producer warp) per SM
        te->configured = true;
    }
    if (e16 && te->ctas16 < 1) e16 = false;
    const int l_a = e16 ? te->l_a / 2 : te->l_a, l_b = e16 ? te->l_b / 2 : te->l_b;
    const long long tiles_a = n2 / l_a, tiles_b = n1 / l_b;
    if (rows * (tiles_a + tiles_b) >= 0x7fffffffLL) return 1;
    const long long resident = e16 ? (long long)te->ctas16 * te->grid : te->grid;      // blocks on the device at once
    long long ring = (long long)((work_bytes - sync_bytes) / row_bytes);
    static const size_t ring_budget = [] { const char *e = getenv("DSC_TMA_RING_MB"); return (size_t)(e && atoi(e) > 0 ? atoi(e) : 64) << 20; }();
    const long long cap = (long long)(ring_budget / row_bytes) > 4 ? (long long)(ring_budget / row_bytes) : 4;
    if (ring > cap) ring = cap;
    if (ring >= rows) ring = 0;
    FourStepSync s{};
    s.ticket = (unsigned *)work;
    s.a_done = s.ticket + 1;
    s.b_done = s.a_done + rows;
    s.tiles_a = (int)tiles_a;
    s.tiles_b = (int)tiles_b;
    s.ring = (int)ring;
    s.rows = (int)rows;
    {
        // a tile is published about five tile times after its ticket was taken (two of them waiting in the loader), while the whole GPU takes
        // ~grid tickets per tile time: put a row's second pass that far behind its first pass
        long long lag = (5LL * resident + tiles_b + tiles_a + tiles_b - 1) / (tiles_a + tiles_b);
        static const long long lag_env = [] { const char *e = getenv("DSC_TMA_LAG"); return e ? atoll(e) : 0LL; }();
        if (lag_env > 0) lag = lag_env;
        if (lag < 1) lag = 1;
        if (ring > 0 && lag > ring / 2) lag = ring / 2;
        if (lag > rows) lag = rows;
        s.lag = (int)lag;
    }
    V *mid = (V *)((char *)work + sync_bytes);
    CUtensorMap map_x, map_w, map_out;
    const unsigned long long wrows = (unsigned long long)(ring ? ring : rows);
    const bool mix_in = rf != nullptr && !FWD;       // bin rows: n2 + 1 columns at a row pitch of n2 (the last one aliases the next row)
    const bool unmix_out = rf != nullptr && FWD;
    if (!encode_3d(&map_x, first.x, sizeof(V), (unsigned long long)(mix_in ? n2 + 1 : n2), (unsigned long long)n1, (unsigned long long)rows,
                   (unsigned long long)n2, (unsigned long long)first.gi.ostride, (unsigned)(mix_in ? l_a / 2 : l_a), (unsigned)te->box_a) ||
        !encode_3d(&map_w, mid, sizeof(V), (unsigned long long)n1, (unsigned long long)n2, wrows,
                   (unsigned long long)n1, (unsigned long long)n, (unsigned)(unmix_out ? l_b / 2 : l_b), (unsigned)te->box_b) ||
        !encode_3d(&map_out, dst, sizeof(V), (unsigned long long)n1, (unsigned long long)n2, (unsigned long long)rows,
                   (unsigned long long)n1, (unsigned long long)dst_row_stride, (unsigned)(unmix_out ? l_b / 2 : l_b), (unsigned)te->box_b))
        return 1;
    TmaArgs a{};
    a.work = mid;
    a.ring = ring;
    for (int i = 0; i < DSC_CUDA_MAX_STAGES; ++i) {
        a.tw_a[i] = e16 ? p->tw1_e16[i] : p->tw1[i];
        a.tw_b[i] = e16 ? p->tw2_e16[i] : p->tw2[i];
    }
    a.tw_lo = p->tw_lo; a.tw_hi = p->tw_hi;
    a.four_shift = p->four_shift; a.four_mask = (1 << p->four_shift) - 1;
    a.do_scale = scale; a.scale = 1.0 / (double)n;
    a.keep_out = keep_out;
    static const bool no_discard = [] { const char *e = getenv("DSC_TMA_NO_DISCARD"); return e != nullptr && *e == '1'; }();
    a.discard_work = !no_discard;
    static const bool no_prefetch = [] { const char *e = getenv("DSC_TMA_PREFETCH"); return e != nullptr && *e == '0'; }();
    a.prefetch = !no_prefetch;
    a.out = dst; a.out_pitch = dst_row_stride;
    if (rf != nullptr) {
        a.twr_lo = p->tw_real_lo; a.twr_hi = p->tw_real_hi;
        a.real_shift = p->real_shift; a.real_mask = (1 << p->real_shift) - 1;
        a.filt = rf->filt;
        a.in = first.x; a.in_pitch = first.gi.ostride;
    }
#if defined(DSC_TMA_EXPERIMENTS)
    static const int debug_skip = [] { const char *e = getenv("DSC_TMA_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
    a.debug_skip = debug_skip;
#endif
    static const bool want_direct = [] { const char *e = getenv("DSC_TMA_DIRECT"); return e != nullptr && *e == '1'; }();
    const bool direct = want_direct && rf == nullptr && !e16 && te->fn_direct != nullptr;
    const cudaError_t me = cudaMemsetAsync(work, 0, sync_bytes, (cudaStream_t)stream);
    if (me != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "memset: %s", cudaGetErrorString(me));
    const long long tiles = rows * (tiles_a + tiles_b);
    const unsigned blocks = (unsigned)(tiles < resident ? tiles : resident);
    if (rf != nullptr) te->fn_real<<<blocks, TMA4_THREADS, te->smem, (cudaStream_t)stream>>>(map_x, map_w, map_out, a, s);
    else if (direct) te->fn_direct<<<blocks, TMA4_THREADS, te->smem, (cudaStream_t)stream>>>(map_x, map_w, map_out, a, s);
    else if (e16) te->fn16<<<blocks, TMA4_THREADS, te->smem16, (cudaStream_t)stream>>>(map_x, map_w, map_out, a, s);
    else te->fn<<<blocks, TMA4_THREADS, te->smem, (cudaStream_t)stream>>>(map_x, map_w, map_out, a, s);
    return check_launch("four_step_tma");
}
// Which line lengths (log2) go to the cluster kernels.  Measured on B200 (profiles/r2_two_pass_lengths.md): they read and
// write exactly the algorithmic bytes, but the all-to-all between the blocks of a cluster couples every line to the
// slowest of C SMs twice, and clusters of 8 leave 19 % of the SMs without a block: pipelined variant 3.28 TB/s at 2^15
// (TMA-fed two-pass launch: 3.0), 2.46 at 2^16 (2.98), 1.97 at 2^17 (2.80), 3.85 at 2^14 (single-pass block: 4.30).
// Default: 2^15 only.  DSC_CLUSTER_LGS=14,15,16,17 selects others (tests do), DSC_NO_CLUSTER=1 none; DSC_CLUSTER_PIPE=0
// selects the one-line-per-cluster launch instead of the persistent pipelined one.
// No DSC_CLUSTER_LGS / DSC_NO_CLUSTER in the environment: 2^15-point batches are decided per device by measurement.
inline bool cluster_autotuned() {
    static const bool on = [] {
        const char *off = getenv("DSC_NO_CLUSTER"), *e = getenv("DSC_CLUSTER_LGS");
        return !(off != nullptr && *off != '\0' && *off != '0') && (e == nullptr || *e == '\0');
    }();
    return on;
}
inline bool cluster_wanted(const int lg_n) {
    static const unsigned mask = [] {
        const char *off = getenv("DSC_NO_CLUSTER");
        if (off != nullptr && *off != '\0' && *off != '0') return 0u;
        const char *e = getenv("DSC_CLUSTER_LGS");
        if (e == nullptr || *e == '\0') return 1u << 15;
        unsigned m = 0;
        for (const char *c = e; *c;) {
            const int v = atoi(c);
            if (v > 0 && v < 32) m |= 1u << v;
            while (*c && *c != ',') ++c;
            if (*c == ',') ++c;
        }
        return m;
    }();
    return (mask >> lg_n) & 1u;
}

// One line per thread-block cluster (fft_cluster.cuh): dense complex lines of 2^14 .. 2^17 points, one pass over HBM.
// Returns 1 when the shape / device is not covered, 0 on success, < 0 on error.
template <typename T, bool FWD>
int cluster_launch(const dsc_cuda_plan *p, const void *x, long long x_row_stride, void *dst, long long dst_row_stride,
                   long long rows, bool scale, void *stream) {
    using V = cx<T>;
    if (p->col_lg_n2 == 0 || !cluster_wanted(p->lg_n)) return 1;
    ClusterEntry *ce = cluster_entry<T, FWD>(p->col_lg_n1, p->col_lg_n2);
    if (ce == nullptr || ce->state < 0) return 1;
    const long long n = p->n, n1 = 1LL << p->col_lg_n1, n2 = 1LL << p->col_lg_n2;
    if ((uintptr_t)x % 16 != 0 || (uintptr_t)dst % 16 != 0 || ((size_t)x_row_stride * sizeof(V)) % 16 != 0 ||
        ((size_t)dst_row_stride * sizeof(V)) % 16 != 0 || x_row_stride < n || dst_row_stride < n || rows <= 0 ||
        rows * ce->blocks >= 0x7fffffffLL)
        return 1;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ce->blocks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(rows * ce->blocks), 1, 1);
    cfg.blockDim = dim3(CLUSTER_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)ce->smem;
    cfg.stream = (cudaStream_t)stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (ce->state == 0) {
        cudaError_t err = cudaFuncSetAttribute((const void *)ce->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ce->smem);
        if (err == cudaSuccess && ce->blocks > 8)
            err = cudaFuncSetAttribute((const void *)ce->fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int clusters = 0;
        if (err == cudaSuccess) err = cudaOccupancyMaxActiveClusters(&clusters, (const void *)ce->fn, &cfg);
        if (err != cudaSuccess || clusters < 1) { cudaGetLastError(); ce->state = -1; return 1; }
        ce->state = 1;
    }
    CUtensorMap map_x, map_out;
    if (!encode_3d(&map_x, x, sizeof(V), (unsigned long long)n2, (unsigned long long)n1, (unsigned long long)rows,
                   (unsigned long long)n2, (unsigned long long)x_row_stride, (unsigned)ce->l, (unsigned)ce->box_a) ||
        !encode_3d(&map_out, dst, sizeof(V), (unsigned long long)n1, (unsigned long long)n2, (unsigned long long)rows,
                   (unsigned long long)n1, (unsigned long long)dst_row_stride, (unsigned)ce->lp, (unsigned)ce->box_b))
        return 1;
    ClusterArgs a{};
    for (int i = 0; i < DSC_CUDA_MAX_STAGES; ++i) { a.tw_a[i] = p->col_tw1[i]; a.tw_b[i] = p->col_tw2[i]; }
    a.tw_lo = p->col_lo; a.tw_hi = p->col_hi;
    a.four_shift = p->col_shift; a.four_mask = (1 << p->col_shift) - 1;
    a.do_scale = scale; a.scale = 1.0 / (double)n;
    // the pipelined variant: persistent clusters, three tile buffers per block, mbarrier-signalled DSMEM stores
    static const bool want_pipe = [] { const char *e = getenv("DSC_CLUSTER_PIPE"); return e == nullptr || *e == '\0' || *e != '0'; }();
    if (want_pipe && ce->state_pipe >= 0 && rows < 0x7fffffffLL) {
        cudaLaunchConfig_t pc = cfg;
        pc.blockDim = dim3(TMA_THREADS, 1, 1);
        pc.dynamicSmemBytes = (size_t)ce->smem_pipe;
        if (ce->state_pipe == 0) {
            cudaError_t err = cudaFuncSetAttribute((const void *)ce->fn_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, ce->smem_pipe);
            if (err == cudaSuccess && ce->blocks > 8)
                err = cudaFuncSetAttribute((const void *)ce->fn_pipe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            int clusters = 0;
            pc.gridDim = dim3((unsigned)(ce->blocks * 1024), 1, 1);
            if (err == cudaSuccess) err = cudaOccupancyMaxActiveClusters(&clusters, (const void *)ce->fn_pipe, &pc);
            if (err != cudaSuccess || clusters < 1) { cudaGetLastError(); ce->state_pipe = -1; }
            else { ce->state_pipe = 1; ce->clusters_pipe = clusters; }
        }
        if (ce->state_pipe == 1) {
            const unsigned clusters = (unsigned)(rows < ce->clusters_pipe ? rows : ce->clusters_pipe);
            pc.gridDim = dim3(clusters * (unsigned)ce->blocks, 1, 1);
            const cudaError_t le = cudaLaunchKernelEx(&pc, ce->fn_pipe, map_x, map_out, a, (unsigned)rows, clusters);
            if (le != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "fft_cluster_pipe: %s", cudaGetErrorString(le));
            return check_launch("fft_cluster_pipe");
        }
    }
    const cudaError_t le = cudaLaunchKernelEx(&cfg, ce->fn, map_x, map_out, a);
    if (le != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "fft_cluster: %s", cudaGetErrorString(le));
    return check_launch("fft_cluster");
}
#endif

// The two passes of the four-step decomposition n = n1*n2 of `rows` lines:
//   A: for every n2, length-n1 transform over stride-n2 data, times W_n^(n2 k1) -> work[row][k1][n2]
//   B: for every k1, length-n2 transform of the contiguous run work[row][k1][:] -> dst[row][k1 + n1 k2]
// `first` carries the source geometry of pass A (element kind, row stride = gi.ostride, limit).
// Preferred: ONE fused launch with a ring of work rows that stays in L2 (four_step_fused).  Shapes the
// fused kernel does not cover fall back to two launches per chunk of rows.
template <typename T, bool FWD>
int four_step(const dsc_cuda_plan *p, FftArgs first, long long rows, void *work, size_t work_bytes,
              void *dst, long long dst_row_stride, bool scale, void *stream, bool keep_out = false,
              const RealFuse *rf = nullptr) {
    using V = cx<T>;
    const long long n = p->n, n1 = 1LL << p->lg_n1, n2 = 1LL << p->lg_n2;
    const size_t row_bytes = (size_t)n * sizeof(V);
    if (rows <= 0) return 0;
    // adjacent real pairs in full, aligned rows ARE dense complex rows: take the bandwidth path
    if (first.in_kind == IN_PAIRS && first.gi_pstride == 1 && first.gi.lstride == 2 && first.gi.estride == 2 * n2 &&
        first.gi.ostride % 2 == 0 && first.in_limit >= 2 * n && (uintptr_t)first.x % sizeof(V) == 0) {
        first.in_kind = IN_COMPLEX;
        first.gi = LineGeom{first.gi.ostride / 2, 1, n2};
        first.in_limit = n;
    }

#if !defined(DSC_EMUL)
    // with the packed-real bin-pair step fused in: the TMA-fed launch or nothing (1 = not covered, the caller sweeps)
    if (rf != nullptr) return four_step_tma_launch<T, FWD>(p, first, rows, work, work_bytes, dst, dst_row_stride, scale, stream, keep_out, rf);
#else
    if (rf != nullptr) return 1;
#endif
#if !defined(DSC_EMUL)
    {
        // dense complex rows up to 2^17 points: one line per thread-block cluster, the transpose through distributed
        // shared memory (fft_cluster.cuh); beyond that the TMA-fed two-pass launch (fft_tma.cuh)
        if (first.in_kind == IN_COMPLEX && first.in_limit >= n && first.seg_shift == 0 && first.gi.lstride == 1 &&
            first.gi.estride == n2 && first.ring_in == 0) {
            // 2^15 points: which launch wins depends on the part -- how many clusters its GPCs co-schedule.  Measured on four
            // B200s: pipelined clusters 3.27 / 3.28 / 2.84 / 2.87 TB/s, the TMA-fed two-pass launch 3.08 on each.  The first
            // batch large enough to time (>= 2^24 points, out of place, not under stream capture) runs both twice -- repeating an
            // out-of-place transform is harmless -- and the faster one serves the process from then on.
            static int choice = 0;                  // per (T, FWD): 0 undecided, 1 clusters, 2 two-pass
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (choice == 0 && cluster_autotuned() && cluster_wanted(p->lg_n) && rows * n >= (1LL << 24) && first.x != dst &&
                cudaStreamIsCapturing((cudaStream_t)stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) {
                cudaEvent_t ev[5];
                for (auto &e : ev) cudaEventCreate(&e);
                int rc_c = 0, rc_t = 0;
                cudaEventRecord(ev[0], (cudaStream_t)stream);
                for (int rep = 0; rep < 2 && rc_c == 0 && rc_t == 0; ++rep) {
                    rc_c = cluster_launch<T, FWD>(p, first.x, first.gi.ostride, dst, dst_row_stride, rows, scale, stream);
                    cudaEventRecord(ev[1 + 2 * rep], (cudaStream_t)stream);
                    if (rc_c == 0) rc_t = four_step_tma_launch<T, FWD>(p, first, rows, work, work_bytes, dst, dst_row_stride, scale, stream, keep_out);
                    cudaEventRecord(ev[2 + 2 * rep], (cudaStream_t)stream);
                }
                float t_c = 0.f, t_t = 0.f;
                if (rc_c == 0 && rc_t == 0 && cudaEventSynchronize(ev[4]) == cudaSuccess &&
                    cudaEventElapsedTime(&t_c, ev[2], ev[3]) == cudaSuccess && cudaEventElapsedTime(&t_t, ev[3], ev[4]) == cudaSuccess)
                    choice = t_c <= t_t ? 1 : 2;
                for (auto &e : ev) cudaEventDestroy(e);
                if (rc_c < 0 || rc_t < 0) return rc_c < 0 ? rc_c : rc_t;
                if (choice != 0) return 0;          // dst holds the transform (written by the last launch)
                cudaGetLastError();
                choice = rc_c > 0 ? 2 : 1;          // one of them does not cover the shape: nothing to decide
            }
            if (!(choice == 2 && cluster_autotuned())) {
                const int rcc = cluster_launch<T, FWD>(p, first.x, first.gi.ostride, dst, dst_row_stride, rows, scale, stream);
                if (rcc <= 0) return rcc;
            }
        }
        const int rc = four_step_tma_launch<T, FWD>(p, first, rows, work, work_bytes, dst, dst_row_stride, scale, stream, keep_out);
        if (rc <= 0) return rc;
    }
#endif

    FftArgs a = first;
    a.inner = n2;
    a.go = LineGeom{n, 1, n2};
    set_stage_tables<T>(a, p->tw1);
    a.tw_lo = p->tw_lo; a.tw_hi = p->tw_hi;
    a.four_shift = p->four_shift; a.four_mask = (1 << p->four_shift) - 1;
    a.strided = 1;
    a.do_scale = 0;

    FftArgs b{};
    b.inner = n1;
    b.gi = LineGeom{n, n2, 1};
    b.go = LineGeom{dst_row_stride, 1, n1};
    b.in_limit = 0;
    b.in_kind = IN_ROWS;
    set_stage_tables<T>(b, p->tw2);
    b.strided = 1;
    b.do_scale = scale; b.scale = 1.0 / (double)n;
    b.keep_out = keep_out;

    FusedEntry *fe = fused_entry<T, FWD>(p->lg_n1, p->lg_n2);
    const size_t sync_bytes = align_up((size_t)(1 + 2 * rows) * sizeof(unsigned), 256);
    // 16 points per thread (64 registers, four 256-thread blocks per SM instead of two): dense complex rows, whole or in
    // segments of at least one thread step.  Measured on B200 against 32 points per thread: 2698 -> 3156 GB/s at 2^15 and
    // 2825 -> 3174 at 2^16, but 2780 -> 2715 at 2^17 and 2791 -> 2488 at 2^18 (a 512-point pass then needs a second exchange):
    // default for passes of at most 256 points; DSC_FUSED_E16=0 / 1 forces it off / on wherever the tables exist.
    static const int f16_env = [] { const char *e = getenv("DSC_FUSED_E16"); return e == nullptr || *e == '\0' ? -1 : atoi(e); }();
    const bool want_f16 = f16_env < 0 ? (p->lg_n1 <= 8 && p->lg_n2 <= 8) : f16_env != 0;
    if (want_f16 && fe != nullptr && fe->fn16 != nullptr && p->tw1_e16[1] != nullptr && p->tw2_e16[1] != nullptr &&
        first.in_kind == IN_COMPLEX && first.in_limit >= n && (first.seg_shift == 0 || first.seg_shift >= p->lg_n - 4) &&
        rows * (n2 / fe->lpb_a16 + n1 / fe->lpb_b16) < 0x7fffffffLL && work != nullptr && work_bytes >= sync_bytes + row_bytes) {
        set_stage_tables<T>(a, p->tw1_e16);
        set_stage_tables<T>(b, p->tw2_e16);
        const size_t l2_budget = 64u << 20;
        long long ring = (long long)((work_bytes - sync_bytes) / row_bytes);
        const long long cap = (long long)(l2_budget / row_bytes) > 4 ? (long long)(l2_budget / row_bytes) : 4;
        if (ring > cap) ring = cap;
        if (ring >= rows) ring = 0;
#if defined(DSC_EMUL)
        fe->grid16 = 3;
#else
        if (fe->grid16 == 0) {
            if (fe->smem16 > 48 * 1024) {
                const cudaError_t err = cudaFuncSetAttribute((const void *)fe->fn16, cudaFuncAttributeMaxDynamicSharedMemorySize, fe->smem16);
                if (err != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "smem attribute: %s", cudaGetErrorString(err));
            }
            int per_sm = 0, dev = 0, sms = 0;
            cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fe->fn16, fe->threads, fe->smem16);
            if (err == cudaSuccess) err = cudaGetDevice(&dev);
            if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (err != cudaSuccess || per_sm < 1) return fail(DSC_CUDA_ELAUNCH, "occupancy query: %s", cudaGetErrorString(err));
            fe->grid16 = per_sm * sms;
        }
#endif
        FourStepSync s{};
        s.ticket = (unsigned *)work;
        s.a_done = s.ticket + 1;
        s.b_done = s.a_done + rows;
        s.tiles_a = (int)(n2 / fe->lpb_a16);
        s.tiles_b = (int)(n1 / fe->lpb_b16);
        s.ring = (int)ring;
        s.rows = (int)rows;
        {
            long long lag = (3LL * fe->grid16 + s.tiles_a + s.tiles_b - 1) / (s.tiles_a + s.tiles_b);
            if (lag < 1) lag = 1;
            if (ring > 0 && lag > ring / 2) lag = ring / 2;
            if (lag > rows) lag = rows;
            s.lag = (int)lag;
        }
        V *mid = (V *)((char *)work + sync_bytes);
        a.out = mid; a.lines = rows * n2; a.ring_out = ring; a.inner_shift = p->lg_n2;
        a.no_limit = 1;
        b.x = mid; b.out = dst; b.lines = rows * n1; b.ring_in = ring; b.inner_shift = p->lg_n1;
#if defined(DSC_EMUL)
        memset(work, 0, sync_bytes);
#else
        const cudaError_t me = cudaMemsetAsync(work, 0, sync_bytes, (cudaStream_t)stream);
        if (me != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "memset: %s", cudaGetErrorString(me));
#endif
        const long long tiles = rows * (s.tiles_a + s.tiles_b);
        const unsigned blocks = (unsigned)(tiles < fe->grid16 ? tiles : fe->grid16);
        DSC_LAUNCH(fe->fn16, blocks, fe->threads, fe->smem16, stream, a, b, s);
        return check_launch("four_step_fused e16");
    }
    if (fe != nullptr && rows * (n2 / fe->lpb_a + n1 / fe->lpb_b) < 0x7fffffffLL &&
        work != nullptr && work_bytes >= sync_bytes + row_bytes) {
        // ring of work rows: as many as fit, but no more than keeps the intermediate inside L2
        const size_t l2_budget = 64u << 20;
        long long ring = (long long)((work_bytes - sync_bytes) / row_bytes);
        const long long cap = (long long)(l2_budget / row_bytes) > 4 ? (long long)(l2_budget / row_bytes) : 4;
        if (ring > cap) ring = cap;
        if (ring >= rows) ring = 0;
#if defined(DSC_EMUL)
        fe->grid = 3;
#else
        if (!fe->configured) {
            if (fe->smem > 48 * 1024) {
                const cudaError_t err = cudaFuncSetAttribute((const void *)fe->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, fe->smem);
                if (err != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "smem attribute: %s", cudaGetErrorString(err));
            }
            // persistent grid: every block that can be resident at once
            int per_sm = 0, dev = 0, sms = 0;
            cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fe->fn, fe->threads, fe->smem);
            if (err == cudaSuccess) err = cudaGetDevice(&dev);
            if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (err != cudaSuccess || per_sm < 1) return fail(DSC_CUDA_ELAUNCH, "occupancy query: %s", cudaGetErrorString(err));
            fe->grid = per_sm * sms;
            fe->configured = true;
        }
#endif
        FourStepSync s{};
        s.ticket = (unsigned *)work;
        s.a_done = s.ticket + 1;
        s.b_done = s.a_done + rows;
        s.tiles_a = (int)(n2 / fe->lpb_a);
        s.tiles_b = (int)(n1 / fe->lpb_b);
        s.ring = (int)ring;
        s.rows = (int)rows;
        // lag (in rows) between a row's first and second pass in ticket order: about three grids' worth of
        // resident blocks (a tile is released half a tile after it ends and looked up a tile before it is
        // needed), so second-pass blocks start on rows that are already complete; below the ring
        {
            const long long resident = fe->grid;
            long long lag = (3 * resident + s.tiles_a + s.tiles_b - 1) / (s.tiles_a + s.tiles_b);
            if (lag < 1) lag = 1;
            // ... and at most half the ring, so a first-pass block that reuses a work row finds the
            // second pass of its previous owner long finished instead of spinning on it
            if (ring > 0 && lag > ring / 2) lag = ring / 2;
            if (lag > rows) lag = rows;
            s.lag = (int)lag;
        }
        V *mid = (V *)((char *)work + sync_bytes);
        a.out = mid; a.lines = rows * n2; a.ring_out = ring; a.inner_shift = p->lg_n2;
        a.no_limit = first.in_limit >= n;        // whole blocks by construction (n2 % lpb_a == 0)
        b.x = mid; b.out = dst; b.lines = rows * n1; b.ring_in = ring; b.inner_shift = p->lg_n1;
#if defined(DSC_EMUL)
        memset(work, 0, sync_bytes);
#else
        const cudaError_t me = cudaMemsetAsync(work, 0, sync_bytes, (cudaStream_t)stream);
        if (me != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "memset: %s", cudaGetErrorString(me));
#endif
        const long long tiles = rows * (s.tiles_a + s.tiles_b);
        const unsigned blocks = (unsigned)(tiles < fe->grid ? tiles : fe->grid);
        DSC_LAUNCH(fe->fn, blocks, fe->threads, fe->smem, stream, a, b, s);
        return check_launch("four_step_fused");
    }

    if (fe != nullptr)
        return fail(DSC_CUDA_ENOMEM, "work buffer too small for the fused four-step launch (n=%lld needs %zu bytes)", n, sync_bytes + row_bytes);
    // two launches per chunk of rows
    if (!work || work_bytes < row_bytes) return fail(DSC_CUDA_ENOMEM, "work buffer holds no line (n=%lld)", n);
    const long long chunk = (long long)(work_bytes / row_bytes);
    const size_t in_es = in_elem_size(first, sizeof(T));
    for (long long r0 = 0; r0 < rows; r0 += chunk) {
        const long long nr = rows - r0 < chunk ? rows - r0 : chunk;
        a.x = (const char *)first.x + (size_t)r0 * (size_t)first.gi.ostride * in_es;
        a.out = work;
        a.lines = nr * n2;
        int rc = launch_lines(c2c_table<T, FWD>(true), p->lg_n1, a, stream);
        if (rc) return rc;
        b.x = work;
        b.out = (V *)dst + (size_t)r0 * (size_t)dst_row_stride;
        b.lines = nr * n1;
        rc = launch_lines(c2c_table<T, FWD>(true), p->lg_n2, b, stream);
        if (rc) return rc;
    }
    return 0;
}

// chunk width (log2) of the column launch: rows of n x Ic points of about 4 MiB, at least one tile wide
inline int columns_chunk_lg(long long n, int lg_inner, int l_max, size_t elem_bytes) {
    int lg_ic = lg_inner;
    while ((1LL << lg_ic) > l_max && ((size_t)n << lg_ic) * elem_bytes > ((size_t)4 << 20)) --lg_ic;
    return lg_ic;
}

// Two-pass transform along a non-last axis of a contiguous (outer, x_n, inner) tensor, inner a power of two wide
// enough for a tile: ONE persistent launch whose two passes are both column passes (four_step_columns).
// Returns DSC_CUDA_EUNSUPPORTED when the shape is not covered (the tensor layer then composes transposes).
struct PostTwiddle {            // optional output twiddle of the column launch (ColumnsGeom::post_twiddle)
    const void *lo, *hi;
    int shift;
    long long total, col_offset;
    int n_peers;                // > 0: the output rows go to peer_out[row block] (ColumnsGeom::peer_out)
    void *const *peer_out;
};

template <typename T, bool FWD>
int four_step_columns_launch(const dsc_cuda_plan *p, const void *x, bool x_real, void *out, long long outer, int x_n,
                             long long inner, void *work, size_t work_bytes, void *stream,
                             const PostTwiddle *post = nullptr) {
    using V = cx<T>;
    const long long n = p->n;
    ColumnsEntry *ce = p->col_lg_n2 ? columns_entry<T, FWD>(p->col_lg_n1, p->col_lg_n2) : nullptr;
    const int lg_inner = pow2_shift(inner);
    if (ce == nullptr || lg_inner < 0) return fail(DSC_CUDA_EUNSUPPORTED, "two-pass transform (n=%lld) along a strided axis of inner extent %lld", n, inner);
    // 16 points per thread where the plan's column tables are its two-pass tables and carry the 16-point copies (passes of at
    // most 256 points): twice the resident warps, like the last-axis register-direct launch.  DSC_COLUMNS_E16=0 disables.
    static const bool no_c16 = [] { const char *e = getenv("DSC_COLUMNS_E16"); return e != nullptr && *e == '0'; }();
    const bool e16 = !no_c16 && ce->fn16 != nullptr && p->lg_n2 != 0 && p->col_lg_n1 == p->lg_n1 && p->col_lg_n2 == p->lg_n2 &&
                     p->tw1_e16[1] != nullptr && p->tw2_e16[1] != nullptr;
    const int ce_l_a = e16 ? ce->l_a16 : ce->l_a, ce_l_b = e16 ? ce->l_b16 : ce->l_b, ce_smem = e16 ? ce->smem16 : ce->smem;
    auto ce_fn = e16 ? ce->fn16 : ce->fn;
    const int l_max = ce_l_a > ce_l_b ? ce_l_a : ce_l_b;
    if (inner < l_max) return fail(DSC_CUDA_EUNSUPPORTED, "two-pass transform (n=%lld): inner extent %lld is narrower than a tile", n, inner);
    const int lg_ic = columns_chunk_lg(n, lg_inner, l_max, sizeof(V));
    const long long ic = 1LL << lg_ic, chunks = inner >> lg_ic, rows = outer * chunks;
    if (rows <= 0) return 0;
    const size_t row_bytes = (size_t)n * (size_t)ic * sizeof(V);
    const size_t sync_bytes = align_up((size_t)(1 + 2 * rows) * sizeof(unsigned), 256);
    const long long tiles_a = (1LL << p->col_lg_n2) * (ic / ce_l_a), tiles_b = (1LL << p->col_lg_n1) * (ic / ce_l_b);
    if (work == nullptr || work_bytes < sync_bytes + row_bytes || rows * (tiles_a + tiles_b) >= 0x7fffffffLL ||
        outer * chunks > 0x7fffffffLL)
        return fail(DSC_CUDA_EUNSUPPORTED, "two-pass transform (n=%lld) along a strided axis: work buffer of %zu bytes holds no row of %zu", n, work_bytes, row_bytes);
#if defined(DSC_EMUL)
    ce->grid = 3;
    ce->grid16 = 3;
#else
    if (e16 ? ce->grid16 == 0 : !ce->configured) {
        if (ce_smem > 48 * 1024) {
            const cudaError_t err = cudaFuncSetAttribute((const void *)ce_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ce_smem);
            if (err != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "smem attribute: %s", cudaGetErrorString(err));
        }
        int per_sm = 0, dev = 0, sms = 0;
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)ce_fn, ce->threads, ce_smem);
        if (err == cudaSuccess) err = cudaGetDevice(&dev);
        if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess || per_sm < 1) return fail(DSC_CUDA_ELAUNCH, "occupancy query: %s", cudaGetErrorString(err));
        if (e16) ce->grid16 = per_sm * sms;
        else { ce->grid = per_sm * sms; ce->configured = true; }
    }
#endif
    const int ce_grid = e16 ? ce->grid16 : ce->grid;
    long long ring = (long long)((work_bytes - sync_bytes) / row_bytes);
    const long long cap = (long long)(((size_t)64 << 20) / row_bytes) > 4 ? (long long)(((size_t)64 << 20) / row_bytes) : 4;
    if (ring > cap) ring = cap;
    if (ring >= rows) ring = 0;
    FourStepSync s{};
    s.ticket = (unsigned *)work;
    s.a_done = s.ticket + 1;
    s.b_done = s.a_done + rows;
    s.tiles_a = (int)tiles_a;
    s.tiles_b = (int)tiles_b;
    s.ring = (int)ring;
    s.rows = (int)rows;
    long long lag = (3LL * ce_grid + tiles_a + tiles_b - 1) / (tiles_a + tiles_b);
    if (lag < 1) lag = 1;
    // at most half the ring -- with a ring of ONE work row that is lag 0 (second pass of row r ticketed before the
    // first pass of row r + 1): raising it back to 1 would ticket A(r + 1) ahead of the B(r) it has to wait for,
    // and a grid smaller than a row's tiles would then spin forever.  decode_ticket handles lag 0.
    if (ring > 0 && lag > ring / 2) lag = ring / 2;
    if (lag > rows) lag = rows;
    s.lag = (int)lag;

    V *mid = (V *)((char *)work + sync_bytes);
    FftArgs a{}, b{};
    a.x = x; a.out = mid; a.ring_out = ring;
    set_stage_tables<T>(a, e16 ? p->tw1_e16 : p->col_tw1);
    a.tw_lo = p->col_lo; a.tw_hi = p->col_hi;
    a.four_shift = p->col_shift; a.four_mask = (1 << p->col_shift) - 1;
    b.x = mid; b.out = out; b.ring_in = ring;
    set_stage_tables<T>(b, e16 ? p->tw2_e16 : p->col_tw2);
    b.do_scale = !FWD; b.scale = 1.0 / (double)n;
    ColumnsGeom g{};
    g.x_ostride = (long long)x_n * inner;
    g.out_ostride = n * inner;
    g.inner = inner;
    g.lg_ic = lg_ic;
    g.chunks = (int)chunks;
    g.x_n = x_n < n ? x_n : (int)n;
    g.x_real = x_real;
    if (post != nullptr) {
        g.post_twiddle = 1;
        g.post_mask = (unsigned)(post->total - 1);
        g.col_offset = post->col_offset;
        b.tw_lo = post->lo; b.tw_hi = post->hi;
        b.four_shift = post->shift; b.four_mask = (1 << post->shift) - 1;
        if (post->n_peers > 0) {
            g.n_peers = post->n_peers;
            g.peer_shift = pow2_shift(n / post->n_peers);
            for (int q = 0; q < post->n_peers; ++q) g.peer_out[q] = post->peer_out[q];
        }
    }
#if defined(DSC_EMUL)
    memset(work, 0, sync_bytes);
#else
    const cudaError_t me = cudaMemsetAsync(work, 0, sync_bytes, (cudaStream_t)stream);
    if (me != cudaSuccess) return fail(DSC_CUDA_ELAUNCH, "memset: %s", cudaGetErrorString(me));
#endif
    const long long tiles = rows * (tiles_a + tiles_b);
    const unsigned blocks = (unsigned)(tiles < ce_grid ? tiles : ce_grid);
    DSC_LAUNCH(ce_fn, blocks, ce->threads, ce_smem, stream, a, b, s, g);
    return check_launch("four_step_columns");
}

template <typename T, bool FWD>
int run_fft(const dsc_cuda_plan *p, const void *x, bool x_real, void *out,
            long long outer, int x_n, long long inner, void *work, size_t work_bytes, void *stream) {
    const long long n = p->n;
    const long long take = x_n < n ? x_n : n;
    if (p->lg_n2 == 0) {
        FftArgs a{};
        a.x = x; a.out = out;
        a.lines = outer * inner;
        a.inner = inner;
        a.gi = LineGeom{(long long)x_n * inner, 1, inner};
        a.go = LineGeom{n * inner, 1, inner};
        a.in_limit = take * inner;
        a.in_kind = x_real ? IN_REAL : IN_COMPLEX;
        set_stage_tables<T>(a, p->tw1);
        a.strided = inner > 1;
        a.do_scale = !FWD; a.scale = 1.0 / (double)n;
        // dense last-axis lines in whole blocks take the bandwidth path
        KernelEntry *fast = get_table<T, FWD, MODE_FAST, false>();
#if !defined(DSC_EMUL)
        if (inner == 1 && !x_real && x_n == n && p->col_lg_n2 != 0) {
            // 2^14-point lines fill a whole SM's shared memory in one block; split over a cluster of two they overlap
            const int rcc = cluster_launch<T, FWD>(p, x, n, out, n, outer, !FWD, stream);
            if (rcc <= 0) return rcc;
        }
#endif
        if (inner == 1 && !x_real && x_n == n && a.lines % fast[p->lg_n].lpb == 0)
            return launch_lines(fast, p->lg_n, a, stream);
        // long columns: both passes of the plan's column decomposition in one launch (whole 64-512-byte
        // segments per access instead of the 8-16 bytes per column a single-pass block of this length touches)
        if (inner > 1 && p->col_lg_n2 != 0 && work != nullptr &&
            four_step_columns_launch<T, FWD>(p, x, x_real, out, outer, x_n, inner, work, work_bytes, stream) == 0)
            return 0;
        return launch_lines(c2c_table<T, FWD>(inner > 1), p->lg_n, a, stream);
    }
    if (inner != 1) return four_step_columns_launch<T, FWD>(p, x, x_real, out, outer, x_n, inner, work, work_bytes, stream);
    FftArgs a{};
    a.x = x;
    a.gi = LineGeom{(long long)x_n, 1, 1LL << p->lg_n2};
    a.in_limit = take;
    a.in_kind = x_real ? IN_REAL : IN_COMPLEX;
    return four_step<T, FWD>(p, a, outer, work, work_bytes, out, n, !FWD, stream);
}

template <typename T>
int run_rfft(const dsc_cuda_plan *p, const void *x, void *out, long long outer, int x_n, long long inner,
             void *work, size_t work_bytes, void *stream) {
    using V = cx<T>;
    const long long n = p->n;                       // complex order; 2n real samples
    const long long take = x_n < 2 * n ? x_n : 2 * n;
    if (p->lg_n2 == 0) {
        FftArgs a{};
        a.x = x; a.out = out;
        a.lines = outer * inner;
        a.inner = inner;
        a.gi = LineGeom{(long long)x_n * inner, 1, 2 * inner};   // REAL elements
        a.gi_pstride = inner;
        a.go = LineGeom{(n + 1) * inner, 1, inner};
        a.in_limit = take * inner;
        set_stage_tables<T>(a, p->tw1);
        a.tw_real = p->tw_real;
        a.strided = inner > 1;
        // dense last-axis lines in whole blocks, no pad / crop: the bandwidth path
        KernelEntry *fast = get_table<T, true, MODE_R2C_FAST, false>();
        if (inner == 1 && x_n == 2 * n && p->lg_n >= real_fast_min_lg<T>() && a.lines % fast[p->lg_n].lpb == 0 &&
            (uintptr_t)x % (2 * sizeof(T)) == 0)
            return launch_lines(fast, p->lg_n, a, stream);
        return launch_lines(get_table<T, true, MODE_R2C, false>(), p->lg_n, a, stream, sizeof(T));
    }
    if (inner != 1) return fail(DSC_CUDA_EUNSUPPORTED, "two-pass rfft (order %lld) along a strided axis", n);
    // packed complex transform into the bin rows (stride n+1), then un-mix the bin pairs in place
    FftArgs a{};
    a.x = x;
    a.gi = LineGeom{(long long)x_n, 2, 2LL << p->lg_n2};     // REAL elements
    a.gi_pstride = 1;
    a.in_limit = take;
    a.in_kind = IN_PAIRS;
    // in chunks of rows that stay in L2 between the transform and the un-mix, which then costs no HBM read
    const long long chunk = l2_chunk_rows((size_t)(n + 1) * sizeof(V));
    for (long long r0 = 0; r0 < outer; r0 += chunk) {
        const long long rows = outer - r0 < chunk ? outer - r0 : chunk;
        FftArgs ac = a;
        ac.x = (const T *)x + (size_t)r0 * x_n;
        V *oc = (V *)out + (size_t)r0 * (n + 1);
        // un-mixing inside the transform's second pass (dense aligned rows): the spectrum is written once, as X
        const RealFuse rf{nullptr};
        int rc = take == 2 * n ? four_step<T, true>(p, ac, rows, work, work_bytes, oc, n + 1, false, stream, false, &rf) : 1;
        if (rc <= 0) { if (rc) return rc; continue; }
        rc = four_step<T, true>(p, ac, rows, work, work_bytes, oc, n + 1, false, stream, true);
        if (rc) return rc;
        const long long items = rows * (n / 2);
        const int blocks = (int)((items + 255) / 256 < 148 * 16 ? (items + 255) / 256 : 148 * 16);
        auto mix = real_mix_rows<true, T>;
        DSC_LAUNCH(mix, blocks, 256, 0, stream, (const V *)nullptr, oc, rows, (int)n,
                   (long long)0, 0, (const V *)p->tw_real_lo, (const V *)p->tw_real_hi, p->real_shift,
                   (1 << p->real_shift) - 1);
        rc = check_launch("real_mix_rows");
        if (rc) return rc;
    }
    return 0;
}

template <typename T>
int run_irfft(const dsc_cuda_plan *p, const void *x, void *out, long long outer, int x_n, long long inner,
              void *work, size_t work_bytes, void *stream, int keep = 0) {
    using V = cx<T>;
    const long long n = p->n;
    const long long take = x_n < n + 1 ? x_n : n + 1;
    if (keep && (p->lg_n2 != 0 || inner != 1)) return fail(DSC_CUDA_EUNSUPPORTED, "fused crop: single-pass orders along the last axis only");
    if (p->lg_n2 == 0) {
        FftArgs a{};
        a.x = x; a.out = out;
        a.lines = outer * inner;
        a.inner = inner;
        a.gi = LineGeom{(long long)x_n * inner, 1, inner};
        a.go = LineGeom{(keep ? (long long)keep : 2 * n) * inner, 1, 2 * inner};            // REAL elements
        a.out_take = keep;
        a.go_pstride = inner;
        a.in_limit = take * inner;
        set_stage_tables<T>(a, p->tw1);
        a.tw_real = p->tw_real;
        a.strided = inner > 1;
        a.do_scale = 1; a.scale = 1.0 / (double)n;               // 2/(2n), dsc_fft.h:232
        KernelEntry *fast = get_table<T, false, MODE_C2R_FAST, false>();
        if (!keep && inner == 1 && x_n == n + 1 && p->lg_n >= real_fast_min_lg<T>() && a.lines % fast[p->lg_n].lpb == 0 &&
            (uintptr_t)out % (2 * sizeof(T)) == 0)
            return launch_lines(fast, p->lg_n, a, stream);
        return launch_lines(get_table<T, false, MODE_C2R, false>(), p->lg_n, a, stream, sizeof(T));
    }
    if (inner != 1) return fail(DSC_CUDA_EUNSUPPORTED, "two-pass irfft (order %lld) along a strided axis", n);
    if (take == n + 1) {
        // the packed points are built inside the inverse transform's first pass (dense aligned bin rows): no packed rows
        const RealFuse rf{nullptr};
        FftArgs a{};
        a.x = x;
        a.gi = LineGeom{(long long)x_n, 1, 1LL << p->lg_n2};
        a.in_limit = take;
        a.in_kind = IN_COMPLEX;
        const int rc = four_step<T, false>(p, a, outer, work, work_bytes, out, n, true, stream, false, &rf);
        if (rc <= 0) return rc;
    }
    // work = [ packed z rows of the chunk | four-step work ]
    const size_t row_bytes = (size_t)n * sizeof(V);
    if (!work || work_bytes < 2 * row_bytes + 4096) return fail(DSC_CUDA_ENOMEM, "work buffer holds no line (order %lld)", n);
    long long chunk = (long long)((work_bytes / 2) / row_bytes);
    if (chunk > l2_chunk_rows(row_bytes)) chunk = l2_chunk_rows(row_bytes);     // the packed rows stay in L2
    if (chunk > outer) chunk = outer;
    V *z = (V *)work;
    char *fs_work = (char *)work + align_up((size_t)chunk * row_bytes, 256);
    const size_t fs_bytes = work_bytes - (size_t)(fs_work - (char *)work);
    for (long long r0 = 0; r0 < outer; r0 += chunk) {
        const long long rows = outer - r0 < chunk ? outer - r0 : chunk;
        const long long items = rows * (n / 2);
        const int blocks = (int)((items + 255) / 256 < 148 * 16 ? (items + 255) / 256 : 148 * 16);
        auto mix = real_mix_rows<false, T>;
        DSC_LAUNCH(mix, blocks, 256, 0, stream, (const V *)x + (size_t)r0 * x_n, z, rows,
                   (int)n, (long long)x_n, (int)take, (const V *)p->tw_real_lo, (const V *)p->tw_real_hi,
                   p->real_shift, (1 << p->real_shift) - 1);
        int rc = check_launch("real_mix_rows");
        if (rc) return rc;
        FftArgs a{};
        a.x = z;
        a.gi = LineGeom{n, 1, 1LL << p->lg_n2};
        a.in_limit = n;
        a.in_kind = IN_COMPLEX;
        rc = four_step<T, false>(p, a, rows, fs_work, fs_bytes, (V *)out + (size_t)r0 * n, n, true, stream);
        if (rc) return rc;
    }
    return 0;
}

template <typename T>
int run_filter(const dsc_cuda_plan *p, const void *x, const void *spectrum, void *out, long long outer, int x_n,
               void *work, size_t work_bytes, void *stream, int keep = 0) {
    using V = cx<T>;
    const long long n = p->n;
    const long long take = x_n < 2 * n ? x_n : 2 * n;
    if (keep && p->lg_n2 != 0) return fail(DSC_CUDA_EUNSUPPORTED, "fused crop: single-pass orders only");
    if (p->lg_n2 == 0) {
        FftArgs a{};
        a.x = x; a.out = out;
        a.lines = outer;
        a.inner = 1;
        a.gi = LineGeom{(long long)x_n, 1, 2};                   // REAL elements
        a.gi_pstride = 1;
        a.go = LineGeom{keep ? (long long)keep : 2 * n, 1, 2};   // REAL elements
        a.out_take = keep;
        a.go_pstride = 1;
        a.in_limit = take;
        set_stage_tables<T>(a, p->tw1);
        a.tw_real = p->tw_real;
        a.filt = spectrum;
        a.do_scale = 1; a.scale = 1.0 / (double)n;
        return launch_lines(get_table<T, true, MODE_FILTER, false>(), p->lg_n, a, stream, sizeof(T));
    }
    // rows in flight: packed spectrum rows of the chunk + four-step work
    const size_t row_bytes = (size_t)n * sizeof(V);
    if (!work || work_bytes < 2 * row_bytes + 4096) return fail(DSC_CUDA_ENOMEM, "work buffer holds no line (order %lld)", n);
    long long chunk = (long long)((work_bytes / 2) / row_bytes);
    if (chunk > l2_chunk_rows(row_bytes)) chunk = l2_chunk_rows(row_bytes);     // the packed rows stay in L2
    if (chunk > outer) chunk = outer;
    V *z = (V *)work;
    char *fs_work = (char *)work + align_up((size_t)chunk * row_bytes, 256);
    const size_t fs_bytes = work_bytes - (size_t)(fs_work - (char *)work);
    for (long long r0 = 0; r0 < outer; r0 += chunk) {
        const long long rows = outer - r0 < chunk ? outer - r0 : chunk;
        FftArgs a{};
        a.x = (const T *)x + (size_t)r0 * x_n;
        a.gi = LineGeom{(long long)x_n, 2, 2LL << p->lg_n2};     // REAL elements
        a.gi_pstride = 1;
        a.in_limit = take;
        a.in_kind = IN_PAIRS;
        // un-mix, spectrum product and mix inside the forward transform's second pass where the shape allows
        const RealFuse rf{spectrum};
        int rc = take == 2 * n ? four_step<T, true>(p, a, rows, fs_work, fs_bytes, z, n, false, stream, true, &rf) : 1;
        if (rc < 0) return rc;
        if (rc > 0) {
            rc = four_step<T, true>(p, a, rows, fs_work, fs_bytes, z, n, false, stream, true);
            if (rc) return rc;
            const long long items = rows * (n / 2);
            const int blocks = (int)((items + 255) / 256 < 148 * 16 ? (items + 255) / 256 : 148 * 16);
            auto pk = filter_pairs_rows<T>;
            DSC_LAUNCH(pk, blocks, 256, 0, stream, z, (const V *)spectrum, rows, (int)n,
                       (const V *)p->tw_real_lo, (const V *)p->tw_real_hi, p->real_shift, (1 << p->real_shift) - 1);
            rc = check_launch("filter_pairs_rows");
            if (rc) return rc;
        }
        FftArgs b{};
        b.x = z;
        b.gi = LineGeom{n, 1, 1LL << p->lg_n2};
        b.in_limit = n;
        b.in_kind = IN_COMPLEX;
        rc = four_step<T, false>(p, b, rows, fs_work, fs_bytes, (V *)out + (size_t)r0 * n, n, true, stream);
        if (rc) return rc;
    }
    return 0;
}

bool plan_ok(const dsc_cuda_plan *p) {
    return p && p->n >= 1 && (p->dtype == DSC_CUDA_F32 || p->dtype == DSC_CUDA_F64);
}

}  // namespace

// ---------------------------------------------------------------------------------------------

namespace {
inline int pointwise_blocks(long long total) {
    const long long b = (total + 255) / 256;
    return (int)(b < 148 * 32 ? b : 148 * 32);
}
template <typename T>
int unary_by_op(int op, const void *x, void *out, long long n, void *stream) {
    using V = cx<T>;
    const int blocks = pointwise_blocks(n);
    switch (op) {
    case DSC_CUDA_OP_ABS: { auto k = pointwise_c2r<T, 0>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)x, (T *)out, n); break; }
    case DSC_CUDA_OP_ANGLE: { auto k = pointwise_c2r<T, 1>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)x, (T *)out, n); break; }
    case DSC_CUDA_OP_REAL: { auto k = pointwise_c2r<T, 2>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)x, (T *)out, n); break; }
    case DSC_CUDA_OP_IMAG: { auto k = pointwise_c2r<T, 3>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)x, (T *)out, n); break; }
    case DSC_CUDA_OP_CONJ: { auto k = pointwise_conj<T>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)x, (V *)out, n); break; }
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_unary: unknown op %d", op);
    }
    return check_launch("pointwise unary");
}
template <typename V>
int binary_by_op(int op, const void *a, const void *b, void *out, long long rows, long long cols, int b_mode, void *stream) {
    const int blocks = pointwise_blocks(rows * cols);
    switch (op) {
    case DSC_CUDA_OP_ADD: { auto k = pointwise_binary<V, 0>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)a, (const V *)b, (V *)out, rows, cols, b_mode); break; }
    case DSC_CUDA_OP_SUB: { auto k = pointwise_binary<V, 1>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)a, (const V *)b, (V *)out, rows, cols, b_mode); break; }
    case DSC_CUDA_OP_MUL: { auto k = pointwise_binary<V, 2>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)a, (const V *)b, (V *)out, rows, cols, b_mode); break; }
    case DSC_CUDA_OP_DIV: { auto k = pointwise_binary<V, 3>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const V *)a, (const V *)b, (V *)out, rows, cols, b_mode); break; }
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_binary: unknown op %d", op);
    }
    return check_launch("pointwise binary");
}
}  // namespace


extern "C" {

const char *dsc_cuda_last_error(void) { return g_err; }

int dsc_cuda_device_count(void) {
#if defined(DSC_EMUL)
    return 1;
#else
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
#endif
}

size_t dsc_cuda_plan_bytes(int n, int fft_type, int dtype) {
    if (n < 1 || (n & (n - 1))) return 0;
    const PlanLayout L = dtype == DSC_CUDA_F64 ? plan_layout<double>(n, fft_type) : plan_layout<float>(n, fft_type);
    return L.ok ? L.total : 0;
}

int dsc_cuda_plan_build(dsc_cuda_plan *plan, int n, int fft_type, int dtype,
                        void *dev_mem, size_t dev_bytes, void *stream) {
    if (!plan || n < 1 || (n & (n - 1))) return fail(DSC_CUDA_EINVAL, "plan length %d is not a power of two", n);
    if (dtype != DSC_CUDA_F32 && dtype != DSC_CUDA_F64) return fail(DSC_CUDA_EINVAL, "plan dtype %d", dtype);
    if (fft_type != DSC_CUDA_FFT_REAL && fft_type != DSC_CUDA_FFT_COMPLEX) return fail(DSC_CUDA_EINVAL, "plan type %d", fft_type);
    const PlanLayout L = dtype == DSC_CUDA_F64 ? plan_layout<double>(n, fft_type) : plan_layout<float>(n, fft_type);
    if (!L.ok) return fail(DSC_CUDA_EUNSUPPORTED, "length 2^%d exceeds the two-pass range", L.lg_n);
    if (!dev_mem || dev_bytes < L.total) return fail(DSC_CUDA_ENOMEM, "plan needs %zu device bytes", L.total);
    memset(plan, 0, sizeof(*plan));
    plan->n = n; plan->lg_n = L.lg_n;
    plan->fft_type = fft_type; plan->dtype = dtype;
    plan->lg_n1 = L.lg_n1; plan->lg_n2 = L.lg_n2;
    plan->four_shift = L.four_shift; plan->real_shift = L.real_shift;
    plan->dev_base = dev_mem; plan->dev_bytes = L.total;
    return dtype == DSC_CUDA_F64 ? build_tables<double>(plan, L, stream) : build_tables<float>(plan, L, stream);
}

size_t dsc_cuda_work_bytes(const dsc_cuda_plan *plan, int64_t lines) {
    if (!plan_ok(plan) || plan->lg_n2 == 0 || lines <= 0) return 0;
    // ring of work rows that stays in L2 (at least 4 rows), plus the per-row flags of the fused launch;
    // REAL plans also stage the packed spectrum of the rows in flight (irfft)
    const size_t es = plan->dtype == DSC_CUDA_F64 ? sizeof(double2) : sizeof(float2);
    const size_t row = (size_t)plan->n * es;
    size_t ring = (64u << 20) / row > 4 ? (64u << 20) / row : 4;
    if (ring > (size_t)lines) ring = (size_t)lines;
    const size_t sync = align_up((size_t)(1 + 2 * lines) * sizeof(unsigned), 256);
    size_t total = sync + ring * row + 256;
    if (plan->fft_type == DSC_CUDA_FFT_REAL) total = 2 * total + 4096;
    return total;
}

size_t dsc_cuda_work_bytes_axis(const dsc_cuda_plan *plan, int64_t outer, int64_t inner) {
    if (inner <= 1) return dsc_cuda_work_bytes(plan, outer);
    if (!plan_ok(plan) || plan->col_lg_n2 == 0 || outer <= 0) return 0;
    const bool f32 = plan->dtype == DSC_CUDA_F32;
    ColumnsEntry *ce = f32 ? columns_entry<float, true>(plan->col_lg_n1, plan->col_lg_n2)
                           : columns_entry<double, true>(plan->col_lg_n1, plan->col_lg_n2);
    const int lg_inner = pow2_shift(inner);
    if (ce == nullptr || lg_inner < 0) return 0;
    const int l_max = ce->l_a > ce->l_b ? ce->l_a : ce->l_b;
    if (inner < l_max) return 0;
    const size_t es = f32 ? sizeof(float2) : sizeof(double2);
    const int lg_ic = columns_chunk_lg(plan->n, lg_inner, l_max, es);
    const size_t rows = (size_t)outer * (size_t)(inner >> lg_ic);
    const size_t row = ((size_t)plan->n << lg_ic) * es;
    size_t ring = ((size_t)64 << 20) / row > 4 ? ((size_t)64 << 20) / row : 4;
    if (ring > rows) ring = rows;
    return align_up((1 + 2 * rows) * sizeof(unsigned), 256) + ring * row + 256;
}

int dsc_cuda_fft(const dsc_cuda_plan *plan, const void *x, int x_dtype, void *out,
                 int64_t outer, int x_n, int64_t inner, int forward,
                 void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || !x || !out || x_n < 1 || outer < 0 || inner < 1) return fail(DSC_CUDA_EINVAL, "dsc_cuda_fft: bad argument");
    const bool x_real = x_dtype == DSC_CUDA_F32 || x_dtype == DSC_CUDA_F64;
    const int x_prec = (x_dtype == DSC_CUDA_F32 || x_dtype == DSC_CUDA_C32) ? DSC_CUDA_F32 : DSC_CUDA_F64;
    if (x_prec != plan->dtype) return fail(DSC_CUDA_EINVAL, "dsc_cuda_fft: input precision does not match the plan");
    if (plan->dtype == DSC_CUDA_F32)
        return forward ? run_fft<float, true>(plan, x, x_real, out, outer, x_n, inner, work, work_bytes, stream)
                       : run_fft<float, false>(plan, x, x_real, out, outer, x_n, inner, work, work_bytes, stream);
    return forward ? run_fft<double, true>(plan, x, x_real, out, outer, x_n, inner, work, work_bytes, stream)
                   : run_fft<double, false>(plan, x, x_real, out, outer, x_n, inner, work, work_bytes, stream);
}

int dsc_cuda_fft_columns_twiddled(const dsc_cuda_plan *plan, const void *x, void *out, int64_t cols, int forward,
                                  int64_t col_offset, const void *tw_lo, const void *tw_hi, int shift, int64_t total,
                                  void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || !x || !out || cols < 1 || !tw_lo || !tw_hi || total < 1 || (total & (total - 1)) || total > (1LL << 31))
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_fft_columns_twiddled: bad argument");
    PostTwiddle post{tw_lo, tw_hi, shift, (long long)total, (long long)col_offset, 0, nullptr};
    if (plan->dtype == DSC_CUDA_F32)
        return forward ? four_step_columns_launch<float, true>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post)
                       : four_step_columns_launch<float, false>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post);
    return forward ? four_step_columns_launch<double, true>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post)
                   : four_step_columns_launch<double, false>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post);
}

int dsc_cuda_fft_columns_twiddled_p2p(const dsc_cuda_plan *plan, const void *x, int64_t cols, int forward,
                                      int64_t col_offset, const void *tw_lo, const void *tw_hi, int shift, int64_t total,
                                      void *const *peer_out, int n_peers, void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || !x || cols < 1 || !tw_lo || !tw_hi || total < 1 || (total & (total - 1)) || total > (1LL << 31) ||
        !peer_out || n_peers < 1 || n_peers > 8 || (n_peers & (n_peers - 1)) || plan->n % n_peers != 0)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_fft_columns_twiddled_p2p: bad argument");
    for (int q = 0; q < n_peers; ++q)
        if (peer_out[q] == nullptr) return fail(DSC_CUDA_EINVAL, "dsc_cuda_fft_columns_twiddled_p2p: null peer pointer");
    PostTwiddle post{tw_lo, tw_hi, shift, (long long)total, (long long)col_offset, n_peers, peer_out};
    void *out = peer_out[0];     // unused by the scattered store; only has to be non-null
    if (plan->dtype == DSC_CUDA_F32)
        return forward ? four_step_columns_launch<float, true>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post)
                       : four_step_columns_launch<float, false>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post);
    return forward ? four_step_columns_launch<double, true>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post)
                   : four_step_columns_launch<double, false>(plan, x, false, out, 1, plan->n, cols, work, work_bytes, stream, &post);
}

int dsc_cuda_fft_segmented(const dsc_cuda_plan *plan, const void *x, void *out, int64_t lines,
                           int64_t seg_len, int64_t seg_stride, int self_seg, const void *self_x, int forward,
                           void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || !x || !out || lines < 0 || seg_len < 1 || seg_stride < seg_len)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_fft_segmented: bad argument");
    const long long n = plan->n;
    const int seg_shift = pow2_shift(seg_len);
    // only what the multi-GPU four-step needs: two-pass plans, power-of-two segments at least one thread step long
    const int lg_e1 = plan->dtype == DSC_CUDA_F32 ? pass_lg_e<float>(plan->lg_n1, plan->lg_n2) : pass_lg_e<double>(plan->lg_n1, plan->lg_n2);
    if (plan->lg_n2 == 0 || seg_shift < 0 || n % seg_len != 0 || seg_len == n || seg_shift < plan->lg_n - lg_e1)
        return fail(DSC_CUDA_EUNSUPPORTED, "dsc_cuda_fft_segmented: n=%lld with segments of %lld", n, (long long)seg_len);
    FftArgs a{};
    a.x = x;
    a.gi = LineGeom{(long long)seg_len, 1, 1LL << plan->lg_n2};     // line r starts r * seg_len into every segment
    a.in_limit = n;
    a.in_kind = IN_COMPLEX;
    a.seg_shift = seg_shift;
    a.seg_extra = (long long)(seg_stride - seg_len);
    a.seg_self = -1;
    if (self_seg >= 0 && self_x != nullptr) {
        const size_t es = plan->dtype == DSC_CUDA_F32 ? sizeof(float2) : sizeof(double2);
        const ptrdiff_t d = (const char *)self_x - (const char *)x;
        if (d % (ptrdiff_t)es != 0) return fail(DSC_CUDA_EINVAL, "dsc_cuda_fft_segmented: self_x is not element-aligned with x");
        a.seg_self = self_seg;
        a.seg_self_delta = (long long)(d / (ptrdiff_t)es);
    }
    if (plan->dtype == DSC_CUDA_F32)
        return forward ? four_step<float, true>(plan, a, lines, work, work_bytes, out, n, false, stream)
                       : four_step<float, false>(plan, a, lines, work, work_bytes, out, n, true, stream);
    return forward ? four_step<double, true>(plan, a, lines, work, work_bytes, out, n, false, stream)
                   : four_step<double, false>(plan, a, lines, work, work_bytes, out, n, true, stream);
}

int dsc_cuda_rfft(const dsc_cuda_plan *plan, const void *x, void *out,
                  int64_t outer, int x_n, int64_t inner, void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || plan->fft_type != DSC_CUDA_FFT_REAL || !x || !out || x_n < 1 || outer < 0 || inner < 1)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_rfft: bad argument");
    return plan->dtype == DSC_CUDA_F32 ? run_rfft<float>(plan, x, out, outer, x_n, inner, work, work_bytes, stream)
                                       : run_rfft<double>(plan, x, out, outer, x_n, inner, work, work_bytes, stream);
}

int dsc_cuda_irfft(const dsc_cuda_plan *plan, const void *x, void *out,
                   int64_t outer, int x_n, int64_t inner, void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || plan->fft_type != DSC_CUDA_FFT_REAL || !x || !out || x_n < 1 || outer < 0 || inner < 1)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_irfft: bad argument");
    return plan->dtype == DSC_CUDA_F32 ? run_irfft<float>(plan, x, out, outer, x_n, inner, work, work_bytes, stream)
                                       : run_irfft<double>(plan, x, out, outer, x_n, inner, work, work_bytes, stream);
}

int dsc_cuda_irfft_keep(const dsc_cuda_plan *plan, const void *x, void *out, int64_t outer, int x_n, int keep,
                        void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || plan->fft_type != DSC_CUDA_FFT_REAL || !x || !out || x_n < 1 || outer < 0 || keep < 1 || keep > 2 * plan->n)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_irfft_keep: bad argument");
    if (keep == 2 * plan->n) keep = 0;
    return plan->dtype == DSC_CUDA_F32 ? run_irfft<float>(plan, x, out, outer, x_n, 1, work, work_bytes, stream, keep)
                                       : run_irfft<double>(plan, x, out, outer, x_n, 1, work, work_bytes, stream, keep);
}

int dsc_cuda_filter_keep(const dsc_cuda_plan *plan, const void *x, const void *spectrum, void *out,
                         int64_t outer, int x_n, int keep, void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || plan->fft_type != DSC_CUDA_FFT_REAL || !x || !spectrum || !out || x_n < 1 || outer < 0 || keep < 1 || keep > 2 * plan->n)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_filter_keep: bad argument");
    if (keep == 2 * plan->n) keep = 0;
    return plan->dtype == DSC_CUDA_F32 ? run_filter<float>(plan, x, spectrum, out, outer, x_n, work, work_bytes, stream, keep)
                                       : run_filter<double>(plan, x, spectrum, out, outer, x_n, work, work_bytes, stream, keep);
}

size_t dsc_cuda_filter_work_bytes(const dsc_cuda_plan *plan, int64_t lines) {
    if (!plan_ok(plan) || plan->lg_n2 == 0 || lines <= 0) return 0;
    // packed spectrum rows of the lines in flight (about as many as keep the GPU busy) + four-step work
    const size_t es = plan->dtype == DSC_CUDA_F64 ? sizeof(double2) : sizeof(float2);
    const size_t row = (size_t)plan->n * es;
    size_t rows = (256u << 20) / row > 8 ? (256u << 20) / row : 8;
    if (rows > (size_t)lines) rows = (size_t)lines;
    return 2 * (rows * row + dsc_cuda_work_bytes(plan, (int64_t)rows)) + 4096;
}

int dsc_cuda_filter(const dsc_cuda_plan *plan, const void *x, const void *spectrum, void *out,
                    int64_t outer, int x_n, void *work, size_t work_bytes, void *stream) {
    if (!plan_ok(plan) || plan->fft_type != DSC_CUDA_FFT_REAL || !x || !spectrum || !out || x_n < 1 || outer < 0)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_filter: bad argument");
    return plan->dtype == DSC_CUDA_F32 ? run_filter<float>(plan, x, spectrum, out, outer, x_n, work, work_bytes, stream)
                                       : run_filter<double>(plan, x, spectrum, out, outer, x_n, work, work_bytes, stream);
}

int dsc_cuda_cmul(const void *a, const void *b, void *out, int dtype,
                  int64_t rows, int64_t cols, int b_rows, void *stream) {
    if (!a || !b || !out || rows < 0 || cols < 0) return fail(DSC_CUDA_EINVAL, "dsc_cuda_cmul: bad argument");
    const long long total = rows * cols;
    if (total == 0) return 0;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    if (dtype == DSC_CUDA_C32)
        DSC_LAUNCH(cmul_rows<float>, blocks, 256, 0, stream, (const float2 *)a, (const float2 *)b, (float2 *)out,
                   (long long)rows, (long long)cols, b_rows);
    else if (dtype == DSC_CUDA_C64)
        DSC_LAUNCH(cmul_rows<double>, blocks, 256, 0, stream, (const double2 *)a, (const double2 *)b, (double2 *)out,
                   (long long)rows, (long long)cols, b_rows);
    else return fail(DSC_CUDA_EINVAL, "dsc_cuda_cmul: dtype %d is not complex", dtype);
    return check_launch("cmul_rows");
}

int dsc_cuda_unary(int op, const void *x, int x_dtype, void *out, int64_t count, void *stream) {
    if (!x || !out || count < 0) return fail(DSC_CUDA_EINVAL, "dsc_cuda_unary: bad argument");
    if (count == 0) return 0;
    if (x_dtype == DSC_CUDA_C32) return unary_by_op<float>(op, x, out, (long long)count, stream);
    if (x_dtype == DSC_CUDA_C64) return unary_by_op<double>(op, x, out, (long long)count, stream);
    return fail(DSC_CUDA_EINVAL, "dsc_cuda_unary: dtype %d is not complex", x_dtype);
}

int dsc_cuda_binary(int op, const void *a, const void *b, void *out, int dtype,
                    int64_t rows, int64_t cols, int b_mode, void *stream) {
    if (!a || !b || !out || rows < 0 || cols < 0 || b_mode < 0 || b_mode > 2) return fail(DSC_CUDA_EINVAL, "dsc_cuda_binary: bad argument");
    if (rows * cols == 0) return 0;
    switch (dtype) {
    case DSC_CUDA_F32: return binary_by_op<float>(op, a, b, out, rows, cols, b_mode, stream);
    case DSC_CUDA_F64: return binary_by_op<double>(op, a, b, out, rows, cols, b_mode, stream);
    case DSC_CUDA_C32: return binary_by_op<float2>(op, a, b, out, rows, cols, b_mode, stream);
    case DSC_CUDA_C64: return binary_by_op<double2>(op, a, b, out, rows, cols, b_mode, stream);
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_binary: unknown dtype %d", dtype);
    }
}

int dsc_cuda_fill_twiddles(void *out, int64_t count, int64_t mult, int64_t denom, int dtype, void *stream) {
    if (!out || count <= 0 || denom <= 0) return fail(DSC_CUDA_EINVAL, "dsc_cuda_fill_twiddles: bad argument");
    const int blocks = (int)((count + 255) / 256 < 1024 ? (count + 255) / 256 : 1024);
    if (dtype == DSC_CUDA_F32 || dtype == DSC_CUDA_C32)
        DSC_LAUNCH(fill_power_twiddles<float>, blocks, 256, 0, stream, (float2 *)out, (long long)count, (long long)mult, (long long)denom);
    else
        DSC_LAUNCH(fill_power_twiddles<double>, blocks, 256, 0, stream, (double2 *)out, (long long)count, (long long)mult, (long long)denom);
    return check_launch("fill_power_twiddles");
}

int dsc_cuda_transpose_twiddle(const void *in, void *out, int64_t rows, int64_t cols, int64_t r0,
                               const void *tw_lo, const void *tw_hi, int shift, int forward,
                               int dtype, void *stream) {
    if (!in || !out || rows <= 0 || cols <= 0 || rows > 0x7fffffff || cols > 0x7fffffff)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_twiddle: bad argument");
    const long long tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    if (tiles > 0x7fffffffLL) return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_twiddle: too many tiles");
    const int mask = (1 << shift) - 1;
    const bool f32 = dtype == DSC_CUDA_F32 || dtype == DSC_CUDA_C32;
    if (f32 && forward) {
        auto k = transpose_twiddle<float, true>;
        DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, (const float2 *)in, (float2 *)out, (int)rows, (int)cols, (long long)r0,
                   (const float2 *)tw_lo, (const float2 *)tw_hi, shift, mask);
    } else if (f32) {
        auto k = transpose_twiddle<float, false>;
        DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, (const float2 *)in, (float2 *)out, (int)rows, (int)cols, (long long)r0,
                   (const float2 *)tw_lo, (const float2 *)tw_hi, shift, mask);
    } else if (forward) {
        auto k = transpose_twiddle<double, true>;
        DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, (const double2 *)in, (double2 *)out, (int)rows, (int)cols, (long long)r0,
                   (const double2 *)tw_lo, (const double2 *)tw_hi, shift, mask);
    } else {
        auto k = transpose_twiddle<double, false>;
        DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, (const double2 *)in, (double2 *)out, (int)rows, (int)cols, (long long)r0,
                   (const double2 *)tw_lo, (const double2 *)tw_hi, shift, mask);
    }
    return check_launch("transpose_twiddle");
}

int dsc_cuda_transpose(const void *in, void *out, int64_t rows, int64_t cols, int elem_bytes, void *stream) {
    if (!in || !out || rows <= 0 || cols <= 0 || rows > 0x7fffffff || cols > 0x7fffffff)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose: bad argument");
    const long long tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    if (tiles > 0x7fffffffLL) return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose: too many tiles");
    if (elem_bytes == 4) { auto k = transpose_plain<float>; DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, (const float *)in, (float *)out, (int)rows, (int)cols); }
    else if (elem_bytes == 8) { auto k = transpose_plain<float2>; DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, (const float2 *)in, (float2 *)out, (int)rows, (int)cols); }
    else if (elem_bytes == 16) { auto k = transpose_plain<double2>; DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, (const double2 *)in, (double2 *)out, (int)rows, (int)cols); }
    else return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose: element size %d", elem_bytes);
    return check_launch("transpose_plain");
}

int dsc_cuda_transpose_cast(const void *in, int in_dtype, void *out, int64_t rows, int64_t cols, int64_t limit, void *stream) {
    if (!in || !out || rows <= 0 || cols <= 0 || rows > 0x7fffffff || cols > 0x7fffffff)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_cast: bad argument");
    const long long tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    if (tiles > 0x7fffffffLL) return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_cast: too many tiles");
    switch (in_dtype) {
        case DSC_CUDA_F32: { auto k = transpose_cast_pad<float, true>; DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, in, (float2 *)out, (int)rows, (int)cols, (long long)limit); break; }
        case DSC_CUDA_C32: { auto k = transpose_cast_pad<float, false>; DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, in, (float2 *)out, (int)rows, (int)cols, (long long)limit); break; }
        case DSC_CUDA_F64: { auto k = transpose_cast_pad<double, true>; DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, in, (double2 *)out, (int)rows, (int)cols, (long long)limit); break; }
        case DSC_CUDA_C64: { auto k = transpose_cast_pad<double, false>; DSC_LAUNCH(k, (unsigned)tiles, 256, 0, stream, in, (double2 *)out, (int)rows, (int)cols, (long long)limit); break; }
        default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_cast: dtype %d", in_dtype);
    }
    return check_launch("transpose_cast_pad");
}

}  // extern "C"

namespace {
template <typename Tin>
int cast_from(const void *x, void *out, int out_dtype, long long n, void *stream) {
    const int blocks = pointwise_blocks(n);
    switch (out_dtype) {
    case DSC_CUDA_F32: { auto k = pointwise_cast<Tin, float>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Tin *)x, (float *)out, n); break; }
    case DSC_CUDA_F64: { auto k = pointwise_cast<Tin, double>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Tin *)x, (double *)out, n); break; }
    case DSC_CUDA_C32: { auto k = pointwise_cast<Tin, float2>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Tin *)x, (float2 *)out, n); break; }
    case DSC_CUDA_C64: { auto k = pointwise_cast<Tin, double2>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Tin *)x, (double2 *)out, n); break; }
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_cast: unknown dtype %d", out_dtype);
    }
    return check_launch("pointwise_cast");
}

template <typename Ta, typename Tb, typename V>
int mixed_by_op(int op, const void *a, const void *b, void *out, long long rows, long long cols, int b_mode, void *stream) {
    const int blocks = pointwise_blocks(rows * cols);
    switch (op) {
    case DSC_CUDA_OP_ADD: { auto k = pointwise_binary_mixed<Ta, Tb, V, 0>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Ta *)a, (const Tb *)b, (V *)out, rows, cols, b_mode); break; }
    case DSC_CUDA_OP_SUB: { auto k = pointwise_binary_mixed<Ta, Tb, V, 1>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Ta *)a, (const Tb *)b, (V *)out, rows, cols, b_mode); break; }
    case DSC_CUDA_OP_MUL: { auto k = pointwise_binary_mixed<Ta, Tb, V, 2>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Ta *)a, (const Tb *)b, (V *)out, rows, cols, b_mode); break; }
    case DSC_CUDA_OP_DIV: { auto k = pointwise_binary_mixed<Ta, Tb, V, 3>; DSC_LAUNCH(k, blocks, 256, 0, stream, (const Ta *)a, (const Tb *)b, (V *)out, rows, cols, b_mode); break; }
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_binary_mixed: unknown op %d", op);
    }
    return check_launch("pointwise binary (mixed dtypes)");
}
// the reference's two-operand conversion table, dsc_dtype.h:73-78, as types
template <typename Ta, typename Tb> struct Promote;
template <> struct Promote<float, float> { using type = float; };
template <> struct Promote<float, double> { using type = double; };
template <> struct Promote<float, float2> { using type = float2; };
template <> struct Promote<float, double2> { using type = double2; };
template <> struct Promote<double, float> { using type = double; };
template <> struct Promote<double, double> { using type = double; };
template <> struct Promote<double, float2> { using type = float2; };
template <> struct Promote<double, double2> { using type = double2; };
template <> struct Promote<float2, float> { using type = float2; };
template <> struct Promote<float2, double> { using type = float2; };
template <> struct Promote<float2, float2> { using type = float2; };
template <> struct Promote<float2, double2> { using type = double2; };
template <> struct Promote<double2, float> { using type = double2; };
template <> struct Promote<double2, double> { using type = double2; };
template <> struct Promote<double2, float2> { using type = double2; };
template <> struct Promote<double2, double2> { using type = double2; };

template <typename Ta>
int mixed_by_b(int op, const void *a, const void *b, int b_dtype, void *out, long long rows, long long cols, int b_mode, void *stream) {
    switch (b_dtype) {
    case DSC_CUDA_F32: return mixed_by_op<Ta, float, typename Promote<Ta, float>::type>(op, a, b, out, rows, cols, b_mode, stream);
    case DSC_CUDA_F64: return mixed_by_op<Ta, double, typename Promote<Ta, double>::type>(op, a, b, out, rows, cols, b_mode, stream);
    case DSC_CUDA_C32: return mixed_by_op<Ta, float2, typename Promote<Ta, float2>::type>(op, a, b, out, rows, cols, b_mode, stream);
    case DSC_CUDA_C64: return mixed_by_op<Ta, double2, typename Promote<Ta, double2>::type>(op, a, b, out, rows, cols, b_mode, stream);
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_binary_mixed: unknown dtype %d", b_dtype);
    }
}

template <typename U>
int gather_typed(const void *in, void *out, const Index4 &g, long long total, void *stream) {
    auto k = gather_strided<U>;
    DSC_LAUNCH(k, pointwise_blocks(total), 256, 0, stream, (const U *)in, (U *)out, g, total);
    return check_launch("gather_strided");
}
template <typename U>
int scatter_typed(void *dst, const void *src, const Index4 &g, long long total, long long src_count, void *stream) {
    auto k = scatter_strided<U>;
    DSC_LAUNCH(k, pointwise_blocks(total), 256, 0, stream, (U *)dst, (const U *)src, g, total, src_count);
    return check_launch("scatter_strided");
}
bool make_index4(Index4 &g, const int shape[4], const int64_t stride[4], int64_t base, long long &total) {
    total = 1;
    for (int d = 0; d < 4; ++d) {
        if (shape[d] < 1) return false;
        g.shape[d] = shape[d];
        g.stride[d] = (long long)stride[d];
        total *= shape[d];
    }
    g.base = (long long)base;
    return true;
}
}  // namespace

extern "C" {

int dsc_cuda_cast(const void *x, int x_dtype, void *out, int out_dtype, int64_t count, void *stream) {
    if (!x || !out || count < 0) return fail(DSC_CUDA_EINVAL, "dsc_cuda_cast: bad argument");
    if (count == 0) return 0;
    switch (x_dtype) {
    case DSC_CUDA_F32: return cast_from<float>(x, out, out_dtype, (long long)count, stream);
    case DSC_CUDA_F64: return cast_from<double>(x, out, out_dtype, (long long)count, stream);
    case DSC_CUDA_C32: return cast_from<float2>(x, out, out_dtype, (long long)count, stream);
    case DSC_CUDA_C64: return cast_from<double2>(x, out, out_dtype, (long long)count, stream);
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_cast: unknown dtype %d", x_dtype);
    }
}

int dsc_cuda_binary_mixed(int op, const void *a, int a_dtype, const void *b, int b_dtype, void *out,
                          int64_t rows, int64_t cols, int b_mode, void *stream) {
    if (!a || !b || !out || rows < 0 || cols < 0 || b_mode < 0 || b_mode > 2) return fail(DSC_CUDA_EINVAL, "dsc_cuda_binary_mixed: bad argument");
    if (rows * cols == 0) return 0;
    switch (a_dtype) {
    case DSC_CUDA_F32: return mixed_by_b<float>(op, a, b, b_dtype, out, rows, cols, b_mode, stream);
    case DSC_CUDA_F64: return mixed_by_b<double>(op, a, b, b_dtype, out, rows, cols, b_mode, stream);
    case DSC_CUDA_C32: return mixed_by_b<float2>(op, a, b, b_dtype, out, rows, cols, b_mode, stream);
    case DSC_CUDA_C64: return mixed_by_b<double2>(op, a, b, b_dtype, out, rows, cols, b_mode, stream);
    default: return fail(DSC_CUDA_EINVAL, "dsc_cuda_binary_mixed: unknown dtype %d", a_dtype);
    }
}

int dsc_cuda_fftfreq(void *out, int dtype, int n, double d, int rfft, void *stream) {
    if (!out || n < 1) return fail(DSC_CUDA_EINVAL, "dsc_cuda_fftfreq: bad argument");
    const int odd = n & 1, half = odd ? (n - 1) >> 1 : n >> 1;
    const int count = rfft ? half + 1 : n;
    const int neg_from = rfft ? count : half + odd;
    const int blocks = (count + 255) / 256 < 1184 ? (count + 255) / 256 : 1184;
    if (dtype == DSC_CUDA_F32) {
        const float factor = 1 / ((float)n * (float)d);
        DSC_LAUNCH(fill_fftfreq<float>, blocks, 256, 0, stream, (float *)out, count, neg_from, factor);
    } else if (dtype == DSC_CUDA_F64) {
        const double factor = 1 / ((double)n * d);
        DSC_LAUNCH(fill_fftfreq<double>, blocks, 256, 0, stream, (double *)out, count, neg_from, factor);
    } else return fail(DSC_CUDA_EINVAL, "dsc_cuda_fftfreq: dtype must be real");
    return check_launch("fill_fftfreq");
}

int dsc_cuda_gather(const void *in, void *out, int elem_bytes, const int shape[4], const int64_t stride[4], int64_t base, void *stream) {
    Index4 g;
    long long total;
    if (!in || !out || !shape || !stride || !make_index4(g, shape, stride, base, total)) return fail(DSC_CUDA_EINVAL, "dsc_cuda_gather: bad argument");
    if (elem_bytes == 4) return gather_typed<float>(in, out, g, total, stream);
    if (elem_bytes == 8) return gather_typed<float2>(in, out, g, total, stream);
    if (elem_bytes == 16) return gather_typed<double2>(in, out, g, total, stream);
    return fail(DSC_CUDA_EINVAL, "dsc_cuda_gather: element size %d", elem_bytes);
}

int dsc_cuda_scatter(void *dst, const void *src, int elem_bytes, const int shape[4], const int64_t stride[4], int64_t base,
                     int64_t src_count, void *stream) {
    Index4 g;
    long long total;
    if (!dst || !src || !shape || !stride || src_count < 1 || !make_index4(g, shape, stride, base, total))
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_scatter: bad argument");
    if (elem_bytes == 4) return scatter_typed<float>(dst, src, g, total, (long long)src_count, stream);
    if (elem_bytes == 8) return scatter_typed<float2>(dst, src, g, total, (long long)src_count, stream);
    if (elem_bytes == 16) return scatter_typed<double2>(dst, src, g, total, (long long)src_count, stream);
    return fail(DSC_CUDA_EINVAL, "dsc_cuda_scatter: element size %d", elem_bytes);
}

int dsc_cuda_transpose_batched(const void *in, void *out, int64_t batches, int64_t rows, int64_t cols, int elem_bytes, void *stream) {
    if (!in || !out || batches <= 0 || rows <= 0 || cols <= 0 || rows > 0x7fffffff || cols > 0x7fffffff)
        return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_batched: bad argument");
    const long long per = ((rows + 31) / 32) * ((cols + 31) / 32);
    if (per * batches > 0x7fffffffLL) return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_batched: too many tiles");
    const unsigned grid = (unsigned)(per * batches);
    if (elem_bytes == 4) { auto k = transpose_batched<float>; DSC_LAUNCH(k, grid, 256, 0, stream, (const float *)in, (float *)out, (int)rows, (int)cols, per); }
    else if (elem_bytes == 8) { auto k = transpose_batched<float2>; DSC_LAUNCH(k, grid, 256, 0, stream, (const float2 *)in, (float2 *)out, (int)rows, (int)cols, per); }
    else if (elem_bytes == 16) { auto k = transpose_batched<double2>; DSC_LAUNCH(k, grid, 256, 0, stream, (const double2 *)in, (double2 *)out, (int)rows, (int)cols, per); }
    else return fail(DSC_CUDA_EINVAL, "dsc_cuda_transpose_batched: element size %d", elem_bytes);
    return check_launch("transpose_batched");
}

}  // extern "C"
