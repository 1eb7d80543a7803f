// dsc.h -- public C ABI of libdsc.so (B200-native build).
//
// This header declares the same 60 entry points, types and layouts that the reference
// library exports (/root/reference/dsc/include/dsc.h:85-428, dsc_dtype.h), so that the
// reference's C++ wrapper (dsc/api/dsc_api.h) compiles against it unchanged and its Python
// ctypes wrapper (python/dsc/_bindings.py) binds the resulting libdsc.so unchanged.
// It is written from the ABI contract, not from the reference's header: struct layouts,
// enum values, argument order and default arguments are the contract; everything else here
// (grouping, helper traits, the declaration macros) is this project's own.
//
// What differs behind the ABI: the FFT entry points run hand-written sm_100a kernels
// (dsc_cuda.h) on a device arena reserved once by dsc_ctx_init; there is no CPU FFT.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <type_traits>

// ---------------------------------------------------------------------------------------
// scalar types (reference: dsc_dtype.h:22-56)

using i8 = int8_t;   using i16 = int16_t;  using i32 = int32_t;  using i64 = int64_t;
using u8 = uint8_t;  using u16 = uint16_t; using u32 = uint32_t; using u64 = uint64_t;
using f32 = float;   using f64 = double;
using size = ptrdiff_t;
using usize = size_t;
using byte = char;

// interleaved {real, imag}; c32 = 2 x f32 (numpy complex64), c64 = 2 x f64 (numpy complex128)
template <typename T> struct dsc_complex_t {
    union {
        T d[2];
        struct { T real, imag; };
    };
};
using c32 = dsc_complex_t<f32>;
using c64 = dsc_complex_t<f64>;
#define dsc_complex(type, re, im) (type{.d = {(re), (im)}})

enum dsc_dtype : u8 { F32 = 0, F64 = 1, C32 = 2, C64 = 3 };
#define DSC_DTYPES ((int) 4)
#define DSC_DEFAULT_TYPE (dsc_dtype::F32)

constexpr static usize DSC_DTYPE_SIZE[DSC_DTYPES] = {sizeof(f32), sizeof(f64), sizeof(c32), sizeof(c64)};
constexpr static const char *DSC_DTYPE_NAMES[DSC_DTYPES] = {"f32", "f64", "c32", "c64"};

// result dtype of a binary op (reference: dsc_dtype.h:73-78)
constexpr static dsc_dtype DSC_DTYPE_CONVERSION_TABLE[DSC_DTYPES][DSC_DTYPES] = {
    {F32, F64, C32, C64},
    {F64, F64, C32, C64},
    {C32, C32, C32, C64},
    {C64, C64, C64, C64},
};

template <typename T> struct dsc_type_mapping;
template <> struct dsc_type_mapping<f32> { static constexpr dsc_dtype value = F32; };
template <> struct dsc_type_mapping<f64> { static constexpr dsc_dtype value = F64; };
template <> struct dsc_type_mapping<c32> { static constexpr dsc_dtype value = C32; };
template <> struct dsc_type_mapping<c64> { static constexpr dsc_dtype value = C64; };

template <typename T> struct dsc_real_of { using type = T; };
template <> struct dsc_real_of<c32> { using type = f32; };
template <> struct dsc_real_of<c64> { using type = f64; };
template <typename T> using real = typename dsc_real_of<T>::type;

template <typename A, typename B> static consteval bool dsc_is_type() noexcept { return std::is_same_v<A, B>; }
template <typename T> static consteval bool dsc_is_complex() noexcept { return dsc_is_type<T, c32>() || dsc_is_type<T, c64>(); }
template <typename T> static consteval bool dsc_is_real() noexcept { return dsc_is_type<T, f32>() || dsc_is_type<T, f64>(); }

template <typename T> static consteval T dsc_pi() noexcept { return (T) 3.14159265358979323846264338327950288L; }
template <typename T> static consteval T dsc_zero() noexcept {
    if constexpr (dsc_is_complex<T>()) return T{.d = {0, 0}};
    else return (T) 0;
}
template <typename T, bool positive = true> static consteval T dsc_inf() noexcept {
    constexpr real<T> v = (positive ? 1 : -1) * std::numeric_limits<real<T>>::infinity();
    if constexpr (dsc_is_complex<T>()) return T{.d = {v, v}};
    else return v;
}

// ---------------------------------------------------------------------------------------
// logging / assertions: failures print to stderr and exit(EXIT_FAILURE), like the
// reference (dsc.h:14-38).  CUDA failures follow the same convention.

#define DSC_LOG_ERR(format, ...)   fprintf(stderr, "%s: " format "\n", __func__, ##__VA_ARGS__)
#define DSC_LOG_INFO(format, ...)  fprintf(stdout, "%s: " format "\n", __func__, ##__VA_ARGS__)
#define DSC_LOG_FATAL(format, ...) do { DSC_LOG_ERR(format, ##__VA_ARGS__); exit(EXIT_FAILURE); } while (0)
#if defined(DSC_DEBUG)
#   define DSC_LOG_DEBUG(format, ...) DSC_LOG_INFO(format, ##__VA_ARGS__)
#else
#   define DSC_LOG_DEBUG(format, ...) ((void) 0)
#endif
#define DSC_ASSERT(x)                                                               \
    do {                                                                            \
        if (!(x)) {                                                                 \
            fprintf(stderr, "DSC_ASSERT: %s:%d %s\n", __FILE__, __LINE__, #x);      \
            exit(EXIT_FAILURE);                                                     \
        }                                                                           \
    } while (0)
#define DSC_INVALID_CASE(format, ...) default: DSC_LOG_FATAL(format, ##__VA_ARGS__)

#define DSC_UNUSED(x)     ((void) (x))
#define DSC_ALIGN(x, y)   (((x) + (y) - 1) & ~((y) - 1))
#define DSC_MAX(x, y)     ((x) > (y) ? (x) : (y))
#define DSC_MIN(x, y)     ((x) < (y) ? (x) : (y))
#define DSC_B_TO_KB(b)    ((f64) (b) / 1024.)
#define DSC_B_TO_MB(b)    ((f64) (b) / (1024. * 1024.))
#define DSC_KB(kb)        ((usize) ((kb) * 1024l))
#define DSC_MB(mb)        ((usize) ((mb) * 1024l * 1024l))

#if defined(__GNUC__)
#   define DSC_INLINE        inline __attribute__((always_inline))
#   define DSC_NOINLINE      __attribute__((noinline))
#   define DSC_STRICTLY_PURE __attribute__((const))
#   define DSC_PURE          __attribute__((pure))
#   define DSC_MALLOC        __attribute__((malloc))
#else
#   define DSC_INLINE        inline
#   define DSC_NOINLINE
#   define DSC_STRICTLY_PURE
#   define DSC_PURE
#   define DSC_MALLOC
#endif
#define DSC_RESTRICT __restrict

#if !defined(DSC_MAX_DIMS)
#   define DSC_MAX_DIMS ((int) 4)
#endif
static_assert(DSC_MAX_DIMS == 4, "tensors are at most 4-dimensional (ABI)");

// "not given" marker for slice fields and for the concat axis (flatten)
#define DSC_VALUE_NONE INT32_MAX

// Position of user dimension `dim` in the right-aligned shape[4]; negative dims count from the end.
#define dsc_tensor_dim(PTR, dim) (((dim) < 0) ? (DSC_MAX_DIMS + (dim)) : (DSC_MAX_DIMS - (PTR)->n_dim + (dim)))
#define DSC_TENSOR_DATA(T, PTR)  T *DSC_RESTRICT PTR##_data = (T *) (PTR)->data
#define dsc_new_like(CTX, PTR)   (dsc_new_tensor((CTX), (PTR)->n_dim, &(PTR)->shape[dsc_tensor_dim(PTR, 0)], (PTR)->dtype))
#define dsc_new_view(CTX, PTR)   (dsc_new_tensor((CTX), (PTR)->n_dim, &(PTR)->shape[dsc_tensor_dim(PTR, 0)], (PTR)->dtype, (PTR)->buffer))

extern "C" {

struct dsc_ctx;
struct dsc_fft_plan;
struct dsc_tensor_buffer;                   // first member: int refs (python/dsc/_bindings.py:38-41)
enum dsc_fft_type : u8;                     // REAL = 0, COMPLEX = 1      (enumerators: dsc_runtime.h)
enum dsc_backend_type : u8;                 // CPU = 0 (host-only context), CUDA = 1

// 64 bytes; mirrored field by field by the Python wrapper (python/dsc/_bindings.py:44-54).
struct dsc_tensor {
    int shape[DSC_MAX_DIMS];                // right-aligned: a 1-D tensor of 4 elements is [1, 1, 1, 4]
    int stride[DSC_MAX_DIMS];               // in ELEMENTS, row-major contiguous
    dsc_tensor_buffer *buffer;              // refcounted payload owner
    void *data;                             // HOST address of the payload (wrappers memcpy through it)
    int ne;                                 // number of elements
    int n_dim;
    dsc_dtype dtype;
    dsc_backend_type backend;
};

// passed BY VALUE through the variadic slice entry points (12 bytes)
struct dsc_slice {
    union {
        int d[3];
        struct { int start, stop, step; };
    };
};

// Smallest power of two >= n (n > 0).  FFT lengths are rounded UP with this.
static DSC_INLINE DSC_STRICTLY_PURE int dsc_pow2_n(const int n) noexcept {
    DSC_ASSERT(n > 0);
    u32 v = (u32) (n - 1);
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return (int) (v + 1);
}

// ---- context ---------------------------------------------------------------------------
// main_mem / scratch_mem size the two host arenas exactly as in the reference; in addition ONE
// device allocation (tensor mirrors + plans + scratch) is made here and none afterwards.
extern dsc_ctx *dsc_ctx_init(usize main_mem, usize scratch_mem) noexcept;
extern void dsc_ctx_free(dsc_ctx *ctx) noexcept;
extern void dsc_ctx_clear(dsc_ctx *ctx) noexcept;
extern usize dsc_used_mem(dsc_ctx *ctx) noexcept;
extern void dsc_print_mem_usage(dsc_ctx *ctx) noexcept;

// Plans are owned by the context: at most DSC_MAX_FFT_PLANS (compile-time, default 16) live at once,
// keyed by (power-of-two n, type, f32|f64); the least recently used one is evicted.
extern dsc_fft_plan *dsc_plan_fft(dsc_ctx *ctx, int n, dsc_fft_type fft_type,
                                  dsc_dtype dtype = dsc_dtype::F64) noexcept;

// ---- tracing (Chrome / Perfetto JSON) ------------------------------------------------------
extern void dsc_traces_record(dsc_ctx *, bool record = true) noexcept;
extern void dsc_dump_traces(dsc_ctx *, const char *filename) noexcept;
extern void dsc_clear_traces(dsc_ctx *) noexcept;

// ---- tensors -------------------------------------------------------------------------------
extern DSC_MALLOC dsc_tensor *dsc_new_tensor(dsc_ctx *ctx, int n_dim, const int *shape, dsc_dtype dtype,
                                             dsc_tensor_buffer *buffer = nullptr) noexcept;
extern DSC_MALLOC dsc_tensor *dsc_view(dsc_ctx *ctx, const dsc_tensor *x) noexcept;
extern void dsc_tensor_free(dsc_ctx *ctx, dsc_tensor *x) noexcept;
extern dsc_tensor *dsc_tensor_1d(dsc_ctx *ctx, dsc_dtype dtype, int dim1) noexcept;
extern dsc_tensor *dsc_tensor_2d(dsc_ctx *ctx, dsc_dtype dtype, int dim1, int dim2) noexcept;
extern dsc_tensor *dsc_tensor_3d(dsc_ctx *ctx, dsc_dtype dtype, int dim1, int dim2, int dim3) noexcept;
extern dsc_tensor *dsc_tensor_4d(dsc_ctx *ctx, dsc_dtype dtype, int dim1, int dim2, int dim3, int dim4) noexcept;
extern dsc_tensor *dsc_wrap_f32(dsc_ctx *ctx, f32 val) noexcept;
extern dsc_tensor *dsc_wrap_f64(dsc_ctx *ctx, f64 val) noexcept;
extern dsc_tensor *dsc_wrap_c32(dsc_ctx *ctx, c32 val) noexcept;
extern dsc_tensor *dsc_wrap_c64(dsc_ctx *ctx, c64 val) noexcept;
extern dsc_tensor *dsc_arange(dsc_ctx *ctx, int n, dsc_dtype dtype = DSC_DEFAULT_TYPE) noexcept;
extern dsc_tensor *dsc_randn(dsc_ctx *ctx, int n_dim, const int *shape, dsc_dtype dtype = DSC_DEFAULT_TYPE) noexcept;
extern dsc_tensor *dsc_cast(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT x, dsc_dtype new_dtype) noexcept;
extern dsc_tensor *dsc_reshape(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, int dimensions...) noexcept;
extern dsc_tensor *dsc_concat(dsc_ctx *ctx, int axis, int tensors...) noexcept;
extern dsc_tensor *dsc_transpose(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, int axes...) noexcept;

// ---- indexing / slicing: results are copies (NumPy semantics, no views) ----------------------
extern dsc_tensor *dsc_tensor_get_idx(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, int indexes...) noexcept;
extern dsc_tensor *dsc_tensor_get_slice(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, int slices...) noexcept;
extern void dsc_tensor_set_idx(dsc_ctx *, dsc_tensor *DSC_RESTRICT xa, const dsc_tensor *DSC_RESTRICT xb, int indexes...) noexcept;
extern void dsc_tensor_set_slice(dsc_ctx *, dsc_tensor *DSC_RESTRICT xa, const dsc_tensor *DSC_RESTRICT xb, int slices...) noexcept;

// ---- element-wise families -------------------------------------------------------------------
#define DSC_DECL_BINARY(name) \
    extern dsc_tensor *name(dsc_ctx *ctx, dsc_tensor *xa, dsc_tensor *xb, dsc_tensor *out = nullptr) noexcept;
DSC_DECL_BINARY(dsc_add) DSC_DECL_BINARY(dsc_sub) DSC_DECL_BINARY(dsc_mul) DSC_DECL_BINARY(dsc_div) DSC_DECL_BINARY(dsc_pow)
#undef DSC_DECL_BINARY

#define DSC_DECL_UNARY(name) \
    extern dsc_tensor *name(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out = nullptr) noexcept;
DSC_DECL_UNARY(dsc_cos)  DSC_DECL_UNARY(dsc_sin)   DSC_DECL_UNARY(dsc_sinc) DSC_DECL_UNARY(dsc_logn) DSC_DECL_UNARY(dsc_log2)
DSC_DECL_UNARY(dsc_log10) DSC_DECL_UNARY(dsc_exp)  DSC_DECL_UNARY(dsc_sqrt) DSC_DECL_UNARY(dsc_abs)
#undef DSC_DECL_UNARY

extern dsc_tensor *dsc_angle(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x) noexcept;
extern dsc_tensor *dsc_conj(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT x) noexcept;   // real input: returns x itself
extern dsc_tensor *dsc_real(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT x) noexcept;   // real input: returns x itself
extern dsc_tensor *dsc_imag(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x) noexcept;
extern dsc_tensor *dsc_i0(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x) noexcept;
extern dsc_tensor *dsc_clip(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out = nullptr,
                            f64 x_min = dsc_inf<f64, false>(), f64 x_max = dsc_inf<f64, true>()) noexcept;

#define DSC_DECL_REDUCE(name) \
    extern dsc_tensor *name(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out = nullptr, \
                            int axis = -1, bool keep_dims = true) noexcept;
DSC_DECL_REDUCE(dsc_sum) DSC_DECL_REDUCE(dsc_mean) DSC_DECL_REDUCE(dsc_max) DSC_DECL_REDUCE(dsc_min)
#undef DSC_DECL_REDUCE

// ---- Fourier transforms (the hot path) ---------------------------------------------------------
// Always out of place.  `out` (optional) must have the result's dtype / n_dim / shape.  `axis` selects
// the dimension; `n` <= 0 means "the axis extent", otherwise the axis is zero-padded / cropped to n.
// Lengths are rounded UP to a power of two.  For dsc_irfft, n counts INPUT BINS.
#define DSC_DECL_FFT(name) \
    extern dsc_tensor *name(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out = nullptr, \
                            int n = -1, int axis = -1) noexcept;
DSC_DECL_FFT(dsc_fft) DSC_DECL_FFT(dsc_ifft) DSC_DECL_FFT(dsc_rfft) DSC_DECL_FFT(dsc_irfft)
#undef DSC_DECL_FFT

extern dsc_tensor *dsc_fftfreq(dsc_ctx *ctx, int n, f64 d = 1., dsc_dtype dtype = DSC_DEFAULT_TYPE) noexcept;
extern dsc_tensor *dsc_rfftfreq(dsc_ctx *ctx, int n, f64 d = 1., dsc_dtype dtype = DSC_DEFAULT_TYPE) noexcept;

// ---- additions of this build (not in the reference; optional for callers) ------------------------
// Fused filter pipeline  out = irfft(rfft(x, n) * B)  with B = rfft(b, n) precomputed (README.md:118-134):
// one launch sequence, the spectrum never leaves the GPU.  Output has 2*order samples per line.
extern dsc_tensor *dsc_fft_filter(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const dsc_tensor *DSC_RESTRICT B,
                                  dsc_tensor *DSC_RESTRICT out = nullptr, int n = -1, int axis = -1) noexcept;
// irfft / dsc_fft_filter along the last axis keeping only the first `keep` samples of every line: the README's
// `irfft(...)[:output_length]` (README.md:130-133, i.e. dsc_irfft followed by dsc_tensor_get_slice, dsc.cpp:950-1007)
// with the crop fused into the inverse kernel's store -- cropped samples are neither written nor downloaded.
extern dsc_tensor *dsc_irfft_keep(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out,
                                  int n, int axis, int keep) noexcept;
extern dsc_tensor *dsc_fft_filter_keep(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const dsc_tensor *DSC_RESTRICT B,
                                       dsc_tensor *DSC_RESTRICT out, int n, int axis, int keep) noexcept;
// Residency policy of FFT results (see DESIGN.md "Host-visible data vs device residency").
//   0 (default): every call uploads its inputs and downloads its outputs -- always coherent with host writes;
//   1: outputs stay valid on the device and are reused as inputs without upload (host writes through
//      raw `data` pointers into an FFT RESULT must be followed by dsc_cuda_touch_host);
//   2: like 1 and outputs are NOT downloaded until dsc_cuda_sync_host(tensor).
extern void dsc_cuda_set_residency(dsc_ctx *ctx, int mode) noexcept;
extern void dsc_cuda_sync_host(dsc_ctx *ctx, dsc_tensor *x) noexcept;
extern void dsc_cuda_touch_host(dsc_ctx *ctx, dsc_tensor *x) noexcept;
// residency 2: start downloading x's payload and return; later calls overlap with the copy (uploads of the next
// transform run in the other PCIe direction); dsc_cuda_sync_host(x) or any host access of x waits for it
extern void dsc_cuda_download_async(dsc_ctx *ctx, dsc_tensor *x) noexcept;
// residency >= 1: upload x's payload to its device mirror now (e.g. a filter spectrum or a window that
// later device ops will read); no effect in strict mode
extern void dsc_cuda_prefetch(dsc_ctx *ctx, dsc_tensor *x) noexcept;
extern usize dsc_cuda_used_mem(dsc_ctx *ctx) noexcept;       // device-arena bytes in use
extern usize dsc_cuda_alloc_calls(dsc_ctx *ctx) noexcept;    // device allocations made so far (stays 1)

}  // extern "C"
