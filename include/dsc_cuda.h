/* dsc_cuda.h -- the thin extern "C" layer between DSC's host C++ and the sm_100a kernels.
 *
 * This is the only header where device pointers and streams appear; dsc.h stays CUDA-free.
 * Everything is plain C: raw pointers, 64-bit counts, `void *stream` (a cudaStream_t).
 * No function here allocates device memory: plans and work buffers are carved by the caller
 * from the device arena that dsc_ctx_init reserves once (dsc_arena.cpp).
 *
 * Reference interfaces replaced (paths relative to /root/reference):
 *   dsc_cuda_plan_bytes  <- dsc_fft_storage             dsc/include/dsc_fft.h:109-135
 *   dsc_cuda_plan_build  <- dsc_init_plan               dsc/include/dsc_fft.h:33-55,137-154
 *   dsc_cuda_fft         <- exec_fft + dsc_complex_fft  dsc/src/dsc.cpp:1958-2007, dsc_fft.h:156-176
 *   dsc_cuda_rfft/irfft  <- exec_rfft + dsc_real_fft    dsc/src/dsc.cpp:2102-2171, dsc_fft.h:178-238
 *   dsc_cuda_cmul        <- binary_op<mul_op> (complex) dsc/src/dsc.cpp:1186-1245, dsc_ops.h:68-78
 *   dsc_cuda_unary       <- dsc_abs / dsc_angle / dsc_real / dsc_imag / dsc_conj  dsc.cpp:1480-1622
 *   dsc_cuda_binary      <- dsc_add / dsc_sub / dsc_mul / dsc_div (same dtype)     dsc.cpp:1186-1310
 *   dsc_cuda_filter      <- README filterFFT            README.md:118-134 (rfft, *, irfft fused)
 *
 * Every entry point returns 0 on success or a negative DSC_CUDA_E* code; the message of
 * the last failure is available from dsc_cuda_last_error().  Launches are asynchronous on
 * `stream`; the caller synchronises.
 */
#ifndef DSC_CUDA_H
#define DSC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same numeric values as dsc_dtype (dsc_dtype.h:51-56) and dsc_fft_type (dsc_fft.h:13-16) */
enum { DSC_CUDA_F32 = 0, DSC_CUDA_F64 = 1, DSC_CUDA_C32 = 2, DSC_CUDA_C64 = 3 };
enum { DSC_CUDA_FFT_REAL = 0, DSC_CUDA_FFT_COMPLEX = 1 };

enum {
    DSC_CUDA_OK = 0,
    DSC_CUDA_EINVAL = -1,     /* bad argument (dtype, length, null pointer) */
    DSC_CUDA_ENOMEM = -2,     /* caller-provided plan / work memory too small */
    DSC_CUDA_ELAUNCH = -3,    /* CUDA runtime error, see dsc_cuda_last_error() */
    DSC_CUDA_EUNSUPPORTED = -4
};

#define DSC_CUDA_MAX_STAGES 5

/* A plan = the radix schedule for a power-of-two length plus its twiddle tables in HBM.
 * `n` is the COMPLEX transform length; a REAL plan of order n serves 2n-sample real
 * transforms and carries one more table (W_2n^k), like the reference's REAL plans. */
typedef struct dsc_cuda_plan {
    int n, lg_n;
    int fft_type;               /* DSC_CUDA_FFT_REAL | DSC_CUDA_FFT_COMPLEX */
    int dtype;                  /* DSC_CUDA_F32 | DSC_CUDA_F64: precision of the tables */
    int lg_n1, lg_n2;           /* n = n1 * n2; lg_n2 == 0: single shared-memory pass */
    int four_shift;             /* split point of the inter-pass twiddle index */
    void *dev_base;             /* block handed to dsc_cuda_plan_build */
    size_t dev_bytes;
    void *tw1[DSC_CUDA_MAX_STAGES];   /* stage tables of the length-n1 transform */
    void *tw2[DSC_CUDA_MAX_STAGES];   /* stage tables of the length-n2 transform */
    void *tw_lo, *tw_hi;        /* W_n^p split tables for the inter-pass twiddle */
    void *tw_real;              /* REAL plans, single pass: W_2n^k, k <= n/2 */
    void *tw_real_lo, *tw_real_hi;  /* REAL plans, two-pass: W_2n^p split the same way */
    int real_shift;
    /* Transforms along a NON-last axis run both passes of a two-pass decomposition as column passes.  Two-pass
     * plans reuse their own tables (the col_* fields alias them); single-pass COMPLEX plans of >= 2^13 points,
     * whose strided single-pass blocks would only touch 8-16 contiguous bytes per column, carry these extra
     * tables.  col_lg_n2 == 0: no column decomposition. */
    int col_lg_n1, col_lg_n2, col_shift;
    void *col_tw1[DSC_CUDA_MAX_STAGES], *col_tw2[DSC_CUDA_MAX_STAGES];
    void *col_lo, *col_hi;
    /* float two-pass plans with both factors <= 512: stage tables of the 16-points-per-thread schedule of the TMA-fed
     * launch (the tw1 / tw2 tables above are laid out for 32 points per thread).  NULL otherwise. */
    void *tw1_e16[DSC_CUDA_MAX_STAGES], *tw2_e16[DSC_CUDA_MAX_STAGES];
} dsc_cuda_plan;

const char *dsc_cuda_last_error(void);
int dsc_cuda_device_count(void);

/* Device bytes a plan for (n, type, dtype) needs.  n must be a power of two >= 1. */
size_t dsc_cuda_plan_bytes(int n, int fft_type, int dtype);

/* Fill `plan` and compute its tables into dev_mem (>= dsc_cuda_plan_bytes, 256-B aligned). */
int dsc_cuda_plan_build(dsc_cuda_plan *plan, int n, int fft_type, int dtype,
                        void *dev_mem, size_t dev_bytes, void *stream);

/* Work bytes needed to transform `lines` lines with this plan (0 for single-pass plans).
 * Any smaller non-zero amount that holds at least one line also works: the launch is
 * chunked.  Two-pass plans keep their intermediate here so it can stay L2-resident. */
size_t dsc_cuda_work_bytes(const dsc_cuda_plan *plan, int64_t lines);

/* The same for a transform along the middle axis of (outer, n, inner): for inner > 1 two-pass plans run both
 * passes as column passes over chunks of the inner extent and keep the chunk's intermediate here.  0 when the
 * shape is not covered (inner not a power of two or narrower than a tile): dsc_cuda_fft then reports
 * DSC_CUDA_EUNSUPPORTED and the caller transposes. */
size_t dsc_cuda_work_bytes_axis(const dsc_cuda_plan *plan, int64_t outer, int64_t inner);

/* fft / ifft along the middle axis of a contiguous (outer, x_n, inner) tensor.
 * x_dtype in {F32,F64,C32,C64}; out is complex of the plan's precision with extent plan->n
 * along the axis.  min(x_n, n) elements are read per line, the rest is zero (pad / crop). */
int dsc_cuda_fft(const dsc_cuda_plan *plan, const void *x, int x_dtype, void *out,
                 int64_t outer, int x_n, int64_t inner, int forward,
                 void *work, size_t work_bytes, void *stream);

/* fft / ifft of `lines` complex lines of plan->n points whose storage is SEGMENTED: line r consists of
 * n / seg_len segments of seg_len contiguous elements, segment s of line r at x + s*seg_stride + r*seg_len
 * -- the receive buffer [peer][line][part] of the multi-GPU four-step's all-to-all, transformed without
 * first un-interleaving it.  self_seg >= 0: that segment is read from self_x instead (same layout and element
 * alignment as x) -- the slab a rank "sends to itself" stays in the send buffer and never moves.
 * Two-pass plans only; seg_len a power of two >= n / 32 (float) or n / 16 (double). */
int dsc_cuda_fft_segmented(const dsc_cuda_plan *plan, const void *x, void *out, int64_t lines,
                           int64_t seg_len, int64_t seg_stride, int self_seg, const void *self_x, int forward,
                           void *work, size_t work_bytes, void *stream);

/* First local step of the multi-GPU four-step, in one launch: x is the rank's natural-order column block
 * [n][cols] (element (i, c) = x_global[i * all_cols + col_offset + c]); out[k][c] = FFT over i of column c,
 * times W_total^((col_offset + c) * k) -- the twiddled, k-major matrix whose row blocks are the all-to-all's
 * send slabs.  tw_lo / tw_hi / shift: the split tables of W_total (dsc_cuda_fill_twiddles).  The inverse
 * (conjugate twiddles) carries the plan's own 1/n like dsc_cuda_fft.  DSC_CUDA_EUNSUPPORTED when the plan has no column decomposition or cols is not a power of two at
 * least one tile wide (the caller then uses dsc_cuda_fft + dsc_cuda_transpose_twiddle).
 * work: dsc_cuda_work_bytes_axis(plan, 1, cols). */
int dsc_cuda_fft_columns_twiddled(const dsc_cuda_plan *plan, const void *x, void *out, int64_t cols, int forward,
                                  int64_t col_offset, const void *tw_lo, const void *tw_hi, int shift, int64_t total,
                                  void *work, size_t work_bytes, void *stream);

/* The same launch with the exchange FUSED into its epilogue: instead of one send buffer, row block q of the k-major
 * result (rows [q n / n_peers, (q+1) n / n_peers), row pitch cols) is stored straight at peer_out[q] -- a pointer into
 * the receive buffer of the GPU that owns those rows, mapped into this process (NVLink peer memory; peer_out[own rank]
 * is local).  There is no separate all-to-all afterwards: the transfer overlaps the transform tile by tile.  The caller
 * synchronises the ranks before the buffers are read.  n_peers a power of two <= 8. */
int dsc_cuda_fft_columns_twiddled_p2p(const dsc_cuda_plan *plan, const void *x, int64_t cols, int forward,
                                      int64_t col_offset, const void *tw_lo, const void *tw_hi, int shift, int64_t total,
                                      void *const *peer_out, int n_peers, void *work, size_t work_bytes, void *stream);

/* rfft: real (outer, x_n, inner) -> complex (outer, n + 1, inner), plan REAL of order n. */
int dsc_cuda_rfft(const dsc_cuda_plan *plan, const void *x, void *out,
                  int64_t outer, int x_n, int64_t inner,
                  void *work, size_t work_bytes, void *stream);

/* irfft: complex (outer, x_n, inner) bins -> real (outer, 2n, inner); min(x_n, n+1) bins read. */
int dsc_cuda_irfft(const dsc_cuda_plan *plan, const void *x, void *out,
                   int64_t outer, int x_n, int64_t inner,
                   void *work, size_t work_bytes, void *stream);

/* Fused filter pipeline (README.md:118-134): out = irfft(rfft(x, 2n) * spectrum) per line of a contiguous
 * (outer, x_n) real tensor; `spectrum` holds n+1 complex bins and is broadcast over lines; out is
 * (outer, 2n) real.  Orders that fit one shared-memory pass run as ONE kernel (the line's spectrum never
 * leaves shared memory); larger orders as fused four-step forward, one bin-pair kernel, fused four-step
 * inverse.  work: dsc_cuda_filter_work_bytes(plan, lines). */
size_t dsc_cuda_filter_work_bytes(const dsc_cuda_plan *plan, int64_t lines);
int dsc_cuda_filter(const dsc_cuda_plan *plan, const void *x, const void *spectrum, void *out,
                    int64_t outer, int x_n, void *work, size_t work_bytes, void *stream);

/* irfft / fused filter along the last axis storing only the first `keep` (<= 2n) real samples of every line; out is
 * (outer, keep).  The crop of README.md:130-133 (`y[:output_length]`, dsc_tensor_get_slice dsc.cpp:950-1007) fused into
 * the inverse kernel's store: the cropped samples are never written.  Orders of one shared-memory pass only
 * (DSC_CUDA_EUNSUPPORTED otherwise: the caller transforms in full and crops with dsc_cuda_gather). */
int dsc_cuda_irfft_keep(const dsc_cuda_plan *plan, const void *x, void *out, int64_t outer, int x_n, int keep,
                        void *work, size_t work_bytes, void *stream);
int dsc_cuda_filter_keep(const dsc_cuda_plan *plan, const void *x, const void *spectrum, void *out,
                         int64_t outer, int x_n, int keep, void *work, size_t work_bytes, void *stream);

/* out = a * b elementwise on complex rows; b has `cols` elements (b_rows == 0, broadcast)
 * or rows*cols (b_rows != 0). */
int dsc_cuda_cmul(const void *a, const void *b, void *out, int dtype,
                  int64_t rows, int64_t cols, int b_rows, void *stream);

/* Elementwise post-processing of device-resident data (one read + one write of the payload each).
 * unary  : x complex (C32/C64) -> ABS, ANGLE, REAL, IMAG give the real dtype, CONJ the same complex dtype
 *          (dsc_abs / dsc_angle / dsc_real / dsc_imag / dsc_conj, dsc/src/dsc.cpp:1480-1622).
 * binary : out = a OP b for F32/F64/C32/C64 (both operands and out of `dtype`), rows x cols elements;
 *          b_mode 0: b is one row of `cols` elements broadcast over the rows, 1: same shape, 2: b is one element
 *          (dsc_add / dsc_sub / dsc_mul / dsc_div, dsc.cpp:1186-1310, dsc_ops.h:46-90).  out may alias a. */
enum { DSC_CUDA_OP_ABS = 0, DSC_CUDA_OP_ANGLE = 1, DSC_CUDA_OP_REAL = 2, DSC_CUDA_OP_IMAG = 3, DSC_CUDA_OP_CONJ = 4 };
enum { DSC_CUDA_OP_ADD = 0, DSC_CUDA_OP_SUB = 1, DSC_CUDA_OP_MUL = 2, DSC_CUDA_OP_DIV = 3 };
int dsc_cuda_unary(int op, const void *x, int x_dtype, void *out, int64_t count, void *stream);
int dsc_cuda_binary(int op, const void *a, const void *b, void *out, int dtype,
                    int64_t rows, int64_t cols, int b_mode, void *stream);

/* Cast between any two of F32 / F64 / C32 / C64 (cast_op, dsc/include/dsc_ops.h:12-44: real -> complex sets imag = 0,
 * complex -> real keeps the real part).  dsc_cast, dsc/src/dsc.cpp:587-597. */
int dsc_cuda_cast(const void *x, int x_dtype, void *out, int out_dtype, int64_t count, void *stream);

/* out = a OP b with operands of DIFFERENT dtypes: both are promoted to out_dtype = table[a_dtype][b_dtype]
 * (dsc_dtype.h:73-78) in registers, where the reference casts them through scratch first (dsc.cpp:44-69).
 * rows / cols / b_mode as in dsc_cuda_binary. */
int dsc_cuda_binary_mixed(int op, const void *a, int a_dtype, const void *b, int b_dtype, void *out,
                          int64_t rows, int64_t cols, int b_mode, void *stream);

/* fftfreq (rfft == 0): out[i] = i / (n d) for i < n - n/2, (i - n) / (n d) above, n values;
 * rfftfreq (rfft != 0): out[i] = i / (n d), n/2 + 1 values.  dtype F32 / F64; the factor 1 / (n d) is formed and
 * applied in that precision like dsc_internal_fftfreq, dsc/src/dsc.cpp:2262-2339. */
int dsc_cuda_fftfreq(void *out, int dtype, int n, double d, int rfft, void *stream);

/* Strided gather / scatter over a right-aligned 4-D index space -- the device side of dsc_transpose (dsc.cpp:764-827),
 * dsc_tensor_get_slice (:950-1007) and dsc_tensor_set_slice (:1108-1169).  shape[4] are the extents of the dense side
 * (row-major), stride[4] / base the element strides / offset of the strided side; elem_bytes in {4, 8, 16}.
 *   gather : out[dense i] = in[base + sum idx_d stride_d]
 *   scatter: dst[base + sum idx_d stride_d] = src[dense i mod src_count]  (the source is recycled when shorter) */
int dsc_cuda_gather(const void *in, void *out, int elem_bytes, const int shape[4], const int64_t stride[4], int64_t base, void *stream);
int dsc_cuda_scatter(void *dst, const void *src, int elem_bytes, const int shape[4], const int64_t stride[4], int64_t base,
                     int64_t src_count, void *stream);

/* out[b][c][r] = in[b][r][c]: tiled transpose of the last two dims of `batches` matrices, elem_bytes in {4, 8, 16}. */
int dsc_cuda_transpose_batched(const void *in, void *out, int64_t batches, int64_t rows, int64_t cols, int elem_bytes, void *stream);

/* ---- building blocks of the multi-GPU four-step (one transform sharded over P GPUs) ----------
 * out[k] = exp(-2 pi i (k * mult mod denom) / denom), k < count: the two sqrt(M)-sized tables
 * lo[p] = W_M^p (mult 1) and hi[q] = W_M^(q << shift) (mult 1 << shift) give W_M^p for any p < M. */
int dsc_cuda_fill_twiddles(void *out, int64_t count, int64_t mult, int64_t denom, int dtype, void *stream);

/* out[c][r] = in[r][c] * W_M^((r0 + r) * c) (conjugated when !forward); in is rows x cols row-major.
 * tw_lo == NULL: plain transpose.  The rows of `out` that belong to peer p are contiguous. */
int dsc_cuda_transpose_twiddle(const void *in, void *out, int64_t rows, int64_t cols, int64_t r0,
                               const void *tw_lo, const void *tw_hi, int shift, int forward,
                               int dtype, void *stream);

/* out[c][r] = in[r][c] for rows x cols elements of elem_bytes in {4, 8, 16}. */
int dsc_cuda_transpose(const void *in, void *out, int64_t rows, int64_t cols, int elem_bytes, void *stream);

/* Transpose with cast to complex and zero padding: out[c][r] = (r*cols + c < limit) ? complex(in[r*cols + c]) : 0.
 * in_dtype in {F32,F64,C32,C64}; out is complex of the same precision. */
int dsc_cuda_transpose_cast(const void *in, int in_dtype, void *out, int64_t rows, int64_t cols, int64_t limit, void *stream);

#ifdef __cplusplus
}
#endif
#endif
