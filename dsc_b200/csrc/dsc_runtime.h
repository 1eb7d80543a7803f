// dsc_runtime.h -- internals of libdsc.so's host runtime: memory pools, tensor buffers with
// device residency, the plan cache and the tracer.  Nothing here is part of the ABI.
//
// Memory model (reference: two host buffers made once by dsc_ctx_init, dsc.cpp:150-180,
// generic + linear allocators in dsc_allocator.cpp):
//   host main arena    best-fit range allocator, tensor headers + payloads (page-locked)
//   host scratch arena bump allocator, reset before every op's temporaries (DSC_CTX_PUSH)
//   device arena       ONE allocation at init; [ range-allocated mirrors + plans | bump scratch ]
// Allocator bookkeeping lives OUT of band in node pools sized at init, because the
// reference's intrusive headers cannot live in device memory; no malloc happens after init.
#pragma once

#include "dsc.h"
#include "dsc_cuda.h"
#include "dsc_device.h"

enum dsc_fft_type : u8 { REAL = 0, COMPLEX = 1 };
enum dsc_backend_type : u8 { CPU = 0, CUDA = 1 };
constexpr static const char *DSC_BACKEND_NAMES[2] = {"CPU", "CUDA"};

#if !defined(DSC_MAX_FFT_PLANS)
#   define DSC_MAX_FFT_PLANS ((int) 16)
#endif
#if !defined(DSC_MAX_TRACES)
#   define DSC_MAX_TRACES ((u64) 1000)
#endif

// ---------------------------------------------------------------------------------------
// Best-fit allocator over an abstract byte range [0, capacity) with address-ordered
// coalescing.  Blocks are described by pool nodes; callers keep the node id.
struct dsc_range_alloc {
    struct node { usize off, size; int prev, next; bool used; };
    node *nodes;
    int capacity_nodes;
    int head;            // first block in address order
    int free_nodes;      // stack of unused node slots (linked through .next)
    usize capacity, used, granule;

    void init(usize bytes, usize granule_, int max_nodes) noexcept;
    void destroy() noexcept;
    void reset() noexcept;
    int alloc(usize bytes) noexcept;        // node id or -1
    void release(int id) noexcept;
    bool is_live(int id, usize off) const noexcept {
        return id >= 0 && id < capacity_nodes && nodes[id].used && nodes[id].off == off;
    }
  private:
    int take_node() noexcept;
    void give_node(int id) noexcept;
};

struct dsc_bump_alloc {
    usize capacity, top;
    void init(usize bytes) noexcept { capacity = bytes; top = 0; }
    void reset() noexcept { top = 0; }
    // offset or (usize)-1
    usize alloc(usize bytes, usize align) noexcept {
        const usize at = DSC_ALIGN(top, align);
        if (at + bytes > capacity) return (usize) -1;
        top = at + bytes;
        return at;
    }
};

// ---------------------------------------------------------------------------------------
// Payload owner.  Only `refs` is visible through the ABI (first member).
enum : int {
    DSC_BUF_DEV_VALID  = 1,   // device mirror holds the current contents
    DSC_BUF_HOST_STALE = 2,   // residency mode 2: host copy not yet downloaded
    DSC_BUF_SCRATCH    = 4,   // lives in the scratch arena (never freed individually)
    DSC_BUF_DOWNLOADING = 8,  // dsc_cuda_download_async: the device -> host copy is in flight on the download stream
};

struct dsc_tensor_buffer {
    int refs;
    int flags;
    int dev_node;                   // block in the device arena, -1 = no mirror
    int busy;                       // operand of the running op: its mirror must not be evicted
    usize nbytes;                   // payload bytes
    dsc_tensor_buffer *dev_prev, *dev_next;   // list of buffers that own a device mirror
    dscdev::Event *downloaded;      // DSC_BUF_DOWNLOADING: completes when the asynchronous download has landed
};

struct dsc_fft_plan {
    dsc_cuda_plan cu;               // n, type, dtype and the device tables
    int last_used;                  // ageing counter of the cache (dsc.cpp:199-213)
    int dev_node;
    // Lengths beyond the kernels' four-step range (> 2^24 complex64 / 2^23 complex128 points) are
    // composed on the host from two cached sub-plans n = h1 * h2, three transposes and this plan's own
    // inter-pass twiddle tables (W_n^p split in two sqrt(n)-sized halves).
    bool huge;
    int lg_h1, lg_h2, h_shift;
    void *h_lo, *h_hi;
};

struct dsc_ctx {
    // host
    byte *main_base, *scratch_base;
    usize main_size, scratch_size;
    dsc_range_alloc main_alloc;
    dsc_bump_alloc scratch_alloc;
    bool use_scratch;               // DSC_CTX_PUSH state: tensors are temporaries
    bool main_pinned;
    // device
    bool has_device;
    byte *dev_base;
    usize dev_size, dev_scratch_size;
    dsc_range_alloc dev_alloc;      // over [0, dev_size - dev_scratch_size)
    dsc_bump_alloc dev_scratch;     // over the tail
    dsc_tensor_buffer *dev_list;    // buffers with a device mirror
    int residency;                  // 0 strict, 1 resident, 2 lazy download
    dsc_fft_plan *fft_plans[DSC_MAX_FFT_PLANS];
    dsc_fft_plan plan_storage[DSC_MAX_FFT_PLANS];
};

// ---- internal services (dsc_core.cpp) ------------------------------------------------------
void *dsc_host_alloc(dsc_ctx *ctx, usize bytes) noexcept;      // from the current default arena
void dsc_host_free(dsc_ctx *ctx, void *ptr) noexcept;          // tolerant of double frees
void dsc_ctx_push(dsc_ctx *ctx) noexcept;                       // temporaries from scratch, scratch reset
void dsc_ctx_pop(dsc_ctx *ctx) noexcept;
void dsc_require_device(dsc_ctx *ctx, const char *who) noexcept;
// device mirror management
void *dsc_dev_ptr(dsc_ctx *ctx, dsc_tensor_buffer *buf) noexcept;                  // allocate the mirror if needed
void dsc_dev_drop(dsc_ctx *ctx, dsc_tensor_buffer *buf) noexcept;                  // release the mirror (syncs host first if stale)
void dsc_host_written(dsc_tensor_buffer *buf) noexcept;                            // host op wrote the payload
void dsc_host_needed(dsc_ctx *ctx, const dsc_tensor *x) noexcept;                  // host op is about to read the payload

// Elementwise arithmetic / spectrum post-processing on the device when an operand already lives there
// (residency >= 1).  binary: op = DSC_CUDA_OP_ADD..DIV, same dtype, xb same shape / one row / one element;
// unary: op = DSC_CUDA_OP_ABS..CONJ on complex input.  Return false when the host loop should run instead.
bool dsc_try_device_binary(dsc_ctx *ctx, int op, const dsc_tensor *xa, const dsc_tensor *xb, dsc_tensor *out) noexcept;
bool dsc_try_device_unary(dsc_ctx *ctx, int op, const dsc_tensor *x, dsc_tensor *out) noexcept;

// The rest of SURVEY.md 8(f) on the device, taken when the operand is device-resident (residency >= 1); each returns
// false when the host loop should run instead:
//   cast      dsc_cast, any dtype pair (dsc.cpp:587-597, cast_op dsc_ops.h:12-44)
//   gather    out (dense) = x viewed through permuted / stepped strides: dsc_transpose (dsc.cpp:764-827) and
//             dsc_tensor_get_slice (:950-1007); shape[] are out's extents, stride[] / base x's element strides
//   scatter   xa's selected elements = xb (recycled when shorter): dsc_tensor_set_slice / set_idx (:1108-1169)
//   fftfreq   dsc_fftfreq / dsc_rfftfreq filled on the device (:2262-2339)
bool dsc_try_device_cast(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out) noexcept;
bool dsc_try_device_gather(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, const int shape[DSC_MAX_DIMS],
                           const i64 stride[DSC_MAX_DIMS], i64 base) noexcept;
bool dsc_try_device_scatter(dsc_ctx *ctx, dsc_tensor *xa, const dsc_tensor *xb, const int shape[DSC_MAX_DIMS],
                            const i64 stride[DSC_MAX_DIMS], i64 base) noexcept;
bool dsc_try_device_fftfreq(dsc_ctx *ctx, dsc_tensor *out, int n, f64 d, bool rfft) noexcept;

// Crop along the last axis of a tensor whose current contents live only on the device (residency 2): download
// just the kept columns [start, start + count) of every row into `out` (the README's y[:output_length] after
// irfft, README.md:133).  Returns false when the generic host path should run.
bool dsc_try_device_crop(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, int start, int count) noexcept;

// ---- tracer (dsc_trace.cpp) ---------------------------------------------------------------------
void dsc_trace_init(u64 max_traces) noexcept;
void dsc_trace_shutdown() noexcept;
void dsc_trace_set_recording(bool on) noexcept;
bool dsc_trace_recording() noexcept;
void dsc_trace_dump(const char *filename) noexcept;
void dsc_trace_clear() noexcept;
void dsc_trace_event(char phase, const char *name, const char *cat, const char *args_json) noexcept;
// a device-side span measured with events on a stream; resolved into an 'X' record at dump time
void dsc_trace_gpu_span(const char *name, const char *cat, int stream_id,
                        dscdev::Event *start, dscdev::Event *stop, const char *args_json) noexcept;
int dsc_trace_describe_tensor(char *dst, int cap, const dsc_tensor *x) noexcept;

// RAII host span: 'B' on construction, 'E' on destruction, like the reference's dsc_trace_tracker
// (dsc_tracing.h:328-361).  Costs one branch when tracing is off or compiled out.
struct dsc_span {
    const char *name, *cat;
    bool live;
    dsc_span(const char *name_, const char *cat_, const char *args_json) noexcept : name(name_), cat(cat_) {
#if defined(DSC_ENABLE_TRACING)
        live = dsc_trace_recording();
        if (live) dsc_trace_event('B', name, cat, args_json);
#else
        live = false; (void) args_json;
#endif
    }
    ~dsc_span() noexcept { if (live) dsc_trace_event('E', name, cat, nullptr); }
};
