// dsc_core.cpp -- context, memory pools, tensors, the FFT plan cache and the FFT entry points
// of libdsc.so.  The numerical work is done by the sm_100a kernels behind dsc_cuda.h; this
// file owns everything around them:
//
//   reference                                       here
//   ----------------------------------------------  -------------------------------------------
//   dsc_ctx_init   dsc/src/dsc.cpp:150-180          + one device arena, page-locked main arena
//   dsc_new_tensor dsc/src/dsc.cpp:342-397          same header/payload layout, + device mirror
//   dsc_plan_fft   dsc/src/dsc.cpp:182-267          same key/ageing/eviction, tables in HBM
//   dsc_internal_fft / exec_fft   :1958-2071        shape rules identical; gather/transform/
//   dsc_internal_rfft / exec_rfft :2102-2244          scatter are ONE launch sequence on the GPU
//
// There is no CPU implementation of the transforms in this library: on a host without a CUDA
// device the context still serves the host-side tensor ops, and every FFT entry point aborts.
#include "dsc_runtime.h"

#include <cstdarg>
#include <cstring>

namespace {

constexpr usize HOST_GRANULE = 64;      // host block granularity
constexpr usize HOST_PREFIX = 32;       // block header in front of every main-arena object
constexpr usize DEV_GRANULE = 256;
constexpr u32 HOST_MAGIC = 0xD5C0B200u;
constexpr usize BUFFER_HEADER = 64;     // sizeof(dsc_tensor_buffer) rounded so the payload is 32-B aligned
static_assert(sizeof(dsc_tensor_buffer) <= BUFFER_HEADER, "buffer header grew");
static_assert(sizeof(dsc_tensor) == 64, "dsc_tensor is 64 bytes in the ABI");

struct host_prefix { u32 magic; int node; };

usize env_size(const char *name, const usize fallback) noexcept {
    const char *v = getenv(name);
    if (v == nullptr || *v == '\0') return fallback;
    return (usize) strtoull(v, nullptr, 10);
}

usize dtype_prec(const dsc_dtype d) noexcept { return (d == F32 || d == C32) ? DSC_CUDA_F32 : DSC_CUDA_F64; }

}  // namespace

// =============================================================================================
// context

dsc_ctx *dsc_ctx_init(const usize main_mem, const usize scratch_mem) noexcept {
    DSC_ASSERT(main_mem > 0);
    DSC_ASSERT(scratch_mem > 0);

    dsc_ctx *ctx = (dsc_ctx *) calloc(1, sizeof(dsc_ctx));
    DSC_ASSERT(ctx != nullptr);

    ctx->main_size = main_mem;
    ctx->scratch_size = scratch_mem;
    ctx->main_base = (byte *) aligned_alloc(4096, DSC_ALIGN(main_mem, (usize) 4096));
    ctx->scratch_base = (byte *) aligned_alloc(4096, DSC_ALIGN(scratch_mem, (usize) 4096));
    if (ctx->main_base == nullptr || ctx->scratch_base == nullptr)
        DSC_LOG_FATAL("error allocating %ldMB of host memory", (long) DSC_B_TO_MB(main_mem + scratch_mem));

    const int max_blocks = (int) env_size("DSC_MAX_BLOCKS", 1 << 18);
    ctx->main_alloc.init(main_mem, HOST_GRANULE, max_blocks);
    ctx->scratch_alloc.init(scratch_mem);

    dsc_trace_init(DSC_MAX_TRACES);

    ctx->has_device = dscdev::device_count() > 0;
    if (ctx->has_device) {
        // Device arena: room for every tensor the host arena can hold plus the scratch mirror,
        // capped by what the GPU has.  This is the only device allocation the context ever makes.
        const usize want = env_size("DSC_DEVICE_MEM", main_mem + scratch_mem);
        const usize avail = dscdev::free_memory();
        usize dev_bytes = DSC_MIN(want, (usize) ((f64) avail * 0.9));
        dev_bytes = dev_bytes / DEV_GRANULE * DEV_GRANULE;
        DSC_ASSERT(dev_bytes >= (1u << 20));
        ctx->dev_base = (byte *) dscdev::arena_alloc(dev_bytes);
        ctx->dev_size = dev_bytes;
        ctx->dev_scratch_size = DSC_MIN(DSC_ALIGN(scratch_mem, DEV_GRANULE), dev_bytes / 2) / DEV_GRANULE * DEV_GRANULE;
        ctx->dev_alloc.init(dev_bytes - ctx->dev_scratch_size, DEV_GRANULE, max_blocks);
        ctx->dev_scratch.init(ctx->dev_scratch_size);
        // Page-lock the main arena so uploads/downloads are asynchronous and run at full PCIe rate.
        // Very large default arenas (python/dsc/context.py:13-26 asks for 10% of RAM) stay pageable.
        const usize pin_limit = env_size("DSC_PIN_LIMIT", (usize) 64 << 30);
        if (main_mem <= pin_limit) ctx->main_pinned = dscdev::host_pin(ctx->main_base, DSC_ALIGN(main_mem, (usize) 4096));
        ctx->residency = (int) env_size("DSC_RESIDENCY", 0);
    }

    DSC_LOG_INFO("created new context %p with %ldMB for main and %ldMB for scratch memory on %s%s",
                 (void *) ctx, (long) DSC_B_TO_MB(main_mem), (long) DSC_B_TO_MB(scratch_mem),
                 ctx->has_device ? dscdev::device_name() : "CPU (no CUDA device: FFT entry points will abort)",
                 ctx->has_device ? (ctx->main_pinned ? " [pinned]" : " [pageable]") : "");
    if (ctx->has_device)
        DSC_LOG_INFO("device arena %ldMB (%ldMB scratch), one allocation",
                     (long) DSC_B_TO_MB(ctx->dev_size), (long) DSC_B_TO_MB(ctx->dev_scratch_size));
    return ctx;
}

void dsc_ctx_free(dsc_ctx *ctx) noexcept {
    DSC_LOG_INFO("freeing context %p: main mem %ldMB, scratch mem %ldMB",
                 (void *) ctx, (long) DSC_B_TO_MB(ctx->main_size), (long) DSC_B_TO_MB(ctx->scratch_size));
    if (ctx->has_device) {
        dscdev::sync_all();
        dsc_trace_shutdown();
        if (ctx->main_pinned) dscdev::host_unpin(ctx->main_base);
        dscdev::arena_free(ctx->dev_base);
        ctx->dev_alloc.destroy();
    } else {
        dsc_trace_shutdown();
    }
    ctx->main_alloc.destroy();
    free(ctx->main_base);
    free(ctx->scratch_base);
    free(ctx);
}

void dsc_ctx_clear(dsc_ctx *ctx) noexcept {
    // Like the reference (dsc.cpp:287-291, dsc_allocator.cpp:134-137) only the scratch arena is
    // actually rewound; live tensors in the main arena stay valid.
    ctx->scratch_alloc.reset();
    if (ctx->has_device) ctx->dev_scratch.reset();
}

usize dsc_used_mem(dsc_ctx *ctx) noexcept { return ctx->main_alloc.used; }

usize dsc_cuda_used_mem(dsc_ctx *ctx) noexcept { return ctx->has_device ? ctx->dev_alloc.used : 0; }

usize dsc_cuda_alloc_calls(dsc_ctx *) noexcept { return dscdev::arena_alloc_calls(); }

void dsc_print_mem_usage(dsc_ctx *ctx) noexcept {
    const usize used = dsc_used_mem(ctx);
    DSC_LOG_INFO("main memory (%s) usage: %ld/%ld MB (%.1f%%)", "CPU",
                 (long) DSC_B_TO_MB(used), (long) DSC_B_TO_MB(ctx->main_size), (f64) used / (f64) ctx->main_size * 1e2);
    if (ctx->has_device)
        DSC_LOG_INFO("device arena (CUDA) usage: %ld/%ld MB", (long) DSC_B_TO_MB(ctx->dev_alloc.used),
                     (long) DSC_B_TO_MB(ctx->dev_alloc.capacity));
}

void dsc_traces_record(dsc_ctx *, const bool record) noexcept { dsc_trace_set_recording(record); }
void dsc_dump_traces(dsc_ctx *, const char *filename) noexcept { dsc_trace_dump(filename); }
void dsc_clear_traces(dsc_ctx *) noexcept { dsc_trace_clear(); }

void dsc_require_device(dsc_ctx *ctx, const char *who) noexcept {
    if (!ctx->has_device)
        DSC_LOG_FATAL("%s needs a CUDA device: this library has no CPU implementation of the FFT path", who);
}

// =============================================================================================
// host objects

void dsc_ctx_push(dsc_ctx *ctx) noexcept {
    ctx->use_scratch = true;
    ctx->scratch_alloc.reset();
}

void dsc_ctx_pop(dsc_ctx *ctx) noexcept { ctx->use_scratch = false; }

void *dsc_host_alloc(dsc_ctx *ctx, const usize bytes) noexcept {
    if (ctx->use_scratch) {
        const usize off = ctx->scratch_alloc.alloc(bytes, 32);
        if (off == (usize) -1) DSC_LOG_FATAL("can't allocate %.2fKB", DSC_B_TO_KB(bytes));
        return ctx->scratch_base + off;
    }
    const int node = ctx->main_alloc.alloc(bytes + HOST_PREFIX);
    if (node < 0) DSC_LOG_FATAL("error allocating %.2fKB", DSC_B_TO_KB(bytes));
    byte *block = ctx->main_base + ctx->main_alloc.nodes[node].off;
    host_prefix *pre = (host_prefix *) block;
    pre->magic = HOST_MAGIC;
    pre->node = node;
    return block + HOST_PREFIX;
}

static bool host_ptr_live(dsc_ctx *ctx, const void *ptr) noexcept {
    const byte *p = (const byte *) ptr;
    if (p < ctx->main_base + HOST_PREFIX || p >= ctx->main_base + ctx->main_size) return false;
    const host_prefix *pre = (const host_prefix *) (p - HOST_PREFIX);
    return pre->magic == HOST_MAGIC &&
           ctx->main_alloc.is_live(pre->node, (usize) ((const byte *) pre - ctx->main_base));
}

void dsc_host_free(dsc_ctx *ctx, void *ptr) noexcept {
    // Pointers into the scratch arena and pointers that were already released are ignored: the
    // Python wrapper's __del__ does free objects twice (reference: dsc_allocator.cpp:152-181).
    if (!host_ptr_live(ctx, ptr)) return;
    host_prefix *pre = (host_prefix *) ((byte *) ptr - HOST_PREFIX);
    pre->magic = 0;
    ctx->main_alloc.release(pre->node);
}

// =============================================================================================
// device mirrors

static void dev_unlink(dsc_ctx *ctx, dsc_tensor_buffer *buf) noexcept {
    if (buf->dev_prev) buf->dev_prev->dev_next = buf->dev_next; else if (ctx->dev_list == buf) ctx->dev_list = buf->dev_next;
    if (buf->dev_next) buf->dev_next->dev_prev = buf->dev_prev;
    buf->dev_prev = buf->dev_next = nullptr;
}

// An asynchronous download (dsc_cuda_download_async) is in flight: wait for THAT copy and retire its event.  Every
// path that frees, rewrites or re-targets the buffer goes through here first, so the DMA never lands in a host block
// that has gone back to the allocator and no stale event is waited on later.  Returns true when there was one.
static bool finish_download(dsc_tensor_buffer *buf) noexcept {
    if (!(buf->flags & DSC_BUF_DOWNLOADING)) return false;
    dscdev::event_wait(buf->downloaded);               // (this copy only: later downloads may still be in flight)
    dscdev::event_release(buf->downloaded);
    buf->downloaded = nullptr;
    buf->flags &= ~(DSC_BUF_HOST_STALE | DSC_BUF_DOWNLOADING);
    return true;
}

static void download_now(dsc_ctx *ctx, dsc_tensor_buffer *buf) noexcept {
    byte *host = (byte *) buf + BUFFER_HEADER;
    if (finish_download(buf)) return;                  // started by dsc_cuda_download_async: just wait for it
    dscdev::stream_sync(0);
    dscdev::copy_d2h(host, ctx->dev_base + ctx->dev_alloc.nodes[buf->dev_node].off, buf->nbytes, 2);
    dscdev::stream_sync(2);
    buf->flags &= ~DSC_BUF_HOST_STALE;
}

void dsc_dev_drop(dsc_ctx *ctx, dsc_tensor_buffer *buf) noexcept {
    if (buf->dev_node < 0) return;
    if (finish_download(buf)) {}                            // the mirror is still being read by the download stream
    else if (buf->flags & DSC_BUF_HOST_STALE) download_now(ctx, buf);
    else if (ctx->residency == 2) dscdev::stream_sync(0);   // a launch may still be reading the mirror
    ctx->dev_alloc.release(buf->dev_node);
    buf->dev_node = -1;
    buf->flags &= ~DSC_BUF_DEV_VALID;
    dev_unlink(ctx, buf);
}

// A block of the device arena; when it is full, the mirrors of tensors that are not operands of the running op are
// only caches (residency >= 1) and are given up first.  -1 when even that does not make room.
static int dev_alloc_evicting(dsc_ctx *ctx, const usize bytes) noexcept {
    int node = ctx->dev_alloc.alloc(bytes);
    if (node >= 0) return node;
    dscdev::sync_all();
    for (dsc_tensor_buffer *b = ctx->dev_list; b != nullptr;) {
        dsc_tensor_buffer *next = b->dev_next;
        if (b->busy == 0) dsc_dev_drop(ctx, b);             // busy != 0 marks operands of the running op
        b = next;
    }
    return ctx->dev_alloc.alloc(bytes);
}

void *dsc_dev_ptr(dsc_ctx *ctx, dsc_tensor_buffer *buf) noexcept {
    if (buf->dev_node < 0) {
        const int node = dev_alloc_evicting(ctx, buf->nbytes);
        if (node < 0)
            DSC_LOG_FATAL("device arena exhausted: %.1fMB requested, %.1fMB of %.1fMB in use (raise dsc_ctx_init sizes or DSC_DEVICE_MEM)",
                          DSC_B_TO_MB(buf->nbytes), DSC_B_TO_MB(ctx->dev_alloc.used), DSC_B_TO_MB(ctx->dev_alloc.capacity));
        buf->dev_node = node;
        buf->flags &= ~DSC_BUF_DEV_VALID;
        buf->dev_prev = nullptr;
        buf->dev_next = ctx->dev_list;
        if (ctx->dev_list) ctx->dev_list->dev_prev = buf;
        ctx->dev_list = buf;
    }
    return ctx->dev_base + ctx->dev_alloc.nodes[buf->dev_node].off;
}

void dsc_host_written(dsc_tensor_buffer *buf) noexcept {
    finish_download(buf);       // an asynchronous download still in flight would land on top of what the host wrote
    buf->flags &= ~(DSC_BUF_DEV_VALID | DSC_BUF_HOST_STALE);
}

void dsc_host_needed(dsc_ctx *ctx, const dsc_tensor *x) noexcept {
    if (x != nullptr && (x->buffer->flags & DSC_BUF_HOST_STALE)) download_now(ctx, x->buffer);
}

namespace {
// device pointer of an operand, uploading it first if only the host copy is current
void *operand_on_device(dsc_ctx *ctx, const dsc_tensor *t) noexcept {
    dsc_tensor_buffer *b = t->buffer;
    b->busy = 1;
    void *d = dsc_dev_ptr(ctx, b);
    if (!(b->flags & DSC_BUF_DEV_VALID)) {
        dscdev::copy_h2d(d, (byte *) b + BUFFER_HEADER, b->nbytes, 0);
        b->flags |= DSC_BUF_DEV_VALID;
    }
    return d;
}

// the result of a device op: valid on the device; downloaded now unless the context downloads lazily
void result_on_device(dsc_ctx *ctx, dsc_tensor *out, void *dout) noexcept {
    out->buffer->flags |= DSC_BUF_DEV_VALID;
    if (ctx->residency == 2) {
        out->buffer->flags |= DSC_BUF_HOST_STALE;
    } else {
        dscdev::stream_sync(0);
        dscdev::copy_d2h((byte *) out->buffer + BUFFER_HEADER, dout, out->buffer->nbytes, 2);
        dscdev::stream_sync(2);
    }
}
}  // namespace

bool dsc_try_device_binary(dsc_ctx *ctx, const int op, const dsc_tensor *xa, const dsc_tensor *xb, dsc_tensor *out) noexcept {
    // Arithmetic between transforms (the spectrum product of dsc.cpp:1273-1284, gains, offsets, ratios) without a
    // round trip through the host: only taken when an operand is already device-resident, so the default
    // (strict) mode never gets here.  Same dtype everywhere; xb same shape, one row, or one element.
    if (!ctx->has_device || ctx->residency < 1) return false;
    const bool mixed = xa->dtype != out->dtype || xb->dtype != out->dtype;    // promoted in registers (dsc_dtype.h:73-78)
    if (!((xa->buffer->flags | xb->buffer->flags) & DSC_BUF_DEV_VALID)) return false;
    if (memcmp(xa->shape, out->shape, sizeof(out->shape)) != 0) return false;
    const bool same = memcmp(xb->shape, out->shape, sizeof(out->shape)) == 0;
    bool row = xb->shape[DSC_MAX_DIMS - 1] == out->shape[DSC_MAX_DIMS - 1];
    for (int d = 0; d < DSC_MAX_DIMS - 1; ++d) row = row && xb->shape[d] == 1;
    const bool scalar = xb->ne == 1;
    if (!same && !row && !scalar) return false;
    if (xa->buffer == out->buffer || xb->buffer == out->buffer) return false;

    const void *da = operand_on_device(ctx, xa), *db = operand_on_device(ctx, xb);
    out->buffer->busy = 1;
    void *dout = dsc_dev_ptr(ctx, out->buffer);
    const i64 cols = out->shape[DSC_MAX_DIMS - 1];
    const i64 rows = cols > 0 ? out->ne / cols : 0;
    const int b_mode = same ? 1 : row ? 0 : 2;
    const int rc = mixed ? dsc_cuda_binary_mixed(op, da, xa->dtype, db, xb->dtype, dout, rows, cols, b_mode, dscdev::stream(0))
                         : dsc_cuda_binary(op, da, db, dout, out->dtype, rows, cols, b_mode, dscdev::stream(0));
    if (rc != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    result_on_device(ctx, out, dout);
    xa->buffer->busy = xb->buffer->busy = out->buffer->busy = 0;
    return true;
}

bool dsc_try_device_unary(dsc_ctx *ctx, const int op, const dsc_tensor *x, dsc_tensor *out) noexcept {
    // |X|, arg X, Re, Im, conj of a spectrum that lives on the device (dsc.cpp:1480-1622)
    if (!ctx->has_device || ctx->residency < 1) return false;
    if (x->dtype != C32 && x->dtype != C64) return false;
    if (!(x->buffer->flags & DSC_BUF_DEV_VALID) || x->buffer == out->buffer) return false;
    const void *dx = operand_on_device(ctx, x);
    out->buffer->busy = 1;
    void *dout = dsc_dev_ptr(ctx, out->buffer);
    if (dsc_cuda_unary(op, dx, x->dtype, dout, x->ne, dscdev::stream(0)) != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    result_on_device(ctx, out, dout);
    x->buffer->busy = out->buffer->busy = 0;
    return true;
}

bool dsc_try_device_cast(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out) noexcept {
    if (!ctx->has_device || ctx->residency < 1) return false;
    if (!(x->buffer->flags & DSC_BUF_DEV_VALID) || x->buffer == out->buffer || x->ne == 0) return false;
    if (out->buffer->flags & DSC_BUF_SCRATCH) return false;        // promotion temporaries of host loops stay on the host
    const void *dx = operand_on_device(ctx, x);
    out->buffer->busy = 1;
    void *dout = dsc_dev_ptr(ctx, out->buffer);
    if (dsc_cuda_cast(dx, x->dtype, dout, out->dtype, x->ne, dscdev::stream(0)) != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    result_on_device(ctx, out, dout);
    x->buffer->busy = out->buffer->busy = 0;
    return true;
}

bool dsc_try_device_gather(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, const int shape[DSC_MAX_DIMS],
                           const i64 stride[DSC_MAX_DIMS], const i64 base) noexcept {
    if (!ctx->has_device || ctx->residency < 1) return false;
    if (!(x->buffer->flags & DSC_BUF_DEV_VALID) || x->buffer == out->buffer || out->ne == 0) return false;
    const void *dx = operand_on_device(ctx, x);
    out->buffer->busy = 1;
    void *dout = dsc_dev_ptr(ctx, out->buffer);
    const int es = (int) DSC_DTYPE_SIZE[x->dtype];
    // a swap of the last two dims of a batch of matrices: the tiled transpose (coalesced on both sides)
    const bool swap_last = base == 0 && stride[3] == (i64) shape[3 - 1] && stride[2] == 1 &&
                           stride[1] == (i64) shape[2] * shape[3] && stride[0] == (i64) shape[1] * shape[2] * shape[3];
    int rc;
    if (swap_last && shape[2] > 1 && shape[3] > 1)
        rc = dsc_cuda_transpose_batched(dx, dout, (i64) shape[0] * shape[1], shape[3], shape[2], es, dscdev::stream(0));
    else
        rc = dsc_cuda_gather(dx, dout, es, shape, (const int64_t *) stride, base, dscdev::stream(0));
    if (rc != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    result_on_device(ctx, out, dout);
    x->buffer->busy = out->buffer->busy = 0;
    return true;
}

bool dsc_try_device_scatter(dsc_ctx *ctx, dsc_tensor *xa, const dsc_tensor *xb, const int shape[DSC_MAX_DIMS],
                            const i64 stride[DSC_MAX_DIMS], const i64 base) noexcept {
    // only when xa's current contents live on the device alone (lazy mode): otherwise the host loop keeps the host
    // copy current and just invalidates the mirror
    if (!ctx->has_device || ctx->residency < 2) return false;
    dsc_tensor_buffer *ba = xa->buffer;
    if (!(ba->flags & DSC_BUF_DEV_VALID) || !(ba->flags & DSC_BUF_HOST_STALE) || ba == xb->buffer || xb->ne == 0) return false;
    finish_download(ba);
    ba->busy = 1;
    void *da = dsc_dev_ptr(ctx, ba);
    const void *db = operand_on_device(ctx, xb);
    if (dsc_cuda_scatter(da, db, (int) DSC_DTYPE_SIZE[xa->dtype], shape, (const int64_t *) stride, base, xb->ne, dscdev::stream(0)) != 0)
        DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    ba->busy = xb->buffer->busy = 0;
    return true;                                                // xa stays device-valid, host-stale
}

bool dsc_try_device_fftfreq(dsc_ctx *ctx, dsc_tensor *out, const int n, const f64 d, const bool rfft) noexcept {
    if (!ctx->has_device || ctx->residency < 1) return false;
    out->buffer->busy = 1;
    void *dout = dsc_dev_ptr(ctx, out->buffer);
    if (dsc_cuda_fftfreq(dout, out->dtype, n, d, rfft ? 1 : 0, dscdev::stream(0)) != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    result_on_device(ctx, out, dout);
    out->buffer->busy = 0;
    return true;
}

bool dsc_try_device_crop(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, const int start, const int count) noexcept {
    dsc_tensor_buffer *b = x->buffer;
    if (!ctx->has_device || !(b->flags & DSC_BUF_HOST_STALE) || b->dev_node < 0) return false;
    const usize es = DSC_DTYPE_SIZE[x->dtype];
    const usize cols = (usize) x->shape[DSC_MAX_DIMS - 1];
    const usize rows = cols > 0 ? (usize) x->ne / cols : 0;
    if (rows == 0 || count <= 0) return false;
    const byte *src = ctx->dev_base + ctx->dev_alloc.nodes[b->dev_node].off + (usize) start * es;
    dscdev::stream_sync(0);                                   // the producer kernel must have finished
    dscdev::copy_d2h_2d(out->data, (usize) count * es, src, cols * es, (usize) count * es, rows, 2);
    dscdev::stream_sync(2);
    dsc_host_written(out->buffer);
    return true;
}

void dsc_cuda_set_residency(dsc_ctx *ctx, const int mode) noexcept {
    DSC_ASSERT(mode >= 0 && mode <= 2);
    if (mode < ctx->residency) {
        // leaving a lazier mode: bring every host copy up to date, forget mirrors if going strict
        for (dsc_tensor_buffer *b = ctx->dev_list; b != nullptr;) {
            dsc_tensor_buffer *next = b->dev_next;
            if (b->flags & DSC_BUF_HOST_STALE) download_now(ctx, b);
            if (mode == 0) dsc_dev_drop(ctx, b);
            b = next;
        }
    }
    ctx->residency = mode;
}

void dsc_cuda_sync_host(dsc_ctx *ctx, dsc_tensor *x) noexcept { dsc_host_needed(ctx, x); }

void dsc_cuda_touch_host(dsc_ctx *, dsc_tensor *x) noexcept { if (x) dsc_host_written(x->buffer); }

void dsc_cuda_download_async(dsc_ctx *ctx, dsc_tensor *x) noexcept {
    // residency 2: start the device -> host copy of a result on the download stream and return; the next calls
    // (uploads and kernels of the following transform) overlap with it -- PCIe is full duplex -- and
    // dsc_cuda_sync_host / any host access of x waits for it.
    if (x == nullptr || !ctx->has_device) return;
    dsc_tensor_buffer *b = x->buffer;
    if (!(b->flags & DSC_BUF_HOST_STALE) || (b->flags & DSC_BUF_DOWNLOADING) || b->dev_node < 0) return;
    dscdev::Event *produced = dscdev::event_record(0);
    dscdev::stream_wait(2, produced);
    dscdev::event_release(produced);
    dscdev::copy_d2h((byte *) b + BUFFER_HEADER, ctx->dev_base + ctx->dev_alloc.nodes[b->dev_node].off, b->nbytes, 2);
    b->downloaded = dscdev::event_record(2);
    b->flags |= DSC_BUF_DOWNLOADING;
}

void dsc_cuda_prefetch(dsc_ctx *ctx, dsc_tensor *x) noexcept {
    if (x == nullptr || !ctx->has_device || ctx->residency < 1 || x->buffer->nbytes == 0) return;
    operand_on_device(ctx, x);
    dscdev::stream_sync(0);          // the caller may overwrite the host payload right away
    x->buffer->busy = 0;
}

// =============================================================================================
// tensors

DSC_MALLOC dsc_tensor *dsc_new_tensor(dsc_ctx *ctx, const int n_dim, const int *shape,
                                      const dsc_dtype dtype, dsc_tensor_buffer *buffer) noexcept {
    DSC_ASSERT((unsigned) n_dim <= (unsigned) DSC_MAX_DIMS);

    i64 ne = 1;
    for (int i = 0; i < n_dim; ++i) ne *= shape[i];
    DSC_ASSERT(ne >= 0 && ne <= INT32_MAX);

    dsc_tensor *t = (dsc_tensor *) dsc_host_alloc(ctx, sizeof(dsc_tensor));
    if (buffer == nullptr) {
        const usize nbytes = (usize) ne * DSC_DTYPE_SIZE[dtype];
        buffer = (dsc_tensor_buffer *) dsc_host_alloc(ctx, BUFFER_HEADER + nbytes);
        buffer->refs = 0;
        buffer->flags = ctx->use_scratch ? DSC_BUF_SCRATCH : 0;
        buffer->dev_node = -1;
        buffer->busy = 0;
        buffer->nbytes = nbytes;
        buffer->dev_prev = buffer->dev_next = nullptr;
        buffer->downloaded = nullptr;
    }
    buffer->refs++;

    t->buffer = buffer;
    t->data = (byte *) buffer + BUFFER_HEADER;
    t->ne = (int) ne;
    t->n_dim = n_dim;
    t->dtype = dtype;
    t->backend = ctx->has_device ? CUDA : CPU;
    const int lead = DSC_MAX_DIMS - n_dim;
    for (int i = 0; i < DSC_MAX_DIMS; ++i) t->shape[i] = i < lead ? 1 : shape[i - lead];
    t->stride[DSC_MAX_DIMS - 1] = 1;
    for (int i = DSC_MAX_DIMS - 2; i >= 0; --i) t->stride[i] = t->stride[i + 1] * t->shape[i + 1];

#if defined(DSC_ENABLE_TRACING)
    if (dsc_trace_recording()) {
        char args[200];
        dsc_trace_describe_tensor(args, sizeof(args), t);
        dsc_trace_event('B', "dsc_new_tensor", "alloc", args);
        dsc_trace_event('E', "dsc_new_tensor", "alloc", nullptr);
    }
#endif
    return t;
}

DSC_MALLOC dsc_tensor *dsc_view(dsc_ctx *ctx, const dsc_tensor *x) noexcept { return dsc_new_view(ctx, x); }

void dsc_tensor_free(dsc_ctx *ctx, dsc_tensor *x) noexcept {
    if (x == nullptr || !host_ptr_live(ctx, x)) return;     // null, scratch temporary or already freed
#if defined(DSC_ENABLE_TRACING)
    if (dsc_trace_recording()) {
        char args[200];
        dsc_trace_describe_tensor(args, sizeof(args), x);
        dsc_trace_event('B', "dsc_tensor_free", "free", args);
        dsc_trace_event('E', "dsc_tensor_free", "free", nullptr);
    }
#endif
    dsc_tensor_buffer *buf = x->buffer;
    if (host_ptr_live(ctx, buf) && --buf->refs == 0) {
        if (buf->dev_node >= 0) {
            buf->flags &= ~DSC_BUF_HOST_STALE;      // nobody can read it any more
            dsc_dev_drop(ctx, buf);
        }
        dsc_host_free(ctx, buf);
    }
    dsc_host_free(ctx, x);
}

dsc_tensor *dsc_tensor_1d(dsc_ctx *ctx, const dsc_dtype dtype, const int dim1) noexcept {
    const int shape[1] = {dim1};
    return dsc_new_tensor(ctx, 1, shape, dtype);
}
dsc_tensor *dsc_tensor_2d(dsc_ctx *ctx, const dsc_dtype dtype, const int dim1, const int dim2) noexcept {
    const int shape[2] = {dim1, dim2};
    return dsc_new_tensor(ctx, 2, shape, dtype);
}
dsc_tensor *dsc_tensor_3d(dsc_ctx *ctx, const dsc_dtype dtype, const int dim1, const int dim2, const int dim3) noexcept {
    const int shape[3] = {dim1, dim2, dim3};
    return dsc_new_tensor(ctx, 3, shape, dtype);
}
dsc_tensor *dsc_tensor_4d(dsc_ctx *ctx, const dsc_dtype dtype, const int dim1, const int dim2,
                          const int dim3, const int dim4) noexcept {
    const int shape[4] = {dim1, dim2, dim3, dim4};
    return dsc_new_tensor(ctx, 4, shape, dtype);
}

// =============================================================================================
// plan cache

static dsc_fft_plan *find_plan(dsc_ctx *ctx, const int n, const dsc_fft_type type, const int prec) noexcept {
    // One pass over the slots: the match is rejuvenated, every other live plan ages by one
    // (dsc.cpp:199-213) -- including on a miss, which is what makes the eviction LRU-like.
    dsc_fft_plan *hit = nullptr;
    for (int i = 0; i < DSC_MAX_FFT_PLANS; ++i) {
        dsc_fft_plan *p = ctx->fft_plans[i];
        if (p == nullptr) continue;
        if (p->cu.n == n && p->cu.fft_type == (int) type && p->cu.dtype == prec) {
            hit = p;
            p->last_used = 0;
        } else {
            p->last_used++;
        }
    }
    return hit;
}

dsc_fft_plan *dsc_plan_fft(dsc_ctx *ctx, const int n, const dsc_fft_type fft_type, const dsc_dtype dtype) noexcept {
    const int fft_n = dsc_pow2_n(n);
    const int prec = (int) dtype_prec(dtype);

    char args[128];
    snprintf(args, sizeof(args), "{\"type\": \"%s\", \"n\": %d, \"order\": %d, \"dtype\": \"%s\"}",
             fft_type == COMPLEX ? "FFT" : "RFFT", n, fft_n, DSC_DTYPE_NAMES[dtype]);
    dsc_span span("dsc_plan_fft", "op;fft;plan", args);

    dsc_require_device(ctx, "dsc_plan_fft");

    dsc_fft_plan *plan = find_plan(ctx, fft_n, fft_type, prec);
    if (plan != nullptr) return plan;

    int slot = -1;
    for (int i = 0; i < DSC_MAX_FFT_PLANS && slot < 0; ++i)
        if (ctx->fft_plans[i] == nullptr) slot = i;
    if (slot < 0) {
        // cache full: the plan that has gone unused the longest gives its tables back to the arena
        int oldest = -1;
        for (int i = 0; i < DSC_MAX_FFT_PLANS; ++i)
            if (ctx->fft_plans[i]->last_used > oldest) { oldest = ctx->fft_plans[i]->last_used; slot = i; }
        dscdev::stream_sync(0);     // no launch may still be reading the evicted tables
        ctx->dev_alloc.release(ctx->fft_plans[slot]->dev_node);
        ctx->fft_plans[slot] = nullptr;
    }

    usize bytes = dsc_cuda_plan_bytes(fft_n, fft_type, prec);
    int lg = 0;
    while ((1 << lg) < fft_n) ++lg;
    const int huge_lg = (int) env_size("DSC_HUGE_LG", 0);          // testing knob: compose above 2^DSC_HUGE_LG
    const bool huge = bytes == 0 || (huge_lg > 0 && lg > huge_lg && fft_type == COMPLEX);
    const usize es = prec == DSC_CUDA_F32 ? sizeof(c32) : sizeof(c64);
    if (huge) {
        if (fft_type != COMPLEX) DSC_LOG_FATAL("real transforms of order %d are outside the supported range", fft_n);
        bytes = (((usize) 1 << ((lg + 1) / 2)) + ((usize) 1 << (lg - (lg + 1) / 2))) * es + 2 * DEV_GRANULE;
    }
    const int node = dev_alloc_evicting(ctx, bytes);
    if (node < 0) DSC_LOG_FATAL("device arena exhausted while planning an FFT of length %d (%.1fMB of tables)", fft_n, DSC_B_TO_MB(bytes));
    byte *mem = ctx->dev_base + ctx->dev_alloc.nodes[node].off;

    plan = &ctx->plan_storage[slot];
    memset(plan, 0, sizeof(*plan));
    if (huge) {
        plan->cu.n = fft_n; plan->cu.lg_n = lg; plan->cu.fft_type = fft_type; plan->cu.dtype = prec;
        plan->huge = true;
        plan->lg_h1 = (lg + 1) / 2; plan->lg_h2 = lg / 2;
        plan->h_shift = (lg + 1) / 2;
        plan->h_lo = mem;
        plan->h_hi = mem + DSC_ALIGN(((usize) 1 << plan->h_shift) * es, DEV_GRANULE);
        int rc = dsc_cuda_fill_twiddles(plan->h_lo, (i64) 1 << plan->h_shift, 1, fft_n, prec, dscdev::stream(0));
        if (rc == 0) rc = dsc_cuda_fill_twiddles(plan->h_hi, (i64) 1 << (lg - plan->h_shift), (i64) 1 << plan->h_shift, fft_n, prec, dscdev::stream(0));
        if (rc != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    } else {
        const int rc = dsc_cuda_plan_build(&plan->cu, fft_n, fft_type, prec, mem, bytes, dscdev::stream(0));
        if (rc != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
    }
    plan->last_used = 0;
    plan->dev_node = node;
    ctx->fft_plans[slot] = plan;
    return plan;
}

// =============================================================================================
// transforms

namespace {

enum xform { XF_FFT, XF_IFFT, XF_RFFT, XF_IRFFT, XF_FILTER };

struct xform_job {
    xform kind;
    const dsc_fft_plan *plan;
    const dsc_tensor *x;
    dsc_tensor *out;
    const dsc_tensor *spectrum;     // XF_FILTER: B = rfft(b), broadcast over lines
    i64 outer, inner;
    int x_n, out_n;
    int keep;                       // XF_IRFFT / XF_FILTER: store only the first `keep` samples per line (0 = all); out_n == keep
    // composed paths (see launch_chunk): two device temporaries and, for huge plans, the sub-plans
    byte *tmp_a, *tmp_b;
    const dsc_fft_plan *sub1, *sub2;
    usize scratch_capacity;             // bytes of device scratch a launch can use as work memory
};

#define CUDA_RC(call) do { const int rc_ = (call); if (rc_ != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error()); } while (0)

// complex two-pass transform along a non-last axis that the launch layer runs as ONE launch of column passes
// (needs a covered shape and a device scratch that holds the work rows of one outer slab; else: transposes)
bool strided_two_pass_direct(const xform_job &j) noexcept {
    if (j.plan->huge || j.plan->cu.lg_n2 == 0 || j.inner <= 1 || (j.kind != XF_FFT && j.kind != XF_IFFT)) return false;
    const usize need = dsc_cuda_work_bytes_axis(&j.plan->cu, 1, j.inner);
    return need > 0 && need <= j.scratch_capacity;
}

bool is_composed(const xform_job &j) noexcept {
    return j.plan->huge || (j.plan->cu.lg_n2 != 0 && j.inner > 1 && !strided_two_pass_direct(j));
}

int one_transform(const xform_job &j, const dsc_cuda_plan *plan, const void *src, const int src_dtype, void *dst,
                  const i64 lines, const int x_n, void *work, const usize work_bytes, void *s) noexcept {
    switch (j.kind) {
        case XF_FFT:  return dsc_cuda_fft(plan, src, src_dtype, dst, lines, x_n, 1, 1, work, work_bytes, s);
        case XF_IFFT: return dsc_cuda_fft(plan, src, src_dtype, dst, lines, x_n, 1, 0, work, work_bytes, s);
        case XF_RFFT: return dsc_cuda_rfft(plan, src, dst, lines, x_n, 1, work, work_bytes, s);
        case XF_IRFFT: return dsc_cuda_irfft(plan, src, dst, lines, x_n, 1, work, work_bytes, s);
        default: return -1;
    }
}

// A transform of more than one shared-memory pass along a NON-last axis: per outer index, transpose the
// (n, inner) slab so the lines become contiguous, transform, transpose back.
void strided_large_slab(const xform_job &j, const byte *src, byte *dst, void *work, const usize work_bytes, void *s) noexcept {
    const int in_es = (int) DSC_DTYPE_SIZE[j.x->dtype], out_es = (int) DSC_DTYPE_SIZE[j.out->dtype];
    CUDA_RC(dsc_cuda_transpose(src, j.tmp_a, j.x_n, j.inner, in_es, s));
    CUDA_RC(one_transform(j, &j.plan->cu, j.tmp_a, j.x->dtype, j.tmp_b, j.inner, j.x_n, work, work_bytes, s));
    CUDA_RC(dsc_cuda_transpose(j.tmp_b, dst, j.inner, j.out_n, out_es, s));
}

// One line of a huge plan n = h1 * h2 (index n = a*h2 + b): transpose (+cast, +zero pad) to [b][a], h2
// transforms of length h1, twiddle W_n^(b k1) + transpose to [k1][b], h1 transforms of length h2,
// transpose to natural order X[k1 + h1 k2].  Each sub-transform may itself be a fused four-step launch.
void huge_line(const xform_job &j, const byte *src, byte *dst, void *work, const usize work_bytes, void *s) noexcept {
    const dsc_fft_plan *p = j.plan;
    const i64 h1 = (i64) 1 << p->lg_h1, h2 = (i64) 1 << p->lg_h2;
    const bool fwd = j.kind == XF_FFT;
    const int cplx = p->cu.dtype == DSC_CUDA_F32 ? DSC_CUDA_C32 : DSC_CUDA_C64;
    const int es = p->cu.dtype == DSC_CUDA_F32 ? (int) sizeof(c32) : (int) sizeof(c64);
    const i64 take = DSC_MIN((i64) j.x_n, (i64) p->cu.n);
    CUDA_RC(dsc_cuda_transpose_cast(src, j.x->dtype, j.tmp_a, h1, h2, take, s));
    CUDA_RC(dsc_cuda_fft(&j.sub1->cu, j.tmp_a, cplx, j.tmp_b, h2, (int) h1, 1, fwd, work, work_bytes, s));
    CUDA_RC(dsc_cuda_transpose_twiddle(j.tmp_b, j.tmp_a, h2, h1, 0, p->h_lo, p->h_hi, p->h_shift, fwd, cplx, s));
    CUDA_RC(dsc_cuda_fft(&j.sub2->cu, j.tmp_a, cplx, j.tmp_b, h1, (int) h2, 1, fwd, work, work_bytes, s));
    CUDA_RC(dsc_cuda_transpose(j.tmp_b, dst, h1, h2, es, s));
}

// One chunk of lines [r0, r0 + rows) on the compute stream.
void launch_chunk(dsc_ctx *ctx, const xform_job &j, const byte *dx, byte *dout, const i64 r0, const i64 rows,
                  void *work, const usize work_bytes) noexcept {
    const usize in_row = (usize) j.x_n * (usize) j.inner * DSC_DTYPE_SIZE[j.x->dtype];
    const usize out_row = (usize) j.out_n * (usize) j.inner * DSC_DTYPE_SIZE[j.out->dtype];
    const void *src = dx + (usize) r0 * in_row;
    void *dst = dout + (usize) r0 * out_row;
    void *s = dscdev::stream(0);
    if (is_composed(j)) {
        for (i64 r = 0; r < rows; ++r) {
            const byte *rs = (const byte *) src + (usize) r * in_row;
            byte *rd = (byte *) dst + (usize) r * out_row;
            if (j.plan->huge) huge_line(j, rs, rd, work, work_bytes, s);
            else strided_large_slab(j, rs, rd, work, work_bytes, s);
        }
        return;
    }
    int rc = 0;
    switch (j.kind) {
        case XF_FFT:
        case XF_IFFT:
            rc = dsc_cuda_fft(&j.plan->cu, src, j.x->dtype, dst, rows, j.x_n, j.inner, j.kind == XF_FFT, work, work_bytes, s);
            break;
        case XF_RFFT:
            rc = dsc_cuda_rfft(&j.plan->cu, src, dst, rows, j.x_n, j.inner, work, work_bytes, s);
            break;
        case XF_IRFFT:
            rc = j.keep ? dsc_cuda_irfft_keep(&j.plan->cu, src, dst, rows, j.x_n, j.keep, work, work_bytes, s)
                        : dsc_cuda_irfft(&j.plan->cu, src, dst, rows, j.x_n, j.inner, work, work_bytes, s);
            break;
        case XF_FILTER:
            // rfft -> spectrum product -> irfft without the spectrum ever leaving the device (one kernel for
            // orders that fit shared memory)
            rc = j.keep ? dsc_cuda_filter_keep(&j.plan->cu, src, dsc_dev_ptr(ctx, j.spectrum->buffer), dst, rows, j.x_n, j.keep, work, work_bytes, s)
                        : dsc_cuda_filter(&j.plan->cu, src, dsc_dev_ptr(ctx, j.spectrum->buffer), dst, rows, j.x_n, work, work_bytes, s);
            break;
    }
    if (rc != 0) DSC_LOG_FATAL("%s", dsc_cuda_last_error());
}

void upload_if_needed(dsc_ctx *ctx, const dsc_tensor *t) noexcept {
    dsc_tensor_buffer *b = t->buffer;
    const bool had_mirror = b->dev_node >= 0;
    void *d = dsc_dev_ptr(ctx, b);
    if (ctx->residency >= 1 && (b->flags & DSC_BUF_DEV_VALID)) return;
    if (had_mirror) {                                          // earlier launches may still read the old contents
        dscdev::Event *readers_done = dscdev::event_record(0);
        dscdev::stream_wait(1, readers_done);
        dscdev::event_release(readers_done);
    }
    dscdev::copy_h2d(d, (byte *) b + BUFFER_HEADER, b->nbytes, 1);
    dscdev::Event *e = dscdev::event_record(1);
    dscdev::stream_wait(0, e);
    dscdev::event_release(e);
    if (ctx->residency >= 1) b->flags |= DSC_BUF_DEV_VALID;
}

// Upload -> transform -> download, pipelined over chunks of lines on three streams so that PCIe
// in, the kernels and PCIe out overlap.  On return the host copy of `out` is current unless the
// context runs in residency mode 2.
void run_job(dsc_ctx *ctx, const xform_job &j) noexcept {
    dsc_tensor_buffer *bx = j.x->buffer, *bo = j.out->buffer;
    DSC_ASSERT(bx != bo);
    bx->busy = bo->busy = 1;                                   // operands of the running op: not evictable
    if (j.spectrum) j.spectrum->buffer->busy = 1;

    finish_download(bo);                                       // `out=` reused while its last download is still in flight
    const bool x_on_device = ctx->residency >= 1 && bx->dev_node >= 0 && (bx->flags & DSC_BUF_DEV_VALID);
    const bool x_mirror_reused = !x_on_device && bx->dev_node >= 0;
    const byte *dx = (const byte *) dsc_dev_ptr(ctx, bx);
    byte *dout = (byte *) dsc_dev_ptr(ctx, bo);
    if (x_mirror_reused) {
        // the upload below overwrites a mirror that an earlier launch on the compute stream may still be reading
        // (residency >= 1 after dsc_cuda_touch_host / a host write): order the upload stream behind it
        dscdev::Event *readers_done = dscdev::event_record(0);
        dscdev::stream_wait(1, readers_done);
        dscdev::event_release(readers_done);
    }
    if (j.spectrum) upload_if_needed(ctx, j.spectrum);

    const usize in_row = (usize) j.x_n * (usize) j.inner * DSC_DTYPE_SIZE[j.x->dtype];
    const usize out_row = (usize) j.out_n * (usize) j.inner * DSC_DTYPE_SIZE[j.out->dtype];
    const byte *hx = (const byte *) j.x->data;
    byte *hout = (byte *) j.out->data;
    const bool download = ctx->residency < 2;

    // chunking: ~32 MiB of traffic per chunk, at least one line; a single chunk when there is no
    // outer dimension to split (lines interleaved through `inner` are not contiguous)
    const usize chunk_bytes = env_size("DSC_CHUNK_BYTES", (usize) 32 << 20);
    i64 rows_per_chunk = (i64) (chunk_bytes / DSC_MAX(DSC_MAX(in_row, out_row), (usize) 1));
    rows_per_chunk = DSC_MAX(rows_per_chunk, (i64) 1);
    if (j.outer * (i64) DSC_MAX(in_row, out_row) < (i64) (2 * chunk_bytes)) rows_per_chunk = j.outer;
    if (x_on_device && !download) rows_per_chunk = j.outer;      // nothing crosses PCIe: nothing to overlap

    // work memory from the device scratch (two-pass intermediates, the filter's spectrum)
    ctx->dev_scratch.reset();
    usize work_bytes = 0;
    void *work = nullptr;
    {
        const i64 lines = DSC_MIN(rows_per_chunk, j.outer) * j.inner;
        usize need = j.kind == XF_FILTER ? dsc_cuda_filter_work_bytes(&j.plan->cu, lines)
                     : strided_two_pass_direct(j) ? dsc_cuda_work_bytes_axis(&j.plan->cu, DSC_MIN(rows_per_chunk, j.outer), j.inner)
                                                  : dsc_cuda_work_bytes(&j.plan->cu, lines);
        // long single-pass columns (2^13, 2^14 points along a non-last axis) also run as column passes when the
        // scratch holds their work rows; without work memory the launch layer uses its single-pass blocks
        if (need == 0 && j.inner > 1 && !j.plan->huge && (j.kind == XF_FFT || j.kind == XF_IFFT)) {
            const usize cols = dsc_cuda_work_bytes_axis(&j.plan->cu, DSC_MIN(rows_per_chunk, j.outer), j.inner);
            if (cols <= ctx->dev_scratch.capacity) need = cols;
        }
        if (need > 0) {
            // less than `need` still works (the launch layer chunks); one line must fit
            work_bytes = DSC_MIN(need, ctx->dev_scratch.capacity) / DEV_GRANULE * DEV_GRANULE;
            const usize off = ctx->dev_scratch.alloc(work_bytes, DEV_GRANULE);
            DSC_ASSERT(off != (usize) -1);
            work = ctx->dev_base + (ctx->dev_size - ctx->dev_scratch_size) + off;
        }
    }

    // composed paths: two device temporaries from the arena for the duration of the op
    int tmp_nodes[2] = {-1, -1};
    xform_job jj = j;
    if (is_composed(j)) {
        usize ta, tb;
        if (j.plan->huge) {
            if (j.inner != 1) DSC_LOG_FATAL("transforms of 2^%d points are supported along the last axis only", j.plan->cu.lg_n);
            const usize es = j.plan->cu.dtype == DSC_CUDA_F32 ? sizeof(c32) : sizeof(c64);
            ta = tb = (usize) j.plan->cu.n * es;
        } else {
            ta = (usize) j.inner * (usize) j.x_n * DSC_DTYPE_SIZE[j.x->dtype];
            tb = (usize) j.inner * (usize) j.out_n * DSC_DTYPE_SIZE[j.out->dtype];
        }
        tmp_nodes[0] = dev_alloc_evicting(ctx, ta);
        tmp_nodes[1] = dev_alloc_evicting(ctx, tb);
        if (tmp_nodes[0] < 0 || tmp_nodes[1] < 0)
            DSC_LOG_FATAL("device arena exhausted: this transform needs %.1fMB of temporaries", DSC_B_TO_MB(ta + tb));
        jj.tmp_a = ctx->dev_base + ctx->dev_alloc.nodes[tmp_nodes[0]].off;
        jj.tmp_b = ctx->dev_base + ctx->dev_alloc.nodes[tmp_nodes[1]].off;
        // the sub-transforms are contiguous batches: size the work buffer for them
        const dsc_cuda_plan *wp = j.plan->huge ? &j.sub1->cu : &j.plan->cu;
        const i64 wl = j.plan->huge ? ((i64) 1 << j.plan->lg_h2) : j.inner;
        usize need = dsc_cuda_work_bytes(wp, wl);
        if (j.plan->huge) need = DSC_MAX(need, dsc_cuda_work_bytes(&j.sub2->cu, (i64) 1 << j.plan->lg_h1));
        if (need > work_bytes) {
            ctx->dev_scratch.reset();
            work_bytes = DSC_MIN(need, ctx->dev_scratch.capacity) / DEV_GRANULE * DEV_GRANULE;
            const usize off = ctx->dev_scratch.alloc(work_bytes, DEV_GRANULE);
            DSC_ASSERT(off != (usize) -1);
            work = ctx->dev_base + (ctx->dev_size - ctx->dev_scratch_size) + off;
        }
    }

    const bool tracing = dsc_trace_recording();
    for (i64 r0 = 0; r0 < j.outer; r0 += rows_per_chunk) {
        const i64 rows = DSC_MIN(rows_per_chunk, j.outer - r0);
        if (!x_on_device) {
            dscdev::Event *t0 = tracing ? dscdev::event_record(1) : nullptr;
            dscdev::copy_h2d((void *) (dx + (usize) r0 * in_row), hx + (usize) r0 * in_row, (usize) rows * in_row, 1);
            dscdev::Event *e = dscdev::event_record(1);
            dscdev::stream_wait(0, e);
            if (tracing) {
                char a[96];
                snprintf(a, sizeof(a), "{\"bytes\": %zu}", (usize) rows * in_row);
                dsc_trace_gpu_span("memcpy_h2d", "gpu;copy", 1, t0, e, a);
            } else {
                dscdev::event_release(e);
            }
        }
        dscdev::Event *k0 = tracing ? dscdev::event_record(0) : nullptr;
        launch_chunk(ctx, jj, dx, dout, r0, rows, work, work_bytes);
        dscdev::Event *k1 = (tracing || download) ? dscdev::event_record(0) : nullptr;
        if (download) {
            dscdev::stream_wait(2, k1);
            dscdev::Event *t0 = tracing ? dscdev::event_record(2) : nullptr;
            dscdev::copy_d2h(hout + (usize) r0 * out_row, dout + (usize) r0 * out_row, (usize) rows * out_row, 2);
            if (tracing) {
                char a[96];
                snprintf(a, sizeof(a), "{\"bytes\": %zu}", (usize) rows * out_row);
                dsc_trace_gpu_span("memcpy_d2h", "gpu;copy", 2, t0, dscdev::event_record(2), a);
            }
        }
        if (tracing) {
            static const char *names[] = {"fft_lines<fwd>", "fft_lines<inv>", "fft_lines<r2c>", "fft_lines<c2r>", "fft_filter"};
            char a[160];
            snprintf(a, sizeof(a), "{\"n\": %d, \"lines\": %lld, \"passes\": %d, \"bytes\": %zu}", j.plan->cu.n,
                     (long long) (rows * j.inner), j.plan->cu.lg_n2 ? 2 : 1, (usize) rows * (in_row + out_row));
            dsc_trace_gpu_span(names[j.kind], "gpu;fft", 0, k0, k1, a);
        } else if (k1) {
            dscdev::event_release(k1);
        }
    }

    if (download) {
        dscdev::stream_sync(2);
        bo->flags &= ~DSC_BUF_HOST_STALE;
    } else {
        dscdev::stream_sync(1);      // the caller may free or overwrite x's host memory right away
        bo->flags |= DSC_BUF_HOST_STALE;
    }
    if (ctx->residency >= 1) {
        bx->flags |= DSC_BUF_DEV_VALID;
        bo->flags |= DSC_BUF_DEV_VALID;
    }
    bx->busy = bo->busy = 0;
    if (j.spectrum) j.spectrum->buffer->busy = 0;
    if (tmp_nodes[0] >= 0) {
        dscdev::stream_sync(0);
        ctx->dev_alloc.release(tmp_nodes[0]);
        ctx->dev_alloc.release(tmp_nodes[1]);
    }
    if (ctx->residency == 0) {
        // strict mode: nothing stays behind on the device, every call starts from host memory
        dscdev::stream_sync(0);
        dsc_dev_drop(ctx, bx);
        dsc_dev_drop(ctx, bo);
        if (j.spectrum) dsc_dev_drop(ctx, j.spectrum->buffer);
    }
}

void split_axis(const dsc_tensor *x, const int axis_idx, i64 *outer, i64 *inner) noexcept {
    *outer = 1; *inner = 1;
    for (int i = 0; i < axis_idx; ++i) *outer *= x->shape[i];
    for (int i = axis_idx + 1; i < DSC_MAX_DIMS; ++i) *inner *= x->shape[i];
}

dsc_tensor *make_out(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, const int axis_idx, const int out_n,
                     const dsc_dtype out_dtype) noexcept {
    int out_shape[DSC_MAX_DIMS];
    for (int i = 0; i < DSC_MAX_DIMS; ++i) out_shape[i] = i != axis_idx ? x->shape[i] : out_n;
    if (out == nullptr) return dsc_new_tensor(ctx, x->n_dim, &out_shape[DSC_MAX_DIMS - x->n_dim], out_dtype);
    DSC_ASSERT(out->dtype == out_dtype);
    DSC_ASSERT(out->n_dim == x->n_dim);
    DSC_ASSERT(memcmp(out_shape, out->shape, DSC_MAX_DIMS * sizeof(out->shape[0])) == 0);
    return out;
}

void fft_trace_args(char *dst, const int cap, const char *type, const int n, const int axis,
                    const dsc_tensor *x, const dsc_tensor *out) noexcept {
    if (!dsc_trace_recording()) { dst[0] = '\0'; return; }
    // always a complete JSON object: a field that does not fit is left out, never cut (the dump is one JSON array)
    char tx[200], to[200];
    dsc_trace_describe_tensor(tx, sizeof(tx), x);
    int o = snprintf(dst, (size_t) cap, "{\"type\": \"%s\", \"order\": %d, \"axis\": %d", type, n, axis);
    if (o < 0 || o >= cap - 1) { dst[0] = '\0'; return; }
    const int lx = (int) strlen(tx);
    if (o + 7 + lx + 1 < cap) o += snprintf(dst + o, (size_t) (cap - o), ", \"x\": %s", tx);
    if (out != nullptr) {
        dsc_trace_describe_tensor(to, sizeof(to), out);
        const int lo = (int) strlen(to);
        if (o + 9 + lo + 1 < cap) o += snprintf(dst + o, (size_t) (cap - o), ", \"out\": %s", to);
    }
    snprintf(dst + o, (size_t) (cap - o), "}");
}

// fft / ifft: shape and dtype rules of dsc_internal_fft (dsc.cpp:2009-2071)
dsc_tensor *dsc_internal_fft(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, int n, const int axis,
                             const bool forward) noexcept {
    DSC_ASSERT(x != nullptr);
    char args[512];
    fft_trace_args(args, sizeof(args), forward ? "FFT" : "IFFT", n, axis, x, out);
    dsc_span span("dsc_internal_fft", "op;fft", args);
    dsc_require_device(ctx, forward ? "dsc_fft" : "dsc_ifft");

    const int axis_idx = dsc_tensor_dim(x, axis);
    DSC_ASSERT((unsigned) axis_idx < (unsigned) DSC_MAX_DIMS);
    const int x_n = x->shape[axis_idx];
    n = n > 0 ? dsc_pow2_n(n) : dsc_pow2_n(x_n);          // non powers of two round UP

    dsc_dtype out_dtype = x->dtype;
    if (x->dtype == F32) out_dtype = C32;
    else if (x->dtype == F64) out_dtype = C64;
    out = make_out(ctx, x, out, axis_idx, n, out_dtype);

    xform_job j{};
    j.scratch_capacity = ctx->dev_scratch.capacity;
    j.kind = forward ? XF_FFT : XF_IFFT;
    j.plan = dsc_plan_fft(ctx, n, COMPLEX, out_dtype);
    if (j.plan->huge) {
        static_assert(DSC_MAX_FFT_PLANS >= 3, "a composed plan and its two sub-plans must fit the cache together");
        j.sub1 = dsc_plan_fft(ctx, 1 << j.plan->lg_h1, COMPLEX, out_dtype);
        j.sub2 = dsc_plan_fft(ctx, 1 << j.plan->lg_h2, COMPLEX, out_dtype);
    }
    j.x = x; j.out = out;
    j.x_n = x_n; j.out_n = n;
    split_axis(x, axis_idx, &j.outer, &j.inner);
    if (x->ne > 0) run_job(ctx, j);
    return out;
}

// rfft / irfft: rules of dsc_internal_rfft (dsc.cpp:2173-2244).  For the inverse, n counts INPUT bins.
dsc_tensor *dsc_internal_rfft(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, const int n, const int axis,
                              const bool forward) noexcept {
    DSC_ASSERT(x != nullptr);
    char args[512];
    fft_trace_args(args, sizeof(args), forward ? "RFFT" : "IRFFT", n, axis, x, out);
    dsc_span span("dsc_internal_rfft", "op;fft", args);
    dsc_require_device(ctx, forward ? "dsc_rfft" : "dsc_irfft");

    const int axis_idx = dsc_tensor_dim(x, axis);
    DSC_ASSERT((unsigned) axis_idx < (unsigned) DSC_MAX_DIMS);
    const int x_n = x->shape[axis_idx];

    int order, out_n;
    dsc_dtype out_dtype;
    if (forward) {
        order = dsc_pow2_n(n > 0 ? n : x_n) >> 1;
        out_n = order + 1;
        if (x->dtype == F32) out_dtype = C32;
        else if (x->dtype == F64) out_dtype = C64;
        else DSC_LOG_FATAL("RFFT input must be real");
    } else {
        order = dsc_pow2_n((n > 0 ? n : x_n) - 1);      // asserts for a single bin, like the reference
        out_n = order << 1;
        if (x->dtype == C32) out_dtype = F32;
        else if (x->dtype == C64) out_dtype = F64;
        else DSC_LOG_FATAL("IRFFT input must be complex");
    }
    DSC_ASSERT(order > 0);
    out = make_out(ctx, x, out, axis_idx, out_n, out_dtype);

    xform_job j{};
    j.scratch_capacity = ctx->dev_scratch.capacity;
    j.kind = forward ? XF_RFFT : XF_IRFFT;
    j.plan = dsc_plan_fft(ctx, order, REAL, x->dtype);
    j.x = x; j.out = out;
    j.x_n = x_n; j.out_n = out_n;
    split_axis(x, axis_idx, &j.outer, &j.inner);
    if (x->ne > 0) run_job(ctx, j);
    return out;
}

}  // namespace

dsc_tensor *dsc_fft(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int n, const int axis) noexcept {
    return dsc_internal_fft(ctx, x, out, n, axis, true);
}
dsc_tensor *dsc_ifft(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int n, const int axis) noexcept {
    return dsc_internal_fft(ctx, x, out, n, axis, false);
}
dsc_tensor *dsc_rfft(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int n, const int axis) noexcept {
    return dsc_internal_rfft(ctx, x, out, n, axis, true);
}
dsc_tensor *dsc_irfft(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int n, const int axis) noexcept {
    return dsc_internal_rfft(ctx, x, out, n, axis, false);
}

namespace {
// The crop of README.md:130-133 fused into the store, where the launch layer can do it (orders of one shared-memory
// pass); else the full transform followed by dsc_tensor_get_slice's device crop.
dsc_tensor *crop_last_axis(dsc_ctx *ctx, dsc_tensor *full, dsc_tensor *out, const int keep) noexcept {
    dsc_slice sl[DSC_MAX_DIMS];
    for (int i = 0; i < full->n_dim; ++i) sl[i] = dsc_slice{DSC_VALUE_NONE, DSC_VALUE_NONE, 1};
    sl[full->n_dim - 1] = dsc_slice{0, keep, 1};
    dsc_tensor *cropped = full->n_dim == 1 ? dsc_tensor_get_slice(ctx, full, 1, sl[0])
                          : full->n_dim == 2 ? dsc_tensor_get_slice(ctx, full, 2, sl[0], sl[1])
                          : full->n_dim == 3 ? dsc_tensor_get_slice(ctx, full, 3, sl[0], sl[1], sl[2])
                                             : dsc_tensor_get_slice(ctx, full, 4, sl[0], sl[1], sl[2], sl[3]);
    dsc_tensor_free(ctx, full);
    if (out == nullptr) return cropped;
    DSC_ASSERT(out->dtype == cropped->dtype && out->ne == cropped->ne);
    dsc_host_needed(ctx, cropped);
    memcpy(out->data, cropped->data, (usize) cropped->ne * DSC_DTYPE_SIZE[cropped->dtype]);
    dsc_host_written(out->buffer);
    dsc_tensor_free(ctx, cropped);
    return out;
}
}  // namespace

dsc_tensor *dsc_irfft_keep(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int n,
                           const int axis, const int keep) noexcept {
    DSC_ASSERT(x != nullptr);
    char args[512];
    fft_trace_args(args, sizeof(args), "IRFFT", n, axis, x, out);
    dsc_span span("dsc_internal_rfft", "op;fft", args);
    dsc_require_device(ctx, "dsc_irfft_keep");
    const int axis_idx = dsc_tensor_dim(x, axis);
    DSC_ASSERT(axis_idx == DSC_MAX_DIMS - 1);               // last axis only
    const int x_n = x->shape[axis_idx];
    const int order = dsc_pow2_n((n > 0 ? n : x_n) - 1);
    DSC_ASSERT(order > 0);
    DSC_ASSERT(keep > 0 && keep <= 2 * order);
    dsc_dtype out_dtype;
    if (x->dtype == C32) out_dtype = F32;
    else if (x->dtype == C64) out_dtype = F64;
    else DSC_LOG_FATAL("IRFFT input must be complex");
    const dsc_fft_plan *plan = dsc_plan_fft(ctx, order, REAL, x->dtype);
    if (plan->cu.lg_n2 != 0 || keep == 2 * order) {
        dsc_tensor *full = dsc_internal_rfft(ctx, x, nullptr, n, axis, false);
        return keep == 2 * order && out == nullptr ? full : crop_last_axis(ctx, full, out, keep);
    }
    out = make_out(ctx, x, out, axis_idx, keep, out_dtype);
    xform_job j{};
    j.scratch_capacity = ctx->dev_scratch.capacity;
    j.kind = XF_IRFFT;
    j.plan = plan;
    j.x = x; j.out = out;
    j.x_n = x_n; j.out_n = keep; j.keep = keep;
    split_axis(x, axis_idx, &j.outer, &j.inner);
    if (x->ne > 0) run_job(ctx, j);
    return out;
}

dsc_tensor *dsc_fft_filter_keep(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const dsc_tensor *DSC_RESTRICT B,
                                dsc_tensor *DSC_RESTRICT out, const int n, const int axis, const int keep) noexcept {
    DSC_ASSERT(x != nullptr);
    DSC_ASSERT(B != nullptr);
    const int axis_idx = dsc_tensor_dim(x, axis);
    DSC_ASSERT(axis_idx == DSC_MAX_DIMS - 1);
    const int x_n = x->shape[axis_idx];
    const int order = dsc_pow2_n(n > 0 ? n : x_n) >> 1;
    DSC_ASSERT(order > 0);
    DSC_ASSERT(keep > 0 && keep <= 2 * order);
    if (x->dtype != F32 && x->dtype != F64) DSC_LOG_FATAL("filter input must be real");
    const dsc_fft_plan *plan = dsc_plan_fft(ctx, order, REAL, x->dtype);
    if (plan->cu.lg_n2 != 0 || keep == 2 * order) {
        dsc_tensor *full = dsc_fft_filter(ctx, x, B, nullptr, n, axis);
        return keep == 2 * order && out == nullptr ? full : crop_last_axis(ctx, full, out, keep);
    }
    char args[512];
    fft_trace_args(args, sizeof(args), "FILTER", n, axis, x, out);
    dsc_span span("dsc_fft_filter", "op;fft", args);
    dsc_require_device(ctx, "dsc_fft_filter_keep");
    DSC_ASSERT(B->dtype == (x->dtype == F32 ? C32 : C64));
    DSC_ASSERT(B->ne == order + 1);
    out = make_out(ctx, x, out, axis_idx, keep, x->dtype);
    xform_job j{};
    j.scratch_capacity = ctx->dev_scratch.capacity;
    j.kind = XF_FILTER;
    j.plan = plan;
    j.x = x; j.out = out; j.spectrum = B;
    j.x_n = x_n; j.out_n = keep; j.keep = keep;
    split_axis(x, axis_idx, &j.outer, &j.inner);
    if (x->ne > 0) run_job(ctx, j);
    return out;
}

// out = irfft(rfft(x, n) * B) along the LAST axis; B holds order+1 bins (e.g. dsc_rfft(b, n)).
dsc_tensor *dsc_fft_filter(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const dsc_tensor *DSC_RESTRICT B,
                           dsc_tensor *DSC_RESTRICT out, const int n, const int axis) noexcept {
    DSC_ASSERT(x != nullptr);
    DSC_ASSERT(B != nullptr);
    char args[512];
    fft_trace_args(args, sizeof(args), "FILTER", n, axis, x, out);
    dsc_span span("dsc_fft_filter", "op;fft", args);
    dsc_require_device(ctx, "dsc_fft_filter");

    const int axis_idx = dsc_tensor_dim(x, axis);
    DSC_ASSERT(axis_idx == DSC_MAX_DIMS - 1);               // last axis only
    const int x_n = x->shape[axis_idx];
    const int order = dsc_pow2_n(n > 0 ? n : x_n) >> 1;
    DSC_ASSERT(order > 0);
    dsc_dtype spec_dtype;
    if (x->dtype == F32) spec_dtype = C32;
    else if (x->dtype == F64) spec_dtype = C64;
    else DSC_LOG_FATAL("filter input must be real");
    DSC_ASSERT(B->dtype == spec_dtype);
    DSC_ASSERT(B->ne == order + 1);
    out = make_out(ctx, x, out, axis_idx, 2 * order, x->dtype);

    xform_job j{};
    j.scratch_capacity = ctx->dev_scratch.capacity;
    j.kind = XF_FILTER;
    j.plan = dsc_plan_fft(ctx, order, REAL, x->dtype);
    j.x = x; j.out = out; j.spectrum = B;
    j.x_n = x_n; j.out_n = 2 * order;
    split_axis(x, axis_idx, &j.outer, &j.inner);
    if (x->ne > 0) run_job(ctx, j);
    return out;
}
