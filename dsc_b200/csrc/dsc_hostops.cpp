// dsc_hostops.cpp -- the tensor entry points that are NOT on the FFT path.
//
// The unchanged Python wrapper binds all 60 symbols at import (python/dsc/_bindings.py:76-767),
// so they must exist; they are not acceleration targets (out of scope) and
// are implemented as plain host loops over the (host-visible) arena with NumPy semantics -- the
// behaviour pinned by the reference's own python/tests/test_ops.py.  Every op that reads a tensor
// first makes its host copy current (dsc_host_needed) and every op that writes one invalidates
// the device mirror (dsc_host_written), which keeps them coherent with the FFT results.
#include "dsc_runtime.h"

#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstring>
#include <random>

namespace {

// ---- dtype dispatch -------------------------------------------------------------------------
template <typename F> DSC_INLINE void by_dtype(const dsc_dtype d, F &&f) noexcept {
    switch (d) {
        case F32: f((f32 *) nullptr); break;
        case F64: f((f64 *) nullptr); break;
        case C32: f((std::complex<f32> *) nullptr); break;
        case C64: f((std::complex<f64> *) nullptr); break;
        DSC_INVALID_CASE("unknown dtype=%d", d);
    }
}

template <typename T> struct is_cplx : std::false_type {};
template <typename R> struct is_cplx<std::complex<R>> : std::true_type {};
template <typename T> struct scalar_of { using type = T; };
template <typename R> struct scalar_of<std::complex<R>> { using type = R; };

template <typename To, typename From> DSC_INLINE To convert(const From v) noexcept {
    if constexpr (is_cplx<To>::value) {
        using R = typename scalar_of<To>::type;
        if constexpr (is_cplx<From>::value) return To((R) v.real(), (R) v.imag());
        else return To((R) v, (R) 0);
    } else {
        if constexpr (is_cplx<From>::value) return (To) v.real();     // complex -> real keeps the real part
        else return (To) v;
    }
}

dsc_dtype real_dtype(const dsc_dtype d) noexcept { return (d == F32 || d == C32) ? F32 : F64; }
bool is_scalar(const dsc_tensor *x) noexcept { return x->n_dim == 1 && x->shape[DSC_MAX_DIMS - 1] == 1; }
const int *user_shape(const dsc_tensor *x) noexcept { return &x->shape[DSC_MAX_DIMS - x->n_dim]; }

// Multi-index walker over a right-aligned 4-D shape; the element offset of the current position
// in a (possibly broadcast or permuted) operand is the dot product with that operand's strides.
struct walker {
    int idx[DSC_MAX_DIMS] = {0, 0, 0, 0};
    const int *shape;
    explicit walker(const int *shape_) noexcept : shape(shape_) {}
    DSC_INLINE void step() noexcept {
        for (int d = DSC_MAX_DIMS - 1; d >= 0; --d) {
            if (++idx[d] < shape[d]) return;
            idx[d] = 0;
        }
    }
    DSC_INLINE int offset(const int *stride) const noexcept {
        return idx[0] * stride[0] + idx[1] * stride[1] + idx[2] * stride[2] + idx[3] * stride[3];
    }
};

void copy_cast(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out) noexcept {
    if (dsc_try_device_cast(ctx, x, out)) return;               // device-resident data is converted where it lies
    dsc_host_needed(ctx, x);
    by_dtype(x->dtype, [&](auto *tx) {
        using Tx = std::remove_pointer_t<decltype(tx)>;
        by_dtype(out->dtype, [&](auto *to) {
            using To = std::remove_pointer_t<decltype(to)>;
            const Tx *src = (const Tx *) x->data;
            To *dst = (To *) out->data;
            for (int i = 0; i < out->ne; ++i) dst[i] = convert<To>(src[i]);
        });
    });
    dsc_host_written(out->buffer);
}

}  // namespace

// =============================================================================================
// creation

#define DSC_DEFINE_WRAP(name, T, DT)                                \
    dsc_tensor *name(dsc_ctx *ctx, const T val) noexcept {          \
        dsc_tensor *out = dsc_tensor_1d(ctx, DT, 1);                \
        *(T *) out->data = val;                                     \
        return out;                                                 \
    }
DSC_DEFINE_WRAP(dsc_wrap_f32, f32, F32)
DSC_DEFINE_WRAP(dsc_wrap_f64, f64, F64)
DSC_DEFINE_WRAP(dsc_wrap_c32, c32, C32)
DSC_DEFINE_WRAP(dsc_wrap_c64, c64, C64)
#undef DSC_DEFINE_WRAP

dsc_tensor *dsc_arange(dsc_ctx *ctx, const int n, const dsc_dtype dtype) noexcept {
    dsc_span span("dsc_arange", "op;arange", nullptr);
    dsc_tensor *out = dsc_tensor_1d(ctx, dtype, n);
    by_dtype(dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        T *dst = (T *) out->data;
        for (int i = 0; i < n; ++i) dst[i] = convert<T>((f64) i);
    });
    return out;
}

dsc_tensor *dsc_randn(dsc_ctx *ctx, const int n_dim, const int *shape, const dsc_dtype dtype) noexcept {
    dsc_span span("dsc_randn", "op;randn", nullptr);
    if (dtype != F32 && dtype != F64) DSC_LOG_FATAL("dtype must be real");
    dsc_tensor *out = dsc_new_tensor(ctx, n_dim, shape, dtype);
    static std::mt19937_64 rng(0x5DC0B200ull);      // one stream per process: successive calls differ
    std::normal_distribution<f64> dist;
    if (dtype == F32) { f32 *d = (f32 *) out->data; for (int i = 0; i < out->ne; ++i) d[i] = (f32) dist(rng); }
    else              { f64 *d = (f64 *) out->data; for (int i = 0; i < out->ne; ++i) d[i] = dist(rng); }
    return out;
}

dsc_tensor *dsc_cast(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT x, const dsc_dtype new_dtype) noexcept {
    dsc_span span("dsc_cast", "op;cast", nullptr);
    if (x->dtype == new_dtype) return x;        // callers rely on pointer identity (python/dsc/tensor.py:325-328)
    dsc_tensor *out = dsc_new_tensor(ctx, x->n_dim, user_shape(x), new_dtype);
    copy_cast(ctx, x, out);
    return out;
}

dsc_tensor *dsc_reshape(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const int dimensions...) noexcept {
    DSC_ASSERT((unsigned) dimensions <= (unsigned) DSC_MAX_DIMS);
    dsc_span span("dsc_reshape", "op;reshape", nullptr);
    int shape[DSC_MAX_DIMS];
    int known = 1, wildcard = -1;
    std::va_list args;
    va_start(args, dimensions);
    for (int i = 0; i < dimensions; ++i) {
        const int d = va_arg(args, int);
        if (d < 0) {
            if (wildcard >= 0) DSC_LOG_FATAL("can only specify one unknown dim");
            wildcard = i;
        } else {
            shape[i] = d;
            known *= d;
        }
    }
    va_end(args);
    if (wildcard >= 0) {
        if (known == 0 || x->ne % known != 0) DSC_LOG_FATAL("cannot reshape %d into %d with an unknown dimension", x->ne, known);
        shape[wildcard] = x->ne / known;
        known = x->ne;
    }
    DSC_ASSERT(x->ne == known);
    return dsc_new_tensor(ctx, dimensions, shape, x->dtype, x->buffer);     // shares the payload
}

dsc_tensor *dsc_concat(dsc_ctx *ctx, const int axis, const int tensors...) noexcept {
    DSC_ASSERT(tensors > 1);
    dsc_span span("dsc_concat", "op;concat", nullptr);
    dsc_tensor **parts = (dsc_tensor **) alloca((usize) tensors * sizeof(dsc_tensor *));
    std::va_list args;
    va_start(args, tensors);
    for (int i = 0; i < tensors; ++i) {
        parts[i] = va_arg(args, dsc_tensor *);
        DSC_ASSERT(parts[i] != nullptr);
        dsc_host_needed(ctx, parts[i]);
    }
    va_end(args);
    const dsc_dtype dtype = parts[0]->dtype;
    const int n_dim = parts[0]->n_dim;
    const usize es = DSC_DTYPE_SIZE[dtype];
    for (int i = 1; i < tensors; ++i) {
        DSC_ASSERT(parts[i]->dtype == dtype);
        DSC_ASSERT(parts[i]->n_dim == n_dim);
    }

    if (axis == DSC_VALUE_NONE) {       // flatten everything into one vector
        int total = 0;
        for (int i = 0; i < tensors; ++i) total += parts[i]->ne;
        dsc_tensor *out = dsc_tensor_1d(ctx, dtype, total);
        byte *dst = (byte *) out->data;
        for (int i = 0; i < tensors; ++i) {
            memcpy(dst, parts[i]->data, (usize) parts[i]->ne * es);
            dst += (usize) parts[i]->ne * es;
        }
        return out;
    }

    const int ax = dsc_tensor_dim(parts[0], axis);
    DSC_ASSERT((unsigned) ax < (unsigned) DSC_MAX_DIMS);
    int shape[DSC_MAX_DIMS];
    memcpy(shape, parts[0]->shape, sizeof(shape));
    for (int i = 1; i < tensors; ++i) {
        for (int d = 0; d < DSC_MAX_DIMS; ++d) {
            if (d == ax) shape[d] += parts[i]->shape[d];
            else DSC_ASSERT(parts[i]->shape[d] == parts[0]->shape[d]);
        }
    }
    dsc_tensor *out = dsc_new_tensor(ctx, n_dim, &shape[DSC_MAX_DIMS - n_dim], dtype);
    // (outer, axis, inner) view: each part contributes a contiguous run of axis_i * inner elements per outer index
    usize outer = 1, inner = 1;
    for (int d = 0; d < ax; ++d) outer *= (usize) shape[d];
    for (int d = ax + 1; d < DSC_MAX_DIMS; ++d) inner *= (usize) shape[d];
    byte *dst = (byte *) out->data;
    for (usize o = 0; o < outer; ++o) {
        for (int i = 0; i < tensors; ++i) {
            const usize run = (usize) parts[i]->shape[ax] * inner * es;
            memcpy(dst, (const byte *) parts[i]->data + o * run, run);
            dst += run;
        }
    }
    return out;
}

dsc_tensor *dsc_transpose(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const int axes...) noexcept {
    DSC_ASSERT(x != nullptr);
    dsc_span span("dsc_transpose", "op;transpose", nullptr);
    if (x->n_dim == 1) return dsc_new_view(ctx, x);
    int perm[DSC_MAX_DIMS];
    if (axes == 0) {
        for (int i = 0; i < x->n_dim; ++i) perm[i] = x->n_dim - 1 - i;
    } else {
        DSC_ASSERT(axes == x->n_dim);
        std::va_list args;
        va_start(args, axes);
        for (int i = 0; i < axes; ++i) {
            perm[i] = va_arg(args, int);
            DSC_ASSERT((unsigned) perm[i] < (unsigned) x->n_dim);
        }
        va_end(args);
    }
    // output dim i takes input dim perm[i]; walk the output in order, gather through permuted strides
    int shape[DSC_MAX_DIMS] = {1, 1, 1, 1}, stride[DSC_MAX_DIMS] = {0, 0, 0, 0};
    for (int i = 0; i < x->n_dim; ++i) {
        const int src = dsc_tensor_dim(x, perm[i]), dst = dsc_tensor_dim(x, i);
        shape[dst] = x->shape[src];
        stride[dst] = x->stride[src];
    }
    dsc_tensor *out = dsc_new_tensor(ctx, x->n_dim, &shape[DSC_MAX_DIMS - x->n_dim], x->dtype);
    {
        // leading dims of extent 1 get the stride a contiguous tensor would have, so the launch layer recognises a
        // plain swap of the last two dims
        i64 st[DSC_MAX_DIMS];
        for (int d = 0; d < DSC_MAX_DIMS; ++d) st[d] = d < DSC_MAX_DIMS - x->n_dim ? (i64) x->ne : (i64) stride[d];
        if (dsc_try_device_gather(ctx, x, out, shape, st, 0)) return out;
    }
    dsc_host_needed(ctx, x);
    const usize es = DSC_DTYPE_SIZE[x->dtype];
    walker w(shape);
    for (int i = 0; i < out->ne; ++i, w.step())
        memcpy((byte *) out->data + (usize) i * es, (const byte *) x->data + (usize) w.offset(stride) * es, es);
    return out;
}

// =============================================================================================
// indexing and slicing (copies, NumPy semantics)

namespace {

struct span1 { int start, step, count; bool collapse; };

// Normalise user slices against x (negative values, missing fields, the "all fields equal" single-index
// convention of the wrappers); dims beyond `given` are taken whole.
void resolve_slices(const dsc_tensor *x, const int given, const dsc_slice *in, span1 *out) noexcept {
    for (int i = 0; i < x->n_dim; ++i) {
        const int dim = x->shape[dsc_tensor_dim(x, i)];
        if (i >= given) { out[i] = span1{0, 1, dim, false}; continue; }
        int start = in[i].start, stop = in[i].stop, step = in[i].step;
        bool single = false;
        if (start == stop && start == step && start != DSC_VALUE_NONE) {
            single = true;
            step = 1;
            if (start < 0) start += dim;
            stop = start + 1;
        }
        DSC_ASSERT(step != 0);
        if (step == DSC_VALUE_NONE) step = 1;
        if (start == DSC_VALUE_NONE) start = step > 0 ? 0 : dim - 1;
        else if (start < 0) start += dim;
        if (stop == DSC_VALUE_NONE) stop = step > 0 ? dim : -1;
        else if (stop < 0 && !single) stop += dim;
        DSC_ASSERT(start >= 0 && start < dim);
        DSC_ASSERT((step > 0 && start < stop && stop <= dim) || (step < 0 && start > stop && stop >= -1));
        const int extent = step > 0 ? stop - start : start - stop;
        const int astep = step > 0 ? step : -step;
        out[i] = span1{start, step, (extent + astep - 1) / astep, single};
    }
}

// visit the element offsets selected by `sp` in row-major order
template <typename F> void for_each_selected(const dsc_tensor *x, const span1 *sp, F &&f) noexcept {
    int shape[DSC_MAX_DIMS] = {1, 1, 1, 1}, start[DSC_MAX_DIMS] = {0, 0, 0, 0}, step[DSC_MAX_DIMS] = {0, 0, 0, 0};
    int total = 1;
    for (int i = 0; i < x->n_dim; ++i) {
        const int d = dsc_tensor_dim(x, i);
        shape[d] = sp[i].count;
        start[d] = sp[i].start * x->stride[d];
        step[d] = sp[i].step * x->stride[d];
        total *= sp[i].count;
    }
    const int base = start[0] + start[1] + start[2] + start[3];
    walker w(shape);
    for (int i = 0; i < total; ++i, w.step()) f(i, base + w.offset(step));
}

void read_slices(std::va_list args, const int n, dsc_slice *dst) noexcept {
    for (int i = 0; i < n; ++i) dst[i] = va_arg(args, dsc_slice);
}

void assign_selected(dsc_ctx *ctx, dsc_tensor *xa, const dsc_tensor *xb, const span1 *sp) noexcept {
    {
        // xa lives on the device only (lazy residency): write the selection there instead of downloading all of xa
        int gshape[DSC_MAX_DIMS] = {1, 1, 1, 1};
        i64 gstride[DSC_MAX_DIMS] = {0, 0, 0, 0}, gbase = 0;
        for (int i = 0; i < xa->n_dim; ++i) {
            const int d = dsc_tensor_dim(xa, i);
            gshape[d] = sp[i].count;
            gstride[d] = (i64) sp[i].step * xa->stride[d];
            gbase += (i64) sp[i].start * xa->stride[d];
        }
        if (dsc_try_device_scatter(ctx, xa, xb, gshape, gstride, gbase)) return;
    }
    dsc_host_needed(ctx, xa);
    dsc_host_needed(ctx, xb);
    const usize es = DSC_DTYPE_SIZE[xa->dtype];
    const byte *src = (const byte *) xb->data;
    byte *dst = (byte *) xa->data;
    const int nb = xb->ne;
    // xb is consumed in order and recycled when shorter (scalar or broadcast row)
    for_each_selected(xa, sp, [&](const int i, const int off) {
        memcpy(dst + (usize) off * es, src + (usize) (i % nb) * es, es);
    });
    dsc_host_written(xa->buffer);
}

}  // namespace

dsc_tensor *dsc_tensor_get_idx(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const int indexes...) noexcept {
    DSC_ASSERT(x != nullptr);
    DSC_ASSERT((unsigned) indexes <= (unsigned) DSC_MAX_DIMS);
    if (indexes > x->n_dim) DSC_LOG_FATAL("too many indexes");
    dsc_span span("dsc_tensor_get_idx", "idx;get", nullptr);
    dsc_host_needed(ctx, x);

    int offset = 0;
    std::va_list args;
    va_start(args, indexes);
    for (int i = 0; i < indexes; ++i) {
        int idx = va_arg(args, int);
        const int d = dsc_tensor_dim(x, i);
        if (idx < 0) idx += x->shape[d];
        DSC_ASSERT((unsigned) idx < (unsigned) x->shape[d]);
        offset += idx * x->stride[d];
    }
    va_end(args);

    // the trailing dims survive; a fully indexed element comes back as a 1-element vector
    const int out_ndim = indexes == x->n_dim ? 1 : x->n_dim - indexes;
    int one = 1;
    const int *out_shape = indexes == x->n_dim ? &one : &x->shape[DSC_MAX_DIMS - out_ndim];
    dsc_tensor *out = dsc_new_tensor(ctx, out_ndim, out_shape, x->dtype);
    const usize es = DSC_DTYPE_SIZE[x->dtype];
    memcpy(out->data, (const byte *) x->data + (usize) offset * es, (usize) out->ne * es);
    return out;
}

dsc_tensor *dsc_tensor_get_slice(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, const int slices...) noexcept {
    DSC_ASSERT(x != nullptr);
    DSC_ASSERT((unsigned) slices <= (unsigned) DSC_MAX_DIMS);
    if (slices > x->n_dim) DSC_LOG_FATAL("too many slices");
    dsc_span span("dsc_tensor_get_slice", "slice;get", nullptr);

    dsc_slice raw[DSC_MAX_DIMS] = {};
    std::va_list args;
    va_start(args, slices);
    read_slices(args, slices, raw);
    va_end(args);
    span1 sp[DSC_MAX_DIMS];
    resolve_slices(x, slices, raw, sp);

    int out_shape[DSC_MAX_DIMS], out_ndim = 0;
    for (int i = 0; i < x->n_dim; ++i)
        if (!sp[i].collapse) out_shape[out_ndim++] = sp[i].count;
    if (out_ndim == 0) { out_shape[0] = 1; out_ndim = 1; }
    dsc_tensor *out = dsc_new_tensor(ctx, out_ndim, out_shape, x->dtype);
    // contiguous crop of the last axis with every other axis whole: if the data only lives on the device,
    // download just the kept columns instead of the whole tensor
    bool last_axis_crop = !sp[x->n_dim - 1].collapse && sp[x->n_dim - 1].step == 1;
    for (int i = 0; i < x->n_dim - 1; ++i)
        last_axis_crop = last_axis_crop && !sp[i].collapse && sp[i].start == 0 && sp[i].step == 1 &&
                         sp[i].count == x->shape[dsc_tensor_dim(x, i)];
    if (last_axis_crop && dsc_try_device_crop(ctx, x, out, sp[x->n_dim - 1].start, sp[x->n_dim - 1].count)) return out;
    {
        // any other selection of device-resident data: gathered on the device
        int gshape[DSC_MAX_DIMS] = {1, 1, 1, 1};
        i64 gstride[DSC_MAX_DIMS] = {0, 0, 0, 0}, gbase = 0;
        for (int i = 0; i < x->n_dim; ++i) {
            const int d = dsc_tensor_dim(x, i);
            gshape[d] = sp[i].count;
            gstride[d] = (i64) sp[i].step * x->stride[d];
            gbase += (i64) sp[i].start * x->stride[d];
        }
        if (dsc_try_device_gather(ctx, x, out, gshape, gstride, gbase)) return out;
    }

    dsc_host_needed(ctx, x);
    const usize es = DSC_DTYPE_SIZE[x->dtype];
    for_each_selected(x, sp, [&](const int i, const int off) {
        memcpy((byte *) out->data + (usize) i * es, (const byte *) x->data + (usize) off * es, es);
    });
    return out;
}

void dsc_tensor_set_idx(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT xa, const dsc_tensor *DSC_RESTRICT xb, const int indexes...) noexcept {
    DSC_ASSERT(xa != nullptr);
    DSC_ASSERT(xb != nullptr);
    DSC_ASSERT((unsigned) indexes <= (unsigned) xa->n_dim);
    DSC_ASSERT(xa->dtype == xb->dtype);
    dsc_span span("dsc_tensor_set_idx", "idx;set", nullptr);

    dsc_slice raw[DSC_MAX_DIMS] = {};
    std::va_list args;
    va_start(args, indexes);
    for (int i = 0; i < indexes; ++i) {
        const int idx = va_arg(args, int);
        raw[i].start = raw[i].stop = raw[i].step = idx;     // single-index convention
        if (idx == DSC_VALUE_NONE) DSC_LOG_FATAL("invalid index");
    }
    va_end(args);
    span1 sp[DSC_MAX_DIMS];
    resolve_slices(xa, indexes, raw, sp);

    // the remaining dims of xa must match xb unless xb is a scalar
    const int rest = xa->n_dim - indexes;
    if (rest == 0) DSC_ASSERT(is_scalar(xb));
    if (!is_scalar(xb)) {
        DSC_ASSERT(xb->n_dim == rest);
        for (int i = 0; i < rest; ++i)
            DSC_ASSERT(xa->shape[dsc_tensor_dim(xa, indexes + i)] == xb->shape[dsc_tensor_dim(xb, i)]);
    }
    assign_selected(ctx, xa, xb, sp);
}

void dsc_tensor_set_slice(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT xa, const dsc_tensor *DSC_RESTRICT xb, const int slices...) noexcept {
    DSC_ASSERT(xa != nullptr);
    DSC_ASSERT(xb != nullptr);
    DSC_ASSERT((unsigned) slices <= (unsigned) xa->n_dim);
    DSC_ASSERT(xa->dtype == xb->dtype);
    dsc_span span("dsc_tensor_set_slice", "slice;set", nullptr);

    dsc_slice raw[DSC_MAX_DIMS] = {};
    std::va_list args;
    va_start(args, slices);
    read_slices(args, slices, raw);
    va_end(args);
    span1 sp[DSC_MAX_DIMS];
    resolve_slices(xa, slices, raw, sp);
    assign_selected(ctx, xa, xb, sp);
}

// =============================================================================================
// binary ops with NumPy broadcasting

namespace {

// device_op: the launch layer's code of the operator, -1 = host only
struct op_add { static constexpr int device_op = DSC_CUDA_OP_ADD; template <typename T> T operator()(T a, T b) const noexcept { return a + b; } };
struct op_sub { static constexpr int device_op = DSC_CUDA_OP_SUB; template <typename T> T operator()(T a, T b) const noexcept { return a - b; } };
struct op_mul { static constexpr int device_op = DSC_CUDA_OP_MUL; template <typename T> T operator()(T a, T b) const noexcept { return a * b; } };
struct op_div { static constexpr int device_op = DSC_CUDA_OP_DIV; template <typename T> T operator()(T a, T b) const noexcept { return a / b; } };
struct op_pow { static constexpr int device_op = -1; template <typename T> T operator()(T a, T b) const noexcept { return std::pow(a, b); } };

template <typename Op>
dsc_tensor *binary(dsc_ctx *ctx, const char *name, dsc_tensor *xa, dsc_tensor *xb, dsc_tensor *out, Op op) noexcept {
    DSC_ASSERT(xa != nullptr);
    DSC_ASSERT(xb != nullptr);
    dsc_span span(name, "op;binary", nullptr);

    int shape[DSC_MAX_DIMS];
    for (int d = 0; d < DSC_MAX_DIMS; ++d) {
        DSC_ASSERT(xa->shape[d] == xb->shape[d] || xa->shape[d] == 1 || xb->shape[d] == 1);
        shape[d] = DSC_MAX(xa->shape[d], xb->shape[d]);
    }
    const int n_dim = DSC_MAX(xa->n_dim, xb->n_dim);
    const dsc_dtype dtype = DSC_DTYPE_CONVERSION_TABLE[xa->dtype][xb->dtype];
    if (out == nullptr) {
        out = dsc_new_tensor(ctx, n_dim, &shape[DSC_MAX_DIMS - n_dim], dtype);
    } else {
        DSC_ASSERT(out->dtype == dtype);
        DSC_ASSERT(out->n_dim == n_dim);
        DSC_ASSERT(memcmp(out->shape, shape, sizeof(shape)) == 0);
    }

    if constexpr (Op::device_op >= 0) {
        if (dsc_try_device_binary(ctx, Op::device_op, xa, xb, out)) return out;   // device-resident data stays there
    }

    // operands are promoted through the scratch arena so nothing needs freeing afterwards
    dsc_ctx_push(ctx);
    const dsc_tensor *a = dsc_cast(ctx, xa, dtype);
    const dsc_tensor *b = dsc_cast(ctx, xb, dtype);
    dsc_ctx_pop(ctx);
    dsc_host_needed(ctx, a);
    dsc_host_needed(ctx, b);

    // broadcast strides: a dimension of extent 1 does not advance
    int sa[DSC_MAX_DIMS], sb[DSC_MAX_DIMS];
    for (int d = 0; d < DSC_MAX_DIMS; ++d) {
        sa[d] = a->shape[d] == 1 ? 0 : a->stride[d];
        sb[d] = b->shape[d] == 1 ? 0 : b->stride[d];
    }
    const bool same = memcmp(a->shape, shape, sizeof(shape)) == 0 && memcmp(b->shape, shape, sizeof(shape)) == 0;
    by_dtype(dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        const T *pa = (const T *) a->data, *pb = (const T *) b->data;
        T *po = (T *) out->data;
        if (same) {
            for (int i = 0; i < out->ne; ++i) po[i] = op(pa[i], pb[i]);
        } else {
            walker w(shape);
            for (int i = 0; i < out->ne; ++i, w.step()) po[i] = op(pa[w.offset(sa)], pb[w.offset(sb)]);
        }
    });
    dsc_host_written(out->buffer);
    return out;
}

}  // namespace

dsc_tensor *dsc_add(dsc_ctx *ctx, dsc_tensor *xa, dsc_tensor *xb, dsc_tensor *out) noexcept { return binary(ctx, "dsc_add", xa, xb, out, op_add()); }
dsc_tensor *dsc_sub(dsc_ctx *ctx, dsc_tensor *xa, dsc_tensor *xb, dsc_tensor *out) noexcept { return binary(ctx, "dsc_sub", xa, xb, out, op_sub()); }
dsc_tensor *dsc_mul(dsc_ctx *ctx, dsc_tensor *xa, dsc_tensor *xb, dsc_tensor *out) noexcept { return binary(ctx, "dsc_mul", xa, xb, out, op_mul()); }
dsc_tensor *dsc_div(dsc_ctx *ctx, dsc_tensor *xa, dsc_tensor *xb, dsc_tensor *out) noexcept { return binary(ctx, "dsc_div", xa, xb, out, op_div()); }
dsc_tensor *dsc_pow(dsc_ctx *ctx, dsc_tensor *xa, dsc_tensor *xb, dsc_tensor *out) noexcept { return binary(ctx, "dsc_pow", xa, xb, out, op_pow()); }

// =============================================================================================
// unary ops

namespace {

dsc_tensor *like_or_check(dsc_ctx *ctx, const dsc_tensor *x, dsc_tensor *out, const dsc_dtype dtype) noexcept {
    if (out == nullptr) return dsc_new_tensor(ctx, x->n_dim, user_shape(x), dtype);
    DSC_ASSERT(out->dtype == dtype);
    DSC_ASSERT(out->n_dim == x->n_dim);
    DSC_ASSERT(memcmp(out->shape, x->shape, sizeof(out->shape)) == 0);
    return out;
}

// same dtype in and out
template <typename Op>
dsc_tensor *unary(dsc_ctx *ctx, const char *name, const dsc_tensor *x, dsc_tensor *out, Op op, const int device_op = -1) noexcept {
    DSC_ASSERT(x != nullptr);
    dsc_span span(name, "op;unary", nullptr);
    out = like_or_check(ctx, x, out, x->dtype);
    if (device_op >= 0 && dsc_try_device_unary(ctx, device_op, x, out)) return out;
    dsc_host_needed(ctx, x);
    by_dtype(x->dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        const T *src = (const T *) x->data;
        T *dst = (T *) out->data;
        for (int i = 0; i < out->ne; ++i) dst[i] = op(src[i]);
    });
    dsc_host_written(out->buffer);
    return out;
}

// complex (or real) in, real out
template <typename Op>
dsc_tensor *unary_to_real(dsc_ctx *ctx, const char *name, const dsc_tensor *x, dsc_tensor *out, Op op, const int device_op = -1) noexcept {
    DSC_ASSERT(x != nullptr);
    dsc_span span(name, "op;unary", nullptr);
    out = like_or_check(ctx, x, out, real_dtype(x->dtype));
    if (device_op >= 0 && dsc_try_device_unary(ctx, device_op, x, out)) return out;
    dsc_host_needed(ctx, x);
    by_dtype(x->dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        using R = typename scalar_of<T>::type;
        const T *src = (const T *) x->data;
        R *dst = (R *) out->data;
        for (int i = 0; i < out->ne; ++i) dst[i] = op(src[i]);
    });
    dsc_host_written(out->buffer);
    return out;
}

struct fn_sinc {
    template <typename T> T operator()(const T v) const noexcept {
        using R = typename scalar_of<T>::type;
        if (v == T(0)) return T(1);
        const T pv = v * (R) 3.14159265358979323846264338327950288L;
        return std::sin(pv) / pv;
    }
};

// modified Bessel function of the first kind, order 0: power series below the switch-over,
// asymptotic expansion above it; both carried in double and rounded to T.
template <typename T> T bessel_i0(const T x) noexcept {
    const f64 ax = std::fabs((f64) x);
    if (ax < 15.0) {
        const f64 q = ax * ax / 4.0;
        f64 term = 1.0, sum = 1.0;
        for (int k = 1; k < 200; ++k) {
            term *= q / ((f64) k * (f64) k);
            sum += term;
            if (term < sum * 1e-17) break;
        }
        return (T) sum;
    }
    // I0(x) ~ e^x / sqrt(2 pi x) * sum_k ((2k-1)!!)^2 / (k! (8x)^k)
    f64 term = 1.0, sum = 1.0;
    for (int k = 1; k < 40; ++k) {
        const f64 next = term * ((2.0 * k - 1.0) * (2.0 * k - 1.0)) / ((f64) k * 8.0 * ax);
        if (next > term) break;
        term = next;
        sum += term;
        if (term < sum * 1e-17) break;
    }
    return (T) (std::exp(ax) / std::sqrt(2.0 * 3.14159265358979323846 * ax) * sum);
}

}  // namespace

#define DSC_DEFINE_UNARY(name, expr)                                                                        \
    dsc_tensor *name(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out) noexcept { \
        return unary(ctx, #name, x, out, [](auto v) noexcept { return expr; });                              \
    }
DSC_DEFINE_UNARY(dsc_cos, std::cos(v))
DSC_DEFINE_UNARY(dsc_sin, std::sin(v))
DSC_DEFINE_UNARY(dsc_sinc, fn_sinc()(v))
DSC_DEFINE_UNARY(dsc_logn, std::log(v))
DSC_DEFINE_UNARY(dsc_log2, std::log(v) / std::log(decltype(v)(2)))
DSC_DEFINE_UNARY(dsc_log10, std::log10(v))
DSC_DEFINE_UNARY(dsc_exp, std::exp(v))
DSC_DEFINE_UNARY(dsc_sqrt, std::sqrt(v))
#undef DSC_DEFINE_UNARY

dsc_tensor *dsc_abs(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out) noexcept {
    return unary_to_real(ctx, "dsc_abs", x, out, [](auto v) noexcept { return std::abs(v); }, DSC_CUDA_OP_ABS);
}

dsc_tensor *dsc_angle(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x) noexcept {
    return unary_to_real(ctx, "dsc_angle", x, nullptr, [](auto v) noexcept { return std::arg(v); }, DSC_CUDA_OP_ANGLE);
}

dsc_tensor *dsc_imag(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x) noexcept {
    return unary_to_real(ctx, "dsc_imag", x, nullptr, [](auto v) noexcept { return std::imag(v); }, DSC_CUDA_OP_IMAG);
}

dsc_tensor *dsc_real(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT x) noexcept {
    DSC_ASSERT(x != nullptr);
    if (x->dtype == F32 || x->dtype == F64) return x;       // identity, same pointer
    return unary_to_real(ctx, "dsc_real", x, nullptr, [](auto v) noexcept { return std::real(v); }, DSC_CUDA_OP_REAL);
}

dsc_tensor *dsc_conj(dsc_ctx *ctx, dsc_tensor *DSC_RESTRICT x) noexcept {
    DSC_ASSERT(x != nullptr);
    if (x->dtype == F32 || x->dtype == F64) return x;       // identity, same pointer
    return unary(ctx, "dsc_conj", x, nullptr, [](auto v) noexcept {
        if constexpr (is_cplx<decltype(v)>::value) return std::conj(v);
        else return v;
    }, DSC_CUDA_OP_CONJ);
}

dsc_tensor *dsc_i0(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x) noexcept {
    DSC_ASSERT(x != nullptr);
    DSC_ASSERT(x->dtype == F32 || x->dtype == F64);
    return unary(ctx, "dsc_i0", x, nullptr, [](auto v) noexcept {
        if constexpr (is_cplx<decltype(v)>::value) return v;
        else return bessel_i0(v);
    });
}

dsc_tensor *dsc_clip(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out,
                     const f64 x_min, const f64 x_max) noexcept {
    // complex values are ordered by their real part and replaced by (bound, 0), as NumPy does
    return unary(ctx, "dsc_clip", x, out, [=](auto v) noexcept {
        using T = decltype(v);
        using R = typename scalar_of<T>::type;
        if constexpr (is_cplx<T>::value) {
            if (v.real() < (R) x_min) return T((R) x_min, 0);
            if (v.real() > (R) x_max) return T((R) x_max, 0);
            return v;
        } else {
            const T lo = v < (R) x_min ? (R) x_min : v;
            return lo > (R) x_max ? (R) x_max : lo;
        }
    });
}

// =============================================================================================
// reductions along one axis

namespace {

template <typename Init, typename Step>
dsc_tensor *reduce(dsc_ctx *ctx, const char *name, const dsc_tensor *x, dsc_tensor *out, const int axis,
                   const bool keep_dims, Init init, Step step) noexcept {
    DSC_ASSERT(x != nullptr);
    dsc_span span(name, "op;unary", nullptr);
    const int ax = dsc_tensor_dim(x, axis);
    DSC_ASSERT((unsigned) ax < (unsigned) DSC_MAX_DIMS);

    int out_shape[DSC_MAX_DIMS], out_ndim = x->n_dim;
    if (keep_dims) {
        memcpy(out_shape, x->shape, sizeof(out_shape));
        out_shape[ax] = 1;
    } else {
        out_ndim = x->n_dim - 1;
        int o = DSC_MAX_DIMS - out_ndim;
        for (int d = 0; d < o; ++d) out_shape[d] = 1;
        for (int d = DSC_MAX_DIMS - x->n_dim; d < DSC_MAX_DIMS; ++d)
            if (d != ax) out_shape[o++] = x->shape[d];
    }
    if (out == nullptr) {
        out = dsc_new_tensor(ctx, out_ndim, &out_shape[DSC_MAX_DIMS - out_ndim], x->dtype);
    } else {
        DSC_ASSERT(out->dtype == x->dtype);
        DSC_ASSERT(out->n_dim == out_ndim);
        DSC_ASSERT(memcmp(out->shape, out_shape, sizeof(out_shape)) == 0);
    }
    dsc_host_needed(ctx, x);

    usize outer = 1, inner = 1;
    for (int d = 0; d < ax; ++d) outer *= (usize) x->shape[d];
    for (int d = ax + 1; d < DSC_MAX_DIMS; ++d) inner *= (usize) x->shape[d];
    const usize n = (usize) x->shape[ax];
    by_dtype(x->dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        const T *src = (const T *) x->data;
        T *dst = (T *) out->data;
        for (usize o = 0; o < outer; ++o)
            for (usize in = 0; in < inner; ++in) {
                T acc = init((T *) nullptr);
                for (usize k = 0; k < n; ++k) acc = step(acc, src[(o * n + k) * inner + in]);
                dst[o * inner + in] = acc;
            }
    });
    dsc_host_written(out->buffer);
    return out;
}

template <typename T> DSC_INLINE bool real_less(const T a, const T b) noexcept {
    if constexpr (is_cplx<T>::value) return a.real() < b.real();    // NumPy orders complex by the real part first
    else return a < b;
}

}  // namespace

dsc_tensor *dsc_sum(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int axis, const bool keep_dims) noexcept {
    return reduce(ctx, "dsc_sum", x, out, axis, keep_dims,
                  [](auto *t) noexcept { return std::remove_pointer_t<decltype(t)>(0); },
                  [](auto acc, auto v) noexcept { return acc + v; });
}

dsc_tensor *dsc_mean(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int axis, const bool keep_dims) noexcept {
    out = dsc_sum(ctx, x, out, axis, keep_dims);
    const f64 inv = 1.0 / (f64) x->shape[dsc_tensor_dim(x, axis)];
    by_dtype(out->dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        using R = typename scalar_of<T>::type;
        T *dst = (T *) out->data;
        for (int i = 0; i < out->ne; ++i) dst[i] *= (R) inv;
    });
    return out;
}

dsc_tensor *dsc_max(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int axis, const bool keep_dims) noexcept {
    return reduce(ctx, "dsc_max", x, out, axis, keep_dims,
                  [](auto *t) noexcept {
                      using T = std::remove_pointer_t<decltype(t)>;
                      using R = typename scalar_of<T>::type;
                      if constexpr (is_cplx<T>::value) return T(-std::numeric_limits<R>::infinity(), -std::numeric_limits<R>::infinity());
                      else return -std::numeric_limits<R>::infinity();
                  },
                  [](auto acc, auto v) noexcept { return real_less(acc, v) ? v : acc; });
}

dsc_tensor *dsc_min(dsc_ctx *ctx, const dsc_tensor *DSC_RESTRICT x, dsc_tensor *DSC_RESTRICT out, const int axis, const bool keep_dims) noexcept {
    return reduce(ctx, "dsc_min", x, out, axis, keep_dims,
                  [](auto *t) noexcept {
                      using T = std::remove_pointer_t<decltype(t)>;
                      using R = typename scalar_of<T>::type;
                      if constexpr (is_cplx<T>::value) return T(std::numeric_limits<R>::infinity(), std::numeric_limits<R>::infinity());
                      else return std::numeric_limits<R>::infinity();
                  },
                  [](auto acc, auto v) noexcept { return real_less(v, acc) ? v : acc; });
}

// =============================================================================================
// frequency grids (dsc.cpp:2262-2339): trivial host fills, kept on the host

dsc_tensor *dsc_fftfreq(dsc_ctx *ctx, const int n, const f64 d, const dsc_dtype dtype) noexcept {
    DSC_ASSERT(n > 0);
    if (dtype != F32 && dtype != F64) DSC_LOG_FATAL("dtype must be real");
    dsc_tensor *out = dsc_tensor_1d(ctx, dtype, n);
    if (dsc_try_device_fftfreq(ctx, out, n, d, false)) return out;
    // [0, 1, .., ceil(n/2)-1, -floor(n/2), .., -1] / (n d)
    const int pos = (n + 1) / 2;
    by_dtype(dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        if constexpr (!is_cplx<T>::value) {
            T *dst = (T *) out->data;
            const T scale = (T) 1 / ((T) n * (T) d);
            for (int i = 0; i < n; ++i) dst[i] = (T) (i < pos ? i : i - n) * scale;
        }
    });
    return out;
}

dsc_tensor *dsc_rfftfreq(dsc_ctx *ctx, const int n, const f64 d, const dsc_dtype dtype) noexcept {
    DSC_ASSERT(n > 0);
    if (dtype != F32 && dtype != F64) DSC_LOG_FATAL("dtype must be real");
    const int bins = n / 2 + 1;
    dsc_tensor *out = dsc_tensor_1d(ctx, dtype, bins);
    if (dsc_try_device_fftfreq(ctx, out, n, d, true)) return out;
    by_dtype(dtype, [&](auto *t) {
        using T = std::remove_pointer_t<decltype(t)>;
        if constexpr (!is_cplx<T>::value) {
            T *dst = (T *) out->data;
            const T scale = (T) 1 / ((T) n * (T) d);
            for (int i = 0; i < bins; ++i) dst[i] = (T) i * scale;
        }
    });
    return out;
}
