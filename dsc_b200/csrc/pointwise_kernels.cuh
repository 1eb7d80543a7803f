// pointwise_kernels.cuh -- elementwise kernels for spectra that already live on the device:
// magnitude / phase / real / imag / conj and the four arithmetic operators with row or scalar broadcast.
// They exist so that typical post-processing between two transforms does not force a round trip through the
// host; each is one read and one write of the payload (HBM-bound), grid-stride, coalesced.
//
// Reference semantics: unary ops /root/reference/dsc/src/dsc.cpp:1480-1622 (std::abs / std::arg / real / imag /
// conj of std::complex<T>), binary ops dsc.cpp:1186-1310 with the functors of dsc/include/dsc_ops.h:46-90.
#pragma once

#include "fft_math.cuh"

namespace dscfft {

template <typename T> DSC_DEV T pw_sqrt(T v);
template <> DSC_DEV float pw_sqrt<float>(float v) { return sqrtf(v); }
template <> DSC_DEV double pw_sqrt<double>(double v) { return sqrt(v); }
template <typename T> DSC_DEV T pw_atan2(T y, T x);
template <> DSC_DEV float pw_atan2<float>(float y, float x) { return atan2f(y, x); }
template <> DSC_DEV double pw_atan2<double>(double y, double x) { return atan2(y, x); }

// complex in, real out: OP 0 = |z|, 1 = arg z, 2 = Re z, 3 = Im z
template <typename T, int OP>
__global__ void pointwise_c2r(const cx<T> *__restrict__ x, T *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const cx<T> v = x[i];
        T r;
        if (OP == 0) r = pw_sqrt<T>(v.x * v.x + v.y * v.y);
        else if (OP == 1) r = pw_atan2<T>(v.y, v.x);
        else if (OP == 2) r = v.x;
        else r = v.y;
        out[i] = r;
    }
}

template <typename T>
__global__ void pointwise_conj(const cx<T> *__restrict__ x, cx<T> *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const cx<T> v = x[i];
        out[i] = mk<T>(v.x, -v.y);
    }
}

// OP 0 add, 1 sub, 2 mul, 3 div on real scalars and on complex pairs
template <int OP> DSC_DEV float pw_apply(float a, float b) { return OP == 0 ? a + b : OP == 1 ? a - b : OP == 2 ? a * b : a / b; }
template <int OP> DSC_DEV double pw_apply(double a, double b) { return OP == 0 ? a + b : OP == 1 ? a - b : OP == 2 ? a * b : a / b; }
template <int OP, typename V> DSC_DEV V pw_apply_cx(V a, V b) {
    V r;
    if (OP == 0) { r.x = a.x + b.x; r.y = a.y + b.y; }
    else if (OP == 1) { r.x = a.x - b.x; r.y = a.y - b.y; }
    else if (OP == 2) { r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; }
    else {
        const auto d = b.x * b.x + b.y * b.y;
        r.x = (a.x * b.x + a.y * b.y) / d;
        r.y = (a.y * b.x - a.x * b.y) / d;
    }
    return r;
}
template <int OP> DSC_DEV float2 pw_apply(float2 a, float2 b) { return pw_apply_cx<OP>(a, b); }
template <int OP> DSC_DEV double2 pw_apply(double2 a, double2 b) { return pw_apply_cx<OP>(a, b); }

// out = a OP b over rows x cols elements; b_mode 0: b is one row (broadcast over rows), 1: same shape,
// 2: b is a single element.  out may alias a.
template <typename V, int OP>
__global__ void pointwise_binary(const V *a, const V *__restrict__ b, V *out, long long rows, long long cols, int b_mode) {
    if (b_mode == 0 && cols >= blockDim.x) {
        // one row of b over many rows of a: blocks walk over rows, threads over columns -- no 64-bit modulo per
        // element, and each thread's slice of b stays in registers / L1
        for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
            const V *ar = a + r * cols;
            V *orow = out + r * cols;
            for (long long c = threadIdx.x; c < cols; c += blockDim.x) orow[c] = pw_apply<OP>(ar[c], b[c]);
        }
        return;
    }
    const long long total = rows * cols, i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    if (b_mode == 1) {
        for (long long i = i0; i < total; i += step) out[i] = pw_apply<OP>(a[i], b[i]);
    } else if (b_mode == 2) {
        const V y = b[0];
        for (long long i = i0; i < total; i += step) out[i] = pw_apply<OP>(a[i], y);
    } else {
        for (long long i = i0; i < total; i += step) out[i] = pw_apply<OP>(a[i], b[i % cols]);
    }
}

// ---- cast (cast_op, /root/reference/dsc/include/dsc_ops.h:12-44): real -> complex sets imag = 0, complex -> real keeps
// the real part, precision changes convert each component
template <typename Tout, typename Tin> struct PwCast;
template <typename A, typename B> struct PwCastReal { static DSC_DEV B run(const A v) { return (B)v; } };
template <> struct PwCast<float, float> : PwCastReal<float, float> {};
template <> struct PwCast<float, double> : PwCastReal<double, float> {};
template <> struct PwCast<double, float> : PwCastReal<float, double> {};
template <> struct PwCast<double, double> : PwCastReal<double, double> {};
template <> struct PwCast<float2, float> { static DSC_DEV float2 run(const float v) { return mk<float>(v, 0.f); } };
template <> struct PwCast<float2, double> { static DSC_DEV float2 run(const double v) { return mk<float>((float)v, 0.f); } };
template <> struct PwCast<double2, float> { static DSC_DEV double2 run(const float v) { return mk<double>((double)v, 0.0); } };
template <> struct PwCast<double2, double> { static DSC_DEV double2 run(const double v) { return mk<double>(v, 0.0); } };
template <> struct PwCast<float, float2> { static DSC_DEV float run(const float2 v) { return v.x; } };
template <> struct PwCast<float, double2> { static DSC_DEV float run(const double2 v) { return (float)v.x; } };
template <> struct PwCast<double, float2> { static DSC_DEV double run(const float2 v) { return (double)v.x; } };
template <> struct PwCast<double, double2> { static DSC_DEV double run(const double2 v) { return v.x; } };
template <> struct PwCast<float2, float2> { static DSC_DEV float2 run(const float2 v) { return v; } };
template <> struct PwCast<float2, double2> { static DSC_DEV float2 run(const double2 v) { return mk<float>((float)v.x, (float)v.y); } };
template <> struct PwCast<double2, float2> { static DSC_DEV double2 run(const float2 v) { return mk<double>((double)v.x, (double)v.y); } };
template <> struct PwCast<double2, double2> { static DSC_DEV double2 run(const double2 v) { return v; } };

template <typename Tin, typename Tout>
__global__ void pointwise_cast(const Tin *__restrict__ x, Tout *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = PwCast<Tout, Tin>::run(x[i]);
}

// a OP b where the operands were promoted to the output dtype first (conversion table dsc_dtype.h:73-78; the
// reference casts both operands through scratch, dsc.cpp:65-68, then runs the same-dtype functor): the casts
// happen in registers.  b_mode as in pointwise_binary.
template <typename Ta, typename Tb, typename V, int OP>
__global__ void pointwise_binary_mixed(const Ta *__restrict__ a, const Tb *__restrict__ b, V *__restrict__ out,
                                       long long rows, long long cols, int b_mode) {
    const long long total = rows * cols, i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long i = i0; i < total; i += step) {
        const long long ib = b_mode == 1 ? i : b_mode == 2 ? 0 : i % cols;
        out[i] = pw_apply<OP>(PwCast<V, Ta>::run(a[i]), PwCast<V, Tb>::run(b[ib]));
    }
}

// ---- fftfreq / rfftfreq (/root/reference/dsc/src/dsc.cpp:2262-2339): out[i] = i * factor for the non-negative half,
// (i - count) * factor for the negative half that starts at `neg_from` (== count for rfftfreq: no negative half);
// the index is converted to T and multiplied in T, like the reference's `i * factor`
template <typename T>
__global__ void fill_fftfreq(T *__restrict__ out, int count, int neg_from, T factor) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        out[i] = (T)(i < neg_from ? i : i - count) * factor;
}

// ---- strided gather / scatter over a (right-aligned) 4-D index space: the device side of dsc_transpose
// (dsc.cpp:764-827), dsc_tensor_get_slice (:950-1007) and dsc_tensor_set_slice (:1108-1169)
struct Index4 {
    int shape[4];            // extents of the dense side, row-major
    long long stride[4];     // element strides of the strided side
    long long base;          // element offset of the strided side
};

template <typename U>
__global__ void gather_strided(const U *__restrict__ in, U *__restrict__ out, const Index4 g, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int i3 = (int)(r % g.shape[3]); r /= g.shape[3];
        const int i2 = (int)(r % g.shape[2]); r /= g.shape[2];
        const int i1 = (int)(r % g.shape[1]); r /= g.shape[1];
        out[i] = in[g.base + r * g.stride[0] + i1 * g.stride[1] + i2 * g.stride[2] + i3 * g.stride[3]];
    }
}

// dst[selected element i] = src[i mod src_count] (the source is recycled when shorter: scalar or broadcast row)
template <typename U>
__global__ void scatter_strided(U *__restrict__ dst, const U *__restrict__ src, const Index4 g, long long total, long long src_count) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int i3 = (int)(r % g.shape[3]); r /= g.shape[3];
        const int i2 = (int)(r % g.shape[2]); r /= g.shape[2];
        const int i1 = (int)(r % g.shape[1]); r /= g.shape[1];
        dst[g.base + r * g.stride[0] + i1 * g.stride[1] + i2 * g.stride[2] + i3 * g.stride[3]] = src[i % src_count];
    }
}

// batched tiled transpose of the last two dims: out[b][c][r] = in[b][r][c] (coalesced on both sides)
template <typename U>
__global__ void transpose_batched(const U *__restrict__ in, U *__restrict__ out, int rows, int cols, long long tiles_per_batch) {
    __shared__ U tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long batch = blockIdx.x / tiles_per_batch;
    const int t = (int)(blockIdx.x % tiles_per_batch);
    const int tiles_c = (cols + 31) / 32;
    const int c0 = (t % tiles_c) * 32, r0 = (t / tiles_c) * 32;
    const U *src = in + batch * (long long)rows * cols;
    U *dst = out + batch * (long long)rows * cols;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + i][tx] = src[(long long)r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (r < rows && c < cols) dst[(long long)c * rows + r] = tile[tx][ty + i];
    }
}

}  // namespace dscfft
