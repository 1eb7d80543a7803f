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

}  // namespace dscfft
