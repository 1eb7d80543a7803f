// fft_math.cuh -- register-resident complex arithmetic and radix-2/4/8/16 butterflies.
//
// Everything here works on values a thread already holds in registers; all array
// indices are compile-time constants after unrolling, so nothing spills to local
// memory.  Forward means e^{-i..} (the reference's sign = +1 branch,
// /root/reference/dsc/include/dsc_fft.h:57-103,162); inverse is the conjugate.
//
// The file is plain CUDA C++ with no intrinsics so that tests/emul can compile
// the very same code for the host and run thread blocks on pthreads.
#pragma once

#if defined(DSC_EMUL)
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif

namespace dscfft {

#define DSC_DEV __device__ __forceinline__

// bar.sync on one of the 16 hardware barriers for a subset of the block's warps
#if defined(DSC_EMUL)
#define dsc_named_barrier(id, count) __syncthreads()   /* every line runs the same sequence: a block barrier is equivalent */
#define dsc_group_barrier(id, count) __syncthreads()
#else
__device__ __forceinline__ void dsc_named_barrier(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// the same for callers that only ever use barriers 1 and 2: with immediate barrier numbers ptxas reserves three hardware
// barriers for the block instead of all sixteen (which would keep a second block off the SM)
__device__ __forceinline__ void dsc_group_barrier(int id, int count) {
    if (id == 1) asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"r"(count) : "memory");
}
#endif

#if defined(DSC_EMUL)
#define dsc_prefetch_l2(p) ((void)(p))
#else
__device__ __forceinline__ void dsc_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// counter += 1 with release semantics at GPU scope: orders the block's earlier writes (observed through a
// block barrier) before the increment WITHOUT the L1 invalidation a full __threadfence() carries -- the
// twiddle tables of a persistent block stay in L1 from tile to tile
#if defined(DSC_EMUL)
#define dsc_signal_release(p) ((void)atomicAdd((p), 1u))
#else
__device__ __forceinline__ void dsc_signal_release(unsigned *p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(1u) : "memory");
}
#endif

// dynamic shared memory of the running block
#if defined(DSC_EMUL)
#define DSC_DYN_SMEM(name) unsigned char *name = dsc_emul::tls.smem
#else
#define DSC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

template <typename T> struct vec2;
template <> struct vec2<float>  { using type = float2; };
template <> struct vec2<double> { using type = double2; };
template <typename T> using cx = typename vec2<T>::type;

template <typename T> DSC_DEV cx<T> mk(T re, T im) { cx<T> r; r.x = re; r.y = im; return r; }
template <typename V> DSC_DEV V cadd(V a, V b) { a.x += b.x; a.y += b.y; return a; }
template <typename V> DSC_DEV V csub(V a, V b) { a.x -= b.x; a.y -= b.y; return a; }

// a * w (FWD) or a * conj(w) (inverse); w is always a FORWARD twiddle e^{-i th}.
template <bool FWD, typename V> DSC_DEV V cmul_tw(V a, V w) {
    V r;
    if (FWD) { r.x = a.x * w.x - a.y * w.y; r.y = a.x * w.y + a.y * w.x; }
    else     { r.x = a.x * w.x + a.y * w.y; r.y = a.y * w.x - a.x * w.y; }
    return r;
}
template <typename V> DSC_DEV V cmul(V a, V b) {
    V r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}

#if defined(DSC_F32X2) && !defined(DSC_EMUL)
// Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): one instruction per complex add, two per complex multiply.
// The FP32 pipe does the same lane-work either way, but the kernels that define DSC_F32X2 are bound by instruction ISSUE
// (two-pass transforms: ~75 instructions per point, 57 % of the issue slots at the power-capped clock), and these halve
// the floating-point instruction count.  Non-template overloads: preferred over the generic templates for float2.
DSC_DEV float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
DSC_DEV float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
DSC_DEV float2 cmul(float2 a, float2 b) {
    return __ffma2_rn(make_float2(a.x, a.x), b, __fmul2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x)));
}
template <bool FWD> DSC_DEV float2 cmul_tw(float2 a, float2 w) {
    if (FWD) return __ffma2_rn(make_float2(a.x, a.x), w, __fmul2_rn(make_float2(a.y, a.y), make_float2(-w.y, w.x)));
    return __ffma2_rn(make_float2(a.x, a.x), make_float2(w.x, -w.y), __fmul2_rn(make_float2(a.y, a.y), make_float2(w.y, w.x)));
}
// a * (c -+ i s): the constant rotations of the radix-8/16/32 butterflies
template <bool FWD> DSC_DEV float2 crot(float2 a, const float c, const float s) {
    return FWD ? __ffma2_rn(a, make_float2(c, c), __fmul2_rn(make_float2(a.y, a.x), make_float2(s, -s)))
               : __ffma2_rn(a, make_float2(c, c), __fmul2_rn(make_float2(a.y, a.x), make_float2(-s, s)));
}
#define DSC_HAVE_CROT 1
#endif

// a * (-i) forward, a * (+i) inverse
template <bool FWD, typename V> DSC_DEV V rot90(V a) {
    V r;
    if (FWD) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; }
    return r;
}

template <typename T> struct consts {
    static constexpr T sqrt1_2 = (T)0.70710678118654752440;
    static constexpr T cos_pi_8 = (T)0.92387953251128675613;
    static constexpr T sin_pi_8 = (T)0.38268343236508977173;
};

// a * w8^q, w8 = e^{-+ i pi/4}
template <bool FWD, int Q, typename T> DSC_DEV cx<T> mul_w8(cx<T> a) {
    constexpr T h = consts<T>::sqrt1_2;
    if (Q == 0) return a;
    if (Q == 2) return rot90<FWD>(a);
#if defined(DSC_HAVE_CROT)
    if constexpr (sizeof(T) == 4) {
        if (Q == 1) return crot<FWD>(a, h, h);          // (1 -+ i) / sqrt 2
        return crot<FWD>(a, -h, h);                     // (-1 -+ i) / sqrt 2
    }
#endif
    if (Q == 1) return FWD ? mk<T>((a.x + a.y) * h, (a.y - a.x) * h)
                           : mk<T>((a.x - a.y) * h, (a.x + a.y) * h);
    /* Q == 3 */ return FWD ? mk<T>((a.y - a.x) * h, -(a.x + a.y) * h)
                            : mk<T>(-(a.x + a.y) * h, (a.x - a.y) * h);
}

// a * w16^q for q in 0..9 (all the 4x4 decomposition needs), w16 = e^{-+ i pi/8}
template <bool FWD, int Q, typename T> DSC_DEV cx<T> mul_w16(cx<T> a) {
    constexpr T c1 = consts<T>::cos_pi_8, s1 = consts<T>::sin_pi_8;
    if (Q % 2 == 0) return mul_w8<FWD, (Q / 2) % 4, T>(Q >= 8 ? mk<T>(-a.x, -a.y) : a);
    // odd q: cos/sin of q*pi/8 from the pi/8 pair
    constexpr T c = (Q == 1) ? c1 : (Q == 3) ? s1 : (Q == 5) ? -s1 : (Q == 7) ? -c1 : (Q == 9) ? -c1 : /*11*/ -s1;
    constexpr T s = (Q == 1) ? s1 : (Q == 3) ? c1 : (Q == 5) ? c1 : (Q == 7) ? s1 : (Q == 9) ? -s1 : /*11*/ -c1;
#if defined(DSC_HAVE_CROT)
    if constexpr (sizeof(T) == 4) return crot<FWD>(a, c, s);
#endif
    return FWD ? mk<T>(a.x * c + a.y * s, a.y * c - a.x * s)
               : mk<T>(a.x * c - a.y * s, a.y * c + a.x * s);
}

template <typename V> DSC_DEV void bfly2(V &a, V &b) {
    const V t = a; a = cadd(t, b); b = csub(t, b);
}

// in-place DFT-4, natural order in and out
template <bool FWD, typename V> DSC_DEV void bfly4(V &x0, V &x1, V &x2, V &x3) {
    const V t0 = cadd(x0, x2), t1 = csub(x0, x2);
    const V t2 = cadd(x1, x3), t3 = rot90<FWD>(csub(x1, x3));
    x0 = cadd(t0, t2); x2 = csub(t0, t2);
    x1 = cadd(t1, t3); x3 = csub(t1, t3);
}

// DFT-R of v[0..R), natural order in and out.  R in {2,4,8,16}.
template <int R, bool FWD, typename T> struct Dft;

template <bool FWD, typename T> struct Dft<1, FWD, T> {
    static DSC_DEV void run(cx<T> (&)[1]) {}
};
template <bool FWD, typename T> struct Dft<2, FWD, T> {
    static DSC_DEV void run(cx<T> (&v)[2]) { bfly2(v[0], v[1]); }
};
template <bool FWD, typename T> struct Dft<4, FWD, T> {
    static DSC_DEV void run(cx<T> (&v)[4]) { bfly4<FWD>(v[0], v[1], v[2], v[3]); }
};
template <bool FWD, typename T> struct Dft<8, FWD, T> {
    // m = 2a + b: two DFT-4 over a, twiddle w8^{b p1}, four DFT-2 over b
    static DSC_DEV void run(cx<T> (&v)[8]) {
        bfly4<FWD>(v[0], v[2], v[4], v[6]);
        bfly4<FWD>(v[1], v[3], v[5], v[7]);
        v[3] = mul_w8<FWD, 1, T>(v[3]);
        v[5] = mul_w8<FWD, 2, T>(v[5]);
        v[7] = mul_w8<FWD, 3, T>(v[7]);
        bfly2(v[0], v[1]); bfly2(v[2], v[3]); bfly2(v[4], v[5]); bfly2(v[6], v[7]);
        // v[2 p1 + p2] holds X[p1 + 4 p2]  ->  natural order
        const cx<T> x1 = v[2], x2 = v[4], x3 = v[6], x4 = v[1], x5 = v[3], x6 = v[5];
        v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6;
    }
};
template <bool FWD, typename T> struct Dft<16, FWD, T> {
    // m = 4a + b: four DFT-4 over a, twiddle w16^{b p1}, four DFT-4 over b
    static DSC_DEV void run(cx<T> (&v)[16]) {
        bfly4<FWD>(v[0], v[4], v[8], v[12]);
        bfly4<FWD>(v[1], v[5], v[9], v[13]);
        bfly4<FWD>(v[2], v[6], v[10], v[14]);
        bfly4<FWD>(v[3], v[7], v[11], v[15]);
        // v[b + 4 p1] = y_b[p1]
        v[5]  = mul_w16<FWD, 1, T>(v[5]);  v[6]  = mul_w16<FWD, 2, T>(v[6]);  v[7]  = mul_w16<FWD, 3, T>(v[7]);
        v[9]  = mul_w16<FWD, 2, T>(v[9]);  v[10] = mul_w16<FWD, 4, T>(v[10]); v[11] = mul_w16<FWD, 6, T>(v[11]);
        v[13] = mul_w16<FWD, 3, T>(v[13]); v[14] = mul_w16<FWD, 6, T>(v[14]); v[15] = mul_w16<FWD, 9, T>(v[15]);
        bfly4<FWD>(v[0], v[1], v[2], v[3]);
        bfly4<FWD>(v[4], v[5], v[6], v[7]);
        bfly4<FWD>(v[8], v[9], v[10], v[11]);
        bfly4<FWD>(v[12], v[13], v[14], v[15]);
        // v[4 p1 + p2] holds X[p1 + 4 p2]: transpose the 4x4 register tile
#define DSC_SWAP(i, j) { const cx<T> s_ = v[i]; v[i] = v[j]; v[j] = s_; }
        DSC_SWAP(1, 4) DSC_SWAP(2, 8) DSC_SWAP(3, 12) DSC_SWAP(6, 9) DSC_SWAP(7, 13) DSC_SWAP(11, 14)
#undef DSC_SWAP
    }
};

// cos / sin of 2 pi q / 32 as compile-time constants (q in 0..31)
template <typename T> struct w32 {
    static constexpr double c_[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                     0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                                     0.19509032201612826785, 0.0};
    static __host__ __device__ constexpr T cosq(int q) {
        q &= 31;
        if (q > 16) q = 32 - q;
        return q <= 8 ? (T)c_[q] : (T)-c_[16 - q];
    }
    static __host__ __device__ constexpr T sinq(int q) {
        q &= 31;
        const bool neg = q > 16;
        if (neg) q = 32 - q;
        const T v = q <= 8 ? (T)c_[8 - q] : (T)c_[q - 8];
        return neg ? -v : v;
    }
};

// a * w32^q, w32 = e^{-+ 2 pi i / 32}
template <bool FWD, int Q, typename T> DSC_DEV cx<T> mul_w32(cx<T> a) {
    if constexpr (Q % 2 == 0) return mul_w16<FWD, Q / 2, T>(a);
    else {
        constexpr T c = w32<T>::cosq(Q), s = w32<T>::sinq(Q);
#if defined(DSC_HAVE_CROT)
        if constexpr (sizeof(T) == 4) return crot<FWD>(a, c, s);
#endif
        return FWD ? mk<T>(a.x * c + a.y * s, a.y * c - a.x * s)
                   : mk<T>(a.x * c - a.y * s, a.y * c + a.x * s);
    }
}

template <bool FWD, typename T> struct Dft<32, FWD, T> {
    // m = 8a + b: eight DFT-4 over a, twiddle w32^{b p1}, four DFT-8 over b
    static DSC_DEV void run(cx<T> (&v)[32]) {
#pragma unroll
        for (int b = 0; b < 8; ++b) bfly4<FWD>(v[b], v[b + 8], v[b + 16], v[b + 24]);
        // v[b + 8 p1] = y_b[p1]
#define DSC_TW(B, P) v[B + 8 * P] = mul_w32<FWD, B * P, T>(v[B + 8 * P]);
        DSC_TW(1, 1) DSC_TW(2, 1) DSC_TW(3, 1) DSC_TW(4, 1) DSC_TW(5, 1) DSC_TW(6, 1) DSC_TW(7, 1)
        DSC_TW(1, 2) DSC_TW(2, 2) DSC_TW(3, 2) DSC_TW(4, 2) DSC_TW(5, 2) DSC_TW(6, 2) DSC_TW(7, 2)
        DSC_TW(1, 3) DSC_TW(2, 3) DSC_TW(3, 3) DSC_TW(4, 3) DSC_TW(5, 3) DSC_TW(6, 3) DSC_TW(7, 3)
#undef DSC_TW
        cx<T> out[32];
#pragma unroll
        for (int p1 = 0; p1 < 4; ++p1) {
            cx<T> g[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) g[b] = v[8 * p1 + b];
            Dft<8, FWD, T>::run(g);
#pragma unroll
            for (int p2 = 0; p2 < 8; ++p2) out[p1 + 4 * p2] = g[p2];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = out[i];
    }
};

}  // namespace dscfft
