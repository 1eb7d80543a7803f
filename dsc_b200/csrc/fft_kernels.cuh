// fft_kernels.cuh -- shared-memory-staged mixed-radix Stockham line transforms.
//
// One thread block transforms LPB lines of N = 2^LG_N complex points.  A line is
// owned by TT = N/E threads; each thread keeps E = 2^LG_E points in registers
// (x[t + c*TT], c < E) and the stages are
//
//     radix-E, radix-E, ..., radix-(N / E^k)        (greedy, big radices first)
//
// Stage s (sub-transform length NS = E^s before it) does, for butterfly j:
//     k = j mod NS;  in_m = x[j + m*N/R] * W_{NS*R}^{k m};  DFT-R;
//     y[(j-k)*R + k + p*NS] = out_p
// which is the Stockham auto-sort recurrence: reads are always x[t + c*TT]
// (conflict-free, coalesced when they come from HBM), the scatter goes through
// shared memory, and after the LAST stage the outputs are again at t + c*TT so
// they leave for HBM coalesced.  HBM traffic is one read and one write of the
// line; twiddles come from small per-stage tables (plan.cu) that stay in L1/L2.
//
// What the reference does per line instead (all fused here):
//   gather + cast + zero-pad/crop   /root/reference/dsc/src/dsc.cpp:1981-1994
//   log2(N) radix-2 levels          /root/reference/dsc/include/dsc_fft.h:57-103
//   1/N on the inverse              /root/reference/dsc/include/dsc_fft.h:168-175
//   real post/pre-processing        /root/reference/dsc/include/dsc_fft.h:199-237
//   scatter                         /root/reference/dsc/src/dsc.cpp:1998-2003
#pragma once

#include "fft_math.cuh"

namespace dscfft {

#ifndef DSC_TW_LADDER
#define DSC_TW_LADDER 1
#endif

enum Mode : int {
    MODE_C2C = 0,   // complex / real-cast / real-pair / staged-row input -> complex out
    MODE_R2C = 1,   // rfft : 2N reals -> N+1 bins, un-mixing fused after the last stage
    MODE_C2R = 2,   // irfft: N+1 bins -> 2N reals, mixing fused before the first stage
    MODE_FAST = 3,  // dense complex lines (inner == 1, no pad/crop, whole blocks): the bandwidth path --
                    // one base pointer per thread, immediate offsets, streaming cache hints
    MODE_R2C_FAST = 4,  // MODE_R2C / MODE_C2R for dense last-axis lines in whole blocks, full-length input: compile-
    MODE_C2R_FAST = 5,  // time geometry, vector accesses, no shared-memory pass outside the butterfly exchanges
    MODE_FILTER = 6,  // rfft -> times a spectrum -> irfft in ONE kernel: the spectrum never leaves shared memory
};
__host__ __device__ constexpr bool mode_is_dense(int mode) { return mode == MODE_FAST || mode == MODE_R2C_FAST || mode == MODE_C2R_FAST; }

// how MODE_C2C reads its input
enum InKind : int {
    IN_COMPLEX = 0,  // cx<T> elements, straight into registers
    IN_REAL = 1,     // T elements, imag = 0 (cast_op, /root/reference/dsc/include/dsc_ops.h:12-23)
    IN_PAIRS = 2,    // two T elements pstride apart form one complex (packed real transform)
    IN_ROWS = 3,     // cx<T> rows, loaded cooperatively through shared memory
};

// Geometry of one launch.  A "line" index L decomposes as (o, in) = (L / inner, L % inner);
// element i of the line sits at   o*ostride + in*lstride + i*estride   (in elements of the
// respective dtype).  For an (outer, n, inner) tensor: ostride = n*inner, lstride = 1,
// estride = inner -- the reference's dsc_axis_iterator order, dsc_iter.h:19-47.
struct LineGeom {
    long long ostride, lstride, estride;
};

struct FftArgs {
    const void *x;
    void *out;
    long long lines;
    long long inner;
    LineGeom gi, go;
    long long in_limit;   // element (i, in) is read iff i*gi.estride + in*gi.lstride < in_limit, else 0
    const void *tw[5];    // tw[s]: stage-s table, (R_s - 1) x NS_s forward twiddles, [m-1][k]
    const void *tw_real;  // W_{2N}^k, k <= N/2            (MODE_R2C / MODE_C2R)
    const void *tw_lo;    // four-step: W_M^p,        p <  2^four_shift
    const void *tw_hi;    // four-step: W_M^(p << four_shift)
    int four_shift;       // 0 = no inter-pass twiddle
    int four_mask;
    long long gi_pstride; // IN_PAIRS / MODE_R2C: distance between the two reals of a pair
    long long go_pstride; // MODE_C2R: same, for the output
    int in_kind;          // InKind (MODE_C2C only)
    int strided;          // thread mapping: 1 = adjacent lines on adjacent lanes
    int inner_shift;      // log2(inner) when inner is a power of two, else -1
    int no_limit;         // every line exists and is read in full: skip the pad/crop predicates
    const void *filt;     // MODE_FILTER: N+1 bins multiplied into every line's spectrum
    long long ring_in;    // input / output row index taken modulo this (0 = off): ring of work rows
    long long ring_out;
    double scale;         // applied to the outputs when do_scale (1/N of the inverse)
    int do_scale;
    int packed_in;        // real pairs are adjacent and vector-aligned in every line: load them as one complex
    int packed_out;       // same for the real output of MODE_C2R / MODE_FILTER
    int seg_shift;        // four-step first pass, dense rows: a row is made of segments of 2^seg_shift elements that lie
    long long seg_extra;  // seg_extra elements further apart than their length (0 / 0: one contiguous row)
    int seg_self;         // segment index whose data lives in ANOTHER buffer (-1: none): the part of the multi-GPU
    long long seg_self_delta;   // exchange that never left this GPU; element offset of that buffer from x
    int keep_out;         // four-step second pass: 1 = a later kernel re-reads the output soon (plain stores, the
                          // rows stay in L2); 0 = streaming stores
    int out_take;         // MODE_C2R / MODE_FILTER: store only the first out_take real samples of every line (0 = all):
                          // the README's irfft(...)[:output_length] crop (README.md:130-133) fused into the store
};

template <int LG_N, int LG_E> struct Sched {
    static constexpr int N = 1 << LG_N;
    static constexpr int E = 1 << LG_E;
    static constexpr int TT = N / E;
    static constexpr int STAGES = LG_E == 0 ? 1 : (LG_N + LG_E - 1) / LG_E;
    static __host__ __device__ constexpr int lg_r(int s) { return (LG_N - s * LG_E) < LG_E ? (LG_N - s * LG_E) : LG_E; }
    // Padded length of one line in shared memory.  In the strided thread mapping adjacent lanes are
    // adjacent LINES at the same position, and when fewer than a full bank phase of lines fit in a block
    // the next lanes are the next position: the line stride must map (line, position) pairs of one phase
    // (16 lanes of 8-byte elements, 8 lanes of 16-byte elements) onto distinct banks, i.e. be congruent
    // to phase/LPB modulo the phase.
    // contiguous_tt > 0: the mapping is known to be "TT consecutive positions, then the next line" (MODE_FAST),
    // and for short lines (TT < phase) a bank phase spans phase/TT lines: the stride must then be congruent to TT.
    static __host__ __device__ constexpr int line_stride(int lpb, int elem_bytes, int contiguous_tt = 0) {
        const int phase = 128 / elem_bytes;
        const int want = (contiguous_tt > 0 && contiguous_tt < phase) ? contiguous_tt : lpb >= phase ? 1 : phase / lpb;
        const int base = N + (N >> LG_E);
        return base + ((want - base % phase) % phase + phase) % phase;
    }
    static DSC_DEV int pad(int i) { return i + (i >> LG_E); }
    // padded index of element t + c*TT given pt = pad(t): when TT is a multiple of E the padding of the
    // two terms separates, so the compiler sees base + compile-time constant (an immediate offset)
    static DSC_DEV int pad_read(int t, int pt, int c) {
        if constexpr (TT % E == 0) return pt + c * (TT + TT / E);
        else return pad(t + c * TT);
    }
    static_assert(STAGES <= 5, "FftArgs::tw / dsc_cuda_plan::tw1 hold 5 stage tables");
};

// Barrier scope of the exchanges of one line.
enum SyncKind : int { SYNC_WARP = 0, SYNC_BLOCK = 1, SYNC_LINE = 2 };
struct LineSync {
    int kind;     // SyncKind
    int id;       // hardware barrier of this line (SYNC_LINE), 1..15
    int count;    // threads of the line
};
DSC_DEV void line_sync(const LineSync &s) {
    if (s.kind == SYNC_WARP) __syncwarp();
    else if (s.kind == SYNC_BLOCK) __syncthreads();
    else dsc_named_barrier(s.id, s.count);
}

template <typename T, int LG_N, int LG_E, bool FWD, int S> struct Stage {
    using Sc = Sched<LG_N, LG_E>;
    using V = cx<T>;
    static constexpr int E = Sc::E, TT = Sc::TT;
    static constexpr int LG_R = Sc::lg_r(S), R = 1 << LG_R, NB = E / R;
    static constexpr int LG_NS = S * LG_E, NS = 1 << LG_NS;
    static constexpr bool LAST = (S == Sc::STAGES - 1);

    struct NoHook { DSC_DEV void operator()() const {} };
    static DSC_DEV void run(V (&v)[E], V *sm, const int t, const FftArgs &a, const LineSync &ls) {
        run(v, sm, t, sm, t, a, ls, NoHook{});
    }
    // (sm_r, t_r): the line buffer and position this thread continues with AFTER this stage's exchange.  The
    // exchange is a free permutation point: a thread may come back as a different (line, position) of the
    // block -- the four-step second pass reads contiguous rows one line per warp and writes with adjacent
    // lanes on adjacent lines.  Needs a block-wide barrier when the two differ.
    // before_scatter(): called once, after this stage's butterflies and before its first write to shared memory
    // (the persistent four-step kernel puts its "line buffers are free again" barrier there).
    template <typename Hook>
    static DSC_DEV void run(V (&v)[E], V *sm, const int t, V *sm_r, const int t_r, const FftArgs &a, const LineSync &ls,
                            Hook &&before_scatter) {
        const V *__restrict__ tw = (const V *)a.tw[S];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int j = t + b * TT;
            const int k = j & (NS - 1);
            V r[R];
#pragma unroll
            for (int m = 0; m < R; ++m) r[m] = v[b + m * NB];
            if constexpr (S > 0) {
                if constexpr (DSC_TW_LADDER && R >= 8) {
                    // W^(k m) for all m from the log2(R) table rows m = 1, 2, 4, ...: w_m = w_hi(m) * w_(m - hi(m)).
                    // At most log2(R) - 1 roundings deep; trades R - 1 - log2(R) loads (the load/store pipe is
                    // the busiest unit of these kernels: 76-82 % on the 4096-point line) for as many complex
                    // multiplies on the FP pipe.  Same-box A/B on B200: complex64 4096 points +4.7 % (0.86 -> 0.90 of
                    // the copy peak), 2^14 +4 %, rfft 2^12-2^14 +4..10 %, batched filter 2^13 +15 %, four-step
                    // lengths +-1 %; complex128 2^10 / 2^12 / 2^13 +6 / +11 / +4 %.  Round-trip error of the
                    // 4096-point float transform 1.8e-7 -> 2.05e-7 (tolerance 1e-5).
                    V w[R];
#pragma unroll
                    for (int m = 1; m < R; ++m) {
                        if ((m & (m - 1)) == 0) w[m] = __ldg(tw + (m - 1) * NS + k);
                        else {
                            int hi = 1;
                            while (hi * 2 <= m) hi *= 2;
                            w[m] = cmul(w[hi], w[m - hi]);
                        }
                        r[m] = cmul_tw<FWD>(r[m], w[m]);
                    }
                } else {
#pragma unroll
                    for (int m = 1; m < R; ++m) r[m] = cmul_tw<FWD>(r[m], __ldg(tw + (m - 1) * NS + k));
                }
            }
            Dft<R, FWD, T>::run(r);
            if constexpr (LAST) {
#pragma unroll
                for (int p = 0; p < R; ++p) v[b + p * NB] = r[p];
            } else {
                if (b == 0) before_scatter();
                // scatter to (j-k)*R + k + p*NS; p*NS is a multiple of E past the first stage, so its
                // padding is a constant too
                const int base = ((j - k) << LG_R) + k;
                const int pbase = Sc::pad(base);
#pragma unroll
                for (int p = 0; p < R; ++p) {
                    if constexpr (NS % E == 0) sm[pbase + p * (NS + NS / E)] = r[p];
                    else if constexpr (NS == 1 && R == E) sm[pbase + p] = r[p];
                    else sm[Sc::pad(base + p * NS)] = r[p];
                }
            }
        }
        if constexpr (!LAST) {
            line_sync(ls);
            const int pt = Sc::pad(t_r);
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = sm_r[Sc::pad_read(t_r, pt, c)];
            if constexpr (S + 1 < Sc::STAGES - 1) line_sync(ls);     // the next stage scatters into the same buffer
            Stage<T, LG_N, LG_E, FWD, S + 1>::run(v, sm_r, t_r, sm_r, t_r, a, ls, NoHook{});
        }
    }
};

// W_M^p from the two sqrt(M)-sized tables; p = n2 * k1 < M <= 2^30 fits 32 bits
template <typename T> DSC_DEV cx<T> four_step_twiddle(const FftArgs &a, const unsigned p) {
    const cx<T> lo = __ldg((const cx<T> *)a.tw_lo + (p & (unsigned)a.four_mask));
    const cx<T> hi = __ldg((const cx<T> *)a.tw_hi + (p >> a.four_shift));
    return cmul(lo, hi);
}

// Streaming accesses for the payload: each element is touched exactly once, so it should not
// displace the twiddle tables from L1 / linger in L2 (ld.global.cs / st.global.cs).
#if defined(DSC_EMUL)
template <typename V> DSC_DEV V ld_stream(const V *p) { return *p; }
#else
// no L1 allocation at all: the payload must not evict the twiddle tables, which are the only data with reuse
DSC_DEV float2 ld_stream(const float2 *p) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
DSC_DEV double2 ld_stream(const double2 *p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
#endif
template <typename V> DSC_DEV void st_stream(V *p, const V v) { __stcs(p, v); }

// Un-mix / mix step of the packed real transform for one bin pair (k, N-k).
// Forward (dsc_fft.h:199-214 with c = -1/2, w = W_2N^k):   X[k], X[N-k] from Z[k], Z[N-k].
// Inverse (same lines with c = +1/2, w = conj W_2N^k):     z[k], z[N-k] from X[k], X[N-k].
template <bool FWD, typename T>
DSC_DEV void real_pair(const cx<T> a, const cx<T> b, const cx<T> w_fwd, cx<T> &ra, cx<T> &rb) {
    const T h1r = (T)0.5 * (a.x + b.x), h1i = (T)0.5 * (a.y - b.y);
    const T c = FWD ? (T)-0.5 : (T)0.5;
    const T h2r = -c * (a.y + b.y), h2i = c * (a.x - b.x);
    const T wr = w_fwd.x, wi = FWD ? w_fwd.y : -w_fwd.y;
    ra = mk<T>(h1r + wr * h2r - wi * h2i,  h1i + wr * h2i + wi * h2r);
    rb = mk<T>(h1r - wr * h2r + wi * h2i, -h1i + wr * h2i + wi * h2r);
}

// rfft un-mix, spectrum product, irfft mix for one bin pair (k, N-k) in a single step:
// (Z[k], Z[N-k]) -> (X[k], X[N-k]) -> times (B[k], B[N-k]) -> packed (z'[k], z'[N-k]).
template <typename T>
DSC_DEV void filter_pair(const cx<T> za, const cx<T> zb, const cx<T> w_fwd, const cx<T> ba, const cx<T> bb,
                         cx<T> &ra, cx<T> &rb) {
    cx<T> xa, xb;
    real_pair<true, T>(za, zb, w_fwd, xa, xb);
    real_pair<false, T>(cmul(xa, ba), cmul(xb, bb), w_fwd, ra, rb);
}
// the self-paired bins: DC + Nyquist (from Z[0]) and, for N >= 2, bin N/2
template <typename T>
DSC_DEV cx<T> filter_dc(const cx<T> z0, const cx<T> b0, const cx<T> bn) {
    // X[0] = (z0.x + z0.y, 0), X[N] = (z0.x - z0.y, 0); only the REAL parts of the products enter the inverse
    const T y0 = (z0.x + z0.y) * b0.x, yn = (z0.x - z0.y) * bn.x;
    return mk<T>((T)0.5 * (y0 + yn), (T)0.5 * (y0 - yn));
}
template <typename T>
DSC_DEV cx<T> filter_mid(const cx<T> zh, const cx<T> bh) {
    const cx<T> y = cmul(mk<T>(zh.x, -zh.y), bh);     // X[N/2] = conj(Z[N/2])
    return mk<T>(y.x, -y.y);                          // z'[N/2] = conj(Y[N/2])
}

// Payload access with a compile-time cache policy.
enum CachePol : int { POL_DEFAULT = 0, POL_STREAM = 1, POL_L2 = 2 };
template <int POL, typename V> DSC_DEV V ld_pol(const V *p) {
    if constexpr (POL == POL_STREAM) return __ldcs(p);
    else if constexpr (POL == POL_L2) return __ldcg(p);      // coherent at L2: data another SM just produced
    else return *p;
}
template <int POL, typename V> DSC_DEV void st_pol(V *p, const V v) {
    if constexpr (POL == POL_STREAM) __stcs(p, v);
    else *p = v;
}

// Body of one thread block: LPB lines starting at line `block * LPB`.
// THREADS = LPB * TT.  Dynamic shared memory: LPB * Sched::LINE * sizeof(cx<T>).
template <typename T, int LG_N, int LG_E, int LPB, bool FWD, int MODE>
DSC_DEV void fft_lines_body(const FftArgs &a, const long long block, unsigned char *smem_raw) {
    using Sc = Sched<LG_N, LG_E>;
    using V = cx<T>;
    constexpr int N = Sc::N, E = Sc::E, TT = Sc::TT, THREADS = LPB * TT;
    constexpr int PAIRS = E > 1 ? E / 2 : 1;   // bin pairs per thread in the real modes
    constexpr bool IS_C2C = MODE == MODE_C2C;
    constexpr int POL_IN = POL_DEFAULT, POL_OUT = POL_DEFAULT;
    V *sm_all = (V *)smem_raw;

    const int tid = threadIdx.x;
    int l, t;
    if (a.strided) { l = tid % LPB; t = tid / LPB; } else { l = tid / TT; t = tid % TT; }
    constexpr int LINE = Sc::line_stride(LPB, (int)sizeof(V), mode_is_dense(MODE) ? TT : 0);
    V *sm = sm_all + l * LINE;
    // A line's exchanges need: a warp barrier when the line lives inside one warp; its own hardware
    // barrier when it is a whole number of warps and the block has few enough lines (lines then
    // progress independently of each other); otherwise the block barrier.
    LineSync ls;
    ls.id = 1 + l; ls.count = TT;
    if (a.strided) ls.kind = SYNC_BLOCK;
    else if (TT <= 32) ls.kind = SYNC_WARP;
    else ls.kind = SYNC_BLOCK;

    const long long line = block * LPB + l;
    const bool active = line < a.lines;
    long long o = 0, in = 0;
    if (active) {
        if (a.inner_shift >= 0) { o = line >> a.inner_shift; in = line & ((1LL << a.inner_shift) - 1); }
        else { o = line / a.inner; in = line - o * a.inner; }
    }
    const long long o_in = a.ring_in ? o % a.ring_in : o;
    const long long o_out = a.ring_out ? o % a.ring_out : o;
    const long long ibase = o_in * a.gi.ostride + in * a.gi.lstride;
    const long long obase = o_out * a.go.ostride + in * a.go.lstride;
    // element offsets are t*estride + c*(TT*estride): one multiply per thread, then additions
    const long long ioff0 = (long long)t * a.gi.estride, istep = (long long)TT * a.gi.estride;
    const long long ooff0 = (long long)t * a.go.estride, ostep = (long long)TT * a.go.estride;
    // read iff offset < lim; "no limit" (dense lines) is encoded as a huge lim by the host
    const long long lim = active ? a.in_limit - in * a.gi.lstride : 0;
    const V zero = mk<T>((T)0, (T)0);
    V v[E];

    // ---------------------------------------------------------------- load
    if constexpr (MODE == MODE_FAST) {
        // every line of the block exists and is dense: a.x / a.out are (lines, N) row-major
        if constexpr (TT >= 4) {
            const V *__restrict__ xp = (const V *)a.x + line * N + t;
#pragma unroll
            for (int c = 0; c < E; ++c) {
                // with fewer than 32 threads per line a warp's accesses are only coalesced ACROSS the c loop:
                // let L1 merge them; full warps per line stream past L1
                if constexpr (TT >= 32) v[c] = ld_stream(xp + c * TT);
                else v[c] = __ldcs(xp + c * TT);
            }
        } else {
            // short lines (a thread owns most of a line): the block's LPB lines are one contiguous run of
            // LPB*N elements -- copy it coalesced into shared memory, then every thread picks up its points
            const V *__restrict__ xb = (const V *)a.x + block * (long long)(LPB * N);
            for (int e = tid; e < LPB * N; e += THREADS) sm_all[(e / N) * LINE + Sc::pad(e % N)] = __ldcs(xb + e);
            __syncthreads();
            const int pt = Sc::pad(t);
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = sm[Sc::pad_read(t, pt, c)];
            __syncthreads();
        }
    } else if constexpr (MODE == MODE_R2C_FAST) {
        // 2N dense reals = N complex, one vector load per point
        const V *__restrict__ xp = (const V *)a.x + line * N + t;
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = ld_stream(xp + c * TT);
    } else if constexpr (MODE == MODE_C2R_FAST) {
        // Every thread builds the packed points z[t + c TT] it needs for the first butterfly itself, from the bin
        // pair (k, N-k) of each: two loads per point (every bin is wanted by two threads of the block; the second
        // request hits L1) instead of a pass through shared memory with two block barriers before the transform
        // can start.  z[k] is the first result of the pair step when k < N/2 and the second one of pair N-k
        // otherwise (dsc_fft.h:199-214, c = +1/2).
        // (instantiated for every length of the table; the launcher only uses it from real_fast_min_lg() up)
        const V *__restrict__ xc = (const V *)a.x + line * (N + 1);
        const V *__restrict__ twr = (const V *)a.tw_real;
#pragma unroll
        for (int c = 0; c < E; ++c) {
            constexpr int HALF = E / 2;
            const bool upper = c >= HALF;
            const int k = t + c * TT;
            const int kk = upper ? N - k : k;
            const V lo = __ldg(xc + kk), hi = __ldg(xc + (N - kk));
            V ra, rb;
            real_pair<false, T>(lo, hi, __ldg(twr + kk), ra, rb);
            v[c] = upper ? rb : ra;
            if (t == 0) {
                if (c == 0) v[c] = mk<T>((T)0.5 * (lo.x + hi.x), (T)0.5 * (lo.x - hi.x));   // DC + Nyquist, real parts only
                if (c == HALF) v[c] = mk<T>(lo.x, -lo.y);                                   // bin N/2
            }
        }
    } else if constexpr (MODE == MODE_C2R) {
        // bins X[0..N] -> packed z[0..N); DC/Nyquist use real parts only (dsc_fft.h:227-228)
        const V *__restrict__ xc = (const V *)a.x + ibase;
        const V *__restrict__ twr = (const V *)a.tw_real;
        auto bin = [&](int i) -> V {
            const long long off = (long long)i * a.gi.estride;
            return off < lim ? xc[off] : zero;
        };
#pragma unroll
        for (int c = 0; c < PAIRS; ++c) {
            const int k = t + c * TT;
            if (c == 0 && k == 0) {
                const V x0 = bin(0), xn = bin(N);
                sm[Sc::pad(0)] = mk<T>((T)0.5 * (x0.x + xn.x), (T)0.5 * (x0.x - xn.x));
                if (N >= 2) { const V xh = bin(N / 2); sm[Sc::pad(N / 2)] = mk<T>(xh.x, -xh.y); }
            } else if (k < N / 2) {
                V za, zb;
                real_pair<false, T>(bin(k), bin(N - k), __ldg(twr + k), za, zb);
                sm[Sc::pad(k)] = za;
                sm[Sc::pad(N - k)] = zb;
            }
        }
        __syncthreads();
        const int pt = Sc::pad(t);
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = sm[Sc::pad_read(t, pt, c)];
        __syncthreads();
    } else if (MODE == MODE_R2C || MODE == MODE_FILTER || a.in_kind == IN_PAIRS) {
        // 2N reals seen as N complex: z[j] = (x[..2j], x[..2j+1]); gi is in REAL elements
        const T *__restrict__ xr = (const T *)a.x + ibase;
        if (a.packed_in) {
            // adjacent reals in aligned rows (last axis): one vector load per pair instead of two scalar loads
            // that each use half of every sector
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const long long off = ioff0 + c * istep;
                if (off + 1 < lim) v[c] = __ldcs((const V *)(xr + off));
                else v[c] = mk<T>(off < lim ? xr[off] : (T)0, (T)0);
            }
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const long long off = ioff0 + c * istep;
                const T re = off < lim ? ld_pol<POL_IN>(xr + off) : (T)0;
                const T im = off + a.gi_pstride < lim ? ld_pol<POL_IN>(xr + off + a.gi_pstride) : (T)0;
                v[c] = mk<T>(re, im);
            }
        }
    } else if (a.in_kind == IN_REAL) {
        const T *__restrict__ xr = (const T *)a.x + ibase;
#pragma unroll
        for (int c = 0; c < E; ++c) {
            const long long off = ioff0 + c * istep;
            v[c] = mk<T>(off < lim ? ld_pol<POL_IN>(xr + off) : (T)0, (T)0);
        }
    } else if (a.in_kind == IN_COMPLEX) {
        const V *__restrict__ xc = (const V *)a.x + ibase + ioff0;
        if (a.no_limit) {       // dense lines, whole blocks: no predicates
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = ld_pol<POL_IN>(xc + c * istep);
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = (ioff0 + c * istep) < lim ? ld_pol<POL_IN>(xc + c * istep) : zero;
        }
    } else {
        // IN_ROWS: the block's LPB rows are loaded cooperatively, coalesced along each row, then every
        // thread picks up the points of its own line.  Row addresses advance without divisions.
        const long long line0 = block * LPB;
        long long ro, rin;
        if (a.inner_shift >= 0) { ro = line0 >> a.inner_shift; rin = line0 & ((1LL << a.inner_shift) - 1); }
        else { ro = line0 / a.inner; rin = line0 - ro * a.inner; }
        for (int row = 0; row < LPB; ++row) {
            const bool row_ok = line0 + row < a.lines;
            const long long ro_in = a.ring_in ? ro % a.ring_in : ro;
            const V *__restrict__ src = (const V *)a.x + ro_in * a.gi.ostride + rin * a.gi.lstride;
            V *dst = sm_all + row * LINE;
            for (int pos = tid; pos < N; pos += THREADS)
                dst[Sc::pad(pos)] = row_ok ? ld_pol<POL_IN>(src + (long long)pos * a.gi.estride) : zero;
            if (++rin == a.inner) { rin = 0; ++ro; }
        }
        __syncthreads();
        const int pt = Sc::pad(t);
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = sm[Sc::pad_read(t, pt, c)];
        __syncthreads();
    }

    // ---------------------------------------------------------------- transform
    if constexpr (MODE == MODE_FILTER) {
        // forward packed transform, then per bin pair un-mix * spectrum * mix in shared memory, then the
        // inverse packed transform: the same registers and the same line buffer all the way
        Stage<T, LG_N, LG_E, true, 0>::run(v, sm, t, a, ls);
        __syncthreads();
        const int ptf = Sc::pad(t);
#pragma unroll
        for (int c = 0; c < E; ++c) sm[Sc::pad_read(t, ptf, c)] = v[c];
        __syncthreads();
        const V *__restrict__ twr = (const V *)a.tw_real;
        const V *__restrict__ bf = (const V *)a.filt;
#pragma unroll
        for (int c = 0; c < PAIRS; ++c) {
            const int k = t + c * TT;
            if (c == 0 && k == 0) {
                sm[Sc::pad(0)] = filter_dc<T>(sm[Sc::pad(0)], __ldg(bf), __ldg(bf + N));
                if (N >= 2) sm[Sc::pad(N / 2)] = filter_mid<T>(sm[Sc::pad(N / 2)], __ldg(bf + N / 2));
            } else if (k < N / 2) {
                V ra, rb;
                filter_pair<T>(sm[Sc::pad(k)], sm[Sc::pad(N - k)], __ldg(twr + k), __ldg(bf + k), __ldg(bf + (N - k)), ra, rb);
                sm[Sc::pad(k)] = ra;
                sm[Sc::pad(N - k)] = rb;
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = sm[Sc::pad_read(t, ptf, c)];
        __syncthreads();
        Stage<T, LG_N, LG_E, false, 0>::run(v, sm, t, a, ls);
    } else {
        Stage<T, LG_N, LG_E, FWD, 0>::run(v, sm, t, a, ls);
    }

    if (a.do_scale) {
        const T s = (T)a.scale;
#pragma unroll
        for (int c = 0; c < E; ++c) { v[c].x *= s; v[c].y *= s; }
    }

    // ---------------------------------------------------------------- store
    if constexpr (MODE == MODE_FAST) {
        if constexpr (TT >= 4) {
            V *__restrict__ op = (V *)a.out + line * N + t;
#pragma unroll
            for (int c = 0; c < E; ++c) st_stream(op + c * TT, v[c]);
        } else {
            __syncthreads();
            const int pt = Sc::pad(t);
#pragma unroll
            for (int c = 0; c < E; ++c) sm[Sc::pad_read(t, pt, c)] = v[c];
            __syncthreads();
            V *__restrict__ ob = (V *)a.out + block * (long long)(LPB * N);
            for (int e = tid; e < LPB * N; e += THREADS) st_stream(ob + e, sm_all[(e / N) * LINE + Sc::pad(e % N)]);
        }
    } else if constexpr (MODE == MODE_C2R_FAST) {
        V *__restrict__ op = (V *)a.out + line * N + t;
#pragma unroll
        for (int c = 0; c < E; ++c) st_stream(op + c * TT, v[c]);
    } else if constexpr (MODE == MODE_R2C_FAST) {
        // Z[k] is in registers; the partner Z[N-k] of each of the thread's first E/2 points comes from shared
        // memory (every thread only overwrites the slots it read last, so no barrier is needed before the writes)
        const int pt = Sc::pad(t);
#pragma unroll
        for (int c = 0; c < E; ++c) sm[Sc::pad_read(t, pt, c)] = v[c];
        line_sync(ls);
        V *__restrict__ oc = (V *)a.out + line * (N + 1);
        const V *__restrict__ twr = (const V *)a.tw_real;
#pragma unroll
        for (int c = 0; c < E / 2; ++c) {
            const int k = t + c * TT;
            if (c == 0 && t == 0) {
                const V z0 = v[0], zh = v[E / 2];          // thread 0 also owns point N/2
                st_stream(oc, mk<T>(z0.x + z0.y, (T)0));
                st_stream(oc + N, mk<T>(z0.x - z0.y, (T)0));
                st_stream(oc + N / 2, mk<T>(zh.x, -zh.y));
            } else {
                V xa, xb;
                real_pair<true, T>(v[c], sm[Sc::pad(N - k)], __ldg(twr + k), xa, xb);
                st_stream(oc + k, xa);
                st_stream(oc + (N - k), xb);
            }
        }
    } else if constexpr (IS_C2C) {
        if (a.four_shift) {   // four-step first pass: times W_M^(in * k1)
            const unsigned q = (unsigned)in;
#pragma unroll
            for (int c = 0; c < E; ++c)
                v[c] = cmul_tw<FWD>(v[c], four_step_twiddle<T>(a, q * (unsigned)(t + c * TT)));
        }
        if (active) {
            V *__restrict__ oc = (V *)a.out + obase + ooff0;
#pragma unroll
            for (int c = 0; c < E; ++c) st_pol<POL_OUT>(oc + c * ostep, v[c]);
        }
    } else if constexpr (MODE == MODE_C2R || MODE == MODE_FILTER) {
        // N complex = 2N reals; go is in REAL elements
        if (active) {
            T *__restrict__ orl = (T *)a.out + obase + ooff0;
            if (a.out_take) {
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    const int r0 = 2 * (t + c * TT);
                    if (r0 < a.out_take) orl[c * ostep] = v[c].x;
                    if (r0 + 1 < a.out_take) orl[c * ostep + a.go_pstride] = v[c].y;
                }
            } else if (a.packed_out) {
#pragma unroll
                for (int c = 0; c < E; ++c) __stcs((V *)(orl + c * ostep), v[c]);
            } else {
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    orl[c * ostep] = v[c].x;
                    orl[c * ostep + a.go_pstride] = v[c].y;
                }
            }
        }
    } else {  // MODE_R2C: Z -> shared memory, then each thread un-mixes its bin pairs
        __syncthreads();
        const int pt = Sc::pad(t);
#pragma unroll
        for (int c = 0; c < E; ++c) sm[Sc::pad_read(t, pt, c)] = v[c];
        __syncthreads();
        if (active) {
            V *__restrict__ oc = (V *)a.out + obase;
            const V *__restrict__ twr = (const V *)a.tw_real;
#pragma unroll
            for (int c = 0; c < PAIRS; ++c) {
                const int k = t + c * TT;
                if (c == 0 && k == 0) {
                    // DC and Nyquist come from Z[0] alone, imag exactly 0 (dsc_fft.h:220-225)
                    const V z0 = sm[Sc::pad(0)];
                    oc[0] = mk<T>(z0.x + z0.y, (T)0);
                    oc[(long long)N * a.go.estride] = mk<T>(z0.x - z0.y, (T)0);
                    if (N >= 2) {
                        const V zh = sm[Sc::pad(N / 2)];
                        oc[(long long)(N / 2) * a.go.estride] = mk<T>(zh.x, -zh.y);
                    }
                } else if (k < N / 2) {
                    V xa, xb;
                    real_pair<true, T>(sm[Sc::pad(k)], sm[Sc::pad(N - k)], __ldg(twr + k), xa, xb);
                    oc[(long long)k * a.go.estride] = xa;
                    oc[(long long)(N - k) * a.go.estride] = xb;
                }
            }
        }
    }
}

// Resident blocks per SM the register allocation must allow.  The float 4096-point bandwidth kernel compiles to
// 56 registers = 4 blocks of 256 threads; capping it at 51 (5 blocks, 40 warps/SM) costs no spills and was
// measured +1.2-1.7 % in same-box A/B runs (0.881 -> 0.892-0.896 of the copy peak); 6 blocks (42 registers)
// spill and lose 15 %.  Shorter float
// lines and all double lines spill a few registers under the same cap and gain nothing, so they keep 4 blocks.
// The dense packed-real kernels would otherwise hoist every load of a thread's 16 / 32 bin pairs and twiddles in
// front of the arithmetic (113-154 registers, one or two blocks per SM): they are capped at the register budget
// of their complex counterparts (no spills).  Same-box A/B: irfft 2^14 float +26 %, 2^13 double +22 %, rfft
// 2^11-2^12 float +5-6 %, 2^12 double +11 %, everything else within +-2 %.
#ifndef DSC_REAL_FAST_BLOCKS
#define DSC_REAL_FAST_BLOCKS 1
#endif
// The generic modes (predicated loads, runtime geometry) have the same habit -- 112-208 registers for the
// contiguous complex shape -- and get the same budget: 4 blocks of <= 256 threads with the standard register tile,
// 2 with the double-size one.
template <typename T> __host__ __device__ constexpr int lines_min_blocks(int mode, int lg_n, int threads = 256) {
    if (mode == MODE_FAST && sizeof(T) == 4 && lg_n == 12) return 5;
    if (!DSC_REAL_FAST_BLOCKS || threads > 256 || mode == MODE_FAST) return 1;
    const bool dense_real = mode == MODE_R2C_FAST || mode == MODE_C2R_FAST;
    // (the shortest lines and the generic double kernels would spill 36-124 bytes under the full cap: the former
    // are left alone, the latter get 3 blocks = 85 registers, <= 8 bytes of spills, +7..27 % in same-box A/B)
    if (!dense_real && lg_n < 9) return 1;
    if (!dense_real && sizeof(T) == 8) return lg_n <= 11 ? 3 : 1;
    return lg_n <= (sizeof(T) == 4 ? 12 : 11) ? 4 : 2;
}

template <typename T, int LG_N, int LG_E, int LPB, bool FWD, int MODE>
__global__ void __launch_bounds__(LPB * (1 << (LG_N - LG_E)), lines_min_blocks<T>(MODE, LG_N, LPB * (1 << (LG_N - LG_E))))
fft_lines(const FftArgs a) {
    DSC_DYN_SMEM(smem_raw);
    fft_lines_body<T, LG_N, LG_E, LPB, FWD, MODE>(a, (long long)blockIdx.x, smem_raw);
}

#if !defined(DSC_EMUL)
// Dense lines so long that ONE block fits an SM (float 2^14, double 2^13: 128 KiB of shared memory per line): with
// one-shot blocks nothing overlaps a line's load with the line before it, and the load of 128 KiB from DRAM at one SM's
// share of the bandwidth is a third of the line's time.  Persistent blocks (one per SM, lines blk, blk + grid, ...) ask L2
// for their NEXT line (one cp.async.bulk.prefetch of the whole line) before they start on the current one: the DRAM
// transfer then runs behind the transform and the line's loads are L2 hits.
template <typename T, int LG_N, int LG_E, int LPB, bool FWD, int MODE>
__global__ void __launch_bounds__(LPB * (1 << (LG_N - LG_E)), 1)
fft_lines_persist(const FftArgs a, const long long blocks) {
    DSC_DYN_SMEM(smem_raw);
    // bytes of a block's input lines: N complex points (2N packed reals), or the N + 1 bins of an inverse real transform
    constexpr size_t BYTES = (size_t)LPB * (size_t)((1 << LG_N) + (MODE == MODE_C2R_FAST ? 1 : 0)) * sizeof(cx<T>);
    for (long long blk = blockIdx.x; blk < blocks; blk += gridDim.x) {
        const long long nxt = blk + gridDim.x;
        if (threadIdx.x == 0 && nxt < blocks) {
            // whole 16-byte granules inside the line (rows of N + 1 complex64 bins start 8 bytes off every other line)
            const unsigned long long lo = ((unsigned long long)a.x + (size_t)nxt * BYTES + 15) & ~15ULL;
            const unsigned long long hi = ((unsigned long long)a.x + (size_t)(nxt + 1) * BYTES) & ~15ULL;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((unsigned)(hi - lo)) : "memory");
        }
        fft_lines_body<T, LG_N, LG_E, LPB, FWD, MODE>(a, blk, smem_raw);
        __syncthreads();            // the next line's first exchange reuses the buffers
    }
}
#endif

// ------------------------------------------------------------------------------------------
// Register-direct tiles for the two passes of the four-step transform.
//
// Both passes are bound by the SM's load/store pipe (shared-memory wavefronts + L1 requests), not by HBM:
// every trip of an element through shared memory costs as much pipe time as its trip to HBM.  So a tile here
// touches shared memory ONLY for the exchanges between butterfly stages (one exchange for float lines up to
// 1024 points: radix-32 x radix-32), and both global sides go straight between registers and memory:
//
//   first pass   L adjacent columns per block, thread = (line l = tid % L, position j = tid / L): a warp's
//                load of element j + c*TT covers L*sizeof(V) contiguous bytes per position (64 B+), the same
//                for its store of k1 = j + c*TT into the work row [k1][n2];
//   second pass  reads its L contiguous work rows one line per (part of a) warp (lane = position), and at
//                the first exchange the threads come back as (l = tid % L, j = tid / L), so the store of
//                k2 = j + c*TT at k1 + n1*k2 again covers L adjacent k1 per position.
//
// The inter-pass twiddle W_n^(q k1), k1 = j + c*TT, is W^(q j) (one two-table lookup per thread) times
// W^(q TT c), which depends on (line, c) only: L*E values per tile, built by the block into shared memory
// once and read back as broadcasts.
// Shared memory of a block: the line buffers (of whichever pass needs more), then -- at TABLE_AT elements -- two
// copies of the first pass's (line, c) twiddle table that alternate from tile to tile (a block starts its
// next tile while slower warps still finish the previous one, whichever pass that was).
template <typename T, int LG_N, int LG_E, int L> struct PassTile {
    using Sc = Sched<LG_N, LG_E>;
    // A bank phase is 128 bytes: PHASE lanes of one access.  Line l starts at l * LINE + rot(l), LINE a multiple
    // of PHASE and rot = the bit reversal of l mod PHASE.  Whatever power-of-two group of lines shares a phase
    // -- 16 adjacent lines at one position (adjacent lanes on adjacent lines), 8 lines x 2 positions, or the
    // 2 / 4 adjacent short lines a half-warp spans while a second-pass tile loads -- the rotations of the
    // group are an arithmetic progression that interleaves exactly with the positions (which advance by one
    // element, or by E + 1 = 1 mod PHASE), so every exchange of the tile is conflict-free.
    static constexpr int PHASE = 128 / (int)sizeof(cx<T>);
    static constexpr int LINE = ((Sc::N + (Sc::N >> LG_E)) + 2 * PHASE - 2) / PHASE * PHASE;   // padded line + largest rotation
    static constexpr int LINES = L * LINE;                  // elements
    static constexpr int TABLE = L * Sc::E;                 // elements per twiddle table copy
    static DSC_DEV int line_base(int l) {
        int r = 0;
#pragma unroll
        for (int b = 1, m = PHASE >> 1; b < PHASE; b <<= 1, m >>= 1) if (l & b) r |= m;
        return l * LINE + r;
    }
};

// LG_M: log2 of the OTHER factor (n = 2^(LG_N + LG_M)); with it the element strides of the dense case are
// compile-time constants and the 2E accesses of a thread use one base pointer plus immediate offsets.
template <typename T, int LG_N, int LG_M, int LG_E, int L, int TABLE_AT, bool FWD, typename Hook>
DSC_DEV void pass_first_tile(const FftArgs &a, const long long tile, const int parity, unsigned char *smem_raw,
                             Hook &&before_scatter) {
    using Sc = Sched<LG_N, LG_E>;
    using V = cx<T>;
    using PT = PassTile<T, LG_N, LG_E, L>;
    constexpr int E = Sc::E, TT = Sc::TT, THREADS = L * TT;
    static_assert(Sc::STAGES >= 2, "the inter-pass table is published by the first exchange barrier");
    constexpr long long STEP = (long long)TT << LG_M;     // elements between a thread's consecutive points
    V *sm_all = (V *)smem_raw;
    const int tid = threadIdx.x, l = tid % L, j = tid / L;
    V *sm = sm_all + PT::line_base(l);
    static_assert(TABLE_AT >= PT::LINES, "twiddle tables overlap the line buffers");
    V *tw_c = sm_all + TABLE_AT + parity * PT::TABLE;     // [c][l]
    const LineSync ls{SYNC_BLOCK, 0, THREADS};

    const long long line0 = tile * L;
    const long long row = line0 >> LG_M;
    const unsigned q0 = (unsigned)(line0 & ((1LL << LG_M) - 1)), q = q0 + (unsigned)l;
    const long long row_in = a.ring_in ? row % a.ring_in : row;
    const long long row_out = a.ring_out ? row % a.ring_out : row;

    V v[E];
    if (a.in_kind == IN_COMPLEX && a.no_limit) {
        // dense complex rows: element (m, q) of the row at m * 2^LG_M + q
        const V *__restrict__ src = (const V *)a.x + row_in * a.gi.ostride + (((long long)j << LG_M) + q);
        if (a.seg_shift == 0) {
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = ld_stream(src + c * STEP);
        } else {
            // segmented rows (the receive buffer of the multi-GPU exchange, [peer][line][part]): a segment is a
            // whole number of this thread's steps long, so its index depends on c alone
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const long long seg = (c * STEP) >> a.seg_shift;
                v[c] = ld_stream(src + c * STEP + seg * a.seg_extra + (seg == a.seg_self ? a.seg_self_delta : 0));
            }
        }
    } else {
        const long long sbase = row_in * a.gi.ostride + (long long)q * a.gi.lstride + (long long)j * a.gi.estride;
        const long long istep = (long long)TT * a.gi.estride;
        const long long tlim = a.in_limit - (long long)q * a.gi.lstride - (long long)j * a.gi.estride;
        if (a.in_kind == IN_COMPLEX) {
            const V *__restrict__ src = (const V *)a.x + sbase;
            const V zero = mk<T>((T)0, (T)0);
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = c * istep < tlim ? __ldcs(src + c * istep) : zero;
        } else if (a.in_kind == IN_REAL) {
            const T *__restrict__ src = (const T *)a.x + sbase;
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = mk<T>(c * istep < tlim ? __ldcs(src + c * istep) : (T)0, (T)0);
        } else {  // IN_PAIRS
            const T *__restrict__ src = (const T *)a.x + sbase;
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const T re = c * istep < tlim ? __ldcs(src + c * istep) : (T)0;
                const T im = c * istep + a.gi_pstride < tlim ? __ldcs(src + c * istep + a.gi_pstride) : (T)0;
                v[c] = mk<T>(re, im);
            }
        }
    }
    // the table lookups travel while the payload does
    V w0 = mk<T>((T)1, (T)0);
    if (a.four_shift) {
        for (int i = tid; i < L * E; i += THREADS) {
            const unsigned ll = (unsigned)(i % L), c = (unsigned)(i / L);
            tw_c[i] = four_step_twiddle<T>(a, (q0 + ll) * (unsigned)TT * c);
        }
        w0 = four_step_twiddle<T>(a, q * (unsigned)j);
    }

    Stage<T, LG_N, LG_E, FWD, 0>::run(v, sm, j, sm, j, a, ls, before_scatter);

    if (a.four_shift) {
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = cmul_tw<FWD>(v[c], c == 0 ? w0 : cmul(w0, tw_c[c * L + l]));
    }
    // work row [k1][n2]: stays in L2 for the second pass
    V *__restrict__ op = (V *)a.out + row_out * a.go.ostride + (((long long)j << LG_M) + q);
#pragma unroll
    for (int c = 0; c < E; ++c) op[c * STEP] = v[c];
}

template <typename T, int LG_N, int LG_M, int LG_E, int L, bool FWD, typename Hook>
DSC_DEV void pass_second_tile(const FftArgs &b, const long long tile, unsigned char *smem_raw, Hook &&before_scatter) {
    using Sc = Sched<LG_N, LG_E>;
    using V = cx<T>;
    constexpr int E = Sc::E, TT = Sc::TT, THREADS = L * TT;
    static_assert(Sc::STAGES >= 2, "the threads change lines at the first exchange");
    using PT = PassTile<T, LG_N, LG_E, L>;
    constexpr long long STEP = (long long)TT << LG_M;
    V *sm_all = (V *)smem_raw;
    const int tid = threadIdx.x;
    const int l1 = tid / TT, j1 = tid % TT;           // while loading: one line per TT consecutive threads
    const int l2 = tid % L, j2 = tid / L;             // after the first exchange: adjacent lanes, adjacent lines
    const LineSync ls{SYNC_BLOCK, 0, THREADS};

    const long long line0 = tile * L;
    const long long row = line0 >> LG_M;
    const long long k0 = line0 & ((1LL << LG_M) - 1);
    const long long row_in = b.ring_in ? row % b.ring_in : row;
    const long long row_out = b.ring_out ? row % b.ring_out : row;

    V v[E];
    // work row [k1][n2], line k1 contiguous; written a moment ago by other SMs: L2 loads
    const V *__restrict__ src = (const V *)b.x + row_in * b.gi.ostride + (((k0 + l1) << LG_N) + j1);
#pragma unroll
    for (int c = 0; c < E; ++c) v[c] = __ldcg(src + c * TT);

    Stage<T, LG_N, LG_E, FWD, 0>::run(v, sm_all + PT::line_base(l1), j1, sm_all + PT::line_base(l2), j2, b, ls, before_scatter);

    if (b.do_scale) {
        const T s = (T)b.scale;
#pragma unroll
        for (int c = 0; c < E; ++c) { v[c].x *= s; v[c].y *= s; }
    }
    // X[k1 + n1 k2]
    V *__restrict__ op = (V *)b.out + row_out * b.go.ostride + (((long long)j2 << LG_M) + (k0 + l2));
    if (b.keep_out) {
#pragma unroll
        for (int c = 0; c < E; ++c) op[c * STEP] = v[c];
    } else {
#pragma unroll
        for (int c = 0; c < E; ++c) __stcs(op + c * STEP, v[c]);
    }
}

// ------------------------------------------------------------------------------------------
// Four-step transform n = n1*n2 as ONE launch.
//
// Per top-level line ("row") there are TA first-pass tiles (length-n1 transforms over stride-n2
// data, times W_n^(n2 k1), into a ring of work rows) and TB second-pass tiles (length-n2 transforms
// of contiguous work rows, stored with stride n1).  Blocks take tickets; ticket order is row-major
// with a row's first-pass tiles `lag` rows ahead of its second-pass tiles, so a tile only ever waits
// for tiles with SMALLER tickets, which are done or being worked on -- no deadlock whatever the
// dispatch order.  A second-pass tile needs all TA first-pass tiles of its row; a first-pass tile that
// reuses a ring slot needs the second pass of the row that last used it.  The intermediate of a
// row is therefore consumed a few microseconds after it is produced and never leaves L2 (.cg loads
// bypass the non-coherent L1; the payload itself streams), so HBM sees one read and one write per
// element although there are two passes.

// points per thread of a four-step pass (must agree with lg_e_for in fft_dispatch.cuh for these lengths)
template <typename T> __host__ __device__ constexpr int pass_lg_e(int lg_n, int lg_other) {
    // register-direct tiles: radix-32 (float) / radix-16 (double) register tiles, so that float lines of up to
    // 1024 points need ONE shared-memory exchange
    (void)lg_other;
    const int e = sizeof(T) == 4 ? 5 : 4;
    return lg_n < e ? lg_n : e;
}

struct FourStepSync {
    unsigned *ticket;    // one counter
    unsigned *a_done;    // per row: first-pass blocks finished
    unsigned *b_done;    // per row: second-pass blocks finished
    int tiles_a, tiles_b;
    int ring;            // work rows (0 = one work row per top-level row, no reuse)
    int rows;
    int lag;             // ticket order: the B blocks of row r come after the A blocks of row r + lag, so
                         // that in steady state a B block finds its row already complete and never spins
};

// ticket -> (role, row, tile): A(0..lag-1), then groups { A(i + lag), B(i) }, then the last B rows
DSC_DEV void decode_ticket(const FourStepSync &s, unsigned ticket, bool &role_a, unsigned &row, unsigned &r) {
    const unsigned ta = (unsigned)s.tiles_a, tb = (unsigned)s.tiles_b, lag = (unsigned)s.lag;
    if (ticket < lag * ta) { role_a = true; row = ticket / ta; r = ticket % ta; return; }
    ticket -= lag * ta;
    const unsigned full = (unsigned)s.rows - lag;
    if (ticket < full * (ta + tb)) {
        const unsigned i = ticket / (ta + tb), w = ticket % (ta + tb);
        if (w < ta) { role_a = true; row = i + lag; r = w; } else { role_a = false; row = i; r = w - ta; }
    } else {
        ticket -= full * (ta + tb);
        role_a = false; row = full + ticket / tb; r = ticket % tb;
    }
}

// Acquire load at GPU scope (no full memory barrier behind it, unlike volatile load + __threadfence()).
#if defined(DSC_EMUL)
DSC_DEV unsigned ld_acquire(const unsigned *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
#else
DSC_DEV unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif
DSC_DEV void spin_until(const unsigned *counter, const unsigned target) {
    while (ld_acquire(counter) < target) __nanosleep(64);
}

// shared-memory layout of the fused launch (elements): [ line buffers, max over the passes | 2 twiddle tables ]
// LGE > 0 overrides the points per thread (log2) of both passes (the 16-point float variant: half the registers, twice the
// resident warps)
template <typename T> __host__ __device__ constexpr int fused_lg_e(int lge, int lg_n, int lg_other) {
    return lge > 0 ? (lg_n < lge ? lg_n : lge) : pass_lg_e<T>(lg_n, lg_other);
}
template <typename T, int LG_N1, int LG_N2, int THREADS, int LGE = 0> __host__ __device__ constexpr int fused_table_at() {
    constexpr int LG_E1 = fused_lg_e<T>(LGE, LG_N1, LG_N2), LG_E2 = fused_lg_e<T>(LGE, LG_N2, LG_N1);
    constexpr int A = PassTile<T, LG_N1, LG_E1, (THREADS >> (LG_N1 - LG_E1))>::LINES;
    constexpr int B = PassTile<T, LG_N2, LG_E2, (THREADS >> (LG_N2 - LG_E2))>::LINES;
    return A > B ? A : B;
}
template <typename T, int LG_N1, int LG_N2, int THREADS, int LGE = 0> __host__ __device__ constexpr int fused_smem_bytes() {
    constexpr int LG_E1 = fused_lg_e<T>(LGE, LG_N1, LG_N2);
    return (fused_table_at<T, LG_N1, LG_N2, THREADS, LGE>() +
            2 * PassTile<T, LG_N1, LG_E1, (THREADS >> (LG_N1 - LG_E1))>::TABLE) * (int)sizeof(cx<T>);
}

// Persistent blocks (the grid is what fits on the GPU at once): a block keeps taking tickets until none is
// left.  Per tile there are two block barriers, back to back around the scatter of the exchange: W ("every
// warp is done with the line buffers of the previous tile") and E (the exchange itself).  There is no
// barrier at the end of a tile: a warp that has stored its outputs goes straight on to load the next tile.
// Everything per tile that is not payload is done by thread 0 off the critical path:
//   * tickets are requested two tiles ahead; the ticket and readiness of tile i+1 are published in shared
//     memory before W of tile i;
//   * the dependency flag of tile i+1 is read at the start of tile i (with the ticket lag it is almost
//     always already satisfied, and then tile i+1 starts without a poll);
//   * tile i is released (release-add on its row counter) after W of tile i+1 -- by then every warp has
//     issued its stores of tile i long ago.
// A flag observed satisfied is followed by a block barrier and then by L2 (.cg) loads of the dependent data,
// which the producer made visible at L2 before it incremented the flag.
// The ticket loop of a persistent block.  first(tile, parity, before_scatter) / second(tile, parity,
// before_scatter) transform one tile of the respective pass and must call before_scatter() exactly once, by
// every thread, after their last read of the previous tile's shared memory and before their first write.
template <typename First, typename Second>
DSC_DEV void run_tickets(const FourStepSync &s, First &&first, Second &&second) {
    __shared__ unsigned ctl_s[2][2];                   // [parity]{ticket, dependency already satisfied}
    const unsigned total = (unsigned)s.rows * (unsigned)(s.tiles_a + s.tiles_b);
    unsigned pending = 0;                              // thread 0: ticket of the tile after the next one
    if (threadIdx.x == 0) {
        ctl_s[0][0] = atomicAdd(s.ticket, 1u);
        ctl_s[0][1] = 0;
        pending = atomicAdd(s.ticket, 1u);
    }
    __syncthreads();
    unsigned *prev_done = nullptr;                     // row counter of the previous tile, not yet released
    for (int par = 0;; par ^= 1) {
        const unsigned ticket = ctl_s[par][0], ready = ctl_s[par][1];
        if (ticket >= total) break;
        unsigned next = 0, flag = 0, target = 0;
        const unsigned *next_flag = nullptr;     // thread 0: the counter `flag` was read from
        if (threadIdx.x == 0) {
            next = pending;
            pending = atomicAdd(s.ticket, 1u);
            if (next < total) {
                unsigned nrow, nr;
                bool na;
                decode_ticket(s, next, na, nrow, nr);
                if (!na) { next_flag = s.a_done + nrow; target = (unsigned)s.tiles_a; }
                else if (s.ring && nrow >= (unsigned)s.ring) { next_flag = s.b_done + (nrow - s.ring); target = (unsigned)s.tiles_b; }
                if (next_flag != nullptr) flag = *(const volatile unsigned *)next_flag;
            }
        }
        unsigned row, r;
        bool role_a;
        decode_ticket(s, ticket, role_a, row, r);
        if (!ready) {
            // about to wait for other tiles: release our own previous tile first (the awaited one may be it)
            __syncthreads();
            if (threadIdx.x == 0) {
                if (prev_done != nullptr) dsc_signal_release(prev_done);
                if (!role_a) spin_until(s.a_done + row, (unsigned)s.tiles_a);
                else if (s.ring && row >= (unsigned)s.ring) spin_until(s.b_done + (row - s.ring), (unsigned)s.tiles_b);
            }
            prev_done = nullptr;
            __syncthreads();
        }
        auto before_scatter = [&]() {
            if (threadIdx.x == 0) {
                // not satisfied when this tile started?  look once more: a short wait here is far cheaper
                // than the poll-and-barrier path at the start of the next tile
                if (flag < target) flag = *(const volatile unsigned *)next_flag;
                ctl_s[par ^ 1][0] = next; ctl_s[par ^ 1][1] = flag >= target;
            }
            __syncthreads();                                                          // W
            if (threadIdx.x == 0 && prev_done != nullptr) dsc_signal_release(prev_done);
        };
        if (role_a) first(row, r, par, before_scatter);
        else second(row, r, par, before_scatter);
        prev_done = (role_a ? s.a_done : s.b_done) + row;
    }
    __syncthreads();
    if (threadIdx.x == 0 && prev_done != nullptr) dsc_signal_release(prev_done);
}

template <typename T, int LG_N1, int LG_N2, int THREADS, bool FWD, int LGE = 0>
__global__ void __launch_bounds__(THREADS, (LGE > 0 && LGE < (sizeof(T) == 4 ? 5 : 4) ? 1024 : 512) / THREADS)
four_step_fused(const FftArgs a, const FftArgs b, const FourStepSync s) {
    constexpr int LG_E1 = fused_lg_e<T>(LGE, LG_N1, LG_N2), LG_E2 = fused_lg_e<T>(LGE, LG_N2, LG_N1);
    constexpr int LPB_A = THREADS >> (LG_N1 - LG_E1), LPB_B = THREADS >> (LG_N2 - LG_E2);
    static_assert(LPB_A >= 1 && LPB_B >= 1, "block too small for one line");
    constexpr int TABLE_AT = fused_table_at<T, LG_N1, LG_N2, THREADS, LGE>();
    DSC_DYN_SMEM(smem_raw);
    run_tickets(s,
        [&](unsigned row, unsigned r, int par, auto &before_scatter) {
            pass_first_tile<T, LG_N1, LG_N2, LG_E1, LPB_A, TABLE_AT, FWD>(a, (long long)row * s.tiles_a + r, par, smem_raw, before_scatter);
        },
        [&](unsigned row, unsigned r, int, auto &before_scatter) {
            pass_second_tile<T, LG_N2, LG_N1, LG_E2, LPB_B, FWD>(b, (long long)row * s.tiles_b + r, smem_raw, before_scatter);
        });
}

// ------------------------------------------------------------------------------------------
// Two-pass transforms along a NON-last axis: both passes are column passes.
//
// The tensor is (outer, n, inner), n = n1*n2, i = i1*n2 + i2.  The inner extent is cut into chunks of Ic
// columns; a "row" of the launch is one (outer, chunk) pair, n x Ic points, small enough for its intermediate
// to stay in L2.  A tile is L adjacent columns (adjacent lanes on adjacent columns: L*sizeof(V) contiguous
// bytes per position on both global sides) at a fixed i2 (first pass) or k1 (second pass):
//   first pass   over i1 at stride n2*inner, times W_n^(i2 k1) -- i2 is one number per tile, so the twiddle
//                is W^(i2 j) per thread times a table of E powers per tile -- into work[k1][i2][Ic];
//   second pass  over i2 at stride Ic of the work row, out at (k1 + n1 k2)*inner.
struct ColumnsGeom {
    long long x_ostride, out_ostride;     // elements between outer slabs of the source / the destination
    long long inner;                      // elements between consecutive points of a column
    int lg_ic, chunks;                    // chunk width (log2) and chunks per outer slab
    int x_n;                              // valid points per source column (pad / crop)
    int x_real;                           // source holds T, not cx<T>
    // optional output twiddle of the second pass: out[k][col] *= W_M^((col_offset + col) k), M = post_mask + 1, looked
    // up through the second-pass FftArgs' tw_lo / tw_hi split tables.  This is the inter-pass twiddle of an OUTER
    // four-step whose columns are spread over several GPUs (the multi-GPU transform): the launch then delivers
    // the twiddled, k-major matrix that goes straight into the all-to-all.
    int post_twiddle;
    unsigned post_mask;
    long long col_offset;
    // optional scattered output of the second pass (the fused compute + collective of the multi-GPU transform): the
    // rows k of the output are cut into blocks of 2^peer_shift rows, and block q is stored at peer_out[q] (row pitch
    // `inner`, like the dense output) -- a pointer into GPU q's receive buffer, mapped into this process (NVLink
    // peer memory).  n_peers == 0: one dense output.
    int n_peers, peer_shift;
    void *peer_out[8];
};

template <typename T, int LG_N, int LG_M, int LG_E, int L, int TABLE_AT, bool FWD, bool FIRST, typename Hook>
DSC_DEV void column_tile(const FftArgs &a, const ColumnsGeom &g, const unsigned row, const unsigned tile, const int parity,
                         unsigned char *smem_raw, Hook &&before_scatter) {
    using Sc = Sched<LG_N, LG_E>;
    using V = cx<T>;
    using PT = PassTile<T, LG_N, LG_E, L>;
    constexpr int E = Sc::E, TT = Sc::TT, THREADS = L * TT;
    static_assert(Sc::STAGES >= 2, "the twiddle table is published by the first exchange barrier");
    V *sm_all = (V *)smem_raw;
    const int tid = threadIdx.x, l = tid % L, j = tid / L;
    V *sm = sm_all + PT::line_base(l);
    V *tw_c = sm_all + TABLE_AT + parity * E;
    const LineSync ls{SYNC_BLOCK, 0, THREADS};

    const unsigned per = (1u << g.lg_ic) / L;                 // tiles per fixed index (i2 or k1)
    const unsigned fixed = tile / per, col0 = (tile % per) * L;
    const unsigned o = row / (unsigned)g.chunks, ch = row % (unsigned)g.chunks;
    const long long ring_row_in = a.ring_in ? row % a.ring_in : row;
    const long long ring_row_out = a.ring_out ? row % a.ring_out : row;
    const long long row_elems = (long long)1 << (LG_N + LG_M + g.lg_ic);

    V v[E];
    if (FIRST) {
        // source column (i2 = fixed): point i1 = j + c TT at i = i1 n2 + i2
        const long long base = (long long)o * g.x_ostride + ((long long)ch << g.lg_ic) + col0 + l;
        const V zero = mk<T>((T)0, (T)0);
#pragma unroll
        for (int c = 0; c < E; ++c) {
            const long long i = ((long long)(j + c * TT) << LG_M) + fixed;
            const long long off = base + i * g.inner;
            if (i >= g.x_n) v[c] = zero;
            else if (g.x_real) v[c] = mk<T>(__ldcs((const T *)a.x + off), (T)0);
            else v[c] = ld_stream((const V *)a.x + off);
        }
        if (tid < E) tw_c[tid] = four_step_twiddle<T>(a, fixed * (unsigned)TT * (unsigned)tid);
    } else {
        // work row [k1 = fixed][i2][Ic]: point i2 = j + c TT
        const V *__restrict__ src = (const V *)a.x + ring_row_in * row_elems +
                                    ((((long long)fixed << LG_N) + j) << g.lg_ic) + col0 + l;
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = __ldcg(src + ((long long)(c * TT) << g.lg_ic));
    }

    Stage<T, LG_N, LG_E, FWD, 0>::run(v, sm, j, sm, j, a, ls, before_scatter);

    if (FIRST) {
        const V w0 = four_step_twiddle<T>(a, fixed * (unsigned)j);
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = cmul_tw<FWD>(v[c], c == 0 ? w0 : cmul(w0, tw_c[c]));
        // work[k1 = j + c TT][i2 = fixed][Ic]
        V *__restrict__ op = (V *)a.out + ring_row_out * row_elems + ((((long long)j << LG_M) + fixed) << g.lg_ic) + col0 + l;
#pragma unroll
        for (int c = 0; c < E; ++c) op[(long long)(c * TT) << (LG_M + g.lg_ic)] = v[c];
    } else {
        if (a.do_scale) {
            const T sc = (T)a.scale;
#pragma unroll
            for (int c = 0; c < E; ++c) { v[c].x *= sc; v[c].y *= sc; }
        }
        if (g.post_twiddle) {
            // exponent (col) * k, k = fixed + (j + c TT) 2^LG_M: a base per thread times the c-th power of one ratio,
            // built from the ratio's 1st, 2nd, 4th, ... powers (each one table lookup, <= log2(E) - 1 products deep)
            const unsigned col = (unsigned)(g.col_offset + ((long long)ch << g.lg_ic) + col0 + l);
            const unsigned k0 = (unsigned)fixed + ((unsigned)j << LG_M);
            // w_c = base * ratio^c with c = 4 a + b:  (base * ratio^(4a)) * ratio^b -- E/4 + 3 values in registers, each
            // at most three products deep from table lookups of base, ratio^1, ^2, ^4, ^8, ^16
            constexpr int GROUPS = E / 4;
            static_assert(E >= 4 && GROUPS <= 8, "column tiles hold 16 or 32 points per thread");
            auto ratio_pow = [&](const int e) { return four_step_twiddle<T>(a, (col * ((unsigned)(e * TT) << LG_M)) & g.post_mask); };
            V hi4[GROUPS], lo3[4];
            hi4[0] = four_step_twiddle<T>(a, (col * k0) & g.post_mask);
            lo3[1] = ratio_pow(1);
            lo3[2] = ratio_pow(2);
            lo3[3] = cmul(lo3[1], lo3[2]);
            const V r4 = ratio_pow(4);
            hi4[1] = cmul(hi4[0], r4);
            if constexpr (GROUPS > 2) {
                const V r8 = ratio_pow(8);
                hi4[2] = cmul(hi4[0], r8);
                hi4[3] = cmul(hi4[2], r4);
                if constexpr (GROUPS > 4) {
                    const V r16 = ratio_pow(16);
                    hi4[4] = cmul(hi4[0], r16);
                    hi4[5] = cmul(hi4[4], r4);
                    hi4[6] = cmul(hi4[4], r8);
                    hi4[7] = cmul(hi4[6], r4);
                }
            }
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = cmul_tw<FWD>(v[c], (c & 3) == 0 ? hi4[c >> 2] : cmul(hi4[c >> 2], lo3[c & 3]));
        }
        // X[k1 + n1 k2], k2 = j + c TT
        if (g.n_peers) {
            // straight into the receive buffers of the GPUs that own the rows: the exchange happens while the transform
            // runs, one row block per peer (a warp writes L * sizeof(V) contiguous bytes per row, like the dense store)
            const long long coff = ((long long)ch << g.lg_ic) + col0 + l;
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const long long k = (long long)fixed + ((long long)(j + c * TT) << LG_M);
                V *__restrict__ pp = (V *)g.peer_out[k >> g.peer_shift];
                pp[(k & ((1LL << g.peer_shift) - 1)) * g.inner + coff] = v[c];
            }
        } else {
            V *__restrict__ op = (V *)a.out + (long long)o * g.out_ostride + ((long long)ch << g.lg_ic) + col0 + l;
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const long long k = (long long)fixed + ((long long)(j + c * TT) << LG_M);
                __stcs(op + k * g.inner, v[c]);
            }
        }
    }
}

template <typename T, int LG_N1, int LG_N2, int THREADS, bool FWD, int LGE = 0>
__global__ void __launch_bounds__(THREADS, (LGE > 0 && LGE < (sizeof(T) == 4 ? 5 : 4) ? 1024 : 512) / THREADS)
four_step_columns(const FftArgs a, const FftArgs b, const FourStepSync s, const ColumnsGeom g) {
    constexpr int LG_E1 = fused_lg_e<T>(LGE, LG_N1, LG_N2), LG_E2 = fused_lg_e<T>(LGE, LG_N2, LG_N1);
    constexpr int L_A = THREADS >> (LG_N1 - LG_E1), L_B = THREADS >> (LG_N2 - LG_E2);
    constexpr int TABLE_AT = fused_table_at<T, LG_N1, LG_N2, THREADS, LGE>();
    DSC_DYN_SMEM(smem_raw);
    run_tickets(s,
        [&](unsigned row, unsigned r, int par, auto &before_scatter) {
            column_tile<T, LG_N1, LG_N2, LG_E1, L_A, TABLE_AT, FWD, true>(a, g, row, r, par, smem_raw, before_scatter);
        },
        [&](unsigned row, unsigned r, int par, auto &before_scatter) {
            column_tile<T, LG_N2, LG_N1, LG_E2, L_B, TABLE_AT, FWD, false>(b, g, row, r, par, smem_raw, before_scatter);
        });
}

// Large packed-real transforms (order N beyond one shared-memory pass): the same bin-pair
// step as a standalone elementwise kernel.
//   FWD : rows of N+1 bins hold Z[0..N) (slot N unused); rewritten in place to X[0..N].
//   !FWD: bins x (row stride x_ostride, `take` valid bins) -> packed z rows of N in `z`.
template <bool FWD, typename T>
__global__ void real_mix_rows(const cx<T> *__restrict__ x, cx<T> *__restrict__ z,
                              long long rows, int n, long long x_ostride, int take,
                              const cx<T> *__restrict__ tw_lo, const cx<T> *__restrict__ tw_hi,
                              int shift, int mask) {
    using V = cx<T>;
    const int half = n / 2;                 // work items per row: k = 0 .. half-1 (k = 0 also does n/2, n)
    int lg_half = 0;
    while ((1 << lg_half) < half) ++lg_half;
    const long long total = rows * (long long)half;
    const V zero = mk<T>((T)0, (T)0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i >> lg_half;             // n is a power of two: no 64-bit division per item
        const int k = (int)(i & (half - 1));
        if (FWD) {
            V *zr = z + row * (long long)(n + 1);
            if (k == 0) {
                const V z0 = zr[0];
                zr[0] = mk<T>(z0.x + z0.y, (T)0);
                zr[n] = mk<T>(z0.x - z0.y, (T)0);
                const V zh = zr[half];
                zr[half] = mk<T>(zh.x, -zh.y);
            } else {
                // W_2N^k through the same two-table split as the four-step twiddle
                const V w = cmul(__ldg(tw_lo + (k & mask)), __ldg(tw_hi + (k >> shift)));
                V xa, xb;
                real_pair<true, T>(zr[k], zr[n - k], w, xa, xb);
                zr[k] = xa;
                zr[n - k] = xb;
            }
        } else {
            const V *xr = x + row * x_ostride;
            V *zr = z + row * (long long)n;
            auto bin = [&](int j) -> V { return j < take ? xr[j] : zero; };
            if (k == 0) {
                const V x0 = bin(0), xn = bin(n), xh = bin(half);
                zr[0] = mk<T>((T)0.5 * (x0.x + xn.x), (T)0.5 * (x0.x - xn.x));
                zr[half] = mk<T>(xh.x, -xh.y);
            } else {
                const V w = cmul(__ldg(tw_lo + (k & mask)), __ldg(tw_hi + (k >> shift)));
                V za, zb;
                real_pair<false, T>(bin(k), bin(n - k), w, za, zb);
                zr[k] = za;
                zr[n - k] = zb;
            }
        }
    }
}

// Filter pipeline for orders beyond one shared-memory pass: between the forward and the inverse four-step
// transforms, ONE elementwise kernel turns packed Z rows into packed z' rows in place (un-mix, spectrum
// product, mix per bin pair) -- instead of un-mix, product and mix as three sweeps.
template <typename T>
__global__ void filter_pairs_rows(cx<T> *__restrict__ z, const cx<T> *__restrict__ spectrum, long long rows, int n,
                                  const cx<T> *__restrict__ tw_lo, const cx<T> *__restrict__ tw_hi, int shift, int mask) {
    using V = cx<T>;
    const int half = n / 2;
    int lg_half = 0;
    while ((1 << lg_half) < half) ++lg_half;
    const long long total = rows * (long long)half;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i >> lg_half;
        const int k = (int)(i & (half - 1));
        V *zr = z + row * (long long)n;
        if (k == 0) {
            zr[0] = filter_dc<T>(zr[0], __ldg(spectrum), __ldg(spectrum + n));
            zr[half] = filter_mid<T>(zr[half], __ldg(spectrum + half));
        } else {
            const V w = cmul(__ldg(tw_lo + (k & mask)), __ldg(tw_hi + (k >> shift)));
            V ra, rb;
            filter_pair<T>(zr[k], zr[n - k], w, __ldg(spectrum + k), __ldg(spectrum + (n - k)), ra, rb);
            zr[k] = ra;
            zr[n - k] = rb;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Twiddle tables, computed on the device in double and rounded once to T
// (the reference evaluates cos/sin in T, dsc_fft.h:42-49; see plan.cu).

// stage table: out[(m-1)*NS + k] = exp(-2 pi i k m / (NS*R)),  m in [1,R), k in [0,NS)
template <typename T>
__global__ void fill_stage_twiddles(cx<T> *out, int ns, int r) {
    const int total = (r - 1) * ns;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int m = i / ns + 1, k = i % ns;
        // angle = -2 pi (k m) / (ns r): reduce the integer product first, exactly
        const long long num = ((long long)k * m) % ((long long)ns * r);
        double s, c;
        sincospi(-2.0 * (double)num / (double)((long long)ns * r), &s, &c);
        out[i] = mk<T>((T)c, (T)s);
    }
}

// out[k] = exp(-2 pi i (k * mult) / denom), k in [0,count)
template <typename T>
__global__ void fill_power_twiddles(cx<T> *out, long long count, long long mult, long long denom) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x) {
        const long long num = (i * mult) % denom;
        double s, c;
        sincospi(-2.0 * (double)num / (double)denom, &s, &c);
        out[i] = mk<T>((T)c, (T)s);
    }
}

// ------------------------------------------------------------------------------------------
// Distributed four-step (one transform sharded over P GPUs): between the two local passes every rank
// holds A[r][c] (r = its block of n2, c = k1) and must hand peer p the columns c in block p.  This kernel
// multiplies by the inter-pass twiddle W_M^((r0 + r) c) and writes the TRANSPOSE out[c][r], so the slab
// for peer p (rows p*cols/P .. of `out`) is contiguous and can go straight into the all-to-all.
template <typename T, bool FWD>
__global__ void transpose_twiddle(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, int rows, int cols,
                                  long long r0, const cx<T> *__restrict__ tw_lo, const cx<T> *__restrict__ tw_hi,
                                  int shift, int mask) {
    using V = cx<T>;
    __shared__ V tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
    const int tiles_c = (cols + 31) / 32;
    const long long tile_id = blockIdx.x;
    const int c0 = (int)(tile_id % tiles_c) * 32, rr0 = (int)(tile_id / tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = rr0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) {
            V v = in[(long long)r * cols + c];
            if (tw_lo != nullptr) {
                const unsigned long long p = (unsigned long long)(r0 + r) * (unsigned long long)c;
                const V w = cmul(__ldg(tw_lo + (unsigned)(p & (unsigned long long)mask)), __ldg(tw_hi + (unsigned)(p >> shift)));
                v = cmul_tw<FWD>(v, w);
            }
            tile[ty + i][tx] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = rr0 + tx;
        if (r < rows && c < cols) out[(long long)c * rows + r] = tile[tx][ty + i];
    }
}

// Plain tiled transpose of rows x cols elements of any 4/8/16-byte type: out[c][r] = in[r][c].
// (Transforms of more than one shared-memory pass along a NON-last axis are done as transpose,
// last-axis transform, transpose.)
template <typename U>
__global__ void transpose_plain(const U *__restrict__ in, U *__restrict__ out, int rows, int cols) {
    __shared__ U tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tiles_c = (cols + 31) / 32;
    const int c0 = (int)(blockIdx.x % tiles_c) * 32, r0 = (int)(blockIdx.x / tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + i][tx] = in[(long long)r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (r < rows && c < cols) out[(long long)c * rows + r] = tile[tx][ty + i];
    }
}

// First step of a transform too long for the four-step plan (> 2^24 points): the line, seen as
// rows x cols, is transposed while being cast to complex and zero-padded past `limit` elements
// (gather + cast + pad of /root/reference/dsc/src/dsc.cpp:1981-1994).
template <typename T, bool IN_REAL>
__global__ void transpose_cast_pad(const void *__restrict__ in, cx<T> *__restrict__ out, int rows, int cols, long long limit) {
    using V = cx<T>;
    __shared__ V tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tiles_c = (cols + 31) / 32;
    const int c0 = (int)(blockIdx.x % tiles_c) * 32, r0 = (int)(blockIdx.x / tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int r = r0 + ty + i, c = c0 + tx;
        if (r < rows && c < cols) {
            const long long idx = (long long)r * cols + c;
            V v = mk<T>((T)0, (T)0);
            if (idx < limit) {
                if (IN_REAL) v.x = ((const T *)in)[idx];
                else v = ((const V *)in)[idx];
            }
            tile[ty + i][tx] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, r = r0 + tx;
        if (r < rows && c < cols) out[(long long)c * rows + r] = tile[tx][ty + i];
    }
}

// ------------------------------------------------------------------------------------------
// Frequency-domain pointwise product (mul_op, /root/reference/dsc/include/dsc_ops.h:68-78),
// b broadcast over rows when b_rows == 0.
template <typename T>
__global__ void cmul_rows(const cx<T> *a, const cx<T> *__restrict__ b, cx<T> *out,   // out may alias a
                          long long rows, long long cols, int b_rows) {
    const long long total = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const cx<T> x = a[i];
        const cx<T> y = b_rows ? b[i] : b[i % cols];
        out[i] = mk<T>(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
    }
}

}  // namespace dscfft
