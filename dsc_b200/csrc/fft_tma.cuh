// fft_tma.cuh -- the two passes of the four-step transform fed by the Tensor Memory Accelerator.
//
// Replaces, for dense complex rows of 2^15 .. 2^20 points (float; double up to 2^18), the register-direct
// tiles of four_step_fused: there the payload of a tile sits in 64 registers per thread while it travels to
// and from memory, the register file is full at 16 warps per SM, and nothing but those 16 warps can hide the
// DRAM / L2 latency (ncu: no unit above 40 %, stalls spread over load-queue throttling and payload latency).
// Here the payload travels by bulk tensor copies (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG) between
// memory and three 64 KiB shared-memory buffers, signalled through mbarriers; the butterfly warps only ever
// touch shared memory and registers; a loader thread and one storer thread per buffer keep loads and stores in flight.
//
//   n = n1 * n2, row x[i1][q] (i1 < n1 stride n2), result X[k1 + n1 k2] = out[k2][k1]
//
//   pass A  tile = L_A adjacent columns q, all i1: box [n1][L_A] of x  -> length-n1 transforms over i1, times
//           W_n^(q k1), TRANSPOSED on the way: the tile leaves as L_A contiguous lines W[q][k1] of the work row
//           -- one linear 64 KiB bulk store;
//   pass B  tile = L_B adjacent k1, all q: box [n2][L_B] of the work row W[q][k1] -> length-n2 transforms over
//           q -> box [n2][L_B] of out (row pitch n1).
//
// Tile layout in shared memory is [position][line] with a line extent of at least 64 bytes, which is both
// what a 2-D tensor box delivers and conflict-free for "adjacent lanes on adjacent lines"; the one exchange
// that turns the block around (pass A, before its last stage: threads come back as consecutive positions of
// ONE line so the final layout is [line][position]) uses an XOR swizzle of the line index by the low
// position bits.  Eight-line tiles (1024-point float passes) swizzle the 64-byte half of each 128-byte bank
// phase as well.
//
// Ticket order, work-row ring and the per-row completion counters are those of four_step_fused
// (FourStepSync, decode_ticket); they are handled by the producer thread alone.
//
// Reference work replaced: /root/reference/dsc/include/dsc_fft.h:57-103 (dsc_fft_pass2), :168-175 (1/N).
#pragma once

#if !defined(DSC_EMUL)

#include <cuda.h>
#include <cstdio>

#include "dsc_cuda.h"
#include "fft_kernels.cuh"

namespace dscfft {

namespace tma {

DSC_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

DSC_DEV void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DSC_DEV void mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSC_DEV void mbar_arrive_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DSC_DEV bool mbar_try_wait(void *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
DSC_DEV bool mbar_test_wait(void *bar, unsigned parity) {       // never suspends
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Watchdog of every wait in these kernels: a wait that has not been satisfied after 2^32 cycles (two seconds) is a
// protocol bug or a lost peer, not load -- say which one and stop the launch instead of hanging the device.
DSC_DEV void wait_timeout(const char *what, const void *bar, unsigned parity) {
    printf("dsc(cuda): block %u thread %u stuck in %s (barrier smem+0x%x, parity %u)\n", blockIdx.x, threadIdx.x, what,
           smem_u32(bar), parity);
    __trap();
}
DSC_DEV void mbar_wait(void *bar, unsigned parity, const char *what = "mbarrier wait") {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > (1LL << 32)) wait_timeout(what, bar, parity);
}
DSC_DEV void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (a following bulk store reads them)
DSC_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// generic-proxy acquire of a flag -> async-proxy (TMA) reads of the data it guards
DSC_DEV void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

DSC_DEV unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
DSC_DEV unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// box (c0.., c1.., c2) of a 3-D tensor -> shared memory, completion on an mbarrier
DSC_DEV void load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, void *bar, unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// box -> L2 only: the later load of the same box finds it there
DSC_DEV void prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
DSC_DEV void store_3d(const CUtensorMap *map, int c0, int c1, int c2, const void *src, unsigned long long pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
        ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)), "l"(pol) : "memory");
}
DSC_DEV void store_linear(void *gdst, const void *src, unsigned bytes, unsigned long long pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(gdst), "r"(smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
}
// register -> global with an L2 eviction hint (the direct-store tiles: work rows evict_last, results evict_first)
DSC_DEV void st_hint(float2 *p, const float2 v, const unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
DSC_DEV void st_hint(double2 *p, const double2 v, const unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
DSC_DEV void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DSC_DEV void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
DSC_DEV void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
DSC_DEV void store_wait_all_but_one() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }

}  // namespace tma

// ---- geometry of one launch -------------------------------------------------------------------------------
constexpr int TMA_TILE_BYTES = 64 * 1024;
constexpr int TMA_BUFFERS = 3;
#if !defined(DSC_TMA_GROUPS)
#define DSC_TMA_GROUPS 2
#endif
constexpr int TMA_GROUPS = DSC_TMA_GROUPS;                 // consumer groups of TMA_GROUP_THREADS threads
constexpr int TMA_GROUP_THREADS = 256;
constexpr int TMA_THREADS = TMA_GROUPS * TMA_GROUP_THREADS + 64;     // + the loader warp and the storer warp (fft_cluster_pipe)
// four_step_tma: the loader warp and one storer warp PER BUFFER.  A storer's chain per tile -- bulk store, wait until it has
// read the buffer, wait until it has completed, proxy fence, release of the row counter -- is ~5000 cycles of pure latency;
// with a single storer that chain, not the memory system and not the butterflies, was the launch's period (a launch whose
// groups pass the tiles through untouched ran no faster than the full transform: tools/micro/two_pass_copy.cu).
constexpr int TMA4_THREADS = TMA_GROUPS * TMA_GROUP_THREADS + 32 * (1 + TMA_BUFFERS);
constexpr int TMA_DONE_RING = 8;                           // a group is never more than three tiles ahead of the storer

template <typename T> __host__ __device__ constexpr int tma_lg_e() { return sizeof(T) == 4 ? 5 : 4; }     // 32 / 16 points per thread
template <typename T> __host__ __device__ constexpr int tma_tile_points() { return TMA_TILE_BYTES / (int)sizeof(cx<T>); }
// lines per tile of a pass of 2^lg points
template <typename T> __host__ __device__ constexpr int tma_lines(int lg) { return tma_tile_points<T>() >> lg; }
// rows per tensor box (box extents are at most 256)
__host__ __device__ constexpr int tma_box_rows(int lg) { return (1 << lg) < 256 ? (1 << lg) : 256; }

struct TmaArgs {
    void *work;                    // ring of work rows W[ring row][q][k1]
    long long ring;                // 0 = one work row per row
    const void *tw_a[DSC_CUDA_MAX_STAGES];     // stage tables of the length-n1 transform
    const void *tw_b[DSC_CUDA_MAX_STAGES];     // stage tables of the length-n2 transform
    const void *tw_lo, *tw_hi;     // W_n^p split tables
    int four_shift, four_mask;
    double scale;                  // applied by pass B when do_scale (1/n of the inverse)
    int do_scale;
    int keep_out;                  // the output is re-read soon by another kernel: keep it in L2
    int discard_work;              // second-pass tiles drop their work-row lines from L2 once loaded (no write-back)
    // fused packed-real launches (REAL != 0): bin-pair step inside the pass that holds both bins of a pair
    const void *twr_lo, *twr_hi;   // W_2n^k split tables, k <= n/2
    int real_shift, real_mask;
    const void *filt;              // REAL == 1: spectrum B[0..n] of the fused filter (nullptr: plain rfft un-mix)
    void *out;                     // REAL == 1: the output rows (for the two bins / the line no box store covers)
    long long out_pitch;           // elements
    const void *in;                // REAL == 2: the bin rows X[0..n] (for the one column no box load covers)
    long long in_pitch;
    int prefetch;                  // first-pass boxes are prefetched into L2 when their ticket is taken, two tiles ahead
    int debug_skip;                // builds with DSC_TMA_EXPERIMENTS only (wrong results): 1 = tiles pass through untouched,
                                   // 2 = only the shared-memory traffic of a tile (two round trips per point), no arithmetic
};

struct TmaTileDesc { unsigned role_a, row, r, exit; };

template <typename T, int TILE = TMA_TILE_BYTES> struct TmaSmem {
    using V = cx<T>;
    // (c, line) inter-pass twiddles of one pass-A tile, per group: lines x points per thread, at most 32 x 32 (float, 64 KiB
    // tiles of 256-point passes), 32 x 16 for the half-size tiles and for double
    static constexpr int TABLE_MAX = (sizeof(T) == 4 && TILE >= TMA_TILE_BYTES) ? 32 * 32 : 32 * 16;
    alignas(1024) unsigned char buf[TMA_BUFFERS][TILE];
    V table[TMA_GROUPS][TABLE_MAX];
    unsigned long long full[TMA_BUFFERS];          // bytes of the tile have landed (producer + TMA -> consumers)
    unsigned long long ready[TMA_BUFFERS];         // the finished tile lies in the buffer (consumers -> storer)
    unsigned long long empty[TMA_BUFFERS];         // the store has read the buffer (storer -> loader)
    unsigned long long posted[TMA_BUFFERS];        // desc[b] names the tile that is on its way (loader -> groups)
    TmaTileDesc desc[TMA_BUFFERS];
    V ladder[2][5 * 32];                           // the stage-1 twiddle rows of the two passes (TmaTile::ladder_fill)
    // direct-store launches: tile t's stores have been issued by every warp of its group (groups -> storer), ring over t
    unsigned long long done[TMA_DONE_RING];
    TmaTileDesc sdesc[TMA_DONE_RING];
};

// W_n^p from the two sqrt(n)-sized tables
template <typename T> DSC_DEV cx<T> tma_twiddle(const TmaArgs &a, const unsigned p) {
    const cx<T> lo = __ldg((const cx<T> *)a.tw_lo + (p & (unsigned)a.four_mask));
    const cx<T> hi = __ldg((const cx<T> *)a.tw_hi + (p >> a.four_shift));
    return cmul(lo, hi);
}

// ---- one tile on one consumer group -----------------------------------------------------------------------
// SPLIT  : the tile's lines are two runs of L/2 adjacent lines, delivered (and stored) as two boxes [position][L/2],
//          the second one behind the first (fused packed-real launches: a run and its mirror image).
// PERMUTE: a TRANSPOSE tile writes its lines with the columns k1 > N/2 moved down by one and k1 = N/2 in the last column,
//          so that the mirror image of an ALIGNED run of columns [a, a + m) of the next pass is the aligned run
//          [N - a - m, N - a) (columns k1 = N - a - m + 1 .. N - a, where "N" stands for N/2).
template <typename T, int LG_N, int L, bool FWD, bool TRANSPOSE, int LGE = tma_lg_e<T>(), bool SPLIT = false, bool PERMUTE = false>
struct TmaTile {
    using V = cx<T>;
    static constexpr int LH = L / 2;
    // element index of (position, line) in the layout the boxes have
    static DSC_DEV int box_at(const int pos, const int l) {
        if constexpr (SPLIT) return (l >= LH ? (LH << LG_N) : 0) + pos * LH + (l & (LH - 1));
        else return pos * L + l;
    }
    static constexpr int LG_E = LGE < LG_N ? LGE : LG_N;
    using Sc = Sched<LG_N, LG_E>;
    static constexpr int N = Sc::N, E = Sc::E, TT = Sc::TT, STAGES = Sc::STAGES;
    static constexpr int SLOTS = 128 / (int)sizeof(V);           // elements of one 128-byte bank phase: 16 / 8
    static constexpr bool NARROW = L < SLOTS;                    // a phase spans two positions
    static_assert(L * TT == TMA_GROUP_THREADS, "a tile is one register tile per thread of the group");
    static_assert(L * 2 >= SLOTS, "lines of at least 64 bytes");
    // narrow tiles: the half flip by position bit LG_E separates the two positions of a phase for the radix-E scatter
    // (E j + p), and bit 0 separates them for every access whose positions differ by one: the readers (j + c TT) and the
    // scatter of a second full-radix stage (((jj - k) E) + k + E p, k = jj mod E)
    static_assert(!NARROW || (STAGES == 2 && LG_N == 2 * LG_E) || (STAGES == 3 && LG_N > 2 * LG_E),
                  "narrow tiles: full-radix first (and second) stage only");
    static constexpr int LG_TT = LG_N - LG_E;
    static constexpr int LG_SLOTS = SLOTS == 16 ? 4 : 3;
    static constexpr int SW_BITS = LG_TT < LG_SLOTS ? LG_TT : LG_SLOTS;

    // element index of (position, line) in the exchange layout.  JFAST: the readers are consecutive positions of one
    // line (and, when a line has fewer threads than a phase has lanes, a few adjacent lines): the low position bits
    // are XOR-ed into the high bits of the line's slot.
    template <bool JFAST> static DSC_DEV int phys(const int pos, const int l) {
        if constexpr (NARROW) {
            // 8 lines of 8 bytes (4 of 16): the two positions of a phase are (j, j+1) for readers and (32 j + p,
            // 32 (j+1) + p) for the first scatter -- flip the half by position bit LG_E so both differ in it, and
            // rotate the line by the next position bits for the consecutive-position readers
            const int row = (pos & ~1) | ((pos ^ (pos >> LG_E)) & 1);
            return row * L + (l ^ ((pos >> 1) & (L - 1)));
        } else if constexpr (JFAST) {
            return pos * L + (l ^ ((pos & ((1 << SW_BITS) - 1)) << (LG_SLOTS - SW_BITS)));
        } else {
            return pos * L + l;
        }
    }

        // lad: optional shared-memory copy of the rows of the stage-1 table a butterfly loads (ladder_rows() of them, row r =
    // table row 2^r - 1 when the ladder is used, else row r), see ladder_fill(); nullptr = read the table in global memory
    static constexpr int LADDER_R1 = STAGES >= 2 ? (1 << Sc::lg_r(1)) : 1;
    static constexpr bool LADDER_POW2 = LADDER_R1 >= 8;
    static constexpr int ladder_rows() {
        int r = 0;
        if (LADDER_POW2) { for (int m = 1; m < LADDER_R1; m *= 2) ++r; } else r = LADDER_R1 - 1;
        return r;
    }
    static constexpr int LADDER_ELEMS = ladder_rows() * E;        // stage 1: NS = E entries per row
    static DSC_DEV void ladder_fill(V *lad, const void *const *tw_all, const int tid, const int threads) {
        if constexpr (STAGES >= 2) {
            const V *__restrict__ tw = (const V *)tw_all[1];
            for (int i = tid; i < LADDER_ELEMS; i += threads) {
                const int r = i / E, k = i % E;
                const int m = LADDER_POW2 ? (1 << r) : r + 1;
                lad[i] = __ldg(tw + (m - 1) * E + k);
            }
        }
    }
    // rel: optional mbarrier a warp arrives on once its threads have read their last-stage inputs, the last time the tile
    // touches its buffer (direct-store tiles: the loader may refill the buffer while the last stage and the stores run)
    static DSC_DEV void release_buffer(const unsigned rel) {      // shared-window address of the mbarrier, 0 = none
        if (rel != 0u) {
            tma::fence_async_smem();           // this thread's exchange writes before the bulk copy that overwrites them
            __syncwarp();
            if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rel) : "memory");
        }
    }
    template <int S>
    static DSC_DEV void stage(V (&v)[E], V *buf, const int l, const int j, const int l_last, const int j_last,
                              const void *const *tw_all, const int bar_id, const V *lad = nullptr, const unsigned rel = 0u) {
        constexpr int LG_R = Sc::lg_r(S), R = 1 << LG_R, NB = E / R;
        constexpr int LG_NS = S * LG_E, NS = 1 << LG_NS;
        constexpr bool LAST = S == STAGES - 1;
        constexpr bool NEXT_JFAST = TRANSPOSE && (S + 1 == STAGES - 1);
        const V *__restrict__ tw = (const V *)tw_all[S];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int jj = j + b * TT;
            const int k = jj & (NS - 1);
            V r[R];
#pragma unroll
            for (int m = 0; m < R; ++m) r[m] = v[b + m * NB];
            if constexpr (S > 0) {
                if constexpr (R >= 8) {
                    V w[R];
#pragma unroll
                    for (int m = 1; m < R; ++m) {
                        if ((m & (m - 1)) == 0) {
                            int r = 0;
                            while ((1 << r) < m) ++r;
                            w[m] = (S == 1 && lad != nullptr) ? lad[r * NS + k] : __ldg(tw + (m - 1) * NS + k);
                        } else {
                            int hi = 1;
                            while (hi * 2 <= m) hi *= 2;
                            w[m] = cmul(w[hi], w[m - hi]);
                        }
                        r[m] = cmul_tw<FWD>(r[m], w[m]);
                    }
                } else {
#pragma unroll
                    for (int m = 1; m < R; ++m)
                        r[m] = cmul_tw<FWD>(r[m], (S == 1 && lad != nullptr) ? lad[(m - 1) * NS + k] : __ldg(tw + (m - 1) * NS + k));
                }
            }
            Dft<R, FWD, T>::run(r);
            if constexpr (LAST) {
#pragma unroll
                for (int p = 0; p < R; ++p) v[b + p * NB] = r[p];
            } else {
                if (b == 0) dsc_group_barrier(bar_id, TMA_GROUP_THREADS);    // every thread has read the previous layout
                const int base = ((jj - k) << LG_R) + k;
#pragma unroll
                for (int p = 0; p < R; ++p) buf[phys<NEXT_JFAST>(base + p * NS, l)] = r[p];
            }
        }
        if constexpr (!LAST) {
            dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
            const int ln = NEXT_JFAST ? l_last : l, jn = NEXT_JFAST ? j_last : j;
#pragma unroll
            for (int c = 0; c < E; ++c) v[c] = buf[phys<NEXT_JFAST>(jn + c * TT, ln)];
            if constexpr (S + 1 == STAGES - 1) release_buffer(rel);
            stage<S + 1>(v, buf, ln, jn, l_last, j_last, tw_all, bar_id, lad, rel);
        }
    }

    // Column q of line ll of first-pass tile u.  SPLIT (fused irfft): the lo run [u LH, u LH + LH) and its mirror image
    // [n2 - u LH - LH + 1, n2 - u LH]; column "n2" (tile 0) is the self-paired column n2/2.
    static DSC_DEV unsigned line_q(const unsigned u, const unsigned ll, const unsigned n_other) {
        if constexpr (SPLIT) {
            if (ll < (unsigned)LH) return u * LH + ll;
            const unsigned q = n_other - u * LH - LH + 1 + (ll - LH);
            return q == n_other ? n_other / 2 : q;
        } else {
            (void)n_other;
            return u * L + ll;
        }
    }

    // buf: the tile as the box delivered it, [position][line]; on return the finished tile, [line][position]
    // (TRANSPOSE, times the inter-pass twiddle W_n^(q k1)) or [position][line] (times the inverse's 1/n).
    // Inter-pass twiddles of a TRANSPOSE tile, W_n^(q k1) with k1 = j + c TT: W^(q j) per thread (returned) and W^(q TT c)
    // in a (c, line) table of the tile.  Depends on the tile's position only, so the launch calls it as soon as the tile's
    // descriptor is known, while the tile itself is still travelling.  The group barrier keeps the previous tile's readers
    // of the table ahead of its new contents.
    static DSC_DEV V prepare(V *table, const TmaArgs &a, const unsigned u, const unsigned n_other, const int gtid, const int bar_id) {
        const int l_last = gtid / TT, j_last = gtid % TT;
        dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
        for (int i = gtid; i < L * E; i += TMA_GROUP_THREADS) {
            const unsigned ll = (unsigned)(i % L), c = (unsigned)(i / L);
            table[i] = tma_twiddle<T>(a, line_q(u, ll, n_other) * (unsigned)TT * c);
        }
        return tma_twiddle<T>(a, line_q(u, (unsigned)l_last, n_other) * (unsigned)j_last);
    }

    // gout != nullptr (direct-store tiles): the finished points leave from the registers -- gout is this thread's first
    // output element, its other points follow at a compile-time pitch (TRANSPOSE: TT elements, the thread's line of the work
    // row; else TT << LG_OS, the result's rows of 2^LG_OS bins) -- and the buffer is released through `rel` right after the
    // last exchange instead of carrying the finished tile to a bulk store.
    template <int LG_OS = 0>
    static DSC_DEV void run(V *buf, V *table, const TmaArgs &a, const void *const *tw_all, const unsigned q0,
                            const int gtid, const int bar_id, const V *lad = nullptr, const bool prepared = false,
                            V w0 = V{}, void *const obase = nullptr, const unsigned line0 = 0u, const int keep = 0,
                            const unsigned rel = 0u) {
        const int l = gtid % L, j = gtid / L;
        const int l_last = TRANSPOSE ? gtid / TT : l, j_last = TRANSPOSE ? gtid % TT : j;
        V v[E];
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = buf[box_at(j + c * TT, l)];
        if constexpr (TRANSPOSE) {
            if (!prepared) w0 = tma_twiddle<T>(a, (q0 + (unsigned)l_last) * (unsigned)j_last);
        }
        stage_first(v, buf, table, a, l, j, l_last, j_last, tw_all, q0, gtid, bar_id, lad, prepared, rel);
        if (obase != nullptr) {
            // address and policy are built here, after the butterflies: nothing 64-bit lives through the transform
            if constexpr (TRANSPOSE) {
                const unsigned long long pol = tma::policy_evict_last();
                V *gout = (V *)obase + (((long long)line0 + l_last) << LG_N) + j_last;
#pragma unroll
                for (int c = 0; c < E; ++c) {
                    const V w = c == 0 ? w0 : cmul(w0, table[c * L + l_last]);
                    tma::st_hint(gout + c * TT, cmul_tw<FWD>(v[c], w), pol);
                }
            } else {
                const unsigned long long pol = keep ? tma::policy_evict_last() : tma::policy_evict_first();
                V *gout = (V *)obase + line0 + l + ((long long)j << LG_OS);
                if (a.do_scale) {
                    const T s = (T)a.scale;
#pragma unroll
                    for (int c = 0; c < E; ++c) { v[c].x *= s; v[c].y *= s; }
                }
#pragma unroll
                for (int c = 0; c < E; ++c) tma::st_hint(gout + ((long long)(c * TT) << LG_OS), v[c], pol);
            }
            return;
        }
        // every thread has read its last-stage inputs: the buffer may take the finished tile
        dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
        if constexpr (TRANSPOSE) {
#pragma unroll
            for (int c = 0; c < E; ++c) {
                const V w = c == 0 ? w0 : cmul(w0, table[c * L + l_last]);
                const int k1 = j_last + c * TT;
                const int col = !PERMUTE ? k1 : k1 < N / 2 ? k1 : k1 == N / 2 ? N - 1 : k1 - 1;
                buf[l_last * N + col] = cmul_tw<FWD>(v[c], w);
            }
        } else {
            if (a.do_scale) {
                const T s = (T)a.scale;
#pragma unroll
                for (int c = 0; c < E; ++c) { v[c].x *= s; v[c].y *= s; }
            }
#pragma unroll
            for (int c = 0; c < E; ++c) buf[box_at(j + c * TT, l)] = v[c];
        }
    }

    // stage 0 with the table build folded in after its first barrier
    static DSC_DEV void stage_first(V (&v)[E], V *buf, V *table, const TmaArgs &a, const int l, const int j,
                                    const int l_last, const int j_last, const void *const *tw_all, const unsigned q0,
                                    const int gtid, const int bar_id, const V *lad = nullptr, const bool prepared = false,
                                    const unsigned rel = 0u) {
        static_assert(STAGES >= 2, "a pass has at least one exchange");
        constexpr int R = 1 << Sc::lg_r(0), NB = E / R;
        static_assert(NB == 1, "the first stage is a full-radix butterfly");
        constexpr bool NEXT_JFAST = TRANSPOSE && (1 == STAGES - 1);
        Dft<R, FWD, T>::run(v);
        dsc_group_barrier(bar_id, TMA_GROUP_THREADS);        // every thread of the group has read the delivered tile
        const int base = j << Sc::lg_r(0);
#pragma unroll
        for (int p = 0; p < R; ++p) buf[phys<NEXT_JFAST>(base + p, l)] = v[p];
        if constexpr (TRANSPOSE) {
            if (!prepared) {
                for (int i = gtid; i < L * E; i += TMA_GROUP_THREADS) {
                    const unsigned ll = (unsigned)(i % L), c = (unsigned)(i / L);
                    table[i] = tma_twiddle<T>(a, (q0 + ll) * (unsigned)TT * c);
                }
            }
        }
        dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
        const int ln = NEXT_JFAST ? l_last : l, jn = NEXT_JFAST ? j_last : j;
#pragma unroll
        for (int c = 0; c < E; ++c) v[c] = buf[phys<NEXT_JFAST>(jn + c * TT, ln)];
        if constexpr (1 == STAGES - 1) release_buffer(rel);
        stage<1>(v, buf, ln, jn, l_last, j_last, tw_all, bar_id, lad, rel);
    }
};

// ---- fused packed-real transforms: the bin-pair step on a finished second-pass tile ------------------------------
// REAL == 1 (rfft and the forward half of the fused filter).  Pass A wrote its work lines with permuted columns
// (TmaTile PERMUTE), so second-pass tile u holds two ALIGNED runs of LH = L/2 work columns:
//   lo: columns [u LH, u LH + LH)           = k1 = u LH + l,                      l = 0 .. LH-1
//   hi: columns [n1 - u LH - LH, n1 - u LH) = k1 = n1 - u LH - LH + 1 + m,        m = 0 .. LH-1   (k1 = "n1" is n1/2)
// and bin k = k1 + n1 k2 of lo line l meets its partner n - k = (n1 - k1) + n1 (n2 - 1 - k2) in hi line LH-1-l at position
// n2 - 1 - k2: both bins of every pair are in the tile, which leaves as X (or as the filter's packed z') instead of Z -- no
// separate sweep over the spectrum (dsc_fft.h:199-214; real_mix_rows / filter_pairs_rows do the same per bin pair).  The
// two self-paired lines share tile 0's pair (lo 0, hi LH-1): k1 = 0 pairs k2 with n2 - k2 (k2 = 0: DC and Nyquist, k2 = n2/2:
// bin n/2), k1 = n1/2 pairs k2 with n2 - 1 - k2.  The hi run is stored one column further right than it was loaded
// (k1 = column + 1); its last column in tile 0 falls outside the box store's tensor and is written by the group itself.
// (The shifted box starts 16 bytes off an aligned run for complex128 -- fine -- but 8 bytes off for complex64, which the
// bulk tensor store does not take: measured, the launch dies with "illegal instruction".)
template <typename T, int LG_N1, int LG_N2, int L>
DSC_DEV void tma_unmix_tile(cx<T> *buf, const TmaArgs &a, const unsigned u, const unsigned row, const int gtid, const int bar_id) {
    using V = cx<T>;
    constexpr int N2 = 1 << LG_N2, LH = L / 2;
    constexpr unsigned n1 = 1u << LG_N1, n = 1u << (LG_N1 + LG_N2);
    V *lo = buf, *hi = buf + N2 * LH;
    const V *__restrict__ t_lo = (const V *)a.twr_lo, *__restrict__ t_hi = (const V *)a.twr_hi;
    const V *__restrict__ flt = (const V *)a.filt;
    // *pa = Z[k], *pb = Z[n - k] -> X[k], X[n - k] (or z'[k], z'[n - k]); the tables and the reference order want k <= n/2
    auto pair = [&](V *pa, V *pb, const unsigned k) {
        const bool up = k > n / 2;
        const unsigned ks = up ? n - k : k;
        const V w = cmul(__ldg(t_lo + (ks & (unsigned)a.real_mask)), __ldg(t_hi + (ks >> a.real_shift)));
        const V za = up ? *pb : *pa, zb = up ? *pa : *pb;
        V ra, rb;
        if (flt != nullptr) filter_pair<T>(za, zb, w, __ldg(flt + ks), __ldg(flt + (n - ks)), ra, rb);
        else real_pair<true, T>(za, zb, w, ra, rb);
        *pa = up ? rb : ra;
        *pb = up ? ra : rb;
    };
    // A thread keeps its line l (the group's 256 threads are a multiple of LH) and takes every K2_STEP-th position:
    // W_2n^k = W_2n^k1 (one lookup per tile) times W_2n^(n1 k2) (one entry of the coarse table per pair; the upper half by
    // W^(n1 (n2 - m)) = -conj W^(n1 m)), valid for every 0 < k < n -- no role swap, one table load per pair.
    constexpr int PAIRS = N2 * LH / TMA_GROUP_THREADS, K2_STEP = TMA_GROUP_THREADS / LH;
    static_assert(PAIRS * TMA_GROUP_THREADS == N2 * LH && K2_STEP * LH == TMA_GROUP_THREADS, "whole pairs per thread");
    const int l = gtid % LH, k2_0 = gtid / LH;
    const bool special = u == 0 && l == 0;
    const unsigned k1 = u * LH + l;
    const int d = LG_N1 - a.real_shift;                      // coarse-table entries per step of n1 (log2)
    constexpr int CH = PAIRS < 4 ? PAIRS : 4;                // spectrum bins are requested CH pairs at a time (L2 hits)
    static_assert(PAIRS % CH == 0, "whole chunks");
    V fa[CH], fb[CH];
    V w1 = mk<T>((T)1, (T)0);
    auto request = [&](const int p0) {
#pragma unroll
        for (int p = 0; p < CH; ++p) {
            const unsigned k = k1 + ((unsigned)(k2_0 + (p0 + p) * K2_STEP) << LG_N1);
            fa[p] = __ldg(flt + k);
            fb[p] = __ldg(flt + (n - k));
        }
    };
    if (!special) {
        w1 = cmul(__ldg(t_lo + (k1 & (unsigned)a.real_mask)), __ldg(t_hi + (k1 >> a.real_shift)));
        if (flt != nullptr) request(0);
    }
    dsc_group_barrier(bar_id, TMA_GROUP_THREADS);            // the finished tile is complete
    if (!special) {
#pragma unroll
        for (int p0 = 0; p0 < PAIRS; p0 += CH) {
            V ra[CH], rb[CH];
#pragma unroll
            for (int p = 0; p < CH; ++p) {
                const int k2 = k2_0 + (p0 + p) * K2_STEP;
                const bool up = k2 > N2 / 2;
                const V wt = __ldg(t_hi + ((unsigned)(up ? N2 - k2 : k2) << d));
                const V w = cmul(w1, up ? mk<T>(-wt.x, wt.y) : wt);
                const V za = lo[k2 * LH + l], zb = hi[(N2 - 1 - k2) * LH + (LH - 1 - l)];
                if (flt != nullptr) filter_pair<T>(za, zb, w, fa[p], fb[p], ra[p], rb[p]);
                else real_pair<true, T>(za, zb, w, ra[p], rb[p]);
            }
            if (flt != nullptr && p0 + CH < PAIRS) request(p0 + CH);
#pragma unroll
            for (int p = 0; p < CH; ++p) {
                const int k2 = k2_0 + (p0 + p) * K2_STEP;
                lo[k2 * LH + l] = ra[p];
                hi[(N2 - 1 - k2) * LH + (LH - 1 - l)] = rb[p];
            }
        }
    } else {
        // tile 0, the pair of self-paired lines (lo 0, hi LH-1)
        for (int k2 = k2_0; k2 < N2; k2 += K2_STEP) {
            if (k2 >= N2 / 2) {
                const int kk = k2 - N2 / 2;                       // line k1 = n1/2
                pair(hi + kk * LH + (LH - 1), hi + (N2 - 1 - kk) * LH + (LH - 1), n1 / 2 + ((unsigned)kk << LG_N1));
            } else if (k2 != 0) {                                 // line k1 = 0
                pair(lo + k2 * LH, lo + (N2 - k2) * LH, (unsigned)k2 << LG_N1);
            } else {
                const V z0 = lo[0], zh = lo[(N2 / 2) * LH];
                if (flt != nullptr) {
                    lo[0] = filter_dc<T>(z0, __ldg(flt), __ldg(flt + n));
                    lo[(N2 / 2) * LH] = filter_mid<T>(zh, __ldg(flt + n / 2));
                } else {
                    lo[0] = mk<T>(z0.x + z0.y, (T)0);
                    ((V *)a.out)[(long long)row * a.out_pitch + n] = mk<T>(z0.x - z0.y, (T)0);
                    lo[(N2 / 2) * LH] = mk<T>(zh.x, -zh.y);
                }
            }
        }
    }
    if constexpr (sizeof(V) == 8) {
        // complex64: the hi run's first column k1 = n1 - u LH - LH + 1 is odd, 8 bytes off the 16-byte granule a bulk tensor
        // store must start on -- the group stores the run itself (rows of LH adjacent bins; k1 = "n1" is line n1/2)
        dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
        V *o = (V *)a.out + (long long)row * a.out_pitch;
        const unsigned k1_0 = n1 - u * LH - LH + 1;
#pragma unroll 4
        for (int i = gtid; i < N2 * LH; i += TMA_GROUP_THREADS) {
            const unsigned k1 = k1_0 + (unsigned)(i % LH);
            st_stream(o + (k1 == n1 ? n1 / 2 : k1) + ((long long)(i / LH) << LG_N1), hi[i]);
        }
    } else if (u == 0) {
        dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
        V *o = (V *)a.out + (long long)row * a.out_pitch + n1 / 2;
        for (int k2 = gtid; k2 < N2; k2 += TMA_GROUP_THREADS) o[(long long)k2 << LG_N1] = hi[k2 * LH + (LH - 1)];
    }
}

// REAL == 2 (irfft): the packed points z[k] are built from the bin pairs (X[k], X[n - k]) inside the inverse transform's
// FIRST pass.  With the bins viewed as [i1][q] (q < n2 + 1, row pitch n2: column n2 of row i1 is column 0 of row i1 + 1, and
// bin n closes the last row), first-pass tile u loads two runs of LH = L/2 columns,
//   lo: q = u LH + l,   hi: q = n2 - u LH - LH + 1 + m,     l, m = 0 .. LH-1,
// and bin k = i1 n2 + q of lo line l meets n - k = (n1 - 1 - i1) n2 + (n2 - q) in hi line LH-1-l at row n1 - 1 - i1.  Tile 0's
// hi line LH-1 (q = "n2") only serves as the partner of column 0; its slot is then given to the one column the runs leave
// out, q = n2/2, whose bins the group fetched itself (xm: rows t and n1 - 1 - t per thread) while the boxes were travelling.
template <typename T, int LG_N1, int LG_N2, int L>
DSC_DEV void tma_mix_tile(cx<T> *buf, const TmaArgs &a, const unsigned u, const int gtid, const int bar_id,
                          const cx<T> xm_a, const cx<T> xm_b) {
    using V = cx<T>;
    constexpr int N1 = 1 << LG_N1, LH = L / 2;
    constexpr unsigned n2 = 1u << LG_N2, n = 1u << (LG_N1 + LG_N2);
    V *lo = buf, *hi = buf + N1 * LH;
    const V *__restrict__ t_lo = (const V *)a.twr_lo, *__restrict__ t_hi = (const V *)a.twr_hi;
    // (X[k], X[n - k]) -> (z[k], z[n - k]); the tables and the reference order want k <= n/2
    auto pair = [&](V xa, V xb, const unsigned k, V &za, V &zb) {
        const bool up = k > n / 2;
        const unsigned ks = up ? n - k : k;
        const V w = cmul(__ldg(t_lo + (ks & (unsigned)a.real_mask)), __ldg(t_hi + (ks >> a.real_shift)));
        if (ks == 0) { xa.y = (T)0; xb.y = (T)0; }            // imaginary parts of DC and Nyquist are ignored (dsc_fft.h:220-228)
        V ra, rb;
        real_pair<false, T>(up ? xb : xa, up ? xa : xb, w, ra, rb);
        za = up ? rb : ra;
        zb = up ? ra : rb;
    };
    // a thread keeps its column q = u LH + l and takes every I1_STEP-th row: W_2n^k = W_2n^q times W_2n^(n2 i1), as in
    // tma_unmix_tile
    constexpr int PAIRS = N1 * LH / TMA_GROUP_THREADS, I1_STEP = TMA_GROUP_THREADS / LH;
    static_assert(PAIRS * TMA_GROUP_THREADS == N1 * LH && I1_STEP * LH == TMA_GROUP_THREADS, "whole pairs per thread");
    {
        const int l = gtid % LH, i1_0 = gtid / LH;
        const unsigned q = u * LH + l;
        const int d = LG_N2 - a.real_shift;
        const V w1 = cmul(__ldg(t_lo + (q & (unsigned)a.real_mask)), __ldg(t_hi + (q >> a.real_shift)));
#pragma unroll
        for (int p = 0; p < PAIRS; ++p) {
            const int i1 = i1_0 + p * I1_STEP;
            const bool up = i1 > N1 / 2;
            const V wt = __ldg(t_hi + ((unsigned)(up ? N1 - i1 : i1) << d));
            const V w = cmul(w1, up ? mk<T>(-wt.x, wt.y) : wt);
            V *pa = lo + i1 * LH + l, *pb = hi + (N1 - 1 - i1) * LH + (LH - 1 - l);
            V xa = *pa, xb = *pb, za, zb;
            if (q == 0 && i1 == 0) { xa.y = (T)0; xb.y = (T)0; }       // imaginary parts of DC and Nyquist are ignored (dsc_fft.h:220-228)
            real_pair<false, T>(xa, xb, w, za, zb);
            *pa = za;
            if (q != 0) *pb = zb;                                      // tile 0: hi line LH-1 is only column 0's partner
        }
    }
    if (u == 0) {
        // column n2/2 into the slot of hi line LH-1: rows t and N1 - 1 - t are a pair
        dsc_group_barrier(bar_id, TMA_GROUP_THREADS);        // the partners of column 0 have been read
        for (int t = gtid; t < N1 / 2; t += TMA_GROUP_THREADS) {
            V za, zb;
            pair(xm_a, xm_b, (unsigned)t * n2 + n2 / 2, za, zb);
            hi[t * LH + (LH - 1)] = za;
            hi[(N1 - 1 - t) * LH + (LH - 1)] = zb;
        }
    }
    dsc_group_barrier(bar_id, TMA_GROUP_THREADS);            // the tile holds z
}

// ---- the persistent launch --------------------------------------------------------------------------------
// One block per SM: warps 0..15 are two butterfly groups, lane 0 of warp 16 is the loader, lane 0 of warps 17..19 the
// storers of buffers 0..2.  The block's tile sequence is t = 0, 1, 2, ...: tile t is transformed by group t % 2 in buffer
// t % 3.
//   loader, tile t   :  ticket and row counter looked up one tile ahead -> wait for the tile's dependency (all first-pass
//                       tiles of the row; the second pass of the row that used the work row before) -> wait empty[t % 3]
//                       (the store of tile t - 3 has read the buffer) -> descriptor + box loads, completion on full[t % 3].
//   group, tile t    :  wait full[t % 3] -> transform in place -> fence.proxy.async -> one arrival per warp on ready[t % 3].
//   storer t % 3     :  wait ready[t % 3] -> bulk store -> wait until it has READ the buffer -> arrive on empty[t % 3]
//                       -> wait until it has COMPLETED -> proxy fence -> release-add on the tile's row counter.
// The launch is paced by this copy pipeline (three buffers circulating between the copy engine and the groups), not by the
// butterflies: DESIGN.md 4a, tools/micro/two_pass_copy.cu.
// LGE / TILE: points per thread (log2) and bytes per tile.  The default (32 points, 64 KiB, one block per SM) leaves an SM
// 16 butterfly warps at 96 registers.  With 16 points per thread the registers allow twice the warps, but two 512-thread
// groups plus the copy warps exceed the 1024-thread block limit; so the 16-point variant runs on 32 KiB tiles (groups stay
// 256 threads) with TWO blocks per SM: 32 butterfly warps, six tiles in flight.  It is the default for float passes of at
// most 256 points (a 512-point pass would need a second exchange, and its half-size tiles 64-byte rows).
// REAL: 0 = complex rows; 1 = forward with the packed-real bin-pair step fused into the second pass (tma_unmix_tile);
// 2 = inverse whose first pass builds the packed points from the bins of the real transform (tma_mix_tile).
// DIRECT (opt-in, measured slower): finished tiles leave from the registers (coalesced st.global, runs of 64 - 256 bytes)
// instead of going back into the buffer for a bulk store, and a tile gives its buffer back right after its last exchange: a
// buffer is held for load + half a transform instead of load + transform + store, but the 32 STG.64 per thread cost the
// load/store pipe more than the freed buffer time returns.  The storer of buffer 0 only publishes row counters (after
// every warp of the group has issued its stores).
template <typename T, int LG_N1, int LG_N2, bool FWD, int LGE = tma_lg_e<T>(), int TILE = TMA_TILE_BYTES, int REAL = 0, bool DIRECT = false>
__global__ void __launch_bounds__(TMA4_THREADS, (TILE * 2 <= TMA_TILE_BYTES ? 2 : 1))
four_step_tma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
              const __grid_constant__ CUtensorMap map_out, const TmaArgs a, const FourStepSync s) {
    using V = cx<T>;
    constexpr int TMA_TILE_BYTES = TILE;           // shadows the default for everything below
    constexpr int L_A = (TILE / (int)sizeof(V)) >> LG_N1, L_B = (TILE / (int)sizeof(V)) >> LG_N2;
    constexpr int BOX_A = tma_box_rows(LG_N1), BOX_B = tma_box_rows(LG_N2);
    using TileA = TmaTile<T, LG_N1, L_A, FWD, true, LGE, REAL == 2, REAL == 1>;
    using TileB = TmaTile<T, LG_N2, L_B, FWD, false, LGE, REAL == 1>;
    constexpr int LH_B = L_B / 2;                  // REAL == 1: second-pass tiles are two runs of LH_B work columns
    constexpr int LH_A = L_A / 2;                  // REAL == 2: first-pass tiles are two runs of LH_A bin columns
    static_assert(REAL != 1 || (FWD && LH_B * (int)sizeof(V) >= 64), "fused packed-real tiles: runs of at least 64 bytes");
    static_assert(REAL != 2 || (!FWD && LH_A * (int)sizeof(V) >= 64 && (1 << LG_N1) / 2 <= TMA_GROUP_THREADS),
                  "fused packed-real tiles: runs of at least 64 bytes, one pair of the middle column per thread");
    using Smem = TmaSmem<T, TILE>;
    static_assert(L_A * TileA::E <= Smem::TABLE_MAX, "inter-pass table");
    static_assert(!DIRECT || REAL == 0, "direct stores: complex rows only");
    constexpr int GROUP_WARPS = TMA_GROUP_THREADS / 32;
    DSC_DYN_SMEM(smem_raw);
    // the tile buffers want 1024-byte alignment (box destinations): round the dynamic window up
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u));

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < TMA_BUFFERS; ++b) {
            tma::mbar_init(&sm.full[b], 1);
            tma::mbar_init(&sm.ready[b], GROUP_WARPS);     // one arrival per warp of the group
            tma::mbar_init(&sm.empty[b], DIRECT ? GROUP_WARPS : 1);
            tma::mbar_init(&sm.posted[b], 1);
        }
        for (int i = 0; i < TMA_DONE_RING; ++i) tma::mbar_init(&sm.done[i], GROUP_WARPS);
        tma::fence_barrier_init();
    }
    static_assert(TileA::LADDER_ELEMS <= 5 * 32 && TileB::LADDER_ELEMS <= 5 * 32, "ladder rows");
    TileA::ladder_fill(sm.ladder[0], a.tw_a, tid, TMA4_THREADS);
    TileB::ladder_fill(sm.ladder[1], a.tw_b, tid, TMA4_THREADS);
    __syncthreads();

    const unsigned total = (unsigned)s.rows * (unsigned)(s.tiles_a + s.tiles_b);
    if (tid >= TMA_GROUPS * TMA_GROUP_THREADS) {
        // ------------------------------------------------------------------ the two copy threads
        const int warp = (tid - TMA_GROUPS * TMA_GROUP_THREADS) / 32;
        if ((tid & 31) != 0) return;
        const unsigned long long pol_stream = tma::policy_evict_first(), pol_keep = tma::policy_evict_last();
        constexpr int ES = sizeof(T) == 4 ? 1 : 2;          // double2 boxes are described in 8-byte elements
        if (warp == 0) {
            // ---- loader: tickets, dependencies, box loads
            // Tickets are taken two tiles ahead, and (a.prefetch) the boxes of a first-pass tile are prefetched into L2 when
            // the tile is looked up, one tile before the load that fills its buffer.
            // (A ticket held here is never awaited by a tile with a smaller ticket, so holding two cannot deadlock.)
            auto prefetch_boxes = [&](const bool pa, const unsigned prow, const unsigned pr) {
                if (!a.prefetch || !pa) return;
                constexpr int ROWS = 1 << LG_N1;
#pragma unroll
                for (int r0 = 0; r0 < ROWS; r0 += BOX_A) {
                    if constexpr (REAL == 2) {
                        tma::prefetch_3d(&map_x, (int)pr * LH_A * ES, r0, (int)prow);
                        tma::prefetch_3d(&map_x, ((1 << LG_N2) - (int)pr * LH_A - LH_A + 1) * ES, r0, (int)prow);
                    } else {
                        tma::prefetch_3d(&map_x, (int)pr * L_A * ES, r0, (int)prow);
                    }
                }
            };
            // One tile ahead of the one being issued, its ticket is decoded and its row counter read once (the value is only
            // looked at an iteration later): the two L2 round trips of a tile -- the ticket and the counter -- overlap with
            // the waits of the tile before instead of adding to every tile's issue.
            struct Pending {
                unsigned ticket, row, r, target, seen;
                bool exit, role_a;
                const unsigned *flag;
            };
            auto lookup = [&](const unsigned ticket) {
                Pending n{ticket, 0u, 0u, 0u, 0u, ticket >= total, false, nullptr};
                if (!n.exit) {
                    decode_ticket(s, ticket, n.role_a, n.row, n.r);
                    // the dependency: every first-pass tile of the row / the second pass of the row that used the
                    // work row before.  The block's own earlier tiles are published by the storers, which never
                    // wait for this thread.
                    if (!n.role_a) { n.flag = s.a_done + n.row; n.target = (unsigned)s.tiles_a; }
                    else if (s.ring && n.row >= (unsigned)s.ring) { n.flag = s.b_done + (n.row - s.ring); n.target = (unsigned)s.tiles_b; }
                    if (n.flag != nullptr) n.seen = ld_acquire(n.flag);
                    prefetch_boxes(n.role_a, n.row, n.r);
                }
                return n;
            };
            unsigned pending = atomicAdd(s.ticket, 1u);
            unsigned pending2 = atomicAdd(s.ticket, 1u);
            Pending cur = lookup(pending);
            unsigned t = 0;
            int exits_posted = 0;
            for (;; ++t) {
                const unsigned ticket = cur.ticket;
                const bool exit = cur.exit;
                const unsigned row = cur.row, r = cur.r;
                const bool role_a = cur.role_a;
                Pending nxt = cur;
                if (!exit) {
                    nxt = lookup(pending2);
                    pending2 = atomicAdd(s.ticket, 1u);      // looked up an iteration from now
                    if (cur.flag != nullptr) {
                        if (cur.seen < cur.target) {
                            const long long t0 = clock64();
                            while (ld_acquire(cur.flag) < cur.target) {
                                __nanosleep(64);
                                if (clock64() - t0 > (1LL << 32)) {
                                    printf("dsc(cuda): block %u loader stuck on the row counter of %s row %u (have %u, need %u), ticket %u\n",
                                           blockIdx.x, role_a ? "second-pass" : "first-pass", role_a ? row - s.ring : row, ld_acquire(cur.flag), cur.target, ticket);
                                    __trap();
                                }
                            }
                        }
                        tma::fence_async_all();
                    }
                }
                cur = nxt;
                // the buffer: the store of tile t - 3 has read it
                const int b = (int)(t % TMA_BUFFERS);
                if (t >= TMA_BUFFERS) tma::mbar_wait(&sm.empty[b], (t / TMA_BUFFERS - 1) & 1, "loader: empty");
                if (exit) {
                    // no more tiles: one more turn for each group and each storer, then leave.  The first TMA_GROUPS exit
                    // tiles are passed on to the storers by the groups; for the others the groups are gone and this thread
                    // arrives in their place.
                    constexpr int EXITS = DIRECT ? TMA_GROUPS : (TMA_GROUPS > TMA_BUFFERS ? TMA_GROUPS : TMA_BUFFERS);
                    sm.desc[b] = TmaTileDesc{0u, 0u, 0u, 1u};
                    tma::mbar_arrive(&sm.posted[b]);
                    tma::mbar_arrive(&sm.full[b]);
                    if (exits_posted >= TMA_GROUPS)
                        for (int i = 0; i < GROUP_WARPS; ++i) tma::mbar_arrive(&sm.ready[b]);
                    if (++exits_posted == EXITS) break;
                    continue;
                }
                sm.desc[b] = TmaTileDesc{role_a ? 1u : 0u, row, r, 0u};
                tma::mbar_arrive(&sm.posted[b]);
                tma::mbar_arrive_expect_tx(&sm.full[b], TMA_TILE_BYTES);
                if (role_a) {
                    constexpr int ROWS = 1 << LG_N1;
#pragma unroll
                    for (int r0 = 0; r0 < ROWS; r0 += BOX_A) {
                        if constexpr (REAL == 2) {
                            // map_x has n2 + 1 columns at a row pitch of n2: the mirror image of columns [0, LH) ends in column n2
                            tma::load_3d(sm.buf[b] + (size_t)r0 * LH_A * sizeof(V), &map_x, (int)r * LH_A * ES, r0, (int)row,
                                         &sm.full[b], pol_stream);
                            tma::load_3d(sm.buf[b] + (size_t)(ROWS + r0) * LH_A * sizeof(V), &map_x,
                                         ((1 << LG_N2) - (int)r * LH_A - LH_A + 1) * ES, r0, (int)row, &sm.full[b], pol_stream);
                        } else {
                            tma::load_3d(sm.buf[b] + (size_t)r0 * L_A * sizeof(V), &map_x, (int)r * L_A * ES, r0, (int)row, &sm.full[b], pol_stream);
                        }
                    }
                } else {
                    const long long wrow = a.ring ? row % a.ring : row;
                    constexpr int ROWS = 1 << LG_N2;
#pragma unroll
                    for (int r0 = 0; r0 < ROWS; r0 += BOX_B) {
                        if constexpr (REAL == 1) {
                            tma::load_3d(sm.buf[b] + (size_t)r0 * LH_B * sizeof(V), &map_w, (int)r * LH_B * ES, r0, (int)wrow,
                                         &sm.full[b], pol_stream);
                            tma::load_3d(sm.buf[b] + (size_t)(ROWS + r0) * LH_B * sizeof(V), &map_w,
                                         ((1 << LG_N1) - (int)r * LH_B - LH_B) * ES, r0, (int)wrow, &sm.full[b], pol_stream);
                        } else {
                            tma::load_3d(sm.buf[b] + (size_t)r0 * L_B * sizeof(V), &map_w, (int)r * L_B * ES, r0, (int)wrow, &sm.full[b], pol_stream);
                        }
                    }
                }
            }
        } else if constexpr (DIRECT) {
            if (warp != 1) return;
            // ---- publisher: a tile's row counter once every warp of its group has issued the tile's stores.  The warps'
            // stores happen before their arrivals, the arrivals before this wait: the release below covers them.
            int exits = 0;
            for (unsigned st = 0;; ++st) {
                const int i = (int)(st % TMA_DONE_RING);
                tma::mbar_wait(&sm.done[i], (st / TMA_DONE_RING) & 1, "publisher: done");
                const TmaTileDesc d = sm.sdesc[i];
                if (d.exit) {
                    if (++exits == TMA_GROUPS) break;
                    continue;
                }
                dsc_signal_release((d.role_a ? s.a_done : s.b_done) + d.row);
            }
        } else {
            // ---- storer of buffer warp - 1: its finished tiles leave; a tile's row counter is published once the copy has
            // completed.  Nothing here waits for another storer, the loader or a group's later tile.
            const int b = warp - 1;
            for (unsigned st = (unsigned)b;; st += TMA_BUFFERS) {
                tma::mbar_wait(&sm.ready[b], (st / TMA_BUFFERS) & 1, "storer: ready");
                const TmaTileDesc d = sm.desc[b];
                if (d.exit) break;
                if (d.role_a) {
                    // L_A contiguous lines W[q0 + l][k1] of the work row
                    const long long wrow = a.ring ? d.row % a.ring : d.row;
                    if constexpr (REAL == 2) {
                        // the two runs of work lines W[q][k1] (TmaTile::line_q); tile 0's last line is q = n2/2
                        V *wr = (V *)a.work + ((wrow << LG_N2) << LG_N1);
                        constexpr unsigned LINE = (unsigned)sizeof(V) << LG_N1;
                        const long long q_hi = (1 << LG_N2) - (long long)d.r * LH_A - LH_A + 1;
                        tma::store_linear(wr + ((long long)d.r * LH_A << LG_N1), sm.buf[b], LH_A * LINE, pol_keep);
                        if (d.r != 0) tma::store_linear(wr + (q_hi << LG_N1), sm.buf[b] + LH_A * LINE, LH_A * LINE, pol_keep);
                        else {
                            tma::store_linear(wr + (q_hi << LG_N1), sm.buf[b] + LH_A * LINE, (LH_A - 1) * LINE, pol_keep);
                            tma::store_linear(wr + ((long long)(1 << LG_N2) / 2 << LG_N1), sm.buf[b] + (2 * LH_A - 1) * LINE, LINE, pol_keep);
                        }
                    } else {
                        V *dst = (V *)a.work + ((wrow << LG_N2) + (long long)d.r * L_A << LG_N1);
                        tma::store_linear(dst, sm.buf[b], TMA_TILE_BYTES, pol_keep);
                    }
                } else {
                    constexpr int ROWS = 1 << LG_N2;
#pragma unroll
                    for (int r0 = 0; r0 < ROWS; r0 += BOX_B) {
                        if constexpr (REAL == 1) {
                            // the hi run holds k1 = column + 1 (TmaTile PERMUTE); k1 = n1 (tile 0) is clipped by the tensor
                            tma::store_3d(&map_out, (int)d.r * LH_B * ES, r0, (int)d.row, sm.buf[b] + (size_t)r0 * LH_B * sizeof(V),
                                          a.keep_out ? pol_keep : pol_stream);
                            if constexpr (sizeof(V) == 16)       // complex64: stored by the group (tma_unmix_tile)
                                tma::store_3d(&map_out, ((1 << LG_N1) - (int)d.r * LH_B - LH_B + 1) * ES, r0, (int)d.row,
                                              sm.buf[b] + (size_t)(ROWS + r0) * LH_B * sizeof(V), a.keep_out ? pol_keep : pol_stream);
                        } else {
                            tma::store_3d(&map_out, (int)d.r * L_B * ES, r0, (int)d.row, sm.buf[b] + (size_t)r0 * L_B * sizeof(V),
                                          a.keep_out ? pol_keep : pol_stream);
                        }
                    }
                }
                tma::store_commit();
                tma::store_wait_read();
                tma::mbar_arrive(&sm.empty[b]);         // the loader may refill the buffer (and overwrite desc[b])
                tma::store_wait_all();
                tma::fence_async_all();                 // async-proxy writes before the generic-proxy release below
                dsc_signal_release((d.role_a ? s.a_done : s.b_done) + d.row);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int group = tid / TMA_GROUP_THREADS, gtid = tid % TMA_GROUP_THREADS;
    const int bar_id = 1 + group;
    for (unsigned t = (unsigned)group;; t += TMA_GROUPS) {
        const int b = (int)(t % TMA_BUFFERS);
        // the descriptor arrives with the load's issue: twiddles that depend on the tile's position only are built while
        // the tile is still on its way
        tma::mbar_wait(&sm.posted[b], (t / TMA_BUFFERS) & 1, "group: posted");
        const TmaTileDesc d = sm.desc[b];
        V w0 = mk<T>((T)1, (T)0);
        if (d.role_a && !d.exit) w0 = TileA::prepare(sm.table[group], a, d.r, 1u << LG_N2, gtid, bar_id);
        V xm_a = V{}, xm_b = V{};
        if constexpr (REAL == 2) {
            // the bins of column n2/2 (tile 0 only), rows gtid and n1 - 1 - gtid: fetched while the boxes travel
            if (d.role_a && !d.exit && d.r == 0 && gtid < (1 << LG_N1) / 2) {
                const V *xr = (const V *)a.in + (long long)d.row * a.in_pitch + (1 << LG_N2) / 2;
                xm_a = __ldg(xr + ((long long)gtid << LG_N2));
                xm_b = __ldg(xr + ((long long)((1 << LG_N1) - 1 - gtid) << LG_N2));
            }
        }
        if constexpr (DIRECT) {
            if (gtid == 0) sm.sdesc[t % TMA_DONE_RING] = d;
        }
        tma::mbar_wait(&sm.full[b], (t / TMA_BUFFERS) & 1, "group: full");
        if (d.exit) {
            if constexpr (DIRECT) { if ((tid & 31) == 0) tma::mbar_arrive(&sm.done[t % TMA_DONE_RING]); }
            else if ((tid & 31) == 0) tma::mbar_arrive(&sm.ready[b]);
            break;
        }
        V *buf = reinterpret_cast<V *>(sm.buf[b]);
        if constexpr (REAL == 2) {
            if (d.role_a) tma_mix_tile<T, LG_N1, LG_N2, L_A>(buf, a, d.r, gtid, bar_id, xm_a, xm_b);
        }
#if defined(DSC_TMA_EXPERIMENTS)
        // timing experiments of DESIGN.md 4a (make DSC_TMA_EXPERIMENTS=1): the product library has no such switch
        if (!DIRECT && REAL == 0 && a.debug_skip != 0) {
            if (a.debug_skip == 2) {
                volatile V *vb = buf;
                constexpr int PTS = TMA_TILE_BYTES / (int)sizeof(V) / TMA_GROUP_THREADS;
                V v[PTS];
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
                    for (int c = 0; c < PTS; ++c) { v[c].x = vb[gtid + c * TMA_GROUP_THREADS].x; v[c].y = vb[gtid + c * TMA_GROUP_THREADS].y; }
                    dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
#pragma unroll
                    for (int c = 0; c < PTS; ++c) { vb[gtid + c * TMA_GROUP_THREADS].x = v[c].x; vb[gtid + c * TMA_GROUP_THREADS].y = v[c].y; }
                    dsc_group_barrier(bar_id, TMA_GROUP_THREADS);
                }
            }
            tma::fence_async_smem();
            __syncwarp();
            if ((tid & 31) == 0) tma::mbar_arrive(&sm.ready[b]);
            continue;
        }
#endif
        if constexpr (DIRECT) {
            if (d.role_a) {
                // this thread's run of the work line W[q0 + l_last][j_last + c TT]
                const long long wrow = a.ring ? d.row % a.ring : d.row;
                TileA::run(buf, sm.table[group], a, a.tw_a, d.r * (unsigned)L_A, gtid, bar_id, sm.ladder[0], true, w0,
                           (V *)a.work + ((wrow << LG_N2) << LG_N1), d.r * (unsigned)L_A, 1, tma::smem_u32(&sm.empty[b]));
            }
        }
        if (DIRECT && d.role_a) {}
        else if (d.role_a) TileA::run(buf, sm.table[group], a, a.tw_a, d.r * (unsigned)L_A, gtid, bar_id, sm.ladder[0], true, w0);
        else {
            constexpr int RUNS = REAL == 1 ? 2 : 1, RUN_BYTES = (REAL == 1 ? LH_B : L_B) * (int)sizeof(V);
            if constexpr (RUN_BYTES >= 128) {
                // The work-row box(es) this tile was loaded from are dead now: nobody else reads these lines, and the slot is
                // only written again a ring turn later.  Drop the (dirty) lines from L2 instead of letting them be
                // written back to HBM when they are evicted -- the ring write-back was 35 - 40 % of all DRAM writes.
                if (a.discard_work) {
                    const long long wrow = a.ring ? d.row % a.ring : d.row;
                    const char *row0 = (const char *)a.work + ((wrow << LG_N2) << LG_N1) * (long long)sizeof(V);
                    constexpr int PER_ROW = RUN_BYTES / 128, LINES = (1 << LG_N2) * PER_ROW;
#pragma unroll
                    for (int run = 0; run < RUNS; ++run) {
                        const long long col = REAL == 1 ? (run == 0 ? (long long)d.r * LH_B : (1 << LG_N1) - (long long)d.r * LH_B - LH_B)
                                                        : (long long)d.r * L_B;
                        const char *base = row0 + col * (long long)sizeof(V);
                        for (int i = gtid; i < LINES; i += TMA_GROUP_THREADS) {
                            const char *p = base + (long long)(i / PER_ROW) * ((long long)sizeof(V) << LG_N1) + (i % PER_ROW) * 128;
                            asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
                        }
                    }
                }
            }
            if constexpr (DIRECT) {
                // bins k1 + n1 (j + c TT) of the result row, k1 = r L_B + l
                TileB::template run<LG_N1>(buf, sm.table[group], a, a.tw_b, 0u, gtid, bar_id, sm.ladder[1], false, V{},
                                           (V *)a.out + (long long)d.row * a.out_pitch, d.r * (unsigned)L_B, a.keep_out,
                                           tma::smem_u32(&sm.empty[b]));
            } else {
                TileB::run(buf, sm.table[group], a, a.tw_b, 0u, gtid, bar_id, sm.ladder[1]);
            }
            if constexpr (REAL == 1) tma_unmix_tile<T, LG_N1, LG_N2, L_B>(buf, a, d.r, d.row, gtid, bar_id);
        }
        if constexpr (DIRECT) {
            __syncwarp();
            if ((tid & 31) == 0) tma::mbar_arrive(&sm.done[t % TMA_DONE_RING]);
        } else {
            tma::fence_async_smem();                    // this thread's writes of the finished tile before the bulk store
            __syncwarp();
            if ((tid & 31) == 0) tma::mbar_arrive(&sm.ready[b]);
        }
    }
}

}  // namespace dscfft

#endif  // !DSC_EMUL
