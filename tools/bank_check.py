"""Shared-memory bank-conflict model of fft_lines<> (8-byte and 16-byte elements).
For every (lg_n, LPB, mapping) reports the worst conflict degree of the stage scatters and reads."""
import sys

def check(lg_n, lg_e_max, lpb, elem_bytes, strided):
    N = 1 << lg_n
    lg_e = min(lg_e_max, lg_n)
    E = 1 << lg_e
    TT = N // E
    stages = 1 if lg_e == 0 else -(-lg_n // lg_e)
    phase = 128 // elem_bytes
    base = N + (N >> lg_e)
    want = 1 if lpb >= phase else phase // lpb
    LINE = base + ((want - base % phase) % phase + phase) % phase
    pad = lambda i: i + (i >> lg_e)
    threads = lpb * TT
    def lt(tid):
        return (tid % lpb, tid // lpb) if strided else (tid // TT, tid % TT)
    worst = 1
    words = elem_bytes // 4
    def degree(addrs):           # addrs: element addresses of one phase of lanes
        banks = {}
        for a in set(addrs):
            b = (a * words) % 32
            banks[b] = banks.get(b, 0) + 1
        return max(banks.values())
    for s in range(stages):
        lg_r = min(lg_e, lg_n - s * lg_e)
        R = 1 << lg_r; NB = E // R; NS = 1 << (s * lg_e)
        last = s == stages - 1
        for w0 in range(0, min(threads, 256), phase):
            lanes = [lt(tid) for tid in range(w0, min(w0 + phase, threads))]
            if not last:
                for b in range(NB):
                    for p in range(R):
                        addrs = []
                        for (l, t) in lanes:
                            j = t + b * TT; k = j & (NS - 1)
                            addrs.append(l * LINE + pad(((j - k) << lg_r) + k + p * NS))
                        worst = max(worst, degree(addrs))
                for c in range(E):
                    worst = max(worst, degree([l * LINE + pad(t + c * TT) for (l, t) in lanes]))
    return worst

for name, lg_e, eb in (("float", 4, 8), ("double", 3, 16)):
    for lg_n in range(4, 15 if eb == 8 else 14):
        tt = 1 << (lg_n - min(lg_e, lg_n))
        lpb_c = max(1, 256 // tt)
        coal = 16 if eb == 8 else 8
        lpb_s = max(1, min(max(coal, 256 // tt), 512 // tt))
        row = [f"{name} N=2^{lg_n}: contiguous LPB={lpb_c} -> {check(lg_n, lg_e, lpb_c, eb, False)}-way",
               f"strided LPB={lpb_s} -> {check(lg_n, lg_e, lpb_s, eb, True)}-way"]
        for thr in (256, 512):
            if tt <= thr:
                row.append(f"fused {thr}thr LPB={thr // tt} -> {check(lg_n, lg_e, thr // tt, eb, True)}-way")
        print("; ".join(row))
