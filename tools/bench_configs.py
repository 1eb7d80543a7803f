"""Device-resident timings of the non-headline BASELINE configs through the drop-in tensor C ABI
(libdsc.so, residency mode 2 so nothing crosses PCIe inside the timed region).  One JSON line each.
  config 1  README filterFFT, 8192 samples, 128 taps, fft_size 16384 (latency per call, fused and unfused)
  config 3  float64 rfft/irfft, ROWS x 262144  (BASELINE: 4096 rows; default here 1024 rows)
  config 4  fused filter rfft(s)*B -> irfft, float32 channels of 2^20 samples (BASELINE: 16384; default 256)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsc_b200 as dsc
from oracle import port

PEAK = 6538.0
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    rows3 = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ch4 = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dsc.init(40 << 30, 2 << 30)
    dsc.set_residency(2)

    # ---- config 1
    s = np.random.default_rng(0).standard_normal(8192).astype(np.float32)
    b = np.random.default_rng(1).standard_normal(128).astype(np.float32)
    ts, tb = dsc.from_numpy(s), dsc.from_numpy(b)
    B = dsc.rfft(tb, n=16384)
    keep = {}

    def unfused():
        keep["y"] = dsc.irfft(dsc.rfft(ts, n=16384) * B)

    def fused():
        keep["y"] = dsc.fft_filter(ts, B, n=16384)
    t_un = timed(unfused, 50)
    t_fu = timed(fused, 200)
    err = rel(keep["y"].numpy(), port.filter_fft(s, b, 16384))
    print(json.dumps({"config": 1, "workload": "README filterFFT 8192 x 128 taps, fft 16384", "us_per_call_fused": t_fu * 1e6,
                      "us_per_call_three_calls": t_un * 1e6, "rel_l2_vs_oracle": err}), flush=True)
    keep.clear()

    # ---- config 3
    n = 262144
    x = np.random.default_rng(3).standard_normal((rows3, n))
    tx = dsc.from_numpy(x)
    X = dsc.rfft(tx)

    def c3():
        keep["X"] = dsc.rfft(tx)
        keep["z"] = dsc.irfft(keep["X"])
    t3 = timed(c3, 5)
    bytes3 = 2 * rows3 * (8 * n + 16 * (n // 2 + 1))
    flops3 = 2 * rows3 * 2.5 * n * 18
    err3 = rel(keep["z"].numpy()[:4], x[:4])
    err3o = rel(keep["X"].numpy()[:2], port.rfft(x[:2]))
    print(json.dumps({"config": 3, "workload": f"float64 rfft+irfft {rows3} x {n}", "ms": t3 * 1e3, "gflops": flops3 / t3 / 1e9,
                      "algorithmic_gbs": bytes3 / t3 / 1e9, "frac_of_measured_hbm": bytes3 / t3 / 1e9 / PEAK,
                      "roundtrip_rel_l2": err3, "rel_l2_vs_oracle": err3o}), flush=True)
    keep.clear()
    del tx, X

    # ---- config 4
    n = 1 << 20
    sig = np.random.default_rng(4).standard_normal((ch4, n)).astype(np.float32)
    taps = np.zeros(n, np.float32)
    taps[:128] = np.random.default_rng(5).standard_normal(128).astype(np.float32)
    tsig = dsc.from_numpy(sig)
    Bt = dsc.rfft(dsc.from_numpy(taps))

    def c4():
        keep["y"] = dsc.fft_filter(tsig, Bt)
    t4 = timed(c4, 5)
    bytes4 = ch4 * 8 * n
    flops4 = ch4 * 2 * 2.5 * n * 20
    err4 = rel(keep["y"].numpy()[0], port.filter_fft(sig[0], taps, n))
    print(json.dumps({"config": 4, "workload": f"fused filter float32 {ch4} x 2^20", "ms": t4 * 1e3, "gflops": flops4 / t4 / 1e9,
                      "algorithmic_gbs": bytes4 / t4 / 1e9, "frac_of_measured_hbm": bytes4 / t4 / 1e9 / PEAK,
                      "rel_l2_vs_oracle": err4}), flush=True)
    keep.clear()
    dsc.shutdown()


if __name__ == "__main__":
    main()
