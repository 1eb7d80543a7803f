"""Subprocess body of tests/test_gpu_two_pass_paths.py: lengths beyond one shared-memory pass against the oracle, through
whichever kernel the environment selects (DSC_NO_TMA / DSC_NO_CLUSTER / DSC_CLUSTER_LGS are read once per process)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsc_b200 import cuda_api  # noqa: E402
from oracle import port  # noqa: E402
from tests.devfft import DevFFT  # noqa: E402
from tests.util import randn, rel_l2  # noqa: E402

TIGHT = {"complex64": 2e-6, "complex128": 5e-15}
# (2^14 complex64 / 2^13 complex128 and the real transforms of twice those lengths are the longest single-pass lines: one
# block per SM, persistent blocks that prefetch their next line into L2 once there are more lines than SMs; DSC_NO_PERSIST=1
# selects the one-shot launch)
CASES = [("complex64", 14, 300), ("complex128", 13, 300), ("complex64", 15, 1500), ("complex64", 15, 1), ("complex64", 16, 9), ("complex64", 16, 700),
         ("complex64", 17, 3), ("complex64", 17, 300), ("complex64", 18, 5), ("complex64", 19, 3), ("complex64", 20, 3),
         ("complex64", 20, 40), ("complex128", 14, 1100), ("complex128", 15, 2), ("complex128", 16, 70), ("complex128", 17, 3),
         ("complex128", 18, 2), ("complex128", 19, 2)]
# packed-real transforms of two-pass orders: (real dtype, log2 of the real length, rows).  float64 rows run with the bin-pair
# step fused into the TMA-fed launch (DSC_NO_REAL_FUSE=1: the separate sweep); float32 rfft rows have an odd pitch and keep
# the sweep, the float32 filter is fused up to 2^20 samples
REAL_CASES = [("float32", 15, 300), ("float64", 14, 300), ("float64", 15, 300), ("float64", 15, 1), ("float64", 16, 2), ("float64", 17, 70), ("float64", 18, 3), ("float64", 18, 41),
              ("float64", 19, 2), ("float32", 16, 5), ("float32", 18, 3), ("float32", 20, 2)]
FILTER_CASES = [("float32", 16, 1), ("float32", 16, 300), ("float32", 17, 3), ("float32", 18, 70), ("float32", 19, 2), ("float32", 20, 37),
                ("float32", 21, 2), ("float64", 16, 3), ("float64", 18, 35), ("float64", 19, 2)]
TIGHT_REAL = {"float32": 2e-6, "float64": 5e-15}


def main():
    dev = DevFFT(cuda_api.LIBDSC, backend="torch")
    bad = 0
    for dtype, lg, rows in REAL_CASES:
        rng = np.random.default_rng(lg * 1000 + rows + 1)
        x = randn(rng, (rows, 1 << lg), dtype)
        X = dev.rfft(x)
        sample = sorted({0, rows // 2, rows - 1})
        Xo = port.rfft(x[sample])
        e_f = rel_l2(X[sample], Xo)
        e_i = rel_l2(dev.irfft(X)[sample], port.irfft(Xo))
        e_b = rel_l2(dev.irfft(X), x)
        ok = max(e_f, e_i, e_b) < TIGHT_REAL[dtype]
        print(f"{dtype} rfft/irfft 2^{lg} x {rows}: rfft vs oracle {e_f:.2e}, irfft vs oracle {e_i:.2e}, round trip {e_b:.2e} "
              f"{'ok' if ok else 'FAILED'}", flush=True)
        bad += not ok
    for dtype, lg, rows in FILTER_CASES:
        rng = np.random.default_rng(lg * 1000 + rows + 2)
        x = randn(rng, (rows, 1 << lg), dtype)
        B = port.rfft(randn(rng, (1 << lg,), dtype))
        y = dev.filter(x, B)
        sample = sorted({0, rows // 2, rows - 1})
        ref = port.irfft(port.rfft(x[sample]) * B)
        e = rel_l2(y[sample], ref)
        ok = e < TIGHT_REAL[dtype]
        print(f"{dtype} filter 2^{lg} x {rows}: vs oracle {e:.2e} {'ok' if ok else 'FAILED'}", flush=True)
        bad += not ok
    for dtype, lg, rows in CASES:
        rng = np.random.default_rng(lg * 1000 + rows)
        x = randn(rng, (rows, 1 << lg), dtype)
        y = dev.fft(x)
        sample = sorted({0, rows // 2, rows - 1})
        e_f = rel_l2(y[sample], port.fft(x[sample]))
        e_b = rel_l2(dev.ifft(y), x)
        ok = e_f < TIGHT[dtype] and e_b < TIGHT[dtype]
        print(f"{dtype} 2^{lg} x {rows}: fft vs oracle {e_f:.2e}, round trip {e_b:.2e} {'ok' if ok else 'FAILED'}", flush=True)
        bad += not ok
    print("ALL OK" if not bad else f"{bad} FAILED")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
