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
CASES = [("complex64", 14, 300), ("complex64", 15, 1500), ("complex64", 15, 1), ("complex64", 16, 9), ("complex64", 16, 700),
         ("complex64", 17, 3), ("complex64", 17, 300), ("complex64", 18, 5), ("complex64", 19, 3), ("complex64", 20, 3),
         ("complex64", 20, 40), ("complex128", 14, 1100), ("complex128", 15, 2), ("complex128", 16, 70), ("complex128", 17, 3),
         ("complex128", 18, 2), ("complex128", 19, 2)]


def main():
    dev = DevFFT(cuda_api.LIBDSC, backend="torch")
    bad = 0
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
