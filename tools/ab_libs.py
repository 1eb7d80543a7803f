"""Same-box A/B of two builds of libdsc.so on the complex64 sweep (device-level C ABI, CUDA events).
usage: python tools/ab_libs.py LIB_A LIB_B [lg ...]   (environment knobs apply to both)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

dev = torch.device("cuda:0")
libs = [cuda_api.CudaApi(os.path.abspath(p)) for p in sys.argv[1:3]]
lgs = [int(v) for v in sys.argv[3:]] or list(range(10, 21))


def bench(api, lg, total=1 << 27, reps=10):
    n = 1 << lg
    rows = total // n
    x = torch.randn(rows, n, dtype=torch.complex64, device=dev)
    y, z = torch.empty_like(x), torch.empty_like(x)
    nb = api.plan_bytes(n, cuda_api.FFT_COMPLEX, 0)
    pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    plan = api.plan_build(n, cuda_api.FFT_COMPLEX, 0, pm.data_ptr(), nb)
    wb = api.work_bytes(plan, rows)
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)

    def step():
        api.fft(plan, x.data_ptr(), 2, y.data_ptr(), rows, n, 1, True, work.data_ptr(), wb)
        api.fft(plan, y.data_ptr(), 2, z.data_ptr(), rows, n, 1, False, work.data_ptr(), wb)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    err = ((z - x).norm() / x.norm()).item()
    return 4 * rows * n * 8 / ms / 1e6, err


for lg in lgs:
    res = []
    for rep in range(2):
        for api in libs:
            res.append(bench(api, lg))
    a = max(res[0][0], res[2][0])
    b = max(res[1][0], res[3][0])
    print(f"2^{lg}: A {a:.0f} GB/s  B {b:.0f} GB/s  ({(b / a - 1) * 100:+.1f} %)  roundtrip {res[0][1]:.2e} / {res[1][1]:.2e}", flush=True)
