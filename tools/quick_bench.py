"""Scratch timing of the device-level entry points (not the contract bench; see bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

api = cuda_api.CudaApi()
dev = torch.device("cuda:0")


def bench(lg, total_elems, prec=0, reps=10, real=False):
    n = 1 << lg
    rows = max(1, total_elems // n)
    cdt = torch.complex64 if prec == 0 else torch.complex128
    es = 8 if prec == 0 else 16
    x = torch.randn(rows, n, dtype=cdt, device=dev)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    nb = api.plan_bytes(n, cuda_api.FFT_COMPLEX, prec)
    pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    plan = api.plan_build(n, cuda_api.FFT_COMPLEX, prec, pm.data_ptr(), nb)
    wb = api.work_bytes(plan, min(rows, max(1, (64 << 20) // (n * es))))
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    code = 2 if prec == 0 else 3

    def step():
        api.fft(plan, x.data_ptr(), code, y.data_ptr(), rows, n, 1, True, work.data_ptr(), wb)
        api.fft(plan, y.data_ptr(), code, z.data_ptr(), rows, n, 1, False, work.data_ptr(), wb)

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
    gb = 2 * 2 * rows * n * es / 1e9
    gflop = 2 * rows * 5 * n * lg / 1e9
    print(f"prec={prec} N=2^{lg} rows={rows}: {ms:.3f} ms/step(fwd+inv)  {gb / ms * 1e3:.0f} GB/s  "
          f"{gflop / ms:.0f} GFLOP/s  roundtrip relL2={err:.2e}", flush=True)


if __name__ == "__main__":
    total = 1 << 27   # 1 GiB of complex64
    for lg in [12] + list(range(4, 21)):
        bench(lg, total)
    for lg in (10, 12, 13, 17):
        bench(lg, 1 << 26, prec=1)
