"""Device-resident timings of single-pass packed-real transforms (rfft / irfft / fused filter) through the
device-level C ABI.  usage: python tools/bench_real.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

api = cuda_api.CudaApi()
dev = torch.device("cuda:0")


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(lg_real, prec=0, total_bytes=1 << 30):
    rdt = torch.float32 if prec == 0 else torch.float64
    cdt = torch.complex64 if prec == 0 else torch.complex128
    nreal = 1 << lg_real
    order = nreal // 2
    rows = total_bytes // (nreal * (4 if prec == 0 else 8))
    x = torch.randn(rows, nreal, dtype=rdt, device=dev)
    X = torch.empty(rows, order + 1, dtype=cdt, device=dev)
    y = torch.empty_like(x)
    nb = api.plan_bytes(order, cuda_api.FFT_REAL, prec)
    pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    plan = api.plan_build(order, cuda_api.FFT_REAL, prec, pm.data_ptr(), nb)
    wb = api.work_bytes(plan, rows)
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    t_r = timed(lambda: api.rfft(plan, x.data_ptr(), X.data_ptr(), rows, nreal, 1, work.data_ptr(), wb))
    t_i = timed(lambda: api.irfft(plan, X.data_ptr(), y.data_ptr(), rows, order + 1, 1, work.data_ptr(), wb))
    err = float(torch.linalg.norm(y[:8] - x[:8]) / torch.linalg.norm(x[:8]))
    nbytes = x.numel() * x.element_size() + X.numel() * X.element_size()
    print(f"real 2^{lg_real} prec={prec} rows={rows}: rfft {t_r:.3f} ms {nbytes / t_r / 1e6:.0f} GB/s | "
          f"irfft {t_i:.3f} ms {nbytes / t_i / 1e6:.0f} GB/s | roundtrip {err:.1e}", flush=True)


if __name__ != "__main__":          # tools/bench_real_ab.py: the one-block-per-SM lengths only
    run(14); run(15); run(13, 1); run(14, 1)
    sys.exit(0)
for lg in (8, 10, 11, 12, 13, 14, 15, 16, 18, 20):
    run(lg)
for lg in (10, 12, 13, 14, 16, 18):
    run(lg, 1)


def run_filter(lg_real, total_bytes=1 << 30):
    nreal = 1 << lg_real
    order = nreal // 2
    rows = total_bytes // (nreal * 4)
    x = torch.randn(rows, nreal, dtype=torch.float32, device=dev)
    B = torch.randn(order + 1, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    nb = api.plan_bytes(order, cuda_api.FFT_REAL, 0)
    pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    plan = api.plan_build(order, cuda_api.FFT_REAL, 0, pm.data_ptr(), nb)
    wb = api.filter_work_bytes(plan, rows)
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    t = timed(lambda: api.filter(plan, x.data_ptr(), B.data_ptr(), y.data_ptr(), rows, nreal, work.data_ptr(), wb))
    print(f"filter 2^{lg_real} float32 rows={rows}: {t:.3f} ms {2 * x.numel() * 4 / t / 1e6:.0f} GB/s", flush=True)


for lg in (10, 12, 13, 14, 15):
    run_filter(lg)
