"""Device-resident timings of the generic (predicated) last-axis kernels: real-input fft and zero-padded complex input.
usage: python tools/bench_generic.py [prec]"""
import os, sys
sys.path.insert(0, "/root/repo")
import torch
from dsc_b200 import cuda_api
api = cuda_api.CudaApi(); dev = torch.device("cuda:0")
def timed(fn, reps=6):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
PREC = int(sys.argv[1]) if len(sys.argv) > 1 else 0
RDT, CDT = (torch.float32, torch.complex64) if PREC == 0 else (torch.float64, torch.complex128)
for lg in ((10, 12, 13) if PREC == 0 else (9, 10, 11)):
    n = 1 << lg
    rows = (1 << (27 - PREC)) // n
    nb = api.plan_bytes(n, cuda_api.FFT_COMPLEX, PREC); pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    plan = api.plan_build(n, cuda_api.FFT_COMPLEX, PREC, pm.data_ptr(), nb)
    xr = torch.randn(rows, n, dtype=RDT, device=dev)            # real input cast
    xp = torch.randn(rows, n - 100, dtype=CDT, device=dev)    # zero-padded complex
    y = torch.empty(rows, n, dtype=CDT, device=dev)
    t1 = timed(lambda: api.fft(plan, xr.data_ptr(), (cuda_api.F32 if PREC == 0 else cuda_api.F64), y.data_ptr(), rows, n, 1, True))
    t2 = timed(lambda: api.fft(plan, xp.data_ptr(), (cuda_api.C32 if PREC == 0 else cuda_api.C64), y.data_ptr(), rows, n - 100, 1, True))
    print(f"2^{lg}: real-input fft {t1:.3f} ms {(xr.numel()*xr.element_size() + y.numel()*y.element_size())/t1/1e6:.0f} GB/s | padded complex {t2:.3f} ms {(xp.numel()*xp.element_size() + y.numel()*y.element_size())/t2/1e6:.0f} GB/s", flush=True)
