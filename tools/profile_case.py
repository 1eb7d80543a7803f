"""Tiny ncu target: a few forward transforms of one shape through the device-level C ABI.
usage: python tools/profile_case.py LG_N ROWS [prec] [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

lg, rows = int(sys.argv[1]), int(sys.argv[2])
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
n = 1 << lg
api = cuda_api.CudaApi()
dev = torch.device("cuda:0")
cdt = torch.complex64 if prec == 0 else torch.complex128
x = torch.randn(rows, n, dtype=cdt, device=dev)
y = torch.empty_like(x)
nb = api.plan_bytes(n, cuda_api.FFT_COMPLEX, prec)
pm = torch.empty(nb, dtype=torch.uint8, device=dev)
plan = api.plan_build(n, cuda_api.FFT_COMPLEX, prec, pm.data_ptr(), nb)
wb = api.work_bytes(plan, rows)
work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
code = 2 if prec == 0 else 3
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    if i == reps - 1:
        e0.record()
    api.fft(plan, x.data_ptr(), code, y.data_ptr(), rows, n, 1, True, work.data_ptr(), wb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"N=2^{lg} rows={rows} prec={prec}: {ms:.3f} ms, {2 * rows * n * (8 if prec == 0 else 16) / ms / 1e6:.0f} GB/s, work={wb >> 20} MiB")
