"""Tiny ncu target: rfft + irfft of one shape through the device-level C ABI.  usage: python tools/profile_real.py LG_REAL ROWS [prec]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

lg, rows = int(sys.argv[1]), int(sys.argv[2])
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
api = cuda_api.CudaApi()
dev = torch.device("cuda:0")
rdt = torch.float32 if prec == 0 else torch.float64
cdt = torch.complex64 if prec == 0 else torch.complex128
nreal = 1 << lg
order = nreal // 2
x = torch.randn(rows, nreal, dtype=rdt, device=dev)
X = torch.empty(rows, order + 1, dtype=cdt, device=dev)
y = torch.empty_like(x)
nb = api.plan_bytes(order, cuda_api.FFT_REAL, prec)
pm = torch.empty(nb, dtype=torch.uint8, device=dev)
plan = api.plan_build(order, cuda_api.FFT_REAL, prec, pm.data_ptr(), nb)
wb = api.work_bytes(plan, rows)
work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
for _ in range(3):
    api.rfft(plan, x.data_ptr(), X.data_ptr(), rows, nreal, 1, work.data_ptr(), wb)
    api.irfft(plan, X.data_ptr(), y.data_ptr(), rows, order + 1, 1, work.data_ptr(), wb)
torch.cuda.synchronize()
print("ok")
