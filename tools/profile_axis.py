"""Tiny ncu target: a few transforms along the middle axis of (outer, n, inner).  usage: python tools/profile_axis.py OUTER LG_N INNER [prec]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

outer, lg, inner = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n = 1 << lg
api = cuda_api.CudaApi()
dev = torch.device("cuda:0")
cdt = torch.complex64 if prec == 0 else torch.complex128
x = torch.randn(outer, n, inner, dtype=cdt, device=dev)
y = torch.empty_like(x)
nb = api.plan_bytes(n, cuda_api.FFT_COMPLEX, prec)
pm = torch.empty(nb, dtype=torch.uint8, device=dev)
plan = api.plan_build(n, cuda_api.FFT_COMPLEX, prec, pm.data_ptr(), nb)
wb = api.work_bytes_axis(plan, outer, inner)
work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
for _ in range(4):
    api.fft(plan, x.data_ptr(), 2 if prec == 0 else 3, y.data_ptr(), outer, n, inner, True, work.data_ptr(), wb)
torch.cuda.synchronize()
print("ok")
