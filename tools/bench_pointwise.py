"""Device-resident timings of the elementwise post-processing kernels (dsc_cuda_unary / dsc_cuda_binary)
against the measured HBM copy peak.  usage: python tools/bench_pointwise.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

api = cuda_api.CudaApi()
dev = torch.device("cuda:0")
PEAK = 6538.0
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    PEAK = float(json.load(open(pk))["hbm_gbs"])


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows, cols = 65536, 4096
z = torch.randn(rows, cols, dtype=torch.complex64, device=dev)
w = torch.randn(rows, cols, dtype=torch.complex64, device=dev)
row = torch.randn(cols, dtype=torch.complex64, device=dev)
zo = torch.empty_like(z)
ro = torch.empty(rows, cols, dtype=torch.float32, device=dev)
n = rows * cols
out = []
for name, op in (("abs", 0), ("angle", 1), ("real", 2), ("conj", 4)):
    dst = zo if name == "conj" else ro
    ms = timed(lambda: api.unary(op, z.data_ptr(), cuda_api.C32, dst.data_ptr(), n))
    nbytes = n * 8 + dst.numel() * dst.element_size()
    out.append({"kernel": f"unary {name} complex64", "ms": ms, "gbs": nbytes / ms / 1e6})
for name, op, b, mode, extra in (("mul row-broadcast", 2, row, 0, 0), ("mul same shape", 2, w, 1, n * 8),
                                 ("add same shape", 0, w, 1, n * 8), ("div scalar", 3, row, 2, 0)):
    ms = timed(lambda: api.binary(op, z.data_ptr(), b.data_ptr(), zo.data_ptr(), cuda_api.C32, rows, cols, mode))
    out.append({"kernel": f"binary {name} complex64", "ms": ms, "gbs": (2 * n * 8 + extra) / ms / 1e6})
ref = torch.abs(z[:64])
assert torch.allclose(ro[:64], z[:64].real)  # last unary run with ro was `real`
for o in out:
    o["frac_of_measured_hbm"] = o["gbs"] / PEAK
    print(json.dumps(o), flush=True)
