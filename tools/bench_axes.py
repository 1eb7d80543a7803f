"""Device-resident timings of transforms along NON-last axes through the device-level C ABI.
usage: python tools/bench_axes.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

api = cuda_api.CudaApi()
dev = torch.device("cuda:0")


def run(outer, n, inner, prec=0, reps=5):
    cdt = torch.complex64 if prec == 0 else torch.complex128
    x = torch.randn(outer, n, inner, dtype=cdt, device=dev)
    y = torch.empty_like(x)
    nb = api.plan_bytes(n, cuda_api.FFT_COMPLEX, prec)
    pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    plan = api.plan_build(n, cuda_api.FFT_COMPLEX, prec, pm.data_ptr(), nb)
    code = 2 if prec == 0 else 3
    wb = api.work_bytes_axis(plan, outer, inner)
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(reps):
        if i == 1:
            e0.record()
        api.fft(plan, x.data_ptr(), code, y.data_ptr(), outer, n, inner, True, work.data_ptr(), wb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps - 1)
    ref = torch.fft.fft(x[:1, :, :64].to(torch.complex128), dim=1)
    err = float(torch.linalg.norm(y[:1, :, :64].to(torch.complex128) - ref) / torch.linalg.norm(ref))
    nbytes = 2 * x.numel() * x.element_size()
    print(f"({outer},{n},{inner}) axis=1 prec={prec}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s  rel={err:.1e}", flush=True)


if __name__ != "__main__":
    for shape in ([(1, 4096, 32768), (32, 4096, 1024), (1, 1024, 131072), (128, 1024, 1024), (1, 2048, 65536)] if os.environ.get("AXES_SINGLE")
                  else [(1, 32768, 4096), (4, 65536, 512), (16, 32768, 256), (2, 65536, 1024)]):
        run(*shape)
    if os.environ.get("AXES_SINGLE"):
        run(1, 4096, 16384, 1)
        run(64, 1024, 1024, 1)
    sys.exit(0)
for shape in [(1, 4096, 32768), (32, 4096, 1024), (1, 1024, 131072), (128, 1024, 1024), (2048, 256, 256),
              (1, 64, 2097152), (1, 8192, 16384), (1, 16384, 8192), (32768, 64, 64), (1048576, 16, 8), (4096, 4096, 3)]:
    run(*shape)
for shape in [(1, 32768, 4096), (4, 65536, 512), (1, 1 << 20, 128), (16, 32768, 256)]:
    run(*shape)
run(1, 4096, 16384, 1)
run(1, 16384, 4096, 1)
run(64, 1024, 1024, 1)
