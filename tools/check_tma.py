"""Correctness + timing of the two-pass (four-step) lengths through the device-level C ABI.
usage: python tools/check_tma.py [prec] [lg ...]    (DSC_NO_TMA=1 selects the register-direct kernel for A/B runs)
Checks fft against torch.fft in float64 on a few rows and the fwd+inv round trip on all rows, then times fwd+inv."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

api = cuda_api.CudaApi(os.path.abspath(os.environ["DSC_LIB"])) if os.environ.get("DSC_LIB") else cuda_api.CudaApi()
dev = torch.device("cuda:0")


def run(lg, prec=0, total=1 << 27, reps=10, rows=None):
    n = 1 << lg
    rows = rows or max(1, total // n)
    cdt = torch.complex64 if prec == 0 else torch.complex128
    es = 8 if prec == 0 else 16
    g = torch.Generator(device=dev).manual_seed(lg)
    x = torch.view_as_complex(torch.randn(rows, n, 2, generator=g, device=dev, dtype=torch.float32 if prec == 0 else torch.float64))
    y = torch.full_like(x, float("nan"))
    z = torch.full_like(x, float("nan"))
    nb = api.plan_bytes(n, cuda_api.FFT_COMPLEX, prec)
    pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    plan = api.plan_build(n, cuda_api.FFT_COMPLEX, prec, pm.data_ptr(), nb)
    wb = api.work_bytes(plan, rows)
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    code = 2 if prec == 0 else 3

    def step():
        api.fft(plan, x.data_ptr(), code, y.data_ptr(), rows, n, 1, True, work.data_ptr(), wb)
        api.fft(plan, y.data_ptr(), code, z.data_ptr(), rows, n, 1, False, work.data_ptr(), wb)

    step()
    torch.cuda.synchronize()
    pick = sorted({0, rows // 2, rows - 1})
    want = torch.fft.fft(x[pick].to(torch.complex128))
    err = float(((y[pick].to(torch.complex128) - want).norm() / want.norm()).item())
    rt = float(((z - x).norm() / x.norm()).item())
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 2 * 2 * rows * n * es / 1e9
    print(f"prec={prec} N=2^{lg} rows={rows}: {ms:.3f} ms fwd+inv  {gb / ms * 1e3:.0f} GB/s  fft relL2={err:.2e} roundtrip={rt:.2e}"
          f"{'  NO_TMA' if os.environ.get('DSC_NO_TMA') else ''}", flush=True)
    return err, rt


if __name__ == "__main__":
    prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    lgs = [int(a) for a in sys.argv[2:]] or (list(range(15, 21)) if prec == 0 else list(range(14, 19)))
    bad = 0
    for lg in lgs:
        tol = 1e-5 if prec == 0 else 1e-12
        # a tiny batch first (fewer tiles than SMs, ring off), then the bandwidth-sized one
        for rows in (1, 3, None):
            e, r = run(lg, prec, rows=rows, reps=2 if rows else 10)
            if not (e < tol and r < tol):
                bad += 1
                print("  ^^^ FAILED", flush=True)
    sys.exit(1 if bad else 0)
