"""Two-pass packed-real transforms with the bin-pair step fused into the TMA-fed launch (fft_tma.cuh REAL): results
against torch.fft in float64 on the same device, and timings.  Run a second time with DSC_NO_REAL_FUSE=1 for the A/B
against the separate bin-pair sweep.  usage: python tools/check_real_fuse.py [quick]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dsc_b200 import cuda_api

api = cuda_api.CudaApi()
dev = torch.device("cuda:0")
quick = "quick" in sys.argv[1:]
only_filter = "filter" in sys.argv[1:]


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rel(a, b):
    return float(torch.linalg.norm(a.to(b.dtype) - b) / torch.linalg.norm(b))


def plan_for(order, prec):
    nb = api.plan_bytes(order, cuda_api.FFT_REAL, prec)
    pm = torch.empty(nb, dtype=torch.uint8, device=dev)
    return api.plan_build(order, cuda_api.FFT_REAL, prec, pm.data_ptr(), nb), pm


def real(lg_real, prec, rows, time_it):
    rdt = torch.float32 if prec == 0 else torch.float64
    cdt = torch.complex64 if prec == 0 else torch.complex128
    nreal = 1 << lg_real
    order = nreal // 2
    g = torch.Generator(device=dev).manual_seed(lg_real * 100 + rows)
    x = torch.randn(rows, nreal, dtype=rdt, device=dev, generator=g)
    X = torch.full((rows, order + 1), float("nan"), dtype=cdt, device=dev)
    y = torch.full_like(x, float("nan"))
    plan, pm = plan_for(order, prec)
    wb = api.work_bytes(plan, rows)
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    api.rfft(plan, x.data_ptr(), X.data_ptr(), rows, nreal, 1, work.data_ptr(), wb)
    api.irfft(plan, X.data_ptr(), y.data_ptr(), rows, order + 1, 1, work.data_ptr(), wb)
    torch.cuda.synchronize()
    k = min(rows, 4)
    sel = sorted({0, rows // 2, rows - 1})[:k]
    ref = torch.fft.rfft(x[sel].double())
    e_f = rel(X[sel], ref)
    e_b = rel(y[sel], x[sel])
    msg = f"real 2^{lg_real} prec={prec} rows={rows}: rfft relL2={e_f:.2e} roundtrip={e_b:.2e}"
    if time_it:
        t_r = timed(lambda: api.rfft(plan, x.data_ptr(), X.data_ptr(), rows, nreal, 1, work.data_ptr(), wb))
        t_i = timed(lambda: api.irfft(plan, X.data_ptr(), y.data_ptr(), rows, order + 1, 1, work.data_ptr(), wb))
        nbytes = x.numel() * x.element_size() + X.numel() * X.element_size()
        msg += f" | rfft {t_r:.3f} ms {nbytes / t_r / 1e6:.0f} GB/s | irfft {t_i:.3f} ms {nbytes / t_i / 1e6:.0f} GB/s"
    print(msg, flush=True)
    tol = 2e-6 if prec == 0 else 1e-14
    return e_f < tol and e_b < tol


def filt(lg_real, rows, time_it, prec=0):
    nreal = 1 << lg_real
    order = nreal // 2
    g = torch.Generator(device=dev).manual_seed(lg_real * 100 + rows + 7)
    x = torch.randn(rows, nreal, dtype=torch.float32 if prec == 0 else torch.float64, device=dev, generator=g)
    B = torch.randn(order + 1, dtype=torch.complex64 if prec == 0 else torch.complex128, device=dev, generator=g)
    y = torch.full_like(x, float("nan"))
    plan, pm = plan_for(order, prec)
    wb = api.filter_work_bytes(plan, rows)
    work = torch.empty(max(wb, 16), dtype=torch.uint8, device=dev)
    api.filter(plan, x.data_ptr(), B.data_ptr(), y.data_ptr(), rows, nreal, work.data_ptr(), wb)
    torch.cuda.synchronize()
    sel = sorted({0, rows // 2, rows - 1})
    Bd = B.to(torch.complex128).clone()
    Bd[0] = complex(Bd[0].real.item(), 0.0)
    Bd[-1] = complex(Bd[-1].real.item(), 0.0)
    ref = torch.fft.irfft(torch.fft.rfft(x[sel].double()) * Bd, n=nreal)
    e = rel(y[sel], ref)
    msg = f"filter 2^{lg_real} prec={prec} rows={rows}: relL2={e:.2e}"
    if time_it:
        t = timed(lambda: api.filter(plan, x.data_ptr(), B.data_ptr(), y.data_ptr(), rows, nreal, work.data_ptr(), wb))
        msg += f" | {t:.3f} ms {2 * x.numel() * x.element_size() / t / 1e6:.0f} GB/s"
    print(msg, flush=True)
    return e < (3e-6 if prec == 0 else 1e-14)


if len(sys.argv) > 5 and sys.argv[1] == "case":      # case filt|real LG PREC ROWS
    fn = filt if sys.argv[2] == "filt" else real
    a = (int(sys.argv[3]), int(sys.argv[5]), False, int(sys.argv[4])) if fn is filt else (int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), False)
    sys.exit(0 if fn(*a) else 1)

ok = True
for lg in (() if only_filter else (15, 16, 17, 18, 19)):                 # double orders 2^14 .. 2^18
    for rows in (1, 3, 37):
        ok &= real(lg, 1, rows, False)
    if not quick:
        ok &= real(lg, 1, (1 << 30) // (8 << lg), True)
for lg in (() if only_filter else (16, 18, 20)):                         # float: rfft rows of odd pitch keep the separate sweep
    ok &= real(lg, 0, 5, False)
for lg in (16, 17, 18, 19, 20, 21):             # float orders 2^15 .. 2^20 (2^20: not fused)
    for rows in (1, 3, 37):
        ok &= filt(lg, rows, False)
    if not quick:
        ok &= filt(lg, (1 << 30) // (4 << lg), True)
print("OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
