"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fft_golden.npz")

# north-star tolerances (BASELINE.json): relative L2 vs DSC's CPU FFT
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.complex64): 1e-5,
       np.dtype(np.float64): 1e-12, np.dtype(np.complex128): 1e-12}


def rel_l2(a, b) -> float:
    a = np.asarray(a)
    b = np.asarray(b)
    den = float(np.linalg.norm(b.ravel()))
    num = float(np.linalg.norm((a.astype(b.dtype) - b).ravel()))
    return num / den if den > 0 else num


def randn(rng, shape, dtype):
    x = rng.standard_normal(shape)
    if np.dtype(dtype).kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


def load_golden():
    """Yield (index, meta, arrays) for every case recorded from the reference library."""
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    for i, m in enumerate(meta):
        arrs = {k: z[f"{k}{i}"] for k in ("x", "b", "y") if f"{k}{i}" in z.files}
        yield i, m, arrs


OPS_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ops_golden.npz")


def load_ops_golden():
    """(meta, arrays) of the cases tests/golden/make_ops_golden.py recorded from the reference: cast, mixed-dtype
    arithmetic, transpose, fftfreq / rfftfreq, get_slice / set_slice."""
    z = np.load(OPS_GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    for i, m in enumerate(meta):
        yield m, {k: z[f"{k}{i}"] for k in ("x", "b", "y") if f"{k}{i}" in z.files}


def use_library(dsc, path):
    """Point the dsc_b200 binding at another build of the same C ABI (tests only: the pthread-emulated
    build where there is no GPU, the product build on the B200).  Drops the current context first."""
    dsc.shutdown()
    dsc._lib = None
    dsc.LIBDSC = path
