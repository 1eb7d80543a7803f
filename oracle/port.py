"""numpy-facing wrapper of the C restatement (oracle/libdsc_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never by dsc_b200/.

Same call shapes as the reference's Python API (python/dsc/tensor.py:684-726):
fft / ifft / rfft / irfft (x, n=-1, axis=-1) on numpy arrays of 1..4 dims.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libdsc_oracle.so")

_CODE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1,
         np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}
_COMPLEX_OF = {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128,
               np.dtype(np.complex64): np.complex64, np.dtype(np.complex128): np.complex128}
_REAL_OF = {np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}

_lib = None


def build() -> str:
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
    return PORT_SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.dsco_pow2_n.argtypes = [C.c_int]
        L.dsco_axis_index.argtypes = [C.c_int, C.c_int]
        L.dsco_fft_len.argtypes = [C.c_int, C.c_int]
        ip = C.POINTER(C.c_int)
        L.dsco_rfft_len.argtypes = [C.c_int, C.c_int, ip, ip]
        L.dsco_irfft_len.argtypes = [C.c_int, C.c_int, ip, ip]
        L.dsco_twiddle_count.restype = C.c_size_t
        L.dsco_twiddle_count.argtypes = [C.c_int, C.c_int]
        for suf in ("f32", "f64"):
            getattr(L, f"dsco_twiddles_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_int]
            getattr(L, f"dsco_cfft_line_{suf}").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            getattr(L, f"dsco_rfft_line_{suf}").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
            getattr(L, f"dsco_cmul_{suf}").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int]
        L.dsco_fft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_long, C.c_int, C.c_int]
        L.dsco_rfft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_long, C.c_int]
        L.dsco_irfft.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_long, C.c_int]
        _lib = L
    return _lib


# -- shape rules ---------------------------------------------------------------------
def pow2_n(n: int) -> int:
    return lib().dsco_pow2_n(n)


def fft_len(x_n: int, n: int = -1) -> int:
    return lib().dsco_fft_len(x_n, n)


def rfft_len(x_n: int, n: int = -1):
    o, m = C.c_int(), C.c_int()
    if lib().dsco_rfft_len(x_n, n, C.byref(o), C.byref(m)) != 0:
        raise ValueError("reference aborts: rfft order 0")
    return o.value, m.value


def irfft_len(x_n: int, n: int = -1):
    o, m = C.c_int(), C.c_int()
    if lib().dsco_irfft_len(x_n, n, C.byref(o), C.byref(m)) != 0:
        raise ValueError("reference aborts: irfft order 0")
    return o.value, m.value


def twiddles(n: int, real_plan: bool, np_real) -> np.ndarray:
    cnt = lib().dsco_twiddle_count(n, int(real_plan))
    tw = np.empty(cnt, dtype=np_real)
    suf = "f32" if tw.dtype == np.float32 else "f64"
    getattr(lib(), f"dsco_twiddles_{suf}")(tw.ctypes.data, n, int(real_plan))
    return tw


def _split(x: np.ndarray, axis: int):
    """(outer, x_n, inner) of a contiguous array and the right-aligned axis rule of dsc.h:81."""
    assert 1 <= x.ndim <= 4
    ax = lib().dsco_axis_index(x.ndim, axis) - (4 - x.ndim)
    if not 0 <= ax < x.ndim:
        raise ValueError("axis out of range")
    outer = int(np.prod(x.shape[:ax], dtype=np.int64))
    inner = int(np.prod(x.shape[ax + 1:], dtype=np.int64))
    return ax, outer, x.shape[ax], inner


def _cfft(x, n, axis, forward):
    x = np.ascontiguousarray(x)
    ax, outer, x_n, inner = _split(x, axis)
    fft_n = fft_len(x_n, n)
    shape = list(x.shape)
    shape[ax] = fft_n
    out = np.empty(shape, dtype=_COMPLEX_OF[x.dtype])
    rc = lib().dsco_fft(x.ctypes.data, _CODE[x.dtype], out.ctypes.data, outer, x_n, inner, n, int(forward))
    if rc != 0:
        raise ValueError("dsco_fft failed")
    return out


def fft(x, n=-1, axis=-1):
    return _cfft(x, n, axis, True)


def ifft(x, n=-1, axis=-1):
    return _cfft(x, n, axis, False)


def rfft(x, n=-1, axis=-1):
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.float32, np.float64):
        raise ValueError("RFFT input must be real")
    ax, outer, x_n, inner = _split(x, axis)
    _, out_n = rfft_len(x_n, n)
    shape = list(x.shape)
    shape[ax] = out_n
    out = np.empty(shape, dtype=_COMPLEX_OF[x.dtype])
    if lib().dsco_rfft(x.ctypes.data, _CODE[x.dtype], out.ctypes.data, outer, x_n, inner, n) != 0:
        raise ValueError("dsco_rfft failed")
    return out


def irfft(x, n=-1, axis=-1):
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.complex64, np.complex128):
        raise ValueError("IRFFT input must be complex")
    ax, outer, x_n, inner = _split(x, axis)
    _, out_n = irfft_len(x_n, n)
    shape = list(x.shape)
    shape[ax] = out_n
    out = np.empty(shape, dtype=_REAL_OF[x.dtype])
    if lib().dsco_irfft(x.ctypes.data, _CODE[x.dtype], out.ctypes.data, outer, x_n, inner, n) != 0:
        raise ValueError("dsco_irfft failed")
    return out


def cmul(a, b):
    """a: (rows, cols) complex, b: (cols,) or (rows, cols) of the same dtype."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    a2 = a.reshape(-1, a.shape[-1])
    out = np.empty_like(a2)
    suf = "f32" if a.dtype == np.complex64 else "f64"
    b_rows = int(b.size == a.size and a2.shape[0] > 1)
    assert b.size in (a2.shape[1], a.size)
    getattr(lib(), f"dsco_cmul_{suf}")(a2.ctypes.data, b.ctypes.data, out.ctypes.data,
                                        a2.shape[0], a2.shape[1], b_rows)
    return out.reshape(a.shape)


def unary(name, x):
    """Spectrum post-processing, numpy restatement of /root/reference/dsc/src/dsc.cpp:1480-1622:
    abs = |z| (std::abs), angle = atan2(im, re) (std::arg), real / imag / conj exact; complex in, the
    real dtype out (conj: same dtype)."""
    x = np.ascontiguousarray(x)
    real_dt = _REAL_OF[x.dtype]
    if name == "abs":
        return np.sqrt(x.real * x.real + x.imag * x.imag).astype(real_dt)
    if name == "angle":
        return np.arctan2(x.imag, x.real).astype(real_dt)
    if name == "real":
        return x.real.astype(real_dt)
    if name == "imag":
        return x.imag.astype(real_dt)
    if name == "conj":
        return np.conj(x)
    raise ValueError(name)


def binary(name, a, b):
    """dsc_add / dsc_sub / dsc_mul / dsc_div for operands of one dtype with NumPy broadcasting
    (/root/reference/dsc/src/dsc.cpp:1186-1310, functors dsc/include/dsc_ops.h:46-90)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    if name == "add":
        return (a + b).astype(a.dtype)
    if name == "sub":
        return (a - b).astype(a.dtype)
    if name == "mul":
        return (a * b).astype(a.dtype)
    if name == "div":
        return (a / b).astype(a.dtype)
    raise ValueError(name)


def filter_fft(s, b, fft_size):
    """README.md:118-134 filterFFT, uncropped: irfft(rfft(s, n) * rfft(b, n))."""
    S = rfft(s, n=fft_size)
    B = rfft(b, n=fft_size)
    return irfft(cmul(S, B))
