"""Raw-ctypes driver for the UNMODIFIED reference library (oracle/_ref/libdsc_ref.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by dsc_b200/.

The reference's Python package cannot travel to the GPU box (its sources must not
be copied into this repo), so this file binds the handful of C entry points the
FFT path needs, straight from the reference's public header:

    dsc_ctx_init / dsc_ctx_free / dsc_ctx_clear   dsc/include/dsc.h:137-148
    dsc_tensor_{1..4}d, dsc_tensor_free           dsc/include/dsc.h:150,182-198
    dsc_fft / dsc_ifft / dsc_rfft / dsc_irfft     dsc/include/dsc.h:392-414
    dsc_mul                                       dsc/include/dsc.h:275-278
    dsc_plan_fft, dsc_used_mem                    dsc/include/dsc.h:139-141,155
    dsc_tensor_get_slice                          dsc/include/dsc.h:248-250
    dsc_traces_record / dsc_dump_traces           dsc/include/dsc.h:162-168

`struct dsc_tensor` layout follows dsc/include/dsc.h:96-108 (64 bytes).
Data moves in and out with memmove on tensor.data exactly as the reference's own
wrapper does (python/dsc/tensor.py:305-323, 371-377).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libdsc_ref.so")
REF_TRACED_SO = os.path.join(_HERE, "_ref", "libdsc_ref_traced.so")

F32, F64, C32, C64 = 0, 1, 2, 3
_NP2DSC = {np.dtype(np.float32): F32, np.dtype(np.float64): F64,
           np.dtype(np.complex64): C32, np.dtype(np.complex128): C64}
_DSC2NP = {v: k for k, v in _NP2DSC.items()}
FFT_REAL, FFT_COMPLEX = 0, 1
VALUE_NONE = 2**31 - 1


class _Tensor(C.Structure):
    _fields_ = [("shape", C.c_int * 4), ("stride", C.c_int * 4),
                ("buffer", C.c_void_p), ("data", C.c_void_p),
                ("ne", C.c_int), ("n_dim", C.c_int),
                ("dtype", C.c_uint8), ("backend", C.c_uint8)]


class _Slice(C.Structure):
    _fields_ = [("start", C.c_int), ("stop", C.c_int), ("step", C.c_int)]


_TP = C.POINTER(_Tensor)


@contextlib.contextmanager
def _quiet_stdout():
    """The library logs with printf on fd 1 (dsc.h:20); keep bench's stdout to one JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def available(path: str = REF_SO) -> bool:
    return os.path.exists(path)


class RefLib:
    """One context of a libdsc-ABI shared object (the reference build by default)."""

    def __init__(self, main_mem: int = 1 << 30, scratch_mem: int = 1 << 28, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        # RTLD_LOCAL: the product library exports the very same symbol names.
        self.lib = lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        lib.dsc_ctx_init.restype = C.c_void_p
        lib.dsc_ctx_init.argtypes = [C.c_size_t, C.c_size_t]
        lib.dsc_ctx_free.argtypes = [C.c_void_p]
        lib.dsc_ctx_clear.argtypes = [C.c_void_p]
        lib.dsc_used_mem.restype = C.c_size_t
        lib.dsc_used_mem.argtypes = [C.c_void_p]
        lib.dsc_tensor_free.argtypes = [C.c_void_p, _TP]
        lib.dsc_plan_fft.restype = C.c_void_p
        lib.dsc_plan_fft.argtypes = [C.c_void_p, C.c_int, C.c_uint8, C.c_uint8]
        for nd in range(1, 5):
            f = getattr(lib, f"dsc_tensor_{nd}d")
            f.restype = _TP
            f.argtypes = [C.c_void_p, C.c_uint8] + [C.c_int] * nd
        for name in ("dsc_fft", "dsc_ifft", "dsc_rfft", "dsc_irfft"):
            f = getattr(lib, name)
            f.restype = _TP
            f.argtypes = [C.c_void_p, _TP, _TP, C.c_int, C.c_int]
        for name in ("dsc_add", "dsc_sub", "dsc_mul", "dsc_div"):
            getattr(lib, name).restype = _TP
            getattr(lib, name).argtypes = [C.c_void_p, _TP, _TP, _TP]
        lib.dsc_abs.restype = _TP
        lib.dsc_abs.argtypes = [C.c_void_p, _TP, _TP]
        for name in ("dsc_angle", "dsc_real", "dsc_imag", "dsc_conj"):
            getattr(lib, name).restype = _TP
            getattr(lib, name).argtypes = [C.c_void_p, _TP]
        lib.dsc_tensor_get_slice.restype = _TP
        lib.dsc_tensor_set_slice.restype = None
        lib.dsc_cast.restype = _TP
        lib.dsc_cast.argtypes = [C.c_void_p, _TP, C.c_uint8]
        lib.dsc_transpose.restype = _TP
        for name in ("dsc_fftfreq", "dsc_rfftfreq"):
            getattr(lib, name).restype = _TP
            getattr(lib, name).argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint8]
        lib.dsc_traces_record.argtypes = [C.c_void_p, C.c_bool]
        lib.dsc_dump_traces.argtypes = [C.c_void_p, C.c_char_p]
        lib.dsc_clear_traces.argtypes = [C.c_void_p]
        with _quiet_stdout():
            self.ctx = lib.dsc_ctx_init(main_mem, scratch_mem)

    # -- tensors -----------------------------------------------------------------
    def new(self, shape, np_dtype) -> "_TP":
        shape = tuple(int(s) for s in shape)
        assert 1 <= len(shape) <= 4
        f = getattr(self.lib, f"dsc_tensor_{len(shape)}d")
        return f(self.ctx, _NP2DSC[np.dtype(np_dtype)], *shape)

    def put(self, a: np.ndarray) -> "_TP":
        a = np.ascontiguousarray(a)
        t = self.new(a.shape, a.dtype)
        C.memmove(t.contents.data, a.ctypes.data, a.nbytes)
        return t

    def get(self, t) -> np.ndarray:
        tc = t.contents
        shape = tuple(tc.shape[4 - tc.n_dim:4])
        out = np.empty(shape, dtype=_DSC2NP[tc.dtype])
        C.memmove(out.ctypes.data, tc.data, out.nbytes)
        return out

    def free(self, t) -> None:
        self.lib.dsc_tensor_free(self.ctx, t)

    # -- numpy-in / numpy-out transforms -------------------------------------------
    def _xform(self, name, x, n, axis, out=None):
        tx = self.put(x)
        to = getattr(self.lib, name)(self.ctx, tx, out, int(n), int(axis))
        res = self.get(to)
        if out is None:
            self.free(to)
        self.free(tx)
        return res

    def fft(self, x, n=-1, axis=-1):
        return self._xform("dsc_fft", x, n, axis)

    def ifft(self, x, n=-1, axis=-1):
        return self._xform("dsc_ifft", x, n, axis)

    def rfft(self, x, n=-1, axis=-1):
        return self._xform("dsc_rfft", x, n, axis)

    def irfft(self, x, n=-1, axis=-1):
        return self._xform("dsc_irfft", x, n, axis)

    def mul(self, a, b):
        ta, tb = self.put(a), self.put(b)
        to = self.lib.dsc_mul(self.ctx, ta, tb, None)
        res = self.get(to)
        for t in (to, ta, tb):
            self.free(t)
        return res

    def binary(self, name, a, b):
        """dsc_add / dsc_sub / dsc_mul / dsc_div with the reference's broadcasting (dsc.cpp:1186-1310)."""
        ta, tb = self.put(a), self.put(b)
        to = getattr(self.lib, f"dsc_{name}")(self.ctx, ta, tb, None)
        res = self.get(to)
        for t in (to, ta, tb):
            self.free(t)
        return res

    def unary(self, name, x):
        """dsc_abs / dsc_angle / dsc_real / dsc_imag / dsc_conj of a complex tensor (dsc.cpp:1480-1622)."""
        tx = self.put(x)
        to = self.lib.dsc_abs(self.ctx, tx, None) if name == "abs" else getattr(self.lib, f"dsc_{name}")(self.ctx, tx)
        res = self.get(to)
        self.free(to)
        self.free(tx)
        return res

    def filter_fft(self, s, b, fft_size):
        """README.md:118-134 `filterFFT`: irfft(rfft(s, n) * rfft(b, n)), uncropped."""
        ts, tb = self.put(s), self.put(b)
        S = self.lib.dsc_rfft(self.ctx, ts, None, int(fft_size), -1)
        B = self.lib.dsc_rfft(self.ctx, tb, None, int(fft_size), -1)
        P = self.lib.dsc_mul(self.ctx, S, B, None)
        y = self.lib.dsc_irfft(self.ctx, P, None, -1, -1)
        res = self.get(y)
        for t in (y, P, B, S, tb, ts):
            self.free(t)
        return res

    def slice_1d(self, x, start, stop, step=1):
        tx = self.put(x)
        nd = x.ndim
        args = [_Slice(VALUE_NONE, VALUE_NONE, 1)] * (nd - 1) + [_Slice(start, stop, step)]
        to = self.lib.dsc_tensor_get_slice(self.ctx, tx, C.c_int(nd), *args)
        res = self.get(to)
        self.free(to)
        self.free(tx)
        return res

    @staticmethod
    def _slice_args(slices):
        """python slices / ints -> dsc_slice structs (the wrappers' single-index convention for ints)."""
        out = []
        for it in slices:
            if isinstance(it, slice):
                out.append(_Slice(VALUE_NONE if it.start is None else it.start, VALUE_NONE if it.stop is None else it.stop,
                                  VALUE_NONE if it.step is None else it.step))
            else:
                out.append(_Slice(int(it), int(it), int(it)))
        return out

    def get_slice(self, x, slices):
        """dsc_tensor_get_slice (dsc.cpp:950-1007) with NumPy-style slices."""
        tx = self.put(x)
        args = self._slice_args(slices)
        to = self.lib.dsc_tensor_get_slice(self.ctx, tx, C.c_int(len(args)), *args)
        res = self.get(to)
        self.free(to)
        self.free(tx)
        return res

    def set_slice(self, xa, xb, slices):
        """dsc_tensor_set_slice (dsc.cpp:1108-1169): returns xa after xa[slices] = xb."""
        ta, tb = self.put(xa), self.put(xb)
        args = self._slice_args(slices)
        self.lib.dsc_tensor_set_slice(self.ctx, ta, tb, C.c_int(len(args)), *args)
        res = self.get(ta)
        self.free(tb)
        self.free(ta)
        return res

    def cast(self, x, np_dtype):
        """dsc_cast (dsc.cpp:587-597)."""
        tx = self.put(x)
        to = self.lib.dsc_cast(self.ctx, tx, _NP2DSC[np.dtype(np_dtype)])
        res = self.get(to)
        if C.addressof(to.contents) != C.addressof(tx.contents):
            self.free(to)
        self.free(tx)
        return res

    def transpose(self, x, axes=()):
        """dsc_transpose (dsc.cpp:764-827); no axes = reversed dims."""
        tx = self.put(x)
        to = self.lib.dsc_transpose(self.ctx, tx, C.c_int(len(axes)), *[C.c_int(int(a)) for a in axes])
        res = self.get(to)
        self.free(to)
        self.free(tx)
        return res

    def fftfreq(self, n, d=1.0, np_dtype=np.float64, rfft=False):
        """dsc_fftfreq / dsc_rfftfreq (dsc.cpp:2262-2339)."""
        f = self.lib.dsc_rfftfreq if rfft else self.lib.dsc_fftfreq
        to = f(self.ctx, int(n), float(d), _NP2DSC[np.dtype(np_dtype)])
        res = self.get(to)
        self.free(to)
        return res

    # -- misc ------------------------------------------------------------------------
    def plan(self, n, fft_type, dsc_dtype):
        return self.lib.dsc_plan_fft(self.ctx, int(n), fft_type, dsc_dtype)

    def used_mem(self) -> int:
        return int(self.lib.dsc_used_mem(self.ctx))

    def clear(self) -> None:
        self.lib.dsc_ctx_clear(self.ctx)

    def close(self) -> None:
        if self.ctx:
            with _quiet_stdout():
                self.lib.dsc_ctx_free(self.ctx)
            self.ctx = None
