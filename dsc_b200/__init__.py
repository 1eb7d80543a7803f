"""dsc_b200 -- Python face of the B200-native libdsc.so (FFT hot path of dspcraft/dsc).

The shared library is a drop-in for the reference's ``libdsc.so``: the reference's own
``python/dsc`` wrapper binds it unchanged.  This package is the
host-side mirror used where that wrapper is not available (its sources are not copied
here): same names, argument meaning and error behaviour for the FFT path --

    init / clear / used_mem          python/dsc/context.py:29-51
    Tensor, from_numpy, .numpy()     python/dsc/tensor.py:159-170, 305-323, 371-377
    fft / ifft / rfft / irfft        python/dsc/tensor.py:684-726   (x, n=-1, axis=-1, out=None)
    mul, Tensor.__mul__, slicing     python/dsc/tensor.py:215-218, 281
    traces_record / dump_traces      python/dsc/profiler.py:14-34 (the C calls, not the web server)

plus what this build adds: ``fft_filter`` (fused rfft * B -> irfft), ``set_residency`` and
``sync_host``.  Everything goes through the C ABI of include/dsc.h with ctypes; there is no
Python or CPU implementation of the transforms -- without the built library importing the
context fails, and without a CUDA device every FFT call aborts inside the library.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Union

import numpy as np

__all__ = ["init", "clear", "shutdown", "used_mem", "device_used_mem", "device_alloc_calls", "Tensor",
           "from_numpy", "fft", "ifft", "rfft", "irfft", "mul", "fft_filter", "plan_fft",
           "traces_record", "dump_traces", "clear_traces", "set_residency", "sync_host", "Dtype", "LIBDSC"]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBDSC = os.path.join(_HERE, "libdsc.so")

VALUE_NONE = 2**31 - 1
FFT_REAL, FFT_COMPLEX = 0, 1


class Dtype:
    F32, F64, C32, C64 = 0, 1, 2, 3


_NP2DSC = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}
_DSC2NP = {v: k for k, v in _NP2DSC.items()}


class _CTensor(C.Structure):
    # struct dsc_tensor, include/dsc.h (64 bytes)
    _fields_ = [("shape", C.c_int * 4), ("stride", C.c_int * 4), ("buffer", C.c_void_p), ("data", C.c_void_p),
                ("ne", C.c_int), ("n_dim", C.c_int), ("dtype", C.c_uint8), ("backend", C.c_uint8)]


class _CSlice(C.Structure):
    _fields_ = [("start", C.c_int), ("stop", C.c_int), ("step", C.c_int)]


_TP = C.POINTER(_CTensor)
_lib = None
_ctx = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBDSC):
        raise ImportError(f"{LIBDSC} is not built (run __graft_entry__.build()); there is no fallback implementation")
    L = C.CDLL(LIBDSC, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    L.dsc_ctx_init.restype = C.c_void_p
    L.dsc_ctx_init.argtypes = [C.c_size_t, C.c_size_t]
    L.dsc_ctx_free.argtypes = [C.c_void_p]
    L.dsc_ctx_clear.argtypes = [C.c_void_p]
    for name in ("dsc_used_mem", "dsc_cuda_used_mem", "dsc_cuda_alloc_calls"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_void_p]
    L.dsc_plan_fft.restype = C.c_void_p
    L.dsc_plan_fft.argtypes = [C.c_void_p, C.c_int, C.c_uint8, C.c_uint8]
    L.dsc_tensor_free.argtypes = [C.c_void_p, _TP]
    L.dsc_view.restype = _TP
    L.dsc_view.argtypes = [C.c_void_p, _TP]
    for nd in range(1, 5):
        f = getattr(L, f"dsc_tensor_{nd}d")
        f.restype = _TP
        f.argtypes = [C.c_void_p, C.c_uint8] + [C.c_int] * nd
    for name in ("dsc_fft", "dsc_ifft", "dsc_rfft", "dsc_irfft"):
        f = getattr(L, name)
        f.restype = _TP
        f.argtypes = [C.c_void_p, _TP, _TP, C.c_int, C.c_int]
    L.dsc_fft_filter.restype = _TP
    L.dsc_fft_filter.argtypes = [C.c_void_p, _TP, _TP, _TP, C.c_int, C.c_int]
    for name in ("dsc_add", "dsc_sub", "dsc_mul", "dsc_div"):
        getattr(L, name).restype = _TP
        getattr(L, name).argtypes = [C.c_void_p, _TP, _TP, _TP]
    L.dsc_abs.restype = _TP
    L.dsc_abs.argtypes = [C.c_void_p, _TP, _TP]
    for name in ("dsc_angle", "dsc_real", "dsc_imag", "dsc_conj"):
        getattr(L, name).restype = _TP
        getattr(L, name).argtypes = [C.c_void_p, _TP]
    L.dsc_tensor_get_slice.restype = _TP
    L.dsc_tensor_set_slice.restype = None
    L.dsc_cast.restype = _TP
    L.dsc_cast.argtypes = [C.c_void_p, _TP, C.c_uint8]
    L.dsc_transpose.restype = _TP
    for name in ("dsc_fftfreq", "dsc_rfftfreq"):
        getattr(L, name).restype = _TP
        getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint8]
    for name in ("dsc_irfft_keep",):
        getattr(L, name).restype = _TP
        getattr(L, name).argtypes = [C.c_void_p, _TP, _TP, C.c_int, C.c_int, C.c_int]
    L.dsc_fft_filter_keep.restype = _TP
    L.dsc_fft_filter_keep.argtypes = [C.c_void_p, _TP, _TP, _TP, C.c_int, C.c_int, C.c_int]
    L.dsc_traces_record.argtypes = [C.c_void_p, C.c_bool]
    L.dsc_dump_traces.argtypes = [C.c_void_p, C.c_char_p]
    L.dsc_clear_traces.argtypes = [C.c_void_p]
    L.dsc_cuda_set_residency.argtypes = [C.c_void_p, C.c_int]
    L.dsc_cuda_sync_host.argtypes = [C.c_void_p, _TP]
    L.dsc_cuda_touch_host.argtypes = [C.c_void_p, _TP]
    L.dsc_cuda_prefetch.argtypes = [C.c_void_p, _TP]
    L.dsc_cuda_download_async.argtypes = [C.c_void_p, _TP]
    _lib = L
    return L


def init(mem_size: int, scratch_size: int) -> None:
    """dsc.init(mem_size, scratch_size): two host arenas plus ONE device arena, made once."""
    global _ctx
    if _ctx is not None:
        raise RuntimeError("context already initialised (the library supports one live context per process)")
    _ctx = _load().dsc_ctx_init(int(mem_size), int(scratch_size))


def _get_ctx():
    if _ctx is None:
        init(1 << 30, 1 << 28)
    return _ctx


def shutdown() -> None:
    global _ctx
    if _ctx is not None:
        _load().dsc_ctx_free(_ctx)
        _ctx = None


def clear() -> None:
    _load().dsc_ctx_clear(_get_ctx())


def used_mem() -> int:
    return int(_load().dsc_used_mem(_get_ctx()))


def device_used_mem() -> int:
    return int(_load().dsc_cuda_used_mem(_get_ctx()))


def device_alloc_calls() -> int:
    return int(_load().dsc_cuda_alloc_calls(_get_ctx()))


class Tensor:
    """Owner of one ``dsc_tensor*``; freed with dsc_tensor_free when collected."""

    def __init__(self, ptr, view: bool = False):
        if not ptr:
            raise RuntimeError("null tensor")
        # an `out=` result is the caller's pointer: wrap a new view so both owners can free theirs
        self._ptr = _load().dsc_view(_get_ctx(), ptr) if view else ptr

    def __del__(self):
        if _ctx is not None and getattr(self, "_ptr", None):
            _load().dsc_tensor_free(_ctx, self._ptr)
            self._ptr = None

    @property
    def c(self):
        return self._ptr

    @property
    def shape(self):
        t = self._ptr.contents
        return tuple(t.shape[4 - t.n_dim:4])

    @property
    def dtype(self):
        return _DSC2NP[self._ptr.contents.dtype]

    @property
    def n_dim(self):
        return self._ptr.contents.n_dim

    def numpy(self) -> np.ndarray:
        """Copy of the payload (the library keeps the host copy of results current unless residency == 2)."""
        sync_host(self)
        t = self._ptr.contents
        out = np.empty(self.shape, dtype=self.dtype)
        if out.nbytes:
            C.memmove(out.ctypes.data, t.data, out.nbytes)
        return out

    def __mul__(self, other: "Tensor") -> "Tensor":
        return mul(self, other)

    def __add__(self, other: "Tensor") -> "Tensor":
        return add(self, other)

    def __sub__(self, other: "Tensor") -> "Tensor":
        return sub(self, other)

    def __truediv__(self, other: "Tensor") -> "Tensor":
        return true_div(self, other)

    @staticmethod
    def _slices(item):
        items = item if isinstance(item, tuple) else (item,)
        args = []
        for it in items:
            if isinstance(it, slice):
                args.append(_CSlice(VALUE_NONE if it.start is None else it.start,
                                    VALUE_NONE if it.stop is None else it.stop,
                                    VALUE_NONE if it.step is None else it.step))
            else:
                args.append(_CSlice(int(it), int(it), int(it)))     # single-index convention
        return args

    def __getitem__(self, item) -> "Tensor":
        args = self._slices(item)
        return Tensor(_load().dsc_tensor_get_slice(_get_ctx(), self._ptr, C.c_int(len(args)), *args))

    def __setitem__(self, item, value) -> None:
        """x[slices] = value (python/dsc/tensor.py:220-237 -> dsc_tensor_set_slice, dsc/src/dsc.cpp:1108-1169)."""
        args = self._slices(item)
        v = value if isinstance(value, Tensor) else from_numpy(np.asarray(value, dtype=self.dtype))
        _load().dsc_tensor_set_slice(_get_ctx(), self._ptr, v._ptr, C.c_int(len(args)), *args)

    def cast(self, np_dtype) -> "Tensor":
        """dsc_cast (dsc/src/dsc.cpp:587-597); returns self when the dtype already matches, like the reference."""
        ptr = _load().dsc_cast(_get_ctx(), self._ptr, _NP2DSC[np.dtype(np_dtype)])
        if C.addressof(ptr.contents) == C.addressof(self._ptr.contents):
            return self
        return Tensor(ptr)


def from_numpy(a: np.ndarray) -> Tensor:
    a = np.ascontiguousarray(a)
    if a.dtype not in _NP2DSC:
        raise TypeError(f"unsupported dtype {a.dtype}")
    if not 1 <= a.ndim <= 4:
        raise ValueError("tensors have 1 to 4 dimensions")
    L = _load()
    ptr = getattr(L, f"dsc_tensor_{a.ndim}d")(_get_ctx(), _NP2DSC[a.dtype], *[int(s) for s in a.shape])
    if a.nbytes:
        C.memmove(ptr.contents.data, a.ctypes.data, a.nbytes)
    return Tensor(ptr)


def _as_tensor(x: Union[Tensor, np.ndarray]) -> Tensor:
    return x if isinstance(x, Tensor) else from_numpy(np.asarray(x))


def _xform(name: str, x, out: Optional[Tensor], n: int, axis: int) -> Tensor:
    x = _as_tensor(x)
    res = getattr(_load(), name)(_get_ctx(), x.c, out.c if out is not None else None, int(n), int(axis))
    return Tensor(res, view=out is not None)


def fft(x, out: Optional[Tensor] = None, n: int = -1, axis: int = -1) -> Tensor:
    return _xform("dsc_fft", x, out, n, axis)


def ifft(x, out: Optional[Tensor] = None, n: int = -1, axis: int = -1) -> Tensor:
    return _xform("dsc_ifft", x, out, n, axis)


def rfft(x, out: Optional[Tensor] = None, n: int = -1, axis: int = -1) -> Tensor:
    return _xform("dsc_rfft", x, out, n, axis)


def irfft(x, out: Optional[Tensor] = None, n: int = -1, axis: int = -1) -> Tensor:
    """n counts INPUT BINS (reference quirk, dsc.cpp:2197-2200)."""
    return _xform("dsc_irfft", x, out, n, axis)


def _binary(name: str, a, b, out: Optional[Tensor]) -> Tensor:
    a, b = _as_tensor(a), _as_tensor(b)
    res = getattr(_load(), name)(_get_ctx(), a.c, b.c, out.c if out is not None else None)
    return Tensor(res, view=out is not None)


def add(a, b, out: Optional[Tensor] = None) -> Tensor:
    return _binary("dsc_add", a, b, out)


def sub(a, b, out: Optional[Tensor] = None) -> Tensor:
    return _binary("dsc_sub", a, b, out)


def mul(a, b, out: Optional[Tensor] = None) -> Tensor:
    return _binary("dsc_mul", a, b, out)


def true_div(a, b, out: Optional[Tensor] = None) -> Tensor:
    return _binary("dsc_div", a, b, out)


def abs(x, out: Optional[Tensor] = None) -> Tensor:   # noqa: A001 (the reference wrapper's name, python/dsc/tensor.py)
    x = _as_tensor(x)
    res = _load().dsc_abs(_get_ctx(), x.c, out.c if out is not None else None)
    return Tensor(res, view=out is not None)


def _unary_new(name: str, x) -> Tensor:
    """Ops that may hand back their argument itself (real / conj of a real tensor, dsc.cpp:1549-1552)."""
    x = _as_tensor(x)
    res = getattr(_load(), name)(_get_ctx(), x.c)
    same = C.cast(res, C.c_void_p).value == C.cast(x.c, C.c_void_p).value
    return x if same else Tensor(res)


def angle(x) -> Tensor:
    return _unary_new("dsc_angle", x)


def real(x) -> Tensor:
    return _unary_new("dsc_real", x)


def imag(x) -> Tensor:
    return _unary_new("dsc_imag", x)


def conj(x) -> Tensor:
    return _unary_new("dsc_conj", x)


def transpose(x, axes=None) -> Tensor:
    """dsc_transpose (dsc/src/dsc.cpp:764-827): axes=None reverses the dims."""
    x = _as_tensor(x)
    axes = tuple(axes) if axes is not None else ()
    return Tensor(_load().dsc_transpose(_get_ctx(), x._ptr, C.c_int(len(axes)), *[C.c_int(int(a)) for a in axes]))


def fftfreq(n: int, d: float = 1.0, dtype=np.float64) -> Tensor:
    return Tensor(_load().dsc_fftfreq(_get_ctx(), int(n), float(d), _NP2DSC[np.dtype(dtype)]))


def rfftfreq(n: int, d: float = 1.0, dtype=np.float64) -> Tensor:
    return Tensor(_load().dsc_rfftfreq(_get_ctx(), int(n), float(d), _NP2DSC[np.dtype(dtype)]))


def irfft_keep(x, keep: int, out: Optional[Tensor] = None, n: int = -1) -> Tensor:
    """irfft along the last axis storing only the first `keep` samples of every line: the README's
    irfft(...)[:output_length] (README.md:130-133) with the crop fused into the inverse kernel's store."""
    x = _as_tensor(x)
    ptr = _load().dsc_irfft_keep(_get_ctx(), x._ptr, out._ptr if out is not None else None, int(n), -1, int(keep))
    return Tensor(ptr, view=out is not None)


def fft_filter_keep(x, B, keep: int, out: Optional[Tensor] = None, n: int = -1) -> Tensor:
    """fft_filter storing only the first `keep` samples of every line (crop fused into the kernel's store)."""
    x, B = _as_tensor(x), _as_tensor(B)
    ptr = _load().dsc_fft_filter_keep(_get_ctx(), x._ptr, B._ptr, out._ptr if out is not None else None, int(n), -1, int(keep))
    return Tensor(ptr, view=out is not None)


def fft_filter(x, B, out: Optional[Tensor] = None, n: int = -1, axis: int = -1) -> Tensor:
    """irfft(rfft(x, n) * B) in one device pipeline; B = rfft(b, n) (README.md:118-134)."""
    x, B = _as_tensor(x), _as_tensor(B)
    res = _load().dsc_fft_filter(_get_ctx(), x.c, B.c, out.c if out is not None else None, int(n), int(axis))
    return Tensor(res, view=out is not None)


def plan_fft(n: int, dtype: int = Dtype.F64, fft_type: int = FFT_COMPLEX) -> int:
    return _load().dsc_plan_fft(_get_ctx(), int(n), fft_type, dtype)


def traces_record(record: bool = True) -> None:
    _load().dsc_traces_record(_get_ctx(), bool(record))


def dump_traces(filename: str) -> None:
    _load().dsc_dump_traces(_get_ctx(), filename.encode())


def clear_traces() -> None:
    _load().dsc_clear_traces(_get_ctx())


def set_residency(mode: int) -> None:
    """0 strict (default), 1 keep results on the device, 2 also defer downloads (include/dsc.h)."""
    _load().dsc_cuda_set_residency(_get_ctx(), int(mode))


def prefetch(x: Tensor) -> None:
    """Upload x to its device mirror now (residency >= 1), so that later ops on it run on the device."""
    _load().dsc_cuda_prefetch(_get_ctx(), x.c)


def download_async(x: Tensor) -> None:
    """Residency 2: start the device -> host copy of x and return; sync_host(x) / x.numpy() waits for it."""
    _load().dsc_cuda_download_async(_get_ctx(), x.c)


def sync_host(x: Tensor) -> None:
    _load().dsc_cuda_sync_host(_get_ctx(), x.c)
