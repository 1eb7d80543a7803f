"""ctypes binding of the device-level C ABI (include/dsc_cuda.h).

Raw pointers in, raw pointers out: callers own the memory (the device arena of a context,
or any device allocation such as a torch tensor's ``data_ptr()``).  Nothing in here
touches the oracle; if the shared library is missing this module raises at load time.
"""
from __future__ import annotations

import ctypes as C
import os

F32, F64, C32, C64 = 0, 1, 2, 3
FFT_REAL, FFT_COMPLEX = 0, 1
MAX_STAGES = 5

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBDSC = os.path.join(_HERE, "libdsc.so")


class Plan(C.Structure):
    """struct dsc_cuda_plan (include/dsc_cuda.h)."""
    _fields_ = [("n", C.c_int), ("lg_n", C.c_int), ("fft_type", C.c_int), ("dtype", C.c_int),
                ("lg_n1", C.c_int), ("lg_n2", C.c_int), ("four_shift", C.c_int),
                ("dev_base", C.c_void_p), ("dev_bytes", C.c_size_t),
                ("tw1", C.c_void_p * MAX_STAGES), ("tw2", C.c_void_p * MAX_STAGES),
                ("tw_lo", C.c_void_p), ("tw_hi", C.c_void_p), ("tw_real", C.c_void_p),
                ("tw_real_lo", C.c_void_p), ("tw_real_hi", C.c_void_p), ("real_shift", C.c_int),
                ("col_lg_n1", C.c_int), ("col_lg_n2", C.c_int), ("col_shift", C.c_int),
                ("col_tw1", C.c_void_p * MAX_STAGES), ("col_tw2", C.c_void_p * MAX_STAGES),
                ("col_lo", C.c_void_p), ("col_hi", C.c_void_p),
                ("tw1_e16", C.c_void_p * MAX_STAGES), ("tw2_e16", C.c_void_p * MAX_STAGES)]


class DscCudaError(RuntimeError):
    pass


class CudaApi:
    """The dsc_cuda_* entry points of one shared object."""

    def __init__(self, path: str = LIBDSC):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the FFT path)")
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        pp = C.POINTER(Plan)
        L.dsc_cuda_last_error.restype = C.c_char_p
        L.dsc_cuda_device_count.restype = C.c_int
        L.dsc_cuda_plan_bytes.restype = C.c_size_t
        L.dsc_cuda_plan_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.dsc_cuda_plan_build.argtypes = [pp, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.dsc_cuda_work_bytes.restype = C.c_size_t
        L.dsc_cuda_work_bytes.argtypes = [pp, C.c_int64]
        L.dsc_cuda_work_bytes_axis.restype = C.c_size_t
        L.dsc_cuda_work_bytes_axis.argtypes = [pp, C.c_int64, C.c_int64]
        L.dsc_cuda_fft.argtypes = [pp, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                   C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.dsc_cuda_fft_segmented.argtypes = [pp, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.dsc_cuda_fft_columns_twiddled.argtypes = [pp, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p,
                                                    C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
        L.dsc_cuda_fft_columns_twiddled_p2p.argtypes = [pp, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                                        C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_int, C.c_void_p,
                                                        C.c_size_t, C.c_void_p]
        L.dsc_cuda_rfft.argtypes = [pp, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                    C.c_void_p, C.c_size_t, C.c_void_p]
        L.dsc_cuda_irfft.argtypes = L.dsc_cuda_rfft.argtypes
        L.dsc_cuda_cmul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                    C.c_int, C.c_void_p]
        L.dsc_cuda_unary.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.dsc_cuda_binary.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                      C.c_int, C.c_void_p]
        L.dsc_cuda_filter_work_bytes.restype = C.c_size_t
        L.dsc_cuda_filter_work_bytes.argtypes = [pp, C.c_int64]
        L.dsc_cuda_filter.argtypes = [pp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.dsc_cuda_fill_twiddles.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        L.dsc_cuda_transpose_twiddle.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise DscCudaError(f"{what} failed ({rc}): {self.lib.dsc_cuda_last_error().decode()}")

    def device_count(self) -> int:
        return self.lib.dsc_cuda_device_count()

    def plan_bytes(self, n: int, fft_type: int, dtype: int) -> int:
        return self.lib.dsc_cuda_plan_bytes(n, fft_type, dtype)

    def plan_build(self, n: int, fft_type: int, dtype: int, dev_ptr: int, dev_bytes: int, stream: int = 0) -> Plan:
        p = Plan()
        self._check(self.lib.dsc_cuda_plan_build(C.byref(p), n, fft_type, dtype, dev_ptr, dev_bytes, stream),
                    "dsc_cuda_plan_build")
        return p

    def work_bytes(self, plan: Plan, lines: int) -> int:
        return self.lib.dsc_cuda_work_bytes(C.byref(plan), lines)

    def work_bytes_axis(self, plan: Plan, outer: int, inner: int) -> int:
        return self.lib.dsc_cuda_work_bytes_axis(C.byref(plan), outer, inner)

    def fft(self, plan, x_ptr, x_dtype, out_ptr, outer, x_n, inner, forward, work_ptr=0, work_bytes=0, stream=0):
        self._check(self.lib.dsc_cuda_fft(C.byref(plan), x_ptr, x_dtype, out_ptr, outer, x_n, inner,
                                          int(forward), work_ptr, work_bytes, stream), "dsc_cuda_fft")

    def fft_segmented(self, plan, x_ptr, out_ptr, lines, seg_len, seg_stride, forward, work_ptr=0, work_bytes=0, stream=0,
                      self_seg=-1, self_ptr=0):
        """Lines stored as n/seg_len segments (segment s of line r at x + s*seg_stride + r*seg_len; segment self_seg
        read from self_ptr instead); returns False when this plan / segment size is not covered (caller
        un-interleaves and uses fft)."""
        rc = self.lib.dsc_cuda_fft_segmented(C.byref(plan), x_ptr, out_ptr, lines, seg_len, seg_stride, self_seg, self_ptr,
                                             int(forward), work_ptr, work_bytes, stream)
        if rc == -4:         # DSC_CUDA_EUNSUPPORTED
            return False
        self._check(rc, "dsc_cuda_fft_segmented")
        return True

    def fft_columns_twiddled(self, plan, x_ptr, out_ptr, cols, forward, col_offset, tw_lo, tw_hi, shift, total,
                             work_ptr=0, work_bytes=0, stream=0):
        """Column transforms of a natural-order block [n][cols] with the outer four-step twiddle applied, k-major
        output; False when the shape is not covered."""
        rc = self.lib.dsc_cuda_fft_columns_twiddled(C.byref(plan), x_ptr, out_ptr, cols, int(forward), col_offset,
                                                    tw_lo, tw_hi, shift, total, work_ptr, work_bytes, stream)
        if rc == -4:         # DSC_CUDA_EUNSUPPORTED
            return False
        self._check(rc, "dsc_cuda_fft_columns_twiddled")
        return True

    def fft_columns_twiddled_p2p(self, plan, x_ptr, cols, forward, col_offset, tw_lo, tw_hi, shift, total, peer_ptrs,
                                 work_ptr=0, work_bytes=0, stream=0):
        """fft_columns_twiddled with the exchange fused into the epilogue: row block q of the result is stored at
        peer_ptrs[q] (peer-mapped receive buffers); False when the shape is not covered."""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        rc = self.lib.dsc_cuda_fft_columns_twiddled_p2p(C.byref(plan), x_ptr, cols, int(forward), col_offset, tw_lo, tw_hi,
                                                        shift, total, arr, len(peer_ptrs), work_ptr, work_bytes, stream)
        if rc == -4:         # DSC_CUDA_EUNSUPPORTED
            return False
        self._check(rc, "dsc_cuda_fft_columns_twiddled_p2p")
        return True

    def rfft(self, plan, x_ptr, out_ptr, outer, x_n, inner, work_ptr=0, work_bytes=0, stream=0):
        self._check(self.lib.dsc_cuda_rfft(C.byref(plan), x_ptr, out_ptr, outer, x_n, inner,
                                           work_ptr, work_bytes, stream), "dsc_cuda_rfft")

    def irfft(self, plan, x_ptr, out_ptr, outer, x_n, inner, work_ptr=0, work_bytes=0, stream=0):
        self._check(self.lib.dsc_cuda_irfft(C.byref(plan), x_ptr, out_ptr, outer, x_n, inner,
                                            work_ptr, work_bytes, stream), "dsc_cuda_irfft")

    def cmul(self, a_ptr, b_ptr, out_ptr, dtype, rows, cols, b_rows, stream=0):
        self._check(self.lib.dsc_cuda_cmul(a_ptr, b_ptr, out_ptr, dtype, rows, cols, int(b_rows), stream),
                    "dsc_cuda_cmul")

    def unary(self, op, x_ptr, x_dtype, out_ptr, count, stream=0):
        """op: 0 abs, 1 angle, 2 real, 3 imag, 4 conj (DSC_CUDA_OP_*)."""
        self._check(self.lib.dsc_cuda_unary(op, x_ptr, x_dtype, out_ptr, count, stream), "dsc_cuda_unary")

    def binary(self, op, a_ptr, b_ptr, out_ptr, dtype, rows, cols, b_mode, stream=0):
        """op: 0 add, 1 sub, 2 mul, 3 div; b_mode: 0 one row, 1 same shape, 2 one element."""
        self._check(self.lib.dsc_cuda_binary(op, a_ptr, b_ptr, out_ptr, dtype, rows, cols, b_mode, stream),
                    "dsc_cuda_binary")

    def fill_twiddles(self, out_ptr, count, mult, denom, dtype, stream=0):
        self._check(self.lib.dsc_cuda_fill_twiddles(out_ptr, count, mult, denom, dtype, stream), "dsc_cuda_fill_twiddles")

    def transpose_twiddle(self, in_ptr, out_ptr, rows, cols, r0, tw_lo, tw_hi, shift, forward, dtype, stream=0):
        self._check(self.lib.dsc_cuda_transpose_twiddle(in_ptr, out_ptr, rows, cols, r0, tw_lo, tw_hi, shift,
                                                        int(forward), dtype, stream), "dsc_cuda_transpose_twiddle")

    def filter_work_bytes(self, plan, lines):
        return self.lib.dsc_cuda_filter_work_bytes(C.byref(plan), lines)

    def filter(self, plan, x_ptr, spectrum_ptr, out_ptr, outer, x_n, work_ptr=0, work_bytes=0, stream=0):
        self._check(self.lib.dsc_cuda_filter(C.byref(plan), x_ptr, spectrum_ptr, out_ptr, outer, x_n,
                                             work_ptr, work_bytes, stream), "dsc_cuda_filter")
