"""numpy-facing driver of the device-level C ABI (dsc_cuda_*), for tests.

backend='numpy': "device" memory is host memory -- used with tests/emul/libdsc_emul.so,
                 the pthread emulation of the kernels (no GPU needed).
backend='torch': device memory comes from torch CUDA tensors -- used with the real
                 dsc_b200/libdsc.so on a B200.
Shape rules are taken from the oracle so that only the kernels are under test here.
"""
import numpy as np

from dsc_b200 import cuda_api
from oracle import port

_CODE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}
_CPLX = {np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128,
         np.dtype(np.complex64): np.complex64, np.dtype(np.complex128): np.complex128}
_REAL = {np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}


class _NumpyMem:
    def alloc(self, nbytes):
        buf = np.zeros(max(nbytes, 256) + 256, dtype=np.uint8)
        off = (-buf.ctypes.data) % 256
        return buf[off:off + max(nbytes, 256)]

    def ptr(self, buf):
        return buf.ctypes.data

    def upload(self, a):
        return np.ascontiguousarray(a).copy()

    def empty(self, shape, dtype, fill=None):
        out = np.empty(shape, dtype=dtype)
        out.view(np.uint8).reshape(-1)[:] = 0xCD       # poison: unwritten output shows up
        return out

    def download(self, buf):
        return buf

    def sync(self):
        pass


class _TorchMem:
    def __init__(self):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda:0")

    def alloc(self, nbytes):
        return self.torch.zeros(max(nbytes, 256), dtype=self.torch.uint8, device=self.dev)

    def ptr(self, buf):
        return buf.data_ptr()

    def upload(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def empty(self, shape, dtype, fill=None):
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        t = self.torch.full((max(nbytes, 1),), 0xCD, dtype=self.torch.uint8, device=self.dev)
        t._np_shape, t._np_dtype = tuple(shape), np.dtype(dtype)
        return t

    def download(self, buf):
        self.torch.cuda.synchronize()
        host = buf.cpu().numpy()
        if hasattr(buf, "_np_shape"):
            n = int(np.prod(buf._np_shape, dtype=np.int64)) * buf._np_dtype.itemsize
            return host[:n].view(buf._np_dtype).reshape(buf._np_shape)
        return host

    def sync(self):
        self.torch.cuda.synchronize()


class DevFFT:
    def __init__(self, lib_path, backend="numpy", work_lines=None):
        self.api = cuda_api.CudaApi(lib_path)
        self.mem = _NumpyMem() if backend == "numpy" else _TorchMem()
        self.plans = {}
        self.work_lines = work_lines      # None: full-size work buffer; k: force chunking by k lines

    def plan(self, n, fft_type, prec):
        key = (n, fft_type, prec)
        if key not in self.plans:
            nbytes = self.api.plan_bytes(n, fft_type, prec)
            assert nbytes > 0, f"plan_bytes({n}) == 0"
            block = self.mem.alloc(nbytes)
            p = self.api.plan_build(n, fft_type, prec, self.mem.ptr(block), nbytes)
            self.mem.sync()
            self.plans[key] = (p, block)
        return self.plans[key][0]

    def _work(self, plan, lines):
        nbytes = self.api.work_bytes(plan, lines if self.work_lines is None else min(lines, self.work_lines))
        if nbytes == 0:
            return None, 0, 0
        w = self.mem.alloc(nbytes)
        return w, self.mem.ptr(w), nbytes

    @staticmethod
    def _split(x, axis):
        ax = (4 + axis if axis < 0 else 4 - x.ndim + axis) - (4 - x.ndim)   # dsc.h:81
        assert 0 <= ax < x.ndim
        outer = int(np.prod(x.shape[:ax], dtype=np.int64))
        inner = int(np.prod(x.shape[ax + 1:], dtype=np.int64))
        return ax, outer, x.shape[ax], inner

    def _cfft(self, x, n, axis, forward):
        x = np.ascontiguousarray(x)
        ax, outer, x_n, inner = self._split(x, axis)
        fft_n = port.fft_len(x_n, n)
        prec = 0 if x.dtype in (np.float32, np.complex64) else 1
        plan = self.plan(fft_n, cuda_api.FFT_COMPLEX, prec)
        shape = list(x.shape); shape[ax] = fft_n
        dx = self.mem.upload(x)
        dout = self.mem.empty(shape, _CPLX[x.dtype])
        if inner > 1 and plan.col_lg_n2:
            nbytes = self.api.work_bytes_axis(plan, outer, inner)
            w = self.mem.alloc(nbytes) if nbytes else None
            wp, wb = (self.mem.ptr(w), nbytes) if nbytes else (0, 0)
        else:
            w, wp, wb = self._work(plan, outer * inner)
        self.api.fft(plan, self.mem.ptr(dx), _CODE[x.dtype], self.mem.ptr(dout), outer, x_n, inner, forward, wp, wb)
        return self.mem.download(dout)

    def fft_segmented(self, segs, forward=True, self_seg=-1):
        """segs: [S][lines][seg_len] complex (line r = concatenation of segs[:, r, :]).  Returns [lines][S*seg_len]
        or None when the library does not cover the shape."""
        segs = np.ascontiguousarray(segs)
        S, lines, seg_len = segs.shape
        n = S * seg_len
        prec = 0 if segs.dtype == np.complex64 else 1
        plan = self.plan(n, cuda_api.FFT_COMPLEX, prec)
        dx = self.mem.upload(segs)
        dout = self.mem.empty((lines, n), segs.dtype)
        w, wp, wb = self._work(plan, lines)
        self_ptr = 0
        if self_seg >= 0:
            # the self segment comes from a second buffer of the same layout; poison it in the first one
            other = segs.copy()
            segs = segs.copy()
            segs[self_seg] = np.nan
            dx = self.mem.upload(segs)
            dself = self.mem.upload(other)
            self_ptr = self.mem.ptr(dself)
        ok = self.api.fft_segmented(plan, self.mem.ptr(dx), self.mem.ptr(dout), lines, seg_len, lines * seg_len, forward, wp, wb,
                                    0, self_seg, self_ptr)
        return self.mem.download(dout) if ok else None

    def fft_columns_twiddled(self, x, col_offset, total, forward=True, peers=0):
        """x: [n][cols] natural-order column block; returns out[k][c] = FFT over i of column c times
        W_total^((col_offset + c) k) (conjugated when not forward), or None when the shape is not covered."""
        x = np.ascontiguousarray(x)
        n, cols = x.shape
        prec = 0 if x.dtype == np.complex64 else 1
        plan = self.plan(n, cuda_api.FFT_COMPLEX, prec)
        lg = int(total).bit_length() - 1
        shift = (lg + 1) // 2
        lo = self.mem.empty((1 << shift,), x.dtype)
        hi = self.mem.empty((1 << (lg - shift),), x.dtype)
        self.api.fill_twiddles(self.mem.ptr(lo), 1 << shift, 1, total, prec)
        self.api.fill_twiddles(self.mem.ptr(hi), 1 << (lg - shift), 1 << shift, total, prec)
        dx = self.mem.upload(x)
        dout = self.mem.empty((n, cols), x.dtype)
        nbytes = self.api.work_bytes_axis(plan, 1, cols)
        w = self.mem.alloc(nbytes) if nbytes else None
        if peers:
            # the fused-exchange variant: row block q goes to its own buffer (here: `peers` separate local buffers
            # standing in for the peers' receive buffers); reassembled for the comparison
            bufs = [self.mem.empty((n // peers, cols), x.dtype) for _ in range(peers)]
            ok = self.api.fft_columns_twiddled_p2p(plan, self.mem.ptr(dx), cols, forward, col_offset, self.mem.ptr(lo),
                                                   self.mem.ptr(hi), shift, total, [self.mem.ptr(b) for b in bufs],
                                                   self.mem.ptr(w) if nbytes else 0, nbytes)
            return np.concatenate([self.mem.download(b) for b in bufs], axis=0) if ok else None
        ok = self.api.fft_columns_twiddled(plan, self.mem.ptr(dx), self.mem.ptr(dout), cols, forward, col_offset,
                                           self.mem.ptr(lo), self.mem.ptr(hi), shift, total,
                                           self.mem.ptr(w) if nbytes else 0, nbytes)
        return self.mem.download(dout) if ok else None

    def fft(self, x, n=-1, axis=-1):
        return self._cfft(x, n, axis, True)

    def ifft(self, x, n=-1, axis=-1):
        return self._cfft(x, n, axis, False)

    def rfft(self, x, n=-1, axis=-1):
        x = np.ascontiguousarray(x)
        ax, outer, x_n, inner = self._split(x, axis)
        order, out_n = port.rfft_len(x_n, n)
        prec = 0 if x.dtype == np.float32 else 1
        plan = self.plan(order, cuda_api.FFT_REAL, prec)
        shape = list(x.shape); shape[ax] = out_n
        dx = self.mem.upload(x)
        dout = self.mem.empty(shape, _CPLX[x.dtype])
        w, wp, wb = self._work(plan, outer * inner)
        self.api.rfft(plan, self.mem.ptr(dx), self.mem.ptr(dout), outer, x_n, inner, wp, wb)
        return self.mem.download(dout)

    def irfft(self, x, n=-1, axis=-1):
        x = np.ascontiguousarray(x)
        ax, outer, x_n, inner = self._split(x, axis)
        order, out_n = port.irfft_len(x_n, n)
        prec = 0 if x.dtype == np.complex64 else 1
        plan = self.plan(order, cuda_api.FFT_REAL, prec)
        shape = list(x.shape); shape[ax] = out_n
        dx = self.mem.upload(x)
        dout = self.mem.empty(shape, _REAL[x.dtype])
        w, wp, wb = self._work(plan, outer * inner)
        self.api.irfft(plan, self.mem.ptr(dx), self.mem.ptr(dout), outer, x_n, inner, wp, wb)
        return self.mem.download(dout)

    def cmul(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        rows = int(np.prod(a.shape[:-1], dtype=np.int64))
        cols = a.shape[-1]
        da, db = self.mem.upload(a), self.mem.upload(b)
        dout = self.mem.empty(a.shape, a.dtype)
        self.api.cmul(self.mem.ptr(da), self.mem.ptr(db), self.mem.ptr(dout), _CODE[a.dtype], rows, cols,
                      b.size == a.size and rows > 1)
        return self.mem.download(dout)

    def filter(self, x, B, n=-1):
        """irfft(rfft(x, n) * B) along the last axis through dsc_cuda_filter."""
        x = np.ascontiguousarray(x)
        x2 = x.reshape(-1, x.shape[-1])
        order, bins = port.rfft_len(x2.shape[1], n)
        prec = 0 if x.dtype == np.float32 else 1
        plan = self.plan(order, cuda_api.FFT_REAL, prec)
        B = np.ascontiguousarray(B, dtype=_CPLX[x.dtype])
        assert B.size == bins
        dx, dB = self.mem.upload(x2), self.mem.upload(B)
        dout = self.mem.empty((x2.shape[0], 2 * order), x.dtype)
        nbytes = self.api.filter_work_bytes(plan, x2.shape[0] if self.work_lines is None else min(x2.shape[0], self.work_lines))
        w = self.mem.alloc(nbytes) if nbytes else None
        self.api.filter(plan, self.mem.ptr(dx), self.mem.ptr(dB), self.mem.ptr(dout), x2.shape[0], x2.shape[1],
                        self.mem.ptr(w) if nbytes else 0, nbytes)
        return self.mem.download(dout).reshape(x.shape[:-1] + (2 * order,))
