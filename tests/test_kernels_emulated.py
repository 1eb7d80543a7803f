"""Functional check of the CUDA sources WITHOUT a GPU.

tests/emul compiles dsc_b200/csrc/*.cu with g++ against a pthread shim of the CUDA
execution model (blocks, threads, __syncthreads, dynamic shared memory), so the index
arithmetic of every kernel -- Stockham scatter and padding, pad/crop predicates, strided
thread mapping, real un-mixing, four-step geometry and chunking -- is exercised here against
the oracle.  This is test infrastructure; the real parity tests are tests/test_gpu_*.py.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import port
from tests.devfft import DevFFT
from tests.util import TOL, randn, rel_l2

EMUL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
# the reference itself is ~3e-7 / 9e-16 away from exact; leave the rest of the budget unused
TIGHT = {"complex64": 1e-6, "float32": 1e-6, "complex128": 4e-15, "float64": 4e-15}


@pytest.fixture(scope="module")
def dev():
    subprocess.run(["make", "-s", "-j8", "-C", EMUL_DIR], check=True)
    return DevFFT(os.path.join(EMUL_DIR, "libdsc_emul.so"))


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("lg", list(range(0, 15)))
def test_c2c_last_axis(dev, dtype, lg):
    if dtype == "complex128" and lg > 13:
        pytest.skip("two-pass for complex128")
    rng = np.random.default_rng(lg)
    rows = 3 if lg < 12 else 1
    x = randn(rng, (rows, 1 << lg), dtype)
    y = dev.fft(x)
    assert y.shape == x.shape and y.dtype == x.dtype
    assert rel_l2(y, port.fft(x)) < TIGHT[dtype]
    z = dev.ifft(y)
    assert rel_l2(z, port.ifft(y)) < TIGHT[dtype]
    assert rel_l2(z, x) < TIGHT[dtype]


@pytest.mark.parametrize("dtype", ["float32", "float64", "complex64", "complex128"])
def test_all_axes_pad_crop(dev, dtype):
    # python/tests/test_ops.py:458-489: every axis of a 4-D tensor, n in {crop, copy, pad}
    rng = np.random.default_rng(5)
    for axis in range(4):
        shape = [3, 4, 2, 5]
        shape[axis] = 16
        x = randn(rng, shape, dtype)
        for n in (8, 16, 32, -1):
            for ax in (axis, axis - 4):
                y = dev.fft(x, n, ax)
                want = port.fft(x, n, ax)
                assert y.shape == want.shape and y.dtype == want.dtype
                assert rel_l2(y, want) < TIGHT[dtype]
            yi = dev.ifft(x, n, axis)
            assert rel_l2(yi, port.ifft(x, n, axis)) < TIGHT[dtype]


def test_non_pow2_lengths(dev):
    rng = np.random.default_rng(6)
    x = randn(rng, (7, 10), "float32")
    assert dev.fft(x).shape == (7, 16)
    assert rel_l2(dev.fft(x), port.fft(x)) < 1e-6
    assert rel_l2(dev.fft(x, 5), port.fft(x, 5)) < 1e-6          # n=5 -> 8, crop to 8
    assert rel_l2(dev.fft(x, axis=0), port.fft(x, axis=0)) < 1e-6  # 7 -> 8 along axis 0
    x = randn(rng, (100, 3), "complex128")
    assert rel_l2(dev.ifft(x, axis=0), port.ifft(x, axis=0)) < 4e-15


def test_many_lines_partial_blocks(dev):
    rng = np.random.default_rng(8)
    for shape in [(37, 64), (301, 16), (5, 512), (1000, 4), (19, 2), (33, 1)]:
        x = randn(rng, shape, "complex64")
        assert rel_l2(dev.fft(x), port.fft(x)) < 1e-6
    x = randn(rng, (64, 37), "complex64")        # strided, 37 lines
    assert rel_l2(dev.fft(x, axis=0), port.fft(x, axis=0)) < 1e-6


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("lg", list(range(0, 14)))
def test_rfft_irfft_last_axis(dev, dtype, lg):
    # order 2^lg <-> 2^(lg+1) real samples
    rng = np.random.default_rng(100 + lg)
    rows = 3 if lg < 12 else 1
    x = randn(rng, (rows, 2 << lg), dtype)
    y = dev.rfft(x)
    want = port.rfft(x)
    assert y.shape == want.shape and y.dtype == want.dtype
    assert rel_l2(y, want) < TIGHT[dtype]
    assert np.all(y[:, 0].imag == 0) and np.all(y[:, -1].imag == 0)   # dsc_fft.h:220-225
    z = dev.irfft(want)
    wz = port.irfft(want)
    assert z.shape == wz.shape and z.dtype == wz.dtype
    assert rel_l2(z, wz) < TIGHT[dtype]
    assert rel_l2(z, x) < TIGHT[dtype]


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("lg", [8, 9, 10, 11, 12])
def test_rfft_irfft_dense_whole_blocks(dev, dtype, lg):
    """16 full-length last-axis lines: whole blocks for every block shape, so the dense packed-real kernels run
    (vector accesses, every thread builds its own packed points from the bin pairs); DC / Nyquist imaginary
    parts of the irfft input are ignored (dsc_fft.h:227-228)."""
    rng = np.random.default_rng(300 + lg)
    x = randn(rng, (16, 2 << lg), dtype)
    y = dev.rfft(x)
    want = port.rfft(x)
    assert rel_l2(y, want) < TIGHT[dtype]
    assert np.all(y[:, 0].imag == 0) and np.all(y[:, -1].imag == 0)
    bins = want.copy()
    bins[:, 0] += 3j
    bins[:, -1] -= 2j
    z = dev.irfft(bins)
    assert rel_l2(z, port.irfft(want)) < TIGHT[dtype]
    assert rel_l2(z, x) < TIGHT[dtype]


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_rfft_axes_and_length_rules(dev, dtype):
    rng = np.random.default_rng(9)
    cdt = np.complex64 if dtype == "float32" else np.complex128
    for axis in range(3):
        shape = [3, 4, 5]
        shape[axis] = 32
        x = randn(rng, shape, dtype)
        for n in (-1, 16, 64, 10):
            y = dev.rfft(x, n, axis)
            want = port.rfft(x, n, axis)
            assert y.shape == want.shape
            assert rel_l2(y, want) < TIGHT[dtype]
            for nb in (-1, 3, want.shape[axis] + 3):       # irfft n counts input bins
                z = dev.irfft(want, nb, axis)
                wz = port.irfft(want, nb, axis)
                assert z.shape == wz.shape
                assert rel_l2(z, wz) < TIGHT[dtype]
    # DC / Nyquist imaginary parts of the input are ignored by irfft (dsc_fft.h:227-228)
    X = randn(rng, (2, 17), cdt)
    assert rel_l2(dev.irfft(X), port.irfft(X)) < TIGHT[dtype]
    # Appendix A
    x = randn(rng, (10,), dtype)
    assert dev.rfft(x).shape == (9,) and dev.rfft(x, 4).shape == (3,)
    X = port.rfft(x)
    assert dev.irfft(X).shape == (16,) and dev.irfft(X, 16).shape == (32,) and dev.irfft(X, 5).shape == (8,)
    x2 = randn(rng, (2,), dtype)
    assert rel_l2(dev.rfft(x2), port.rfft(x2)) < TIGHT[dtype]
    assert rel_l2(dev.irfft(port.rfft(x2)), port.irfft(port.rfft(x2))) < TIGHT[dtype]


@pytest.mark.parametrize("dtype,lg", [("complex64", 15), ("complex64", 17), ("complex128", 14), ("complex128", 16)])
def test_two_pass_c2c(dev, dtype, lg):
    rng = np.random.default_rng(lg)
    x = randn(rng, (2, 1 << lg), dtype)
    y = dev.fft(x)
    assert rel_l2(y, port.fft(x)) < TIGHT[dtype]
    z = dev.ifft(y)
    assert rel_l2(z, x) < TIGHT[dtype]
    # pad and crop through the first pass predicate; real input cast
    xs = x[:, : (1 << lg) - 1000]
    assert rel_l2(dev.fft(xs), port.fft(xs)) < TIGHT[dtype]
    xr = xs.real.copy()
    assert rel_l2(dev.fft(xr), port.fft(xr)) < TIGHT[dtype]


@pytest.mark.parametrize("dtype,lg,segs", [("complex64", 15, 2), ("complex64", 15, 8), ("complex64", 16, 4),
                                            ("complex128", 14, 2), ("complex128", 15, 4)])
def test_two_pass_segmented_rows(dev, dtype, lg, segs):
    """The receive buffer of the multi-GPU exchange, [peer][line][part], transformed where it lies."""
    rng = np.random.default_rng(lg * 10 + segs)
    n = 1 << lg
    x = randn(rng, (3, n), dtype)
    parts = np.ascontiguousarray(x.reshape(3, segs, n // segs).transpose(1, 0, 2))
    y = dev.fft_segmented(parts)
    assert y is not None
    assert rel_l2(y, port.fft(x)) < TIGHT[dtype]
    z = dev.fft_segmented(np.ascontiguousarray(y.reshape(3, segs, n // segs).transpose(1, 0, 2)), forward=False)
    assert rel_l2(z, x) < TIGHT[dtype]
    # one segment served from a second buffer (the slab a rank keeps instead of sending it to itself)
    assert rel_l2(dev.fft_segmented(parts, self_seg=segs - 1), port.fft(x)) < TIGHT[dtype]
    assert dev.fft_segmented(parts[:, :, :64].copy()) is None      # single-pass length: not covered, caller copies


@pytest.mark.parametrize("dtype,lg,outer,inner", [("complex64", 15, 2, 64), ("complex128", 14, 2, 32),
                                                   ("complex64", 13, 2, 128), ("complex64", 14, 1, 64),
                                                   ("complex128", 13, 1, 64)])
def test_two_pass_along_strided_axis(dev, dtype, lg, outer, inner):
    """Long columns (a NON-last axis): one launch, both passes of a two-pass decomposition as column tiles --
    lengths beyond one shared-memory pass, and the single-pass lengths 2^13 / 2^14 whose plans carry a column
    decomposition as well."""
    rng = np.random.default_rng(lg + inner)
    n = 1 << lg
    x = randn(rng, (outer, n, inner), dtype)
    y = dev.fft(x, axis=1)
    assert rel_l2(y, port.fft(x, axis=1)) < TIGHT[dtype]
    assert rel_l2(dev.ifft(y, axis=1), x) < TIGHT[dtype]
    xs = x[:, : n - 300, :]                              # zero-padded columns, and real input cast
    assert rel_l2(dev.fft(xs, n=n, axis=1), port.fft(xs, n=n, axis=1)) < TIGHT[dtype]
    xr = np.ascontiguousarray(xs.real)
    assert rel_l2(dev.fft(xr, n=n, axis=1), port.fft(xr, n=n, axis=1)) < TIGHT[dtype]


@pytest.mark.parametrize("dtype,lg,cols", [("complex64", 15, 64), ("complex64", 13, 128), ("complex128", 14, 32)])
def test_columns_with_outer_twiddle(dev, dtype, lg, cols):
    """First local step of the multi-GPU four-step in one launch: column transforms of a natural-order block,
    times the outer twiddle, k-major."""
    rng = np.random.default_rng(lg + cols)
    n, all_cols, off = 1 << lg, 4 * cols, 2 * cols
    x = randn(rng, (n, cols), dtype)
    y = dev.fft_columns_twiddled(x, off, n * all_cols)
    assert y is not None
    k = np.arange(n, dtype=np.float64)[:, None]
    c = (off + np.arange(cols, dtype=np.float64))[None, :]
    want = port.fft(x, axis=0).astype(np.complex128) * np.exp(-2j * np.pi * ((k * c) % (n * all_cols)) / (n * all_cols))
    assert rel_l2(y, want.astype(dtype)) < TIGHT[dtype] * 2
    yi = dev.fft_columns_twiddled(x, off, n * all_cols, forward=False)
    wanti = np.fft.ifft(x.astype(np.complex128), axis=0) * np.exp(2j * np.pi * ((k * c) % (n * all_cols)) / (n * all_cols))
    assert rel_l2(yi, wanti.astype(dtype)) < TIGHT[dtype] * 2
    assert dev.fft_columns_twiddled(x[:1024].copy(), off, 1024 * all_cols) is None      # no column decomposition
    # the exchange fused into the epilogue: the same numbers, row blocks scattered over per-peer buffers
    for peers in (2, 8):
        assert np.array_equal(dev.fft_columns_twiddled(x, off, n * all_cols, peers=peers), y)


def test_two_pass_chunked_work_buffer():
    d = DevFFT(os.path.join(EMUL_DIR, "libdsc_emul.so"), work_lines=2)
    rng = np.random.default_rng(3)
    x = randn(rng, (5, 1 << 15), "complex64")
    assert rel_l2(d.fft(x), port.fft(x)) < 1e-6
    xr = randn(rng, (3, 1 << 16), "float32")
    d1 = DevFFT(os.path.join(EMUL_DIR, "libdsc_emul.so"), work_lines=1)
    X = d1.rfft(xr)
    assert rel_l2(X, port.rfft(xr)) < 1e-6
    assert rel_l2(d1.irfft(X), xr) < 1e-6


def test_columns_one_row_work_buffer():
    """A work buffer that holds ONE row of the column launch with several rows to do (ring == 1): the second pass of a
    row must be ticketed before the first pass of the next row, or the persistent grid spins forever
    (fft_launch.cu, four_step_columns_launch: lag 0)."""
    from dsc_b200 import cuda_api
    d = DevFFT(os.path.join(EMUL_DIR, "libdsc_emul.so"))
    rng = np.random.default_rng(31)
    outer, n, inner = 2, 8192, 128
    x = randn(rng, (outer, n, inner), "complex64")
    plan = d.plan(n, cuda_api.FFT_COMPLEX, 0)
    nbytes = d.api.work_bytes_axis(plan, 1, inner)           # sized for one outer slab only
    assert 0 < nbytes < d.api.work_bytes_axis(plan, outer, inner)
    w = d.mem.alloc(nbytes)
    dx, dout = d.mem.upload(x), d.mem.empty(x.shape, x.dtype)
    d.api.fft(plan, d.mem.ptr(dx), 2, d.mem.ptr(dout), outer, n, inner, True, d.mem.ptr(w), nbytes)
    assert rel_l2(d.mem.download(dout), port.fft(x, axis=1)) < 1e-6


@pytest.mark.parametrize("dtype,lg", [("float32", 15), ("float64", 14)])
def test_two_pass_real(dev, dtype, lg):
    rng = np.random.default_rng(lg)
    x = randn(rng, (2, 2 << lg), dtype)
    y = dev.rfft(x)
    want = port.rfft(x)
    assert y.shape == want.shape
    assert rel_l2(y, want) < TIGHT[dtype]
    assert np.all(y[:, 0].imag == 0) and np.all(y[:, -1].imag == 0)
    z = dev.irfft(want)
    assert rel_l2(z, port.irfft(want)) < TIGHT[dtype]
    xs = x[:, :-777]                                   # zero-padded real input
    assert rel_l2(dev.rfft(xs, n=2 << lg), port.rfft(xs, n=2 << lg)) < TIGHT[dtype]
    Xs = want[:, :-5]                                  # fewer bins than order+1
    assert rel_l2(dev.irfft(Xs, n=want.shape[1]), port.irfft(Xs, n=want.shape[1])) < TIGHT[dtype]
    bins = want.copy()                                 # imaginary parts of DC / Nyquist are ignored (dsc_fft.h:227-228)
    bins[:, 0] += 3j
    bins[:, -1] -= 2j
    assert rel_l2(dev.irfft(bins), port.irfft(want)) < TIGHT[dtype]


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_fused_filter(dev, dtype):
    """irfft(rfft(x, n) * B) as one kernel (single pass) and as four-step + bin-pair kernel (large order)."""
    rng = np.random.default_rng(21)
    for rows, xn, n in [(3, 100, 256), (1, 8192, 16384), (5, 2, 2), (2, 64, 4), (2, 3000, 1 << 16)]:
        if dtype == "float64" and n > 1 << 15:
            n = 1 << 15
        x = randn(rng, (rows, xn), dtype)
        b = randn(rng, (min(xn, 37),), dtype)
        B = port.rfft(b, n)
        want = port.irfft(port.cmul(port.rfft(x, n), B))
        got = dev.filter(x, B, n)
        assert got.shape == want.shape and got.dtype == want.dtype
        assert rel_l2(got, want) < TIGHT[dtype] * 3, (rows, xn, n, rel_l2(got, want))


def test_cmul(dev):
    rng = np.random.default_rng(1)
    for dt in ("complex64", "complex128"):
        a = randn(rng, (5, 129), dt)
        b = randn(rng, (129,), dt)
        assert rel_l2(dev.cmul(a, b), a * b) < TOL[np.dtype(dt)] / 10
        b2 = randn(rng, (5, 129), dt)
        assert rel_l2(dev.cmul(a, b2), a * b2) < TOL[np.dtype(dt)] / 10
