"""Kernel-level parity on the B200: the same cases as tests/test_kernels_emulated.py, but
through the real dsc_b200/libdsc.so (dsc_cuda_* C ABI) with device memory."""
import numpy as np
import pytest

from dsc_b200 import cuda_api
from oracle import port
from tests.devfft import DevFFT
from tests.util import load_golden, randn, rel_l2, TOL
# the emulated suite's cases run unchanged against the `dev` fixture defined below
from tests.test_kernels_emulated import (  # noqa: F401
    TIGHT, test_columns_with_outer_twiddle, test_all_axes_pad_crop, test_c2c_last_axis, test_cmul, test_fused_filter, test_many_lines_partial_blocks,
    test_non_pow2_lengths, test_rfft_axes_and_length_rules, test_rfft_irfft_dense_whole_blocks, test_rfft_irfft_last_axis,
    test_two_pass_along_strided_axis, test_two_pass_c2c, test_two_pass_real, test_two_pass_segmented_rows)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return DevFFT(cuda_api.LIBDSC, backend="torch")


@pytest.mark.parametrize("dtype,lg,outer,inner", [("complex64", 16, 3, 256), ("complex64", 18, 1, 64),
                                                   ("complex128", 15, 2, 128), ("complex64", 14, 4, 1024)])
def test_long_columns_gpu(dev, dtype, lg, outer, inner):
    """Larger shapes of tests/test_kernels_emulated.py::test_two_pass_along_strided_axis (several chunks per slab,
    several slabs, ring reuse)."""
    rng = np.random.default_rng(lg + inner)
    x = randn(rng, (outer, 1 << lg, inner), dtype)
    y = dev.fft(x, axis=1)
    assert rel_l2(y, port.fft(x, axis=1)) < TIGHT[dtype]
    assert rel_l2(dev.ifft(y, axis=1), x) < TIGHT[dtype]


def test_two_pass_chunked_work_buffer_gpu():
    d = DevFFT(cuda_api.LIBDSC, backend="torch", work_lines=2)
    rng = np.random.default_rng(3)
    x = randn(rng, (5, 1 << 15), "complex64")
    assert rel_l2(d.fft(x), port.fft(x)) < 1e-6
    xr = randn(rng, (3, 1 << 16), "float32")
    d1 = DevFFT(cuda_api.LIBDSC, backend="torch", work_lines=1)
    X = d1.rfft(xr)
    assert rel_l2(X, port.rfft(xr)) < 1e-6
    assert rel_l2(d1.irfft(X), xr) < 1e-6


GOLD = [(i, m, a) for i, m, a in load_golden() if m["op"] in ("fft", "ifft", "rfft", "irfft", "mul")]


@pytest.mark.parametrize("i,meta,arrs", GOLD, ids=[f"{i}-{m['op']}" for i, m, _ in GOLD])
def test_golden_vectors(dev, i, meta, arrs):
    """Outputs recorded from the unmodified reference library (tests/golden/make_golden.py)."""
    op = meta["op"]
    got = dev.cmul(arrs["x"], arrs["b"]) if op == "mul" else getattr(dev, op)(arrs["x"], meta["n"], meta["axis"])
    want = arrs["y"]
    assert got.shape == want.shape and got.dtype == want.dtype
    assert rel_l2(got, want) < TOL[want.dtype]      # north-star tolerance: 1e-5 / 1e-12


# the many-row cases wrap the ring of work rows of the persistent four-step launch several times
@pytest.mark.parametrize("dtype,lg,rows", [("complex64", 20, 3), ("complex64", 19, 3), ("complex64", 18, 5),
                                           ("complex64", 16, 9), ("complex64", 15, 1500), ("complex128", 19, 2),
                                           ("complex128", 17, 3), ("complex128", 14, 1100),
                                           ("complex64", 12, 4096), ("complex128", 12, 513)])
def test_large_vs_oracle(dev, dtype, lg, rows):
    rng = np.random.default_rng(lg)
    x = randn(rng, (rows, 1 << lg), dtype)
    y = dev.fft(x)
    assert rel_l2(y, port.fft(x)) < TIGHT[dtype]
    assert rel_l2(dev.ifft(y), x) < TIGHT[dtype]


@pytest.mark.parametrize("dtype,lg,rows", [("float32", 20, 3), ("float64", 18, 3), ("float32", 14, 257)])
def test_large_real_vs_oracle(dev, dtype, lg, rows):
    rng = np.random.default_rng(lg)
    x = randn(rng, (rows, 1 << lg), dtype)
    X = dev.rfft(x)
    assert rel_l2(X, port.rfft(x)) < TIGHT[dtype]
    assert rel_l2(dev.irfft(X), x) < TIGHT[dtype]
