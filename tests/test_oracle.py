"""Pins the CPU restatement (oracle/dsc_fft_oracle.c) before anything trusts it.

(1) golden vectors recorded from the unmodified reference library,
(2) the live reference library when oracle/_ref exists,
(3) NumPy float64, the oracle of the reference's own test (python/tests/test_ops.py:458-489).
"""
import numpy as np
import pytest

from oracle import port, ref_harness
from tests.util import TOL, load_golden, randn, rel_l2

GOLD = list(load_golden())


@pytest.mark.parametrize("i,meta,arrs", GOLD, ids=[f"{i}-{m['op']}" for i, m, _ in GOLD])
def test_port_matches_golden(i, meta, arrs):
    op = meta["op"]
    if op == "plan_sweep":
        pytest.skip("plan-cache fixture is for the product library")
    if op == "filter":
        got = port.filter_fft(arrs["x"], arrs["b"], meta["n"])
    elif op == "mul":
        got = port.cmul(arrs["x"], arrs["b"])
    elif op.startswith(("unary:", "binary:")):
        kind, name = op.split(":")
        got = port.unary(name, arrs["x"]) if kind == "unary" else port.binary(name, arrs["x"], arrs["b"])
        want = arrs["y"]
        assert got.shape == want.shape and got.dtype == want.dtype
        if name in ("real", "imag", "conj", "add", "sub"):
            assert np.array_equal(got, want)
        else:   # libm / fused-multiply-add differences of one rounding
            assert rel_l2(got, want) < TOL[want.dtype] / 50
        return
    else:
        got = getattr(port, op)(arrs["x"], meta["n"], meta["axis"])
    want = arrs["y"]
    assert got.shape == want.shape and got.dtype == want.dtype
    if op == "mul":
        # whether gcc fuses a*b - c*d differs between the reference's same-shape and
        # broadcast loops (dsc.cpp:1186-1245); one rounding either way
        assert rel_l2(got, want) < TOL[want.dtype] / 50
        return
    # same compiler flags on both sides (oracle/Makefile) => bit-identical
    assert np.array_equal(got, want), f"rel-L2 {rel_l2(got, want):.3e}"


def test_shape_rules_appendix_a():
    # SURVEY.md Appendix A, all probed against the compiled reference
    assert port.pow2_n(10) == 16 and port.pow2_n(16) == 16 and port.pow2_n(1) == 1
    assert port.fft_len(10) == 16 and port.fft_len(10, 5) == 8
    assert port.rfft_len(10) == (8, 9) and port.rfft_len(10, 4) == (2, 3)
    assert port.irfft_len(9) == (8, 16) and port.irfft_len(9, 9) == (8, 16)
    assert port.irfft_len(9, 16) == (16, 32) and port.irfft_len(9, 5) == (4, 8)
    with pytest.raises(ValueError):
        port.rfft_len(1)
    with pytest.raises(ValueError):
        port.irfft_len(1)


def test_twiddle_layout():
    # dsc_fft.h:42-49: block for size t at real offset 2*(t/2-1); real plans carry one more block
    tw = port.twiddles(8, False, np.float64)
    assert tw.size == 14
    k = np.arange(4)
    np.testing.assert_allclose(tw[6:14:2], np.cos(-2 * np.pi * k / 8), atol=1e-15)
    np.testing.assert_allclose(tw[7:14:2], np.sin(-2 * np.pi * k / 8), atol=1e-15)
    assert port.twiddles(8, True, np.float32).size == 30


@pytest.mark.parametrize("dtype", ["complex64", "complex128", "float32", "float64"])
@pytest.mark.parametrize("lg", [1, 5, 10, 14])
def test_port_vs_numpy(dtype, lg):
    rng = np.random.default_rng(lg)
    x = randn(rng, (3, 1 << lg), dtype)
    want = np.fft.fft(x.astype(np.complex128 if x.dtype.kind == "c" else np.float64), axis=-1)
    got = port.fft(x)
    slack = 40  # the reference itself sits at 2e-7 / 9e-16 from float64
    assert rel_l2(got, want) < TOL[x.dtype] / 10 * slack / 40
    back = port.ifft(got)
    assert rel_l2(back, x.astype(back.dtype)) < TOL[x.dtype]
    if x.dtype.kind == "f":
        r = port.rfft(x)
        assert rel_l2(r, np.fft.rfft(x.astype(np.float64), axis=-1)) < TOL[x.dtype] / 10
        assert np.all(r[..., 0].imag == 0) and np.all(r[..., -1].imag == 0)
        assert rel_l2(port.irfft(r), x) < TOL[x.dtype]


@pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built")
def test_port_vs_live_reference():
    ref = ref_harness.RefLib(main_mem=1 << 29, scratch_mem=1 << 27)
    rng = np.random.default_rng(7)
    try:
        for dtype, shape, n, axis in [("complex64", (5, 4096), -1, -1), ("complex128", (2, 1 << 15), -1, 1),
                                      ("float32", (7, 100, 3), 256, 1), ("complex64", (1 << 17,), -1, 0),
                                      ("float64", (6, 3, 40), 16, -1)]:
            x = randn(rng, shape, dtype)
            for op in ("fft", "ifft"):
                assert np.array_equal(getattr(port, op)(x, n, axis), getattr(ref, op)(x, n, axis))
        for dtype, shape, n, axis in [("float32", (3, 1 << 15), -1, -1), ("float64", (1 << 16,), -1, 0),
                                      ("float32", (50, 6), 128, 0)]:
            x = randn(rng, shape, dtype)
            r = ref.rfft(x, n, axis)
            assert np.array_equal(port.rfft(x, n, axis), r)
            assert np.array_equal(port.irfft(r, -1, axis), ref.irfft(r, -1, axis))
    finally:
        ref.close()
