"""Parity of the drop-in C ABI (dsc_fft / dsc_ifft / dsc_rfft / dsc_irfft through libdsc.so) on
the B200 against the oracle and the golden vectors, plus size-independent properties at the
BASELINE sizes."""
import numpy as np
import pytest

from oracle import port
from tests import api_cases
from tests.util import use_library
from tests.util import randn, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dsc():
    import dsc_b200
    from dsc_b200 import cuda_api
    use_library(dsc_b200, cuda_api.LIBDSC)
    dsc_b200.init(6 << 30, 1 << 30)
    yield dsc_b200
    dsc_b200.shutdown()


CASES = [api_cases.check_golden, api_cases.check_shapes_appendix_a, api_cases.check_out_param,
         api_cases.check_vs_oracle_sweep, api_cases.check_filter_pipeline, api_cases.check_plan_cache,
         api_cases.check_memory_accounting, api_cases.check_residency_modes, api_cases.check_device_pointwise,
         api_cases.check_device_ops,
         api_cases.check_traces,
         api_cases.check_composed_paths]


@pytest.mark.parametrize("case", CASES, ids=[c.__name__ for c in CASES])
def test_api(dsc, case):
    case(dsc)


def test_oracle_sweep_large(dsc):
    api_cases.check_vs_oracle_sweep(dsc, max_lg=17)
    api_cases.check_plan_cache(dsc, max_lg=16)


def test_config2_properties(dsc):
    """BASELINE configs[1] at full width (complex64, 4096 points) on a batch the oracle cannot cover:
    round trip, linearity, Parseval, and an oracle check on sampled rows."""
    rng = np.random.default_rng(2)
    rows = 16384
    x = randn(rng, (rows, 4096), "complex64")
    tx = dsc.from_numpy(x)
    ty = dsc.fft(tx)
    y = ty.numpy()
    sample = rng.choice(rows, 64, replace=False)
    assert rel_l2(y[sample], port.fft(x[sample])) < 1e-5
    tz = dsc.ifft(ty)
    assert rel_l2(tz.numpy(), x) < 1e-5                                    # ifft(fft(x)) == x
    e_t = np.sum(np.abs(x.astype(np.complex128)) ** 2, axis=1)
    e_f = np.sum(np.abs(y.astype(np.complex128)) ** 2, axis=1) / 4096
    assert np.max(np.abs(e_f / e_t - 1)) < 1e-5                            # Parseval per line
    a = np.float32(0.75)
    x2 = randn(rng, (256, 4096), "complex64")
    lin = dsc.fft(a * x[:256] + x2).numpy()
    assert rel_l2(lin, a * y[:256] + dsc.fft(x2).numpy()) < 1e-5           # linearity


def test_config3_properties(dsc):
    """BASELINE configs[2] shape per line (float64, 262144 samples), reduced batch."""
    rng = np.random.default_rng(3)
    x = randn(rng, (8, 262144), "float64")
    X = dsc.rfft(x)
    assert X.shape == (8, 131073) and X.dtype == np.complex128
    assert rel_l2(X.numpy()[:2], port.rfft(x[:2])) < 1e-12
    assert rel_l2(dsc.irfft(X).numpy(), x) < 1e-12
    Xn = X.numpy()
    assert np.all(Xn[:, 0].imag == 0) and np.all(Xn[:, -1].imag == 0)


def test_config4_properties(dsc):
    """BASELINE configs[3] per channel (float32, 2^20 samples): fused filter vs the unfused
    three-call pipeline and vs the oracle on one channel."""
    rng = np.random.default_rng(4)
    s = randn(rng, (4, 1 << 20), "float32")
    b = np.zeros(1 << 20, np.float32)
    b[:128] = np.random.default_rng(5).standard_normal(128).astype(np.float32)
    B = dsc.rfft(b)
    fused = dsc.fft_filter(s, B).numpy()
    unfused = dsc.irfft(dsc.rfft(s) * B).numpy()
    assert rel_l2(fused, unfused) < 1e-5
    assert rel_l2(fused[0], port.filter_fft(s[0], b, 1 << 20)) < 1e-5


def test_huge_single_transform(dsc):
    """2^25 complex64 points: past the four-step plan range, composed from 2^13 x 2^12 sub-plans.  The
    oracle needs ~1 s for this size; sampled bins are also checked against a float64 DFT."""
    rng = np.random.default_rng(25)
    n = 1 << 25
    x = randn(rng, (n,), "complex64")
    y = dsc.fft(x).numpy()
    want = port.fft(x)
    assert rel_l2(y, want) < 1e-5
    ks = [0, 1, 12345, n // 2, n - 1]
    t = np.arange(n, dtype=np.float64)
    for k in ks:
        ref = np.sum(x.astype(np.complex128) * np.exp(-2j * np.pi * ((k * t) % n) / n))
        assert abs(y[k] - ref) / abs(ref) < 1e-4
    assert rel_l2(dsc.ifft(dsc.from_numpy(y)).numpy(), x) < 1e-5
