"""Error behaviour of the drop-in library: no error codes, a message on stderr and exit(EXIT_FAILURE)
(reference convention, dsc/include/dsc.h:14-28) -- including the one this build adds: FFT entry points on a
host without a CUDA device abort loudly instead of falling back to a CPU path."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, lib=None):
    env = dict(os.environ, PYTHONPATH=ROOT)
    prog = "import os, sys, numpy as np\nimport dsc_b200 as dsc\n"
    if lib:
        prog += f"dsc.LIBDSC = {lib!r}\n"
    prog += textwrap.dedent(code)
    return subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=env, timeout=300)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:       # noqa: BLE001
        return False


@pytest.mark.skipif(_has_gpu(), reason="this box has a GPU")
def test_product_library_aborts_without_a_device():
    """The nvcc-built product library loads on a GPU-less host, serves host ops, and refuses to transform."""
    r = _run("""
        dsc.init(1 << 24, 1 << 22)
        a = dsc.from_numpy(np.arange(8, dtype=np.float32))
        print("host op ok", (a * a).numpy()[3])
        sys.stdout.flush()
        dsc.fft(a)
        print("NOT REACHED")
    """)
    assert "host op ok 9.0" in r.stdout
    assert "NOT REACHED" not in r.stdout
    assert r.returncode != 0
    assert "needs a CUDA device" in r.stderr and "no CPU implementation" in r.stderr


EMUL = os.path.join(ROOT, "tests", "emul", "libdsc_emul.so")
CASES = {
    "rfft of complex input": ("dsc.rfft(np.zeros(8, np.complex64))", "RFFT input must be real"),
    "irfft of real input": ("dsc.irfft(np.zeros(9, np.float32))", "IRFFT input must be complex"),
    "out= with the wrong shape": ("dsc.fft(np.zeros((2, 8), np.complex64), out=dsc.from_numpy(np.zeros((2, 4), np.complex64)))", "DSC_ASSERT"),
    "out= with the wrong dtype": ("dsc.fft(np.zeros(8, np.float32), out=dsc.from_numpy(np.zeros(8, np.complex128)))", "DSC_ASSERT"),
    "axis out of range": ("dsc.fft(np.zeros((2, 8), np.float32), axis=5)", "DSC_ASSERT"),
    "rfft of a single sample": ("dsc.rfft(np.zeros(1, np.float32))", "DSC_ASSERT"),
    "irfft of a single bin": ("dsc.irfft(np.zeros(1, np.complex64))", "DSC_ASSERT"),
    "arena exhausted": ("dsc.from_numpy(np.zeros(1 << 23, np.float64))", "error allocating"),
}


def _check_fatal(snippet, needle, lib):
    r = _run(f"""
        dsc.init(1 << 24, 1 << 22)
        {snippet}
        print("NOT REACHED")
    """, lib=lib)
    assert r.returncode != 0 and "NOT REACHED" not in r.stdout, r.stdout + r.stderr
    assert needle in r.stderr, r.stderr


@pytest.mark.parametrize("name", list(CASES))
def test_fatal_conditions_emulated(name):
    subprocess.run(["make", "-s", "-j8", "-C", os.path.dirname(EMUL)], check=True)
    _check_fatal(*CASES[name], lib=EMUL)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_fatal_conditions_gpu(name):
    _check_fatal(*CASES[name], lib=None)
