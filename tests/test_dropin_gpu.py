"""Drop-in proof ON THE B200: the reference's unchanged wrappers on top of dsc_b200/libdsc.so.

tools/stage_reference_wrappers.py (run by build() in the container that has /root/reference) copies the reference's
python/dsc package, its python/tests/test_ops.py and a README-filterFFT binary compiled against the reference's
dsc_api.h into the git-ignored baseline/_ref/, which travels with the snapshot.  Here they run against the real CUDA
library (the CPU container runs the same files over the emulated build: tests/test_dropin_reference.py).
  _bindings.py:31-35 loads ./libdsc.so next to itself;  test_ops.py:458-523 are the FFT / fftfreq tests."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
LIB = os.path.join(ROOT, "dsc_b200", "libdsc.so")

pytestmark = pytest.mark.gpu


def test_reference_python_suite_on_cuda_library(tmp_path):
    src = os.path.join(STAGED, "python")
    assert os.path.isdir(os.path.join(src, "dsc")), "baseline/_ref/python missing: run tools/stage_reference_wrappers.py where /root/reference exists"
    assert os.path.exists(LIB), "dsc_b200/libdsc.so not built"
    shutil.copytree(src, tmp_path / "python")
    os.symlink(LIB, tmp_path / "python" / "dsc" / "libdsc.so")
    env = dict(os.environ, PYTHONPATH=str(tmp_path / "python"))
    r = subprocess.run([sys.executable, "-m", "pytest", str(tmp_path / "python" / "tests" / "test_ops.py"), "-q", "-x",
                        "-p", "no:cacheprovider"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "16 passed" in r.stdout
    # the library that served the run is the CUDA one: its init line names the device
    log = r.stdout + r.stderr
    assert "no CUDA device" not in log


def test_reference_cpp_api_on_cuda_library():
    exe = os.path.join(STAGED, "cpp", "filter_readme")
    assert os.path.exists(exe), "baseline/_ref/cpp/filter_readme missing: run tools/stage_reference_wrappers.py"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "samples=8319" in r.stdout
