"""The product library (nvcc build, dsc_b200/libdsc.so) must load without a GPU and export every
entry point that include/dsc.h and include/dsc_cuda.h declare.  No compute is attempted here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dsc_b200", "libdsc.so")


@pytest.fixture(scope="module")
def lib_path():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "dsc_b200", "csrc")], check=True)
    return LIB


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"//[^\n]*", "", text)
    names = set(re.findall(r"\b(dsc_[a-z0-9_]+)\s*\(", text))
    # macro-generated families in dsc.h
    for fam in re.findall(r"DSC_DECL_[A-Z]+\((dsc_[a-z0-9_]+)\)", text):
        names.add(fam)
    return names - {"dsc_pow2_n", "dsc_tensor_dim", "dsc_new_like", "dsc_new_view", "dsc_inf", "dsc_zero", "dsc_pi",
                    "dsc_is_type", "dsc_is_complex", "dsc_is_real", "dsc_complex", "dsc_new_tensor_", "dsc_complex_t",
                    "dsc_cuda_plan", "dsc_dtype", "dsc_fft_type"}


def test_library_loads_and_exports_everything(lib_path):
    lib = ctypes.CDLL(lib_path)
    wanted = _declared("dsc.h") | _declared("dsc_cuda.h")
    assert len(wanted) >= 60 + 9
    missing = [n for n in sorted(wanted) if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cufft_or_oracle_linked(lib_path):
    out = subprocess.run(["ldd", lib_path], capture_output=True, text=True).stdout
    assert "cufft" not in out.lower()
    assert "oracle" not in out.lower() and "dsc_ref" not in out.lower()


def test_product_sources_never_touch_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "dsc_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|libdsc_oracle|libdsc_ref|cufft", text):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
