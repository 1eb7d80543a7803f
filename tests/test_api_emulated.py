"""Host logic of libdsc.so (shape rules, plan cache, arenas, residency, tracer, chunked
transfers) exercised WITHOUT a GPU through tests/emul/libdsc_emul.so: the product's host
runtime compiled unchanged, with host memory standing in for the device and the kernels
running on pthreads.  The B200 run of the same cases is tests/test_gpu_api.py."""
import os
import subprocess

import pytest

from tests import api_cases
from tests.util import use_library

EMUL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
EMUL_SO = os.path.join(EMUL_DIR, "libdsc_emul.so")


@pytest.fixture(scope="module")
def dsc():
    subprocess.run(["make", "-s", "-j8", "-C", EMUL_DIR], check=True)
    import dsc_b200
    use_library(dsc_b200, EMUL_SO)
    os.environ["DSC_CHUNK_BYTES"] = str(1 << 16)       # force the multi-chunk transfer pipeline
    dsc_b200.init(1 << 28, 1 << 26)
    yield dsc_b200
    dsc_b200.shutdown()
    os.environ.pop("DSC_CHUNK_BYTES", None)


CASES = [api_cases.check_golden, api_cases.check_shapes_appendix_a, api_cases.check_out_param,
         api_cases.check_vs_oracle_sweep, api_cases.check_filter_pipeline, api_cases.check_plan_cache,
         api_cases.check_memory_accounting, api_cases.check_residency_modes, api_cases.check_device_pointwise,
         api_cases.check_device_ops,
         api_cases.check_traces,
         api_cases.check_composed_paths]


@pytest.mark.parametrize("case", CASES, ids=[c.__name__ for c in CASES])
def test_api(dsc, case):
    case(dsc)
