"""Every implementation of the lengths beyond one shared-memory pass, against the oracle on the B200:
   default      TMA-fed two-pass launch (fft_tma.cuh), thread-block clusters for 2^15-point lines (fft_cluster.cuh)
   clusters     one line per cluster for 2^14 .. 2^17 (2, 4, 8 and 16 blocks, distributed shared memory)
   registers    the register-direct persistent launch (four_step_fused), no TMA, no clusters
   tma-e16      the TMA-fed launch with 16 points per thread on 32 KiB tiles, two blocks per SM, for every pair of passes
                that has the tables (default only for passes of at most 256 points)
   tma-e32      ... with 32 points per thread on 64 KiB tiles everywhere
   tma-direct   finished tiles stored from the registers, buffers released after the last exchange (opt-in variant)
   registers-e16  the register-direct launch with 16 points per thread, four blocks per SM, for every pair of passes that
                has the tables (default only for passes of at most 256 points); "registers" forces 32 points
   real-sweep   packed-real transforms with the bin-pair step as a separate sweep instead of fused into the TMA-fed launch
   real-fused-f32  the float32 filter through the fused launch too (default for float64 only: slower for float32)
   one-shot     the longest single-pass lines (one block per SM) as one-shot blocks instead of persistent blocks with an
                L2 prefetch of their next line
The selection is made through environment variables the library reads once, hence one subprocess per variant."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "two_pass_worker.py")

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,env", [("default", {}), ("clusters", {"DSC_CLUSTER_LGS": "14,15,16,17"}),
                                      ("clusters-unpipelined", {"DSC_CLUSTER_LGS": "14,15,16,17", "DSC_CLUSTER_PIPE": "0"}),
                                      ("registers", {"DSC_NO_TMA": "1", "DSC_NO_CLUSTER": "1", "DSC_FUSED_E16": "0"}),
                                      ("tma-e16", {"DSC_TMA_E16": "1", "DSC_NO_CLUSTER": "1"}),
                                      ("tma-e32", {"DSC_TMA_E16": "0", "DSC_NO_CLUSTER": "1"}),
                                      ("tma-direct", {"DSC_TMA_DIRECT": "1", "DSC_TMA_E16": "0", "DSC_NO_CLUSTER": "1"}),
                                      ("registers-e16", {"DSC_NO_TMA": "1", "DSC_NO_CLUSTER": "1", "DSC_FUSED_E16": "1"}),
                                      ("one-shot", {"DSC_NO_PERSIST": "1"}),
                                      ("real-sweep", {"DSC_NO_REAL_FUSE": "1"}),
                                      ("real-fused-f32", {"DSC_REAL_FUSE_F32": "1"})])
def test_two_pass_paths(name, env):
    e = dict(os.environ)
    for k in ("DSC_NO_TMA", "DSC_NO_CLUSTER", "DSC_CLUSTER_LGS", "DSC_TMA_E16", "DSC_CLUSTER_PIPE", "DSC_NO_REAL_FUSE", "DSC_REAL_FUSE_F32",
              "DSC_TMA_DIRECT", "DSC_FUSED_E16", "DSC_TMA_DEBUG_SKIP", "DSC_TMA_LAG", "DSC_NO_PERSIST", "DSC_COLUMNS_E16"):
        e.pop(k, None)
    e.update(env)
    r = subprocess.run([sys.executable, WORKER], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
