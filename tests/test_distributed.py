"""Multi-rank path of the sharded four-step (BASELINE configs[4] shape, small sizes).

CPU: world_size 2 and 4 on gloo with the emulated kernels (host logic, layouts, the one all-to-all).
GPU: the same worker on NCCL with the real library when at least 2 devices are visible."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _run(world, backend, lib, lg, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, backend, lib, str(lg)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK world=%d" % world in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("world,lg", [(2, 10), (2, 13), (4, 12)])
def test_sharded_four_step_gloo(world, lg):
    subprocess.run(["make", "-s", "-j8", "-C", EMUL_DIR], check=True)
    _run(world, "gloo", os.path.join(EMUL_DIR, "libdsc_emul.so"), lg, 29600 + world * 10 + lg)


@pytest.mark.gpu
@pytest.mark.parametrize("lg", [20, 24, 26])
def test_sharded_four_step_nccl(lg):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from dsc_b200 import cuda_api
    _run(world, "nccl", cuda_api.LIBDSC, lg, 29700 + lg)


@pytest.mark.gpu
def test_sharded_four_step_single_gpu():
    """P = 1 degenerates to the local four-step with an explicit transpose: same kernels, no exchange."""
    from dsc_b200 import cuda_api
    _run(1, "nccl", cuda_api.LIBDSC, 22, 29790)
