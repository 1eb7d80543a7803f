"""Worker of tests/test_distributed.py: one rank of the sharded four-step on the gloo backend with the
pthread-emulated kernels (CPU), or on NCCL with the real library (B200).  Prints 'OK <err>' on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dsc_b200 import cuda_api  # noqa: E402
from dsc_b200.distributed import ShardedFFT  # noqa: E402


def main():
    backend, lib, lg = sys.argv[1], sys.argv[2], int(sys.argv[3])
    dist.init_process_group(backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        device = torch.device("cuda", torch.cuda.current_device())
    else:
        device = torch.device("cpu")
    api = cuda_api.CudaApi(lib)
    n = 1 << lg
    rng = np.random.default_rng(lg)                       # same vector on every rank
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    f = ShardedFFT(n, api=api, device=device)
    local = f.scatter_input(torch.from_numpy(x))
    out = f.forward(local)
    X = f.gather_output(out).cpu().numpy()
    # the natural-order entry (steps 1 and 2 as one launch of column passes where the shape is covered)
    out_nat = f.forward_natural(f.scatter_input_natural(torch.from_numpy(x)))
    nat_err = float(torch.linalg.norm(out_nat - out) / torch.linalg.norm(out))
    assert nat_err < 2e-6, nat_err
    mode = f.last_mode
    back = f.gather_output(f.forward(f.scatter_input(torch.from_numpy(X)), inverse=True)).cpu().numpy()
    # the inverse through the natural-order entry too (on GPUs: the fused exchange, conjugate twiddles)
    back_nat = f.gather_output(f.forward_natural(f.scatter_input_natural(torch.from_numpy(X)), inverse=True)).cpu().numpy()
    assert float(np.linalg.norm(back_nat - back) / np.linalg.norm(back)) < 2e-6
    if rank == 0:
        from oracle import port
        want = port.fft(x)                                 # DSC's CPU FFT (C restatement, pinned bit-exact)
        err = float(np.linalg.norm(X - want) / np.linalg.norm(want))
        rt = float(np.linalg.norm(back - x) / np.linalg.norm(x))
        print(f"OK world={world} lg={lg} err={err:.3e} roundtrip={rt:.3e} natural-order path: {mode} (p2p unavailable: {f.p2p_error})", flush=True)
        assert err < 1e-5 and rt < 1e-5
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
