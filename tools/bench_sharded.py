"""BASELINE configs[4]: ONE complex64 FFT of 2^LG points sharded over the ranks of this job
(four-step, all-to-all over NVLink).  Launch: python -m torch.distributed.run --nproc-per-node P tools/bench_sharded.py [LG]
Prints one JSON line on rank 0: time (max over ranks, CUDA events), GFLOP/s (5 N log2 N), per-GPU HBM GB/s
of the algorithmic bytes, all-to-all GB/s per GPU, and a sampled-bin check against a float64 DFT."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsc_b200.distributed import ShardedFFT  # noqa: E402


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    n = 1 << lg
    f = ShardedFFT(n, device=dev)
    g = torch.Generator(device=dev).manual_seed(6 + rank)
    local = torch.view_as_complex(torch.randn(f.rows, f.N1, 2, generator=g, device=dev, dtype=torch.float32))

    local_cols = local.t().contiguous()                 # the same shard in natural order: [N1][N2/P]
    for _ in range(2):
        out = f.forward_natural(local_cols)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = f.forward_natural(local_cols)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    # sampled check: a few output bins against a float64 DFT accumulated over all ranks.
    # local[j][n1] = x[n1*N2 + n2], n2 = rank*rows + j ; out[i][k2] = X[k1 + N1*k2], k1 = rank*cols + i
    ks = [(0, 0), (f.cols // 2, 3), (f.cols - 1, f.N2 - 1), (1, f.N2 // 2)] if rank == 0 else []
    probes = torch.tensor([[rank * f.cols + i + f.N1 * k2 for i, k2 in ks]], device=dev, dtype=torch.int64).reshape(-1)
    n_probe = torch.tensor([probes.numel()], device=dev)
    dist.broadcast(n_probe, 0)
    if rank != 0:
        probes = torch.empty(int(n_probe.item()), device=dev, dtype=torch.int64)
    dist.broadcast(probes, 0)
    acc = torch.zeros(probes.numel(), 2, device=dev, dtype=torch.float64)
    n1_idx = torch.arange(f.N1, device=dev, dtype=torch.int64) * f.N2
    for j0 in range(0, f.rows, 64):
        j1 = min(j0 + 64, f.rows)
        nn = n1_idx[None, :] + (rank * f.rows + torch.arange(j0, j1, device=dev, dtype=torch.int64))[:, None]
        blk = local[j0:j1].to(torch.complex128)
        for pi, k in enumerate(probes.tolist()):
            ph = ((nn * k) % n).to(torch.float64) * (-2.0 * torch.pi / n)
            w = torch.complex(torch.cos(ph), torch.sin(ph))
            s = (blk * w).sum()
            acc[pi, 0] += s.real
            acc[pi, 1] += s.imag
    dist.all_reduce(acc)
    err = None
    if rank == 0:
        got = torch.stack([out[i, k2] for i, k2 in ks]).to(torch.complex128)
        want = torch.complex(acc[:, 0], acc[:, 1])
        err = float(((got - want).abs() / want.abs()).max().item())

    if rank == 0:
        flops = 5.0 * n * lg
        line = {
            "workload": f"single complex64 FFT, 2^{lg} points, four-step N1={f.N1} x N2={f.N2}, {world} GPU(s)",
            "n_gpus": world, "ms": ms, "gflops": flops / ms / 1e6,
            "hbm_gbs_per_gpu_algorithmic": 16.0 * n / world / ms / 1e6,
            "alltoall_gbs_per_gpu_each_way": 8.0 * n / world * (world - 1) / world / ms / 1e6 if world > 1 else 0.0,
            "sampled_bin_max_rel_err_vs_f64_dft": err, "reps": reps,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
