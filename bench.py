#!/usr/bin/env python
"""bench.py -- headline benchmark of the FFT hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- batched complex64 FFT along the last
axis, 65536 x 4096 points, one step = forward + inverse over the whole batch.
Metric: GFLOP/s counted as 5*N*log2(N) per transform (the reference's own convention,
benchmarks/python/bench_fft.py:44), whole job over all N GPUs.  N > 1 (launched by
torch.distributed.run, one rank per GPU): every rank transforms its own batch -- lines are
independent, so there is no collective on the data path (weak scaling); the only NCCL traffic is
the barrier and the max-over-ranks of the timings.

  value         device-resident: inputs already in HBM, dsc_cuda_fft (device-level C ABI) on the
                current stream, timed with CUDA events, max over ranks.
  roofline      per-launch duration of fft_lines<> measured with CUDA events inside the timed loop
                against the algorithmic bytes (16 B per point: one read + one write).
  e2e           the same step through the drop-in tensor C ABI (dsc_fft / dsc_ifft of libdsc.so)
                with HOST buffers: uploads and downloads inside the timed region.
  cpu_baseline  the unmodified reference (oracle/_ref/libdsc_ref.so, else the C port) on the
                box's host cores, bounded sample, N = 1 / rank 0 only.
--impl reference prints the reference-arm line (CPU, all host cores) for the same metric/config.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LG_N = 12
N_POINTS = 1 << LG_N
BATCH = 65536
FLOP_PER_TRANSFORM = 5.0 * N_POINTS * LG_N
BYTES_PER_TRANSFORM = 16.0 * N_POINTS            # complex64: 8 B read + 8 B written per point
METRIC = "batched FFT GFLOP/s (5N*log2N)"
WORKLOAD = "complex64 fft+ifft, last axis, 65536 x 4096 (BASELINE configs[1])"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# reference / CPU baseline (oracle side: the only place bench.py touches oracle/)

def _cpu_worker(args):
    """One process = one single-threaded reference context (the reference has no threading)."""
    lines, reps, seed, use_ref = args
    import numpy as np
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((lines, N_POINTS)) + 1j * rng.standard_normal((lines, N_POINTS))).astype(np.complex64)
    if use_ref:
        from oracle.ref_harness import RefLib
        ref = RefLib(main_mem=max(4 * x.nbytes, 1 << 26) + (1 << 24), scratch_mem=1 << 24)
        tx = ref.put(x)

        def step():
            ty = ref.lib.dsc_fft(ref.ctx, tx, None, -1, -1)
            tz = ref.lib.dsc_ifft(ref.ctx, ty, None, -1, -1)
            ref.free(tz)
            ref.free(ty)
    else:
        from oracle import port

        def step():
            port.ifft(port.fft(x))
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return times


def cpu_reference_run(steps, warmup, lines_per_core=1024):
    """Whole-host throughput of the reference CPU FFT: one process per core, each doing fwd+inv over
    `lines_per_core` lines per step.  Returns (gflops, seconds_per_step, cores, kind, sample)."""
    from oracle import ref_harness
    use_ref = ref_harness.available()
    if not use_ref:
        from oracle import port
        port.lib()      # build / load once in the parent so workers do not race on make
    cores = len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(lines_per_core, warmup + steps, 1000 + i, use_ref) for i in range(cores)])
    # a step ends when the slowest core is done
    per_step = [max(r[i] for r in res) for i in range(warmup, warmup + steps)]
    sec = statistics.mean(per_step)
    gflops = cores * lines_per_core * 2 * FLOP_PER_TRANSFORM / sec / 1e9
    kind = "reference" if use_ref else "port"
    sample = f"{cores} processes x {lines_per_core} lines x 4096 points, fwd+inv, mean of {steps} steps after {warmup} warm-ups"
    return gflops, sec, cores, kind, sample


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    gflops, sec, cores, kind, sample = cpu_reference_run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex64 (f32 arithmetic)", "data": "synthetic N(0,1), seeded",
        "config": {"workload": WORKLOAD, "note": "reference CPU FFT, single-threaded library, one process per host core; "
                                                 "each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:     # noqa: BLE001
            log(f"bench: NVML unavailable ({e}); clocks not sampled")
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:       # noqa: BLE001  older binding name
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:           # noqa: BLE001
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# our arm

def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("fft_lines_c64_4096_bytes_per_launch")
        except Exception:       # noqa: BLE001
            return None
    return None


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from dsc_b200 import cuda_api

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process: the workers are forked
        gfl, sec, cores, kind, sample = cpu_reference_run(steps=5, warmup=2, lines_per_core=2048)
        cpu = {"value": gfl, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample}
        log(f"bench: cpu baseline {gfl:.1f} GFLOP/s on {cores} cores ({kind})")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    api = cuda_api.CudaApi()      # raises if libdsc.so is missing: no fallback
    rows = BATCH
    # ---- device-resident leg ----------------------------------------------------------------
    g = torch.Generator(device=dev).manual_seed(2 + rank)
    x = torch.view_as_complex(torch.randn(rows, N_POINTS, 2, generator=g, device=dev, dtype=torch.float32))
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    nb = api.plan_bytes(N_POINTS, cuda_api.FFT_COMPLEX, cuda_api.F32)
    plan_mem = torch.empty(nb, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    plan = api.plan_build(N_POINTS, cuda_api.FFT_COMPLEX, cuda_api.F32, plan_mem.data_ptr(), nb, stream)

    def step(events=None):
        if events is not None:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
        api.fft(plan, x.data_ptr(), cuda_api.C32, y.data_ptr(), rows, N_POINTS, 1, True, 0, 0, stream)
        if events is not None:
            e[1].record()
        api.fft(plan, y.data_ptr(), cuda_api.C32, z.data_ptr(), rows, N_POINTS, 1, False, 0, 0, stream)
        if events is not None:
            e[2].record()
            events.append(e)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launch_events = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        t0.record()
        for _ in range(args.steps):
            step(launch_events)
        t1.record()
        barrier()
    ms_total = max_over_ranks(t0.elapsed_time(t1))
    ms_per_step = ms_total / args.steps
    value = world * rows * 2 * FLOP_PER_TRANSFORM / (ms_per_step * 1e-3) / 1e9

    fwd_ms = [e[0].elapsed_time(e[1]) for e in launch_events]
    inv_ms = [e[1].elapsed_time(e[2]) for e in launch_events]
    launch_ms = statistics.mean(fwd_ms + inv_ms)
    algo_bytes = rows * BYTES_PER_TRANSFORM
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()

    # parity spot check inside the bench (device result vs the oracle on a few rows)
    err = float(((z[:64] - x[:64]).norm() / x[:64].norm()).item())

    # ---- end-to-end leg: drop-in tensor C ABI with host buffers -------------------------------
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                              "ms_per_step": ms_per_step, "note": "profiling run (--no-e2e): not a bench line",
                              "roofline": {"achieved": achieved, "peak": peak, "frac": achieved / peak, "launch_ms": launch_ms}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    del y, z
    x_host = torch.view_as_real(x[: rows]).cpu().numpy().view(np.complex64).reshape(rows, N_POINTS)
    del x
    torch.cuda.empty_cache()
    import dsc_b200 as dsc
    tensor_bytes = rows * N_POINTS * 8
    dsc.init(4 * tensor_bytes + (1 << 28), 1 << 26)
    tx = dsc.from_numpy(x_host)
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_leg(residency):
        """fft -> ifft through the tensor C ABI, x in (pinned) host memory, result read on the host.
        residency 0: the library default, every call uploads and downloads (y crosses PCIe twice);
        residency 2: the intermediate y stays on the device; x is marked host-dirty before every
        step so its upload is inside the timed region, and z is downloaded by numpy()/sync_host."""
        dsc.set_residency(residency)
        lib, ctx = dsc._load(), dsc._get_ctx()

        def one_step(prev):
            """Upload x, fft, ifft, and the device -> host read of the result -- every step, inside the timed
            region.  Lazy mode double-buffers the results: the download of step i is started asynchronously
            (dsc_cuda_download_async) and awaited after step i+1 has been issued, so it overlaps that step's
            upload in the other PCIe direction.  Strict mode downloads inside dsc_ifft itself."""
            if residency:
                lib.dsc_cuda_touch_host(ctx, tx.c)        # fresh host data: forces the upload
            ty = dsc.fft(tx)
            tz = dsc.ifft(ty)
            del ty
            if residency:
                dsc.download_async(tz)
            if prev is not None:
                dsc.sync_host(prev)                        # the previous step's result is now on the host
            return tz

        tz = None
        for _ in range(2):
            tz = one_step(tz)                              # at most two results (z) alive, plus x and y
        dsc.sync_host(tz)
        barrier()
        tz = None
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            tz = one_step(tz)
        dsc.sync_host(tz)                                  # the last result, still inside the timed region
        torch.cuda.synchronize()
        sec = max_over_ranks((time.perf_counter() - w0) / e2e_steps)
        zr = tz.numpy()[:64]
        err_ = float(np.linalg.norm(zr - x_host[:64]) / np.linalg.norm(x_host[:64]))
        del tz
        dsc.set_residency(0)
        return sec, err_

    strict_sec, strict_err = e2e_leg(0)
    e2e_sec, e2e_err = e2e_leg(2)
    del tx
    dsc.shutdown()
    e2e_value = world * rows * 2 * FLOP_PER_TRANSFORM / e2e_sec / 1e9
    e2e_strict_value = world * rows * 2 * FLOP_PER_TRANSFORM / strict_sec / 1e9

    if world > 1:
        dist.barrier()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "complex64 (f32 arithmetic)", "data": "synthetic N(0,1), seeded, generated on device",
            "config": {"workload": WORKLOAD, "batch_per_gpu": rows, "n": N_POINTS, "step": "fft then ifft, out of place",
                       "l2": "2 GiB per tensor >> 126 MB L2, no flush needed", "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "roundtrip_rel_l2": err, "e2e_roundtrip_rel_l2": e2e_err},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": tensor_bytes, "d2h_bytes_per_step": tensor_bytes,
                    "ms_per_step": e2e_sec * 1e3, "steps": e2e_steps,
                    "api": "dsc_fft + dsc_ifft (libdsc.so tensor C ABI), x in the pinned host arena and uploaded every step, "
                           "result z downloaded every step (started with dsc_cuda_download_async, awaited after the next step is issued: "
                           "full-duplex PCIe), intermediate y kept on the device (dsc_cuda_set_residency(2))",
                    "strict": {"value": e2e_strict_value, "ms_per_step": strict_sec * 1e3, "h2d_bytes_per_step": 2 * tensor_bytes,
                               "d2h_bytes_per_step": 2 * tensor_bytes, "roundtrip_rel_l2": strict_err,
                               "api": "same calls with the library default (residency 0): every call uploads its input and "
                                      "downloads its output, so y crosses PCIe twice"}},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "kernel": "fft_lines<float,12,4,1,{fwd,inv},MODE_FAST>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_per_launch(),
                         "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": launch_ms,
                         "launch_ms_fwd": statistics.mean(fwd_ms), "launch_ms_inv": statistics.mean(inv_ms), "peak_source": peak_src},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: relaunch under torchrun, one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29517",
                                   os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
