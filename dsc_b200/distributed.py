"""One transform sharded over P GPUs: the four-step decomposition whose transpose is an all-to-all.

N = N1 * N2 points, P ranks (one process per GPU, ``torch.distributed``; NCCL over NVLink on a B200
box).  Index split  n = n1*N2 + n2,  k = k1 + N1*k2:

    X[k1 + N1 k2] = sum_{n2} W_N2^{n2 k2} [ W_N^{n2 k1} sum_{n1} x[n1 N2 + n2] W_N1^{n1 k1} ]

  layout in  : rank p owns the column block n2 in [p N2/P, (p+1) N2/P), stored n2-major:
               local[n2_local][n1] = x[n1*N2 + n2]          (so the length-N1 lines are contiguous)
  step 1     : N2/P local transforms of length N1          (dsc_cuda_fft, sm_100a kernels)
  step 2     : times W_N^{n2 k1} and transpose to [k1][n2_local] (dsc_cuda_transpose_twiddle): the slab
               for peer q, k1 in block q, is then contiguous
  step 3     : all-to-all of the slabs                      (the ONE exchange; NCCL.  The slab a rank would send to
               itself stays in the send buffer and is read from there by step 4)
  step 4     : N1/P local transforms of length N2, read straight from the receive buffer [peer][k1_local][n2_local]
               (dsc_cuda_fft_segmented; short N2: un-interleave to [k1_local][n2] first)
  layout out : rank p owns k1 in block p: local[k1_local][k2] = X[k1 + N1*k2]   (block-transposed order)

The reference has no counterpart (single process, dsc/src/dsc.cpp:2082-2088 only marks where parallelism
was intended); parity is checked against its CPU FFT on sizes the oracle can do (tests/test_distributed.py).
torch supplies device memory, streams and the process group -- plumbing; every transform, twiddle and
transpose runs in this repo's kernels.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import cuda_api


def _ilog2(v: int) -> int:
    assert v > 0 and v & (v - 1) == 0, "power of two expected"
    return v.bit_length() - 1


class _PeerBuffers:
    """One receive buffer per rank, allocated in symmetric memory and mapped into every process
    (torch.distributed._symmetric_memory: plumbing only -- the stores into it are this repo's kernels)."""

    def __init__(self, nbytes: int, device, group, rank: int, world: int):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.ptrs) == world
        self.local_ptr = self.buf.data_ptr()
        assert self.ptrs[rank] == self.local_ptr or True

    def barrier(self, channel: int) -> None:
        # device-side barrier over the signal pads, ordered on the current stream
        self.hdl.barrier(channel=channel)


class ShardedFFT:
    def __init__(self, n_total: int, api: cuda_api.CudaApi | None = None, device=None, group=None,
                 dtype=torch.complex64, p2p=None):
        """p2p: None = use peer-mapped receive buffers when available, True = require them, False = NCCL all-to-all."""
        self.p2p_error = None
        self.group = group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.api = api or cuda_api.CudaApi()
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype
        self.prec = cuda_api.F32 if dtype == torch.complex64 else cuda_api.F64
        self.code = cuda_api.C32 if dtype == torch.complex64 else cuda_api.C64
        lg = _ilog2(n_total)
        self.N = n_total
        self.N1 = 1 << ((lg + 1) // 2)
        self.N2 = 1 << (lg // 2)
        assert self.N1 % self.P == 0 and self.N2 % self.P == 0, "P must divide both factors"
        self.rows = self.N2 // self.P          # local n2 lines in step 1
        self.cols = self.N1 // self.P          # local k1 lines in step 4
        self.plan1, self._mem1 = self._plan(self.N1)
        self.plan2, self._mem2 = (self.plan1, None) if self.N2 == self.N1 else self._plan(self.N2)
        # W_N^p through two sqrt(N)-sized tables: p = (p >> shift) << shift | (p & mask)
        self.shift = (lg + 1) // 2
        self.tw_lo = torch.empty(1 << self.shift, dtype=dtype, device=self.device)
        self.tw_hi = torch.empty(1 << (lg - self.shift), dtype=dtype, device=self.device)
        self.api.fill_twiddles(self.tw_lo.data_ptr(), self.tw_lo.numel(), 1, self.N, self.prec, self._stream())
        self.api.fill_twiddles(self.tw_hi.data_ptr(), self.tw_hi.numel(), 1 << self.shift, self.N, self.prec, self._stream())
        wb = max(self.api.work_bytes(self.plan1, self.rows), self.api.work_bytes(self.plan2, self.cols),
                 self.api.work_bytes_axis(self.plan1, 1, self.rows), 256)
        self.work = torch.empty(wb, dtype=torch.uint8, device=self.device)
        # Fused compute + collective: receive buffers in symmetric memory (every rank's buffer mapped into every process
        # over NVLink), so the first local step stores its row blocks straight into the owners' buffers and no separate
        # all-to-all runs.  None when symmetric memory is not available (CPU / gloo, one rank, or an old driver): the
        # NCCL all-to-all path below is used instead.
        self.p2p = None
        if p2p is not False and self.P > 1 and self.device.type == "cuda" and self.P <= 8:
            try:
                self.p2p = _PeerBuffers(self.N1 * self.rows * torch.empty(0, dtype=dtype).element_size(), self.device, group,
                                        self.rank, self.P)
            except Exception as e:      # noqa: BLE001
                if p2p is True:
                    raise
                self.p2p_error = f"{type(e).__name__}: {e}"
        # measurement hooks (bench.py): when `events` is a list, every exchange appends a (start, stop) pair of CUDA
        # events recorded on the current stream around the all-to-all; `last_mode` names the exchange path taken
        self.events = None
        self.last_mode = None

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0

    def _plan(self, n):
        nb = self.api.plan_bytes(n, cuda_api.FFT_COMPLEX, self.prec)
        assert nb > 0, f"no plan for length {n}"
        mem = torch.empty(nb, dtype=torch.uint8, device=self.device)
        return self.api.plan_build(n, cuda_api.FFT_COMPLEX, self.prec, mem.data_ptr(), nb, self._stream()), mem

    # ---- layout helpers (host side, for tests and small inputs) -----------------------------------
    def scatter_input(self, x_full: torch.Tensor) -> torch.Tensor:
        """local[n2_local][n1] of this rank from the natural-order vector (every rank passes the same x)."""
        m = x_full.reshape(self.N1, self.N2)                     # [n1][n2]
        blk = m[:, self.rank * self.rows:(self.rank + 1) * self.rows]
        return blk.t().contiguous().to(self.device)

    def scatter_input_natural(self, x_full: torch.Tensor) -> torch.Tensor:
        """The rank's column block in NATURAL order, local[n1][n2_local] = x[n1*N2 + n2] (input of forward_natural)."""
        m = x_full.reshape(self.N1, self.N2)
        return m[:, self.rank * self.rows:(self.rank + 1) * self.rows].contiguous().to(self.device)

    def gather_output(self, local_out: torch.Tensor) -> torch.Tensor:
        """Natural-order X on every rank from the block-transposed shards (all_gather + transpose)."""
        if self.P > 1:
            parts = [torch.empty_like(local_out) for _ in range(self.P)]
            dist.all_gather(parts, local_out.contiguous(), group=self.group)
            full = torch.cat(parts, dim=0)                       # [k1][k2]
        else:
            full = local_out
        return full.t().contiguous().reshape(-1)                 # index k1 + N1*k2

    # ---- the transform -------------------------------------------------------------------------------
    def forward_natural(self, local_cols: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        """local_cols: [N1, N2/P], the rank's column block of x in natural order.  Steps 1 and 2 are ONE launch
        (dsc_cuda_fft_columns_twiddled: both passes of the length-N1 transform as column passes, the outer
        twiddle applied on the way out, k1-major output); shapes it does not cover are transposed and go through
        forward()."""
        assert local_cols.shape == (self.N1, self.rows) and local_cols.dtype == self.dtype and local_cols.is_contiguous()
        api, s = self.api, self._stream()
        if self.p2p is not None:
            out = self._forward_p2p(local_cols, not inverse)
            if out is not None:
                return out
        send = torch.empty(self.N1, self.rows, dtype=self.dtype, device=self.device)
        if api.fft_columns_twiddled(self.plan1, local_cols.data_ptr(), send.data_ptr(), self.rows, not inverse,
                                    self.rank * self.rows, self.tw_lo.data_ptr(), self.tw_hi.data_ptr(), self.shift,
                                    self.N, self.work.data_ptr(), self.work.numel(), s):
            return self._exchange_and_finish(send, not inverse)
        return self.forward(local_cols.t().contiguous(), inverse)

    def _forward_p2p(self, local_cols: torch.Tensor, fwd: bool):
        """Steps 1-3 as ONE launch: the column transforms' epilogue stores row block q (k1 in block q) into rank q's
        receive buffer at [this rank][k1_local][n2_local] through NVLink peer memory, so the exchange overlaps the
        transform tile by tile.  Two cross-rank barriers bracket the launch: every rank has finished reading its buffer
        (the previous transform's last step) before anyone writes into it, and every rank's writes have landed before
        the buffers are read."""
        api, s = self.api, self._stream()
        es = local_cols.element_size()
        slab_bytes = self.cols * self.rows * es
        peer_ptrs = [self.p2p.ptrs[q] + self.rank * slab_bytes for q in range(self.P)]
        ev = None
        if self.events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self.p2p.barrier(0)
        if ev is not None:
            ev[0].record()
        ok = api.fft_columns_twiddled_p2p(self.plan1, local_cols.data_ptr(), self.rows, fwd, self.rank * self.rows,
                                          self.tw_lo.data_ptr(), self.tw_hi.data_ptr(), self.shift, self.N, peer_ptrs,
                                          self.work.data_ptr(), self.work.numel(), s)
        if not ok:
            return None
        self.p2p.barrier(1)
        if ev is not None:
            ev[1].record()
            self.events.append(ev)
        self.last_mode = "fused: column-pass epilogue stores into peer-mapped receive buffers (NVLink), no separate all-to-all"
        out = torch.empty(self.cols, self.N2, dtype=self.dtype, device=self.device)
        if api.fft_segmented(self.plan2, self.p2p.local_ptr, out.data_ptr(), self.cols, self.rows, self.cols * self.rows, fwd,
                             self.work.data_ptr(), self.work.numel(), s, -1, 0):
            return out
        # short second transforms (one shared-memory pass): un-interleave [peer][k1_local][n2_local] first
        recv = self.p2p.buf.view(self.dtype).view(self.P, self.cols, self.rows)
        b = torch.empty(self.cols, self.N2, dtype=self.dtype, device=self.device)
        b.view(self.cols, self.P, self.rows).copy_(recv.permute(1, 0, 2))
        api.fft(self.plan2, b.data_ptr(), self.code, out.data_ptr(), self.cols, self.N2, 1, fwd,
                self.work.data_ptr(), self.work.numel(), s)
        return out

    def forward(self, local_in: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        """local_in: [N2/P, N1] (n2-major shard).  Returns [N1/P, N2]: X[k1 + N1 k2] for this rank's k1 block."""
        assert local_in.shape == (self.rows, self.N1) and local_in.dtype == self.dtype and local_in.is_contiguous()
        api, s = self.api, self._stream()
        fwd = not inverse
        a = torch.empty_like(local_in)
        api.fft(self.plan1, local_in.data_ptr(), self.code, a.data_ptr(), self.rows, self.N1, 1, fwd,
                self.work.data_ptr(), self.work.numel(), s)
        # twiddle + transpose: send[k1][n2_local]; rows [q*cols, (q+1)*cols) go to peer q
        send = torch.empty(self.N1, self.rows, dtype=self.dtype, device=self.device)
        api.transpose_twiddle(a.data_ptr(), send.data_ptr(), self.rows, self.N1, self.rank * self.rows,
                              self.tw_lo.data_ptr(), self.tw_hi.data_ptr(), self.shift, fwd, self.code, s)
        return self._exchange_and_finish(send, fwd)

    def _exchange_and_finish(self, send: torch.Tensor, fwd: bool) -> torch.Tensor:
        """send[k1][n2_local] (twiddled) -> all-to-all -> N1/P local transforms of length N2."""
        api, s = self.api, self._stream()
        inverse = not fwd
        if self.P > 1:
            recv = torch.empty_like(send)                        # [q][k1_local][n2_local]
            slab = self.cols * self.rows
            in_place = self.plan2.lg_n2 != 0 and self.device.type == "cuda"
            ev = None
            if self.events is not None and self.device.type == "cuda":
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            self.last_mode = "blocking all_to_all (P-1 slabs, own slab stays)" if in_place else "blocking all_to_all_single"
            if in_place:
                # the slab a rank would send to itself stays where it is: the exchange moves the P-1 others only,
                # and the second transform reads that segment from the send buffer
                flat_s, flat_r = send.view(-1), recv.view(-1)
                empty = flat_s[:0]
                ins = [empty if q == self.rank else flat_s[q * slab:(q + 1) * slab] for q in range(self.P)]
                outs = [flat_r[:0] if q == self.rank else flat_r[q * slab:(q + 1) * slab] for q in range(self.P)]
                dist.all_to_all(outs, ins, group=self.group)
            else:
                dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
            if ev is not None:
                ev[1].record()
                self.events.append(ev)
            out = torch.empty(self.cols, self.N2, dtype=self.dtype, device=self.device)
            # line k1_local is P segments of N2/P points, one per source rank: transformed where it lies when the
            # plan is a two-pass one (the first pass reads segmented rows), else un-interleaved first
            if api.fft_segmented(self.plan2, recv.data_ptr(), out.data_ptr(), self.cols, self.rows, slab, fwd,
                                 self.work.data_ptr(), self.work.numel(), s,
                                 self.rank if in_place else -1, send.data_ptr() if in_place else 0):
                return out
            if in_place:
                recv.view(self.P, slab)[self.rank].copy_(send.view(self.P, slab)[self.rank])
            b = torch.empty(self.cols, self.N2, dtype=self.dtype, device=self.device)
            b.view(self.cols, self.P, self.rows).copy_(recv.view(self.P, self.cols, self.rows).permute(1, 0, 2))
        else:
            b = send                                             # [k1][n2] already
        out = torch.empty_like(b)
        api.fft(self.plan2, b.data_ptr(), self.code, out.data_ptr(), self.cols, self.N2, 1, fwd,
                self.work.data_ptr(), self.work.numel(), s)
        if inverse:
            # each local pass scaled by its own 1/length; together that is 1/N
            pass
        return out
