"""Parity cases for the tensor-level C ABI (dsc_fft / dsc_ifft / dsc_rfft / dsc_irfft, plan cache,
tracer, residency), shared by the emulated (CPU) and the real (B200) runs.  Each function takes
the bound `dsc_b200` module; the oracle is the C restatement in oracle/ (pinned bit-exact to the
reference) and the golden vectors recorded from the reference itself."""
import json
import os
import tempfile

import numpy as np

from oracle import port
from tests.util import TOL, load_golden, randn, rel_l2


def check_golden(dsc):
    n = 0
    for i, m, a in load_golden():
        op = m["op"]
        if op in ("fft", "ifft", "rfft", "irfft"):
            got = getattr(dsc, op)(a["x"], n=m["n"], axis=m["axis"]).numpy()
        elif op == "mul":
            got = dsc.mul(a["x"], a["b"]).numpy()
        elif op == "filter":
            S = dsc.rfft(a["x"], n=m["n"])
            B = dsc.rfft(a["b"], n=m["n"])
            got = dsc.irfft(S * B).numpy()
        else:
            continue
        want = a["y"]
        assert got.shape == want.shape and got.dtype == want.dtype, (i, m, got.shape, want.shape)
        assert rel_l2(got, want) < TOL[want.dtype], (i, m, rel_l2(got, want))
        n += 1
    assert n > 150


def check_shapes_appendix_a(dsc):
    rng = np.random.default_rng(11)
    x = randn(rng, (10,), "float32")
    y = dsc.fft(x)
    assert y.shape == (16,) and y.dtype == np.complex64
    assert rel_l2(y.numpy(), np.fft.fft(x.astype(np.float64), n=16)) < 1e-5
    assert dsc.fft(x, n=5).shape == (8,)
    assert rel_l2(dsc.fft(x, n=5).numpy(), np.fft.fft(x[:8].astype(np.float64))) < 1e-5
    assert dsc.rfft(x).shape == (9,) and dsc.rfft(x, n=4).shape == (3,)
    X = dsc.rfft(x)
    assert dsc.irfft(X).shape == (16,) and dsc.irfft(X, n=9).shape == (16,)
    assert dsc.irfft(X, n=16).shape == (32,) and dsc.irfft(X, n=5).shape == (8,)
    assert rel_l2(dsc.irfft(X, n=16).numpy(), np.fft.irfft(X.numpy().astype(np.complex128), n=32)) < 1e-5
    assert dsc.ifft(x).dtype == np.complex64                      # real input is cast, then inverted
    x3 = randn(rng, (4, 8, 16), "float64")
    for axis in range(-3, 3):
        assert rel_l2(dsc.fft(x3, axis=axis).numpy(), np.fft.fft(x3, axis=axis)) < 1e-12


def check_out_param(dsc):
    rng = np.random.default_rng(12)
    x = randn(rng, (6, 64), "complex64")
    out = dsc.from_numpy(np.zeros((6, 64), np.complex64))
    res = dsc.fft(x, out=out)
    assert rel_l2(out.numpy(), port.fft(x)) < 1e-6
    assert np.array_equal(res.numpy(), out.numpy())
    xr = randn(rng, (3, 32), "float64")
    out = dsc.from_numpy(np.zeros((3, 17), np.complex128))
    dsc.rfft(xr, out=out)
    assert rel_l2(out.numpy(), port.rfft(xr)) < 1e-14


def check_vs_oracle_sweep(dsc, max_lg=12):
    rng = np.random.default_rng(13)
    for dtype in ("float32", "float64", "complex64", "complex128"):
        for shape, n, axis in [((5, 128), -1, -1), ((16, 7), -1, 0), ((3, 4, 32, 2), 64, 2), ((2, 1 << max_lg), -1, 1),
                               ((9, 33), 16, -1)]:
            x = randn(rng, shape, dtype)
            for op in ("fft", "ifft"):
                got = getattr(dsc, op)(x, n=n, axis=axis).numpy()
                want = getattr(port, op)(x, n, axis)
                assert got.shape == want.shape and got.dtype == want.dtype
                assert rel_l2(got, want) < TOL[want.dtype] / 5
            if np.dtype(dtype).kind == "f":
                X = dsc.rfft(x, n=n, axis=axis)
                want = port.rfft(x, n, axis)
                assert X.shape == want.shape
                assert rel_l2(X.numpy(), want) < TOL[want.dtype] / 5
                back = dsc.irfft(X, axis=axis).numpy()
                wb = port.irfft(want, -1, axis)
                assert back.shape == wb.shape and rel_l2(back, wb) < TOL[wb.dtype] / 5


def check_filter_pipeline(dsc):
    # BASELINE configs[0]: README filterFFT (README.md:118-134)
    s = np.random.default_rng(0).standard_normal(8192).astype(np.float32)
    b = np.random.default_rng(1).standard_normal(128).astype(np.float32)
    want = port.filter_fft(s, b, 16384)
    B = dsc.rfft(b, n=16384)
    unfused = dsc.irfft(dsc.rfft(s, n=16384) * B)
    assert rel_l2(unfused.numpy(), want) < 1e-5
    fused = dsc.fft_filter(s, B, n=16384)
    assert fused.shape == (16384,)
    assert rel_l2(fused.numpy(), want) < 1e-5
    conv = np.convolve(s.astype(np.float64), b.astype(np.float64))
    assert rel_l2(fused[:8319].numpy(), conv) < 1e-5            # the README's [:output_length] crop
    # batched channels, broadcast spectrum
    S = randn(np.random.default_rng(2), (5, 1000), "float32")
    Bs = dsc.rfft(b, n=2048)
    got = dsc.fft_filter(S, Bs, n=2048).numpy()
    for r in range(5):
        assert rel_l2(got[r], port.filter_fft(S[r], b, 2048)) < 1e-5


def check_plan_cache(dsc, max_lg=12):
    # > DSC_MAX_FFT_PLANS distinct plans: results stay right, device usage plateaus (eviction frees)
    rng = np.random.default_rng(14)
    usage = []
    for rnd in range(2):
        for lg in range(1, max_lg + 1):
            for dtype in ("complex64", "complex128"):
                x = randn(rng, (2, 1 << lg), dtype)
                assert rel_l2(dsc.fft(x).numpy(), port.fft(x)) < TOL[x.dtype] / 5
        usage.append(dsc.device_used_mem())
    assert 2 * max_lg > 16
    assert usage[0] == usage[1], usage
    assert dsc.device_alloc_calls() == 1, "the device arena is the only device allocation"


def check_memory_accounting(dsc):
    base_h, base_d = dsc.used_mem(), dsc.device_used_mem()
    rng = np.random.default_rng(15)
    for _ in range(3):
        x = dsc.from_numpy(randn(rng, (8, 256), "complex64"))
        y = dsc.fft(x)
        z = dsc.ifft(y)
        del x, y, z
    assert dsc.used_mem() == base_h
    assert dsc.device_used_mem() == base_d


def _host_write(dsc, t, values):
    """What a raw-pointer user of the C ABI does: announce the write (dsc_cuda_touch_host waits for a download that
    is still in flight and invalidates the device mirror), then store through tensor.data."""
    import ctypes
    values = np.ascontiguousarray(values, dtype=t.dtype)
    dsc._load().dsc_cuda_touch_host(dsc._get_ctx(), t.c)
    ctypes.memmove(t.c.contents.data, values.ctypes.data, values.nbytes)


def check_residency_modes(dsc):
    rng = np.random.default_rng(16)
    x = randn(rng, (32, 512), "complex64")
    want = port.ifft(port.fft(x))
    try:
        for mode in (1, 2):
            dsc.set_residency(mode)
            y = dsc.fft(x)
            z = dsc.ifft(y)                   # y never re-uploaded; in mode 2 never downloaded either
            assert rel_l2(z.numpy(), want) < 1e-6
            assert rel_l2(y.numpy(), port.fft(x)) < 1e-6
            prod = y * y                      # host op on a device-resident tensor syncs it first
            assert rel_l2(prod.numpy(), port.fft(x) ** 2) < 1e-5
            del y, z, prod
            # the README's three-call filter with the spectra staying on the device (row-broadcast product)
            sig = randn(rng, (5, 1000), "float32")
            taps = randn(rng, (37,), "float32")
            S, B = dsc.rfft(sig, n=2048), dsc.rfft(taps, n=2048)
            yf = dsc.irfft(S * B)
            for r in range(5):
                assert rel_l2(yf.numpy()[r], port.filter_fft(sig[r], taps, 2048)) < 1e-5
            # README.md:133 y[:output_length]: in mode 2 only the kept columns are downloaded
            crop = yf[:, :1036].numpy() if mode == 2 else yf.numpy()[:, :1036]
            for r in range(5):
                assert rel_l2(crop[r], port.filter_fft(sig[r], taps, 2048)[:1036]) < 1e-5
            assert rel_l2(yf[:, 100:300].numpy(), yf.numpy()[:, 100:300]) == 0.0
            del S, B, yf
            # double-buffered results: the download of one result overlaps the next transform (no-op in mode 1)
            xs = [randn(rng, (16, 1024), "complex64") for _ in range(3)]
            pending = []
            for xi in xs:
                zi = dsc.ifft(dsc.fft(xi))
                dsc.download_async(zi)
                pending.append(zi)
            for xi, zi in zip(xs, pending):
                assert rel_l2(zi.numpy(), xi) < 1e-6
            del pending
            # a result whose asynchronous download is still in flight is (a) freed, (b) reused as out=, (c) rewritten
            # by the host: each must retire the pending copy first (no DMA into freed memory, no stale event)
            big = randn(rng, (64, 4096), "complex64")
            tb = dsc.from_numpy(big)
            z1 = dsc.ifft(dsc.fft(tb))
            dsc.download_async(z1)
            del z1                                                   # (a)
            z2 = dsc.ifft(dsc.fft(tb))
            dsc.download_async(z2)
            other = randn(rng, (64, 4096), "complex64")
            z2 = dsc.ifft(dsc.fft(other), out=z2)                    # (b)
            assert rel_l2(z2.numpy(), other) < 1e-6
            z3 = dsc.ifft(dsc.fft(tb))
            dsc.download_async(z3)
            _host_write(dsc, z3, np.ones((64, 4096), np.complex64))  # (c): retires the pending copy, then writes
            assert np.all(dsc.fft(z3).numpy()[:, 0] == 4096.0)
            # the input mirror is overwritten by the next upload while the previous launch may still read it
            for rep in range(4):
                _host_write(dsc, tb, other * (rep + 1))
                zi = dsc.ifft(dsc.fft(tb))
                assert rel_l2(zi.numpy(), other * (rep + 1)) < 1e-6
            del tb, z2, z3, zi
    finally:
        dsc.set_residency(0)


def check_device_pointwise(dsc):
    """|X|, arg X, Re, Im, conj and add/sub/mul/div against what the unmodified reference
    returned for the same inputs -- through the host loops (strict mode) and through the device kernels
    (operands made device-resident first; in mode 2 the results are downloaded lazily)."""
    gold = [(m, a) for _, m, a in load_golden() if m["op"].startswith(("unary:", "binary:"))]
    assert len(gold) >= 58
    fn = {"abs": dsc.abs, "angle": dsc.angle, "real": dsc.real, "imag": dsc.imag, "conj": dsc.conj,
          "add": dsc.add, "sub": dsc.sub, "mul": dsc.mul, "div": dsc.true_div}
    exact = ("real", "imag", "conj", "add", "sub")
    try:
        for mode in (0, 1, 2):
            dsc.set_residency(mode)
            for meta, arrs in gold:
                kind, name = meta["op"].split(":")
                x = dsc.from_numpy(arrs["x"])
                dsc.prefetch(x)
                if kind == "unary":
                    got = fn[name](x)
                else:
                    b = dsc.from_numpy(arrs["b"])
                    got = fn[name](x, b)
                want = arrs["y"]
                g = got.numpy()
                assert g.shape == want.shape and g.dtype == want.dtype, (meta, mode)
                if name in exact:
                    assert np.array_equal(g, want), (meta, mode)
                else:
                    assert rel_l2(g, want) < TOL[want.dtype] * 1e-1, (meta, mode, rel_l2(g, want))
            # post-processing chained behind a transform: nothing goes back to the host in between
            z = randn(np.random.default_rng(5), (4, 256), "complex64")
            Z = dsc.fft(z)
            mag, ph = dsc.abs(Z), dsc.angle(Z)
            ref = port.fft(z)
            assert rel_l2(mag.numpy(), np.abs(ref)) < 1e-5
            assert rel_l2(mag.numpy() * np.exp(1j * ph.numpy()), ref) < 1e-5
            del Z, mag, ph
    finally:
        dsc.set_residency(0)


def _py_slices(enc):
    return tuple(slice(*e) if isinstance(e, list) else int(e) for e in enc)


def check_device_ops(dsc):
    """SURVEY.md 8(f) ranks 1, 3 and 4 against what the unmodified reference returned (tests/golden/ops_golden.npz):
    dsc_cast, mixed-dtype add/sub/mul/div, dsc_transpose, dsc_fftfreq / dsc_rfftfreq, dsc_tensor_get_slice /
    set_slice -- through the host loops (strict mode) and through the device kernels (operands device-resident;
    in mode 2 results stay on the device until read)."""
    from tests.util import load_ops_golden
    gold = list(load_ops_golden())
    assert len(gold) >= 280
    fn = {"add": dsc.add, "sub": dsc.sub, "mul": dsc.mul, "div": dsc.true_div}
    try:
        for mode in (0, 1, 2):
            dsc.set_residency(mode)
            for meta, arrs in gold:
                op = meta["op"]
                want = arrs["y"]
                if op in ("fftfreq", "rfftfreq"):
                    got = getattr(dsc, op)(meta["n"], meta["d"], np.dtype(meta["dtype"]))
                    g = got.numpy()
                    assert g.shape == want.shape and g.dtype == want.dtype, (meta, mode)
                    assert np.array_equal(g, want), (meta, mode, np.max(np.abs(g - want)))
                    continue
                x = dsc.from_numpy(arrs["x"])
                dsc.prefetch(x)                              # device-resident in modes 1 and 2, a no-op in strict mode
                exact = True
                if op == "cast":
                    got = x.cast(np.dtype(meta["to"]))
                elif op.startswith("binary:"):
                    name = op.split(":")[1]
                    got = fn[name](x, dsc.from_numpy(arrs["b"]))
                    exact = name in ("add", "sub")
                elif op == "transpose":
                    got = dsc.transpose(x, meta["axes"] or None)
                elif op == "get_slice":
                    got = x[_py_slices(meta["slices"])]
                elif op == "set_slice":
                    if mode == 2:
                        # make the device copy the only current one, as after a transform in lazy mode
                        x = dsc.add(x, dsc.from_numpy(np.zeros(1, dtype=arrs["x"].dtype)))
                    x[_py_slices(meta["slices"])] = dsc.from_numpy(arrs["b"])
                    got = x
                else:
                    raise AssertionError(op)
                g = got.numpy()
                assert g.shape == want.shape and g.dtype == want.dtype, (meta, mode, g.shape, want.shape)
                if exact:
                    assert np.array_equal(g, want), (meta, mode)
                else:
                    assert rel_l2(g, want) < TOL[want.dtype] * 1e-1, (meta, mode, rel_l2(g, want))
            # spectrum post-processing chained behind a transform, nothing returning to the host in between:
            # magnitude of the upper half of a transposed spectrum, cast to double
            z = randn(np.random.default_rng(6), (8, 128), "complex64")
            Z = dsc.fft(z)
            half = dsc.transpose(Z)[64:, :]
            mag = dsc.abs(half).cast(np.float64)
            ref = np.abs(port.fft(z).T[64:, :]).astype(np.float64)
            assert mag.numpy().dtype == np.float64 and rel_l2(mag.numpy(), ref) < 1e-5
            del Z, half, mag
            # irfft / fused filter with the crop fused into the store (README.md:130-133), single-pass and two-pass orders
            for n_fft, rows in ((2048, 5), (1 << 16, 2)):
                sig = randn(np.random.default_rng(7), (rows, n_fft // 2 - 24), "float32")
                taps = randn(np.random.default_rng(8), (37,), "float32")
                keep = sig.shape[1] + 36
                B = dsc.rfft(taps, n=n_fft)
                y1 = dsc.irfft_keep(dsc.rfft(sig, n=n_fft) * B, keep)
                y2 = dsc.fft_filter_keep(sig, B, keep, n=n_fft)
                assert y1.shape == (rows, keep) and y2.shape == (rows, keep)
                for r in range(rows):
                    w = port.filter_fft(sig[r], taps, n_fft)[:keep]
                    assert rel_l2(y1.numpy()[r], w) < 1e-5 and rel_l2(y2.numpy()[r], w) < 1e-5
                del B, y1, y2
    finally:
        dsc.set_residency(0)


def check_traces(dsc):
    rng = np.random.default_rng(17)
    x = dsc.from_numpy(randn(rng, (4, 64), "float32"))
    dsc.clear_traces()
    dsc.traces_record(True)
    y = dsc.fft(x)
    r = dsc.rfft(x)
    dsc.traces_record(False)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "traces.json")
        dsc.dump_traces(path)
        events = json.load(open(path))
    dsc.clear_traces()
    names = [(e["name"], e["cat"], e["ph"]) for e in events]
    # reference names / categories (dsc_tracing.h:20-21, 90, 94)
    assert ("dsc_internal_fft", "op;fft", "B") in names and ("dsc_internal_fft", "op;fft", "E") in names
    assert ("dsc_internal_rfft", "op;fft", "B") in names
    assert ("dsc_plan_fft", "op;fft;plan", "B") in names
    fft_b = next(e for e in events if e["name"] == "dsc_internal_fft" and e["ph"] == "B")
    assert fft_b["args"]["type"] == "FFT" and fft_b["args"]["order"] == -1 and fft_b["args"]["axis"] == -1
    assert fft_b["args"]["x"]["shape"] == "[4, 64]" and fft_b["args"]["x"]["dtype"] == "f32"
    plan_b = next(e for e in events if e["name"] == "dsc_plan_fft" and e["ph"] == "B")
    assert plan_b["args"] == {"type": "FFT", "n": 64, "order": 64, "dtype": "c32"}
    # new: device spans as complete events with a duration
    gpu = [e for e in events if e["cat"] == "gpu;fft"]
    assert gpu and all(e["ph"] == "X" and e["dur"] >= 0 and e["args"]["n"] in (64, 32) for e in gpu)
    assert any(e["cat"] == "gpu;copy" for e in events)
    del y, r
    # two 4-D tensor descriptions in one args object must still dump as valid JSON
    x4 = dsc.from_numpy(randn(rng, (100, 100, 16, 16), "complex64"))
    o4 = dsc.ifft(x4)
    dsc.traces_record(True)
    o4 = dsc.ifft(x4, out=o4)
    dsc.traces_record(False)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "traces4.json")
        dsc.dump_traces(path)
        events = json.load(open(path))
    dsc.clear_traces()
    b4 = next(e for e in events if e["name"] == "dsc_internal_fft" and e["ph"] == "B")
    assert b4["args"]["x"]["shape"] == "[100, 100, 16, 16]" and b4["args"]["out"]["shape"] == "[100, 100, 16, 16]"
    del x4, o4


def check_composed_paths(dsc):
    """Coverage beyond the kernels' direct range: (a) lengths of more than one shared-memory pass along a
    NON-last axis (transpose, transform, transpose); (b) the host-composed plan used for lengths past the
    four-step range, forced at a small size through the DSC_HUGE_LG testing knob."""
    rng = np.random.default_rng(18)
    x = randn(rng, (1 << 15, 3), "complex64")
    got = dsc.fft(x, axis=0).numpy()
    assert rel_l2(got, port.fft(x, axis=0)) < 1e-6
    # ... and as ONE launch of two column passes when the inner extent is a power of two wide enough for a tile
    xw = randn(rng, (2, 1 << 15, 64), "complex64")
    yw = dsc.fft(xw, axis=1)
    assert rel_l2(yw.numpy(), port.fft(xw, axis=1)) < 1e-6
    assert rel_l2(dsc.ifft(yw, axis=-2).numpy(), xw) < 1e-6
    del yw
    xr = randn(rng, (2, 1 << 16, 2), "float32")
    X = dsc.rfft(xr, axis=1)
    want = port.rfft(xr, -1, 1)
    assert X.shape == want.shape and rel_l2(X.numpy(), want) < 1e-6
    assert rel_l2(dsc.irfft(X, axis=1).numpy(), xr) < 1e-6
    os.environ["DSC_HUGE_LG"] = "11"
    try:
        x = randn(rng, (2, 1 << 14), "complex64")
        y = dsc.fft(x)
        assert rel_l2(y.numpy(), port.fft(x)) < 1e-6
        assert rel_l2(dsc.ifft(y).numpy(), x) < 1e-6
        xs = randn(rng, (10000,), "float32")               # real input, zero-padded to 16384
        assert rel_l2(dsc.fft(xs).numpy(), port.fft(xs)) < 1e-6
    finally:
        os.environ.pop("DSC_HUGE_LG", None)
