"""Generate tests/golden/fft_golden.npz from the UNMODIFIED reference library.

Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

Every case stores the seeded input, the call (op, n, axis) and the bytes the reference
(oracle/_ref/libdsc_ref.so, flags in oracle/Makefile) returned for it.  The fixture is what
pins the C restatement (oracle/dsc_fft_oracle.c) and, on the GPU box where /root/reference
does not exist, the CUDA path.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_harness import RefLib  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fft_golden.npz")

REAL = ("float32", "float64")
CPLX = ("complex64", "complex128")


def randn(rng, shape, dtype):
    x = rng.standard_normal(shape)
    if np.dtype(dtype).kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


def cases():
    """(op, dtype, shape, n, axis) -- mirrors python/tests/test_ops.py:458-489 (all axes,
    crop / copy / pad) plus what that test leaves unpinned."""
    out = []
    # every dtype, 1-D, pow2 and non-pow2 lengths (Appendix A: len 10 -> 16, n=5 -> 8)
    for dt in REAL + CPLX:
        for op in ("fft", "ifft"):
            out += [(op, dt, (16,), -1, -1), (op, dt, (10,), -1, -1), (op, dt, (10,), 5, -1),
                    (op, dt, (1,), -1, -1), (op, dt, (2,), -1, 0), (op, dt, (4, 64), -1, -1)]
    # all axes of a 4-D tensor, crop / copy / pad, like test_fft
    for axis in range(4):
        shape = [3, 4, 2, 5]
        shape[axis] = 16
        for n in (8, 16, 32):
            out.append(("fft", "float64", tuple(shape), n, axis))
            out.append(("ifft", "complex128", tuple(shape), n, axis))
            out.append(("fft", "complex64", tuple(shape), n, axis - 4))
            out.append(("rfft", "float64", tuple(shape), n, axis))
            out.append(("rfft", "float32", tuple(shape), n, axis - 4))
    # longer single lines, every radix schedule boundary up to 2^13
    for lg in range(1, 14):
        out.append(("fft", "complex64", (1 << lg,), -1, -1))
        out.append(("rfft", "float32", (2 << lg,), -1, -1))
        if lg <= 11:
            out.append(("ifft", "complex128", (1 << lg,), -1, -1))
            out.append(("rfft", "float64", (2 << lg,), -1, -1))
    out.append(("fft", "complex64", (3, 4096), -1, -1))
    out.append(("fft", "complex64", (1 << 16,), -1, -1))   # first two-pass size
    # rfft length rules
    for dt in REAL:
        out += [("rfft", dt, (10,), -1, -1), ("rfft", dt, (10,), 4, -1), ("rfft", dt, (2,), -1, -1),
                ("rfft", dt, (6, 32), 64, 1), ("rfft", dt, (32, 6), -1, 0)]
    return out


def irfft_cases():
    """irfft n counts INPUT BINS (dsc.cpp:2197-2200): (dtype, bins-shape, n, axis)."""
    out = []
    for dt in CPLX:
        out += [(dt, (9,), -1, -1), (dt, (9,), 9, -1), (dt, (9,), 16, -1), (dt, (9,), 5, -1),
                (dt, (2,), -1, -1), (dt, (4, 33), -1, -1), (dt, (17, 3), -1, 0), (dt, (3, 17, 2), 9, 1),
                (dt, (513,), -1, -1), (dt, (4097,), -1, -1)]
    out.append(("complex64", (8193,), -1, -1))
    return out


def main():
    ref = RefLib(main_mem=1 << 28, scratch_mem=1 << 26)
    rng = np.random.default_rng(20261018)
    blob, meta = {}, []

    def add(op, x, n, axis, y):
        i = len(meta)
        meta.append({"op": op, "n": int(n), "axis": int(axis)})
        blob[f"x{i}"] = x
        blob[f"y{i}"] = y

    for op, dt, shape, n, axis in cases():
        x = randn(rng, shape, dt)
        add(op, x, n, axis, getattr(ref, op)(x, n, axis))
    for dt, shape, n, axis in irfft_cases():
        x = randn(rng, shape, dt)
        add("irfft", x, n, axis, ref.irfft(x, n, axis))

    # deterministic sanity vectors: impulse, constant, single tone
    for dt in CPLX:
        imp = np.zeros(64, dtype=dt); imp[3] = 1
        add("fft", imp, -1, -1, ref.fft(imp))
        add("fft", np.ones(64, dtype=dt), -1, -1, ref.fft(np.ones(64, dtype=dt)))
        tone = np.exp(2j * np.pi * 5 * np.arange(256) / 256).astype(dt)
        add("fft", tone, -1, -1, ref.fft(tone))

    # BASELINE configs[0]: README filterFFT (README.md:118-134), seeds per SURVEY.md 8(d)
    s = np.random.default_rng(0).standard_normal(8192).astype(np.float32)
    b = np.random.default_rng(1).standard_normal(128).astype(np.float32)
    i = len(meta)
    meta.append({"op": "filter", "n": 16384, "axis": -1})
    blob[f"x{i}"] = s
    blob[f"b{i}"] = b
    blob[f"y{i}"] = ref.filter_fft(s, b, 16384)

    # complex product used between rfft and irfft
    for dt in CPLX:
        a = randn(rng, (5, 129), dt)
        w = randn(rng, (129,), dt)
        i = len(meta)
        meta.append({"op": "mul", "n": -1, "axis": -1})
        blob[f"x{i}"] = a
        blob[f"b{i}"] = w
        blob[f"y{i}"] = ref.mul(a, w)

    # plan-cache behaviour (Appendix A): used_mem after sweeps of >16 distinct plans
    base = ref.used_mem()
    sweep = []
    for rnd in range(2):
        for lg in range(1, 21):
            x = randn(rng, (1 << lg,), "complex64")
            ref.fft(x)
        sweep.append(ref.used_mem() - base)
    meta.append({"op": "plan_sweep", "n": 20, "axis": -1, "used_after_round": sweep})

    # spectrum post-processing and arithmetic, appended last so that the earlier cases keep
    # their indices and random streams: |X|, arg X, Re, Im, conj; add/sub/mul/div with same-shape, row and scalar b
    for dt in CPLX:
        z = randn(rng, (7, 33), dt)
        for name in ("abs", "angle", "real", "imag", "conj"):
            i = len(meta)
            meta.append({"op": f"unary:{name}", "n": -1, "axis": -1})
            blob[f"x{i}"] = z
            blob[f"y{i}"] = ref.unary(name, z)
    for dt in REAL + CPLX:
        a = randn(rng, (6, 40), dt)
        for bshape in ((6, 40), (40,), (1,)):
            b = randn(rng, bshape, dt)
            for name in ("add", "sub", "mul", "div"):
                i = len(meta)
                meta.append({"op": f"binary:{name}", "n": -1, "axis": -1})
                blob[f"x{i}"] = a
                blob[f"b{i}"] = b
                blob[f"y{i}"] = ref.binary(name, a, b)

    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    ref.close()
    print(f"{len(meta)} cases -> {OUT} ({os.path.getsize(OUT) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
