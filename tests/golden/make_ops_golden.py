"""Generate tests/golden/ops_golden.npz from the UNMODIFIED reference library: the ops either side of the FFT path
that SURVEY.md 8(f) lists (cast and mixed-dtype arithmetic, transpose, fftfreq / rfftfreq, get_slice / set_slice).

    make -C oracle ref && python tests/golden/make_ops_golden.py

Kept separate from fft_golden.npz so that the existing cases keep their indices and random streams."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_harness import RefLib  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ops_golden.npz")
DTYPES = ("float32", "float64", "complex64", "complex128")


def randn(rng, shape, dtype):
    x = rng.standard_normal(shape)
    if np.dtype(dtype).kind == "c":
        x = x + 1j * rng.standard_normal(shape)
    return x.astype(dtype)


def enc_slices(slices):
    return [[s.start, s.stop, s.step] if isinstance(s, slice) else int(s) for s in slices]


def main():
    ref = RefLib(main_mem=1 << 27, scratch_mem=1 << 25)
    rng = np.random.default_rng(20261019)
    blob, meta = {}, []

    def add(m, **arrays):
        i = len(meta)
        meta.append(m)
        for k, v in arrays.items():
            blob[f"{k}{i}"] = v

    # cast: every ordered pair of dtypes (dsc.cpp:587-597)
    for a in DTYPES:
        x = randn(rng, (5, 37), a)
        for b in DTYPES:
            add({"op": "cast", "to": b}, x=x, y=ref.cast(x, b))
    # mixed-dtype arithmetic: promotion table dsc_dtype.h:73-78 (note f64 (op) c32 -> c32)
    for a in DTYPES:
        for b in DTYPES:
            if a == b:
                continue
            xa = randn(rng, (6, 24), a)
            for bshape in ((6, 24), (24,), (1,)):
                xb = randn(rng, bshape, b)
                for name in ("add", "sub", "mul", "div"):
                    add({"op": f"binary:{name}"}, x=xa, b=xb, y=ref.binary(name, xa, xb))
    # transpose (dsc.cpp:764-827)
    for dt in DTYPES:
        for shape, axes in (((7, 45), ()), ((3, 33, 40), ()), ((3, 33, 40), (0, 2, 1)), ((2, 3, 5, 70), (0, 1, 3, 2)),
                            ((2, 3, 5, 7), (3, 1, 0, 2)), ((4, 6, 8), (1, 0, 2))):
            x = randn(rng, shape, dt)
            add({"op": "transpose", "axes": list(axes)}, x=x, y=ref.transpose(x, axes))
    # fftfreq / rfftfreq (dsc.cpp:2262-2339)
    for dt in ("float32", "float64"):
        for n, d in ((8, 1.0), (9, 1.0), (1000, 0.25), (1001, 1e-3), (1, 2.0), (2, 0.5), (4097, 1 / 48000.0)):
            add({"op": "fftfreq", "n": n, "d": d, "dtype": dt}, y=ref.fftfreq(n, d, dt))
            add({"op": "rfftfreq", "n": n, "d": d, "dtype": dt}, y=ref.fftfreq(n, d, dt, rfft=True))
    # get_slice / set_slice (dsc.cpp:950-1007, 1108-1169)
    sel = [((40,), (slice(3, 31, 2),)), ((40,), (slice(None, 17, None),)), ((6, 50), (slice(None), slice(0, 33, None))),
           ((6, 50), (slice(1, 5, 2), slice(49, 3, -3))), ((4, 5, 30), (2, slice(None), slice(2, 28, 5))),
           ((3, 4, 5, 16), (slice(None), 1, slice(0, 5, 2), slice(-9, None, None)))]
    for dt in DTYPES:
        for shape, slices in sel:
            x = randn(rng, shape, dt)
            got = ref.get_slice(x, slices)
            add({"op": "get_slice", "slices": enc_slices(slices)}, x=x, y=got)
            # the reference compares xb's dims with the selection's dims one by one from the front, indexed dims counting
            # as extent 1 (dsc.cpp:1128-1148): shape the value like that
            it = iter(got.shape)
            vshape = tuple(1 if not isinstance(sl, slice) else next(it) for sl in slices) + tuple(it)
            val = randn(rng, vshape, dt)
            add({"op": "set_slice", "slices": enc_slices(slices)}, x=x, b=val, y=ref.set_slice(x, val, slices))
            one = randn(rng, (1,), dt)
            add({"op": "set_slice", "slices": enc_slices(slices)}, x=x, b=one, y=ref.set_slice(x, one, slices))

    blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    ref.close()
    print(f"{len(meta)} cases -> {OUT} ({os.path.getsize(OUT) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
