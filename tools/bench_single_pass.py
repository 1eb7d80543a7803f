"""fwd+inv GB/s of single-pass lengths through the device-level C ABI, with the round trip checked.
usage: python tools/bench_single_pass.py [prec] [lg ...]   (DSC_NO_PERSIST=1 for the A/B of the persistent launch)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from check_tma import run

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for lg in [int(v) for v in sys.argv[2:]] or [12, 13, 14]:
    run(lg, prec, rows=None)
    run(lg, prec, rows=149, reps=2)
