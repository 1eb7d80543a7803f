"""rfft / irfft at the longest single-pass real lengths (one block per SM): python tools/bench_real_ab.py
(DSC_NO_PERSIST=1 for the A/B of the persistent launch)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_real  # noqa: F401  (runs its short list on import, see below)
