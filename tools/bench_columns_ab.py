"""Same-process timing of long columns (two-pass transforms along a non-last axis): python tools/bench_columns_ab.py
(DSC_COLUMNS_E16=0 selects 32 points per thread for the A/B)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_axes import run  # noqa: F401  (bench_axes runs its own list on import)
