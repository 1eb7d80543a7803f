"""usage: python tools/prof_rows.py LG ROWS [prec]  -- one (length, batch) case of tools/check_tma.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from check_tma import run
run(int(sys.argv[1]), int(sys.argv[3]) if len(sys.argv) > 3 else 0, rows=int(sys.argv[2]), reps=2)
