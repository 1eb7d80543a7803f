"""One two-pass length, a few launches: the command line profiled under ncu.  usage: python tools/prof_one.py LG [prec] [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from check_tma import run

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
run(lg, prec, reps=reps)
