"""Timing of the generic (polyhedral) numeric assembly kernel on n^3 hexes with the fast paths off: exp_generic.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mimpy_b200 import _lib, device
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = _lib.Context(0)
r = bench.run_generic_path(ctx, bench.measured_peak()[0], n)
print("generic kernel %d^3: %.3f ms, %.1f Mcells/s, frac %.3f" % (n, r["asm_ms"], r["asm_mcells_per_s"], r["asm_roofline"]["frac"]))
