"""Iteration count of the bench solve at n^3 under the current environment knobs: exp_iters.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mimpy_b200 import _lib, device
import bench
ctx = _lib.Context(0)
if len(sys.argv) > 1 and sys.argv[1] == "c2":
    r = bench.run_c2(ctx, 6546.6, 1e-8)
    print("C2 env=%s iterations=%d solve_s=%.3f saddle=%.2e" % ({k: v for k, v in os.environ.items() if k.startswith("MIMPY_B200_AUX")},
                                                                r["pcg_iters"], r["solve_s"], r["saddle_rel_residual"]))
    sys.exit(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pb = bench.build_problem(ctx, n, 0, 1)
dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
vals = device.numeric(dm, plan, 1)
hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1)
sol, info = hyb.solve(rhs, rtol=1e-8, maxit=3000, check_every=5, norm="saddle", saddle_vals=vals, raise_on_noconv=False)
print("n=%d env=%s iterations=%d ms=%.1f" % (n, {k: v for k, v in os.environ.items() if k.startswith("MIMPY_B200_AUX")}, info.iterations, info.ms_total))
