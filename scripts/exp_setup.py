"""Breakdown of the face-system set-up at n^3: exp_setup.py [n]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mimpy_b200 import _lib, device
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = _lib.Context(0)
pb = bench.build_problem(ctx, n, 0, 1)
dm, plan = pb["dm"], pb["plan"]
def timed(label, fn, reps=3):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ctx.on_stream():
            e0.record(ctx.stream); out = fn(); e1.record(ctx.stream)
        ctx.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); del out
    print("%-46s %.3f ms" % (label, min(ts)))
timed("HybridSystem(auxmg), whole constructor", lambda: device.HybridSystem(dm, plan, None, precond="auxmg", method=1))
timed("HybridSystem(jacobi), whole constructor", lambda: device.HybridSystem(dm, plan, None, precond="jacobi", method=1))
hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1)
timed("setup() only (fused set-up + auxmg_setup)", lambda: hyb.setup(None))
hj = device.HybridSystem(dm, plan, None, precond="jacobi", method=1)
timed("setup() only, jacobi (fused set-up alone)", lambda: hj.setup(None))
