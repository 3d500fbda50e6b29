"""SELL SpMV of the face system: indexed vs column patterns.  exp_spmv.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mimpy_b200 import _lib, device
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = _lib.Context(0)
pb = bench.build_problem(ctx, n, 0, 1)
dm, plan = pb["dm"], pb["plan"]
hyb = device.HybridSystem(dm, plan, None, precond="jacobi", method=1)
F = plan.flux_dof
pid = plan.t.get("sell_pat_id")
print("n=%d F=%d npat=%d" % (n, F, getattr(plan, "sell_npat", -1)))
if pid is not None:
    esc = int((pid == 255).sum())
    print("escaped rows: %d (%.2f %%)" % (esc, 100. * esc / pid.numel()))
    cnt = torch.bincount(pid.to(torch.int64), minlength=256)
    top = torch.sort(cnt, descending=True)
    print("top pattern shares:", [round(float(v) / pid.numel(), 4) for v in top.values[:8]])
with ctx.on_stream():
    xs = torch.ones(F, dtype=torch.float64, device=ctx.device); ys = torch.empty_like(xs)
def timeit(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ctx.on_stream():
        for i in range(reps + 3):
            if i == 3: e0.record(ctx.stream)
            fn()
        e1.record(ctx.stream)
    ctx.sync(); return e0.elapsed_time(e1) / reps
P = _lib.ptr
t_idx = timeit(lambda: ctx.call("mimpy_b200_spmv_sell", F, P(hyb.sell_ptr), P(hyb.sell_cols), P(hyb.a_vals), P(xs), P(ys)))
print("indexed  %.3f ms" % t_idx)
if getattr(plan, "sell_npat", 0) > 0:
    t_pat = timeit(lambda: ctx.call("mimpy_b200_spmv_sell_pat", F, P(hyb.sell_ptr), P(hyb.sell_cols), P(plan.sell_pat_id), P(plan.sell_pat_table), plan.sell_npat, P(hyb.a_vals), P(xs), P(ys)))
    print("patterns %.3f ms" % t_pat)
