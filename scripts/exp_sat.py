"""Device time of the saturation update of BASELINE config 4 in its three forms (twophase.cu): exp_sat.py [n] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mimpy_b200 import _lib
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = _lib.Context(0)
tp, _, _ = bench.build_c4(ctx, n)
tp.initialize_system()
for step in (1, 2):
    tp.update_pressure(step)
    if step == 1:
        tp.find_upwinding_direction()
    tp.update_saturation(step)
ctx.sync()
for form in ("", "MIMPY_B200_SAT_NOHEX", "MIMPY_B200_SAT_PERFACE", "MIMPY_B200_SAT_FUSED"):
    for k in ("MIMPY_B200_SAT_PERFACE", "MIMPY_B200_SAT_FUSED", "MIMPY_B200_SAT_NOHEX"):
        os.environ.pop(k, None)
    if form:
        os.environ[form.split("=")[0]] = form.split("=")[1] if "=" in form else "1"
    ts = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ctx.on_stream():
            e0.record(ctx.stream)
            tp.update_saturation(3 + r)
            e1.record(ctx.stream)
        ctx.sync()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("form %-24s min %.1f us  median %.1f us" % (form or "two-pass (default)", min(ts[2:]), float(np.median(ts[2:]))))
