"""Timing of the numeric assembly kernel variants: exp_asm.py [n] [all]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mimpy_b200 import _lib, device
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = _lib.Context(0); dev = ctx.device; nc = n ** 3
dm = device.DeviceMesh.hex(ctx, n, n, n, 1., 1., 1., cell_k=bench.lognormal_k(nc, 256))
with ctx.on_stream():
    fc = dm.face_centroids; eps = 1e-9
    mask = torch.zeros(dm.nf, dtype=torch.uint8, device=dev)
    mask[(fc[:, 1] < eps) | (fc[:, 1] > 1 - eps) | (fc[:, 2] < eps) | (fc[:, 2] > 1 - eps)] = 1
plan = device.symbolic(dm, mask)
vals = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
def timeit(label, **kw):
    for _ in range(2): device.numeric(dm, plan, 1, vals=vals, **kw)
    ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with ctx.on_stream():
            e0.record(ctx.stream); device.numeric(dm, plan, 1, vals=vals, **kw); e1.record(ctx.stream)
        ctx.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("%-40s %.3f ms  (%.0f Mcells/s, %.1f%% of 6546.6 GB/s algorithmic)" % (label, min(ts), nc / min(ts) / 1e3, 700. * nc / min(ts) / 1e6 / 6546.6 * 100))
ref = device.to_host(ctx, device.numeric(dm, plan, 1, generic_kernel=True)) if n <= 128 else None
timeit("structured-hex kernel, MIMPY_B200_ASM_VARIANT=%s" % os.environ.get("MIMPY_B200_ASM_VARIANT", "(default: pipelined brick)"))
if len(sys.argv) > 2:
    timeit("direct", debug_flags=4)
    timeit("generic warp-group kernel", generic_kernel=True)
if ref is not None:
    out = device.to_host(ctx, device.numeric(dm, plan, 1, vals=vals))
    print("max |brick - generic| / max = %.2e" % (abs(out - ref).max() / abs(ref).max()))
