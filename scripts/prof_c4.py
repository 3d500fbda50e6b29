"""Where an IMPES step of BASELINE config 4 goes: every library call and the torch glue of TwoPhase.update_pressure /
update_saturation timed with a stream synchronisation on both sides (so the shares add up; the unsynchronised step is
printed beside them).  usage: prof_c4.py [n] [steps]"""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mimpy_b200 import _lib, device
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = _lib.Context(0)
tp, _, _ = bench.build_c4(ctx, n)
tp.initialize_system()
ctx.sync()


def run(nsteps, first):
    out = []
    for step in range(first, first + nsteps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tp.update_pressure(step)
        if step == 1 or step % 10 == 0:
            tp.find_upwinding_direction()
        tp.update_saturation(step)
        ctx.sync()
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    return out


plain = run(steps, 1)
print("unsynchronised steps (ms):", ["%.2f" % t for t in plain], "pcg iterations", tp.last_solver_info.iterations)

acc = collections.OrderedDict()
depth = [0]


def timed(name, fn):
    def w(*a, **kw):
        if depth[0]:
            return fn(*a, **kw)
        depth[0] += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            return fn(*a, **kw)
        finally:
            torch.cuda.synchronize()
            e = acc.setdefault(name, [0, 0.])
            e[0] += 1
            e[1] += (time.perf_counter() - t0) * 1e3
            depth[0] -= 1
    return w


orig_call = ctx.call
ctx.call = lambda name, *a, **kw: timed(name, lambda: orig_call(name, *a, **kw))()
device.numeric = timed("device.numeric", device.numeric)
device.HybridSystem.setup = timed("HybridSystem.setup", device.HybridSystem.setup)
device.HybridSystem.solve = timed("HybridSystem.solve", device.HybridSystem.solve)
synced = run(steps, steps + 1)
print("synchronised steps (ms):", ["%.2f" % t for t in synced])
tot = sum(synced)
named = 0.
for name, (cnt, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("  %-44s %4d calls  %8.2f ms/step  %5.1f %%" % (name, cnt, ms / steps, 100. * ms / tot))
    named += ms
print("  %-44s            %8.2f ms/step  %5.1f %%  (torch glue + Python between the calls)" % ("rest", (tot - named) / steps, 100. * (tot - named) / tot))
info = tp.last_solver_info
print("last solve:", {k: getattr(info, k) for k in dir(info) if not k.startswith("_") and not callable(getattr(info, k))})
