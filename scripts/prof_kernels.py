"""Small driver for ncu: runs the numeric assembly / hybrid set-up / PCG kernels a few times.
usage: python scripts/prof_kernels.py [n] [what]   what in {asm, pcg, all, bench}
bench = the path bench.py times: brick assembly, fused set-up, auxmg PCG (10 iterations, no graph replay)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mimpy_b200 import _lib, device
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
what = sys.argv[2] if len(sys.argv) > 2 else "all"
ctx = _lib.Context(0)
dev = ctx.device
nc = n ** 3
dm = device.DeviceMesh.hex(ctx, n, n, n, 1., 1., 1., cell_k=bench.lognormal_k(nc, 256))
with ctx.on_stream():
    fc = dm.face_centroids
    eps = 1e-9
    mask = torch.zeros(dm.nf, dtype=torch.uint8, device=dev)
    mask[(fc[:, 1] < eps) | (fc[:, 1] > 1 - eps) | (fc[:, 2] < eps) | (fc[:, 2] > 1 - eps)] = 1
    d1 = torch.nonzero(fc[:, 0] < eps).flatten().to(torch.int32)
    d0 = torch.nonzero(fc[:, 0] > 1 - eps).flatten().to(torch.int32)
    faces = torch.cat([d1, d0]).cpu().numpy()
    vals_d = torch.cat([-dm.face_areas[d1.long()], torch.zeros(d0.numel(), dtype=torch.float64, device=dev)]).cpu().numpy()
plan = device.symbolic(dm, mask)
rhs = device.build_rhs(dm, plan, faces, vals_d, True, 0., None)
vals = None
for it in range(3):
    if what in ("asm", "all"):
        vals = device.numeric(dm, plan, 1, vals=vals)
    if what in ("pcg", "all"):
        m_e, _, _ = device.local_matrices(dm, plan, 1)
        hyb = device.HybridSystem(dm, plan, m_e)
        sol, info = hyb.solve(rhs, rtol=1e-30, maxit=20, check_every=10, raise_on_noconv=False, use_graph=False)
if what == "bench":
    for it in range(2):
        vals = device.numeric(dm, plan, 1, vals=vals)
        hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1)
        sol, info = hyb.solve(rhs, rtol=1e-30, maxit=10, check_every=10, raise_on_noconv=False, use_graph=False)
        del hyb
ctx.sync()
print("done", n, what)
