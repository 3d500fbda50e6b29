"""BASELINE config 2 (SURVEY 8d "C2"): SPE10-shaped 60x220x85 hex mesh, log-normal anisotropic K,
single-phase pressure solve.  usage: run_c2.py [jacobi_too]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mimpy_b200 import _lib, device

nx, ny, nz = 60, 220, 85
dims = (365.76, 670.56, 51.816)
nc = nx * ny * nz
rng = np.random.default_rng(10)
g = rng.standard_normal((nc, 3))
k = np.zeros((nc, 9))
k[:, 0] = k[:, 4] = np.exp(2.0 * g[:, 0])
k[:, 8] = 0.1 * np.exp(2.0 * g[:, 1])
ctx = _lib.Context(0)
dm = device.DeviceMesh.hex(ctx, nx, ny, nz, *dims, cell_k=k)
fc = device.to_host(ctx, dm.face_centroids); areas = device.to_host(ctx, dm.face_areas)
eps = 1e-6
mask = ((fc[:, 1] < eps) | (fc[:, 1] > dims[1] - eps) | (fc[:, 2] < eps) | (fc[:, 2] > dims[2] - eps)).astype(np.uint8)
plan = device.symbolic(dm, mask)
xm = np.nonzero(fc[:, 0] < eps)[0]; xp = np.nonzero(fc[:, 0] > dims[0] - eps)[0]
rhs = device.build_rhs(dm, plan, np.concatenate([xm, xp]).astype(np.int32),
                       np.concatenate([-areas[xm], np.zeros(len(xp))]), True, 0., None)
vals = torch.empty(plan.nnz, dtype=torch.float64, device=ctx.device)
def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ctx.on_stream():
            e0.record(ctx.stream); out = fn(); e1.record(ctx.stream)
        ctx.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), out
t_asm, _ = timed(lambda: device.numeric(dm, plan, 1, vals=vals))
print("C2 %dx%dx%d = %d cells: assembly %.3f ms (%.0f Mcells/s, %.1f%% of 6546.6 GB/s algorithmic)"
      % (nx, ny, nz, nc, t_asm, nc / t_asm / 1e3, 700. * nc / t_asm / 1e6 / 6546.6 * 100))
m_e, _, _ = device.local_matrices(dm, plan, 1)
F = plan.flux_dof
sols = {}
for pc in (("auxmg", "jacobi") if len(sys.argv) > 1 else ("auxmg",)):
    t_set, hyb = timed(lambda: device.HybridSystem(dm, plan, m_e, precond=pc), reps=2)
    t0 = time.perf_counter()
    sol, info = hyb.solve(rhs, rtol=1e-8, maxit=40000, check_every=5 if pc == "auxmg" else 100, raise_on_noconv=False)
    ctx.sync(); wall = time.perf_counter() - t0
    ax = device.spmv(ctx, plan.indptr, plan.indices, vals, sol, n=plan.ndof)
    res = float(torch.linalg.norm(ax - rhs) / torch.linalg.norm(rhs))
    sols[pc] = device.to_host(ctx, sol)
    print("   %-7s set-up %.1f ms, %d iterations, PCG %.1f ms (%.3f ms/iter), solve wall %.3f s, saddle rel. residual %.2e, mean p %.6f"
          % (pc, t_set, info.iterations, info.ms_total, info.ms_total / max(info.iterations, 1), wall, res, sols[pc][F:].mean()))
    del hyb
if len(sols) == 2:
    print("   |p_aux - p_jac|/|p_jac| = %.2e" % (np.linalg.norm(sols["auxmg"][F:] - sols["jacobi"][F:]) / np.linalg.norm(sols["jacobi"][F:])))
