"""Jacobi vs auxiliary-space multigrid PCG on the C3 problem: exp_precond.py n [n ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mimpy_b200 import _lib, device
import bench

ctx = _lib.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [32]:
    pb = bench.build_problem(ctx, n, 0, 1)
    dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
    m_e, _, _ = device.local_matrices(dm, plan, 1)
    out = {}
    for pc in ("jacobi", "auxmg"):
        hyb = device.HybridSystem(dm, plan, m_e, precond=pc)
        sol, info = hyb.solve(rhs, rtol=1e-8, maxit=20000, check_every=5 if pc == "auxmg" else 50, raise_on_noconv=False)
        sol, info = hyb.solve(rhs, rtol=1e-8, maxit=20000, check_every=5 if pc == "auxmg" else 50, raise_on_noconv=False)
        out[pc] = device.to_host(ctx, sol)
        print("n=%d %-7s iters %5d  pcg %.1f ms (%.3f ms/iter)  rel.res %.2e" % (
            n, pc, info.iterations, info.ms_total, info.ms_total / max(info.iterations, 1), info.residual / info.rhs_norm))
        del hyb
    F = plan.flux_dof
    print("   |p_aux - p_jac| / |p_jac| = %.2e" % (np.linalg.norm(out["auxmg"][F:] - out["jacobi"][F:]) / np.linalg.norm(out["jacobi"][F:])))
