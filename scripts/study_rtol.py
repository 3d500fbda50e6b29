#!/usr/bin/env python
"""Which PCG stopping test gives |x - x_spsolve| / |x_spsolve| <= 1e-8 on the C3 family?

For n^3 hexes (log-normal diagonal K, seed 256, method 1, Dirichlet x-/x+, no-flow elsewhere --
bench.py's workload) the bench path (DeviceMesh.hex, fused hybrid set-up, auxmg PCG, check_every 5)
is run for k = 5, 10, ... iterations; each line gives the face residual relative to ||h|| and to the
saddle right-hand side, the preconditioned residual sqrt(rz/rz0), the residual of the ORIGINAL
saddle system and the distance to the reference solution (SuperLU spsolve of the assembled saddle
CSR for n <= 32; the solve run to stagnation beyond).

    python scripts/study_rtol.py 16 24 256
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, kmax):
    import torch
    from scipy import sparse
    from scipy.sparse.linalg import spsolve

    import bench
    from mimpy_b200 import _lib, device

    ctx = _lib.default_context(0)
    pb = bench.build_problem(ctx, n, 0, 1)
    dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
    vals = device.numeric(dm, plan, 1)
    H = lambda t: device.to_host(ctx, t)
    ref = None
    if n <= 32:
        nd = plan.ndof
        a = sparse.csr_matrix((H(vals)[:plan.nnz], H(plan.indices)[:plan.nnz], H(plan.indptr)[:nd + 1]), shape=(nd, nd))
        t0 = time.perf_counter()
        ref = torch.as_tensor(spsolve(a.tocsc(), H(rhs)), device=ctx.device)
        print("n=%d spsolve %.1f s" % (n, time.perf_counter() - t0), flush=True)
    hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1)
    bnorm = float(torch.linalg.norm(rhs))
    if ref is None:
        ref, info = hyb.solve(rhs, rtol=0., maxit=kmax + 100, check_every=5, raise_on_noconv=False)
        print("n=%d reference = %d iterations, face residual %.3e" % (n, info.iterations, info.residual))
    refn = float(torch.linalg.norm(ref))
    print("n=%d ||b_saddle||=%.4e" % (n, bnorm))
    print("%5s %11s %11s %11s %11s %11s %11s %8s" % ("iters", "r/|h|", "r/|b|", "sqrt(rz/rz0)", "saddle_res", "err_x", "err_p", "ms"))
    for k in range(5, kmax + 1, 5):
        sol, info = hyb.solve(rhs, rtol=0., maxit=k, check_every=5, raise_on_noconv=False)
        with ctx.on_stream():
            ax = device.spmv(ctx, plan.indptr, plan.indices, vals, sol, n=plan.ndof)
            sres = float(torch.linalg.norm(ax - rhs)) / bnorm
            err = float(torch.linalg.norm(sol - ref)) / refn
            F = plan.flux_dof
            errp = float(torch.linalg.norm(sol[F:] - ref[F:]) / torch.linalg.norm(ref[F:]))
        print("%5d %11.3e %11.3e %11.3e %11.3e %11.3e %11.3e %8.1f" % (
            info.iterations, info.residual / info.rhs_norm, info.residual / bnorm,
            np.sqrt(abs(info.rz) / info.rz0), sres, err, errp, info.ms_total), flush=True)
    del hyb


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [16, 24]
    for n in sizes:
        run(n, 90 if n <= 64 else 120)
