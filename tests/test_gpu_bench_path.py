"""The EXACT code path bench.py times (DeviceMesh.hex with even nx -> pipelined brick assembly,
HybridSystem(dm, plan, None, precond="auxmg", method=1) = fused hex set-up, PCG with check_every=5 and
the bench's stopping test) against SciPy's SuperLU on the assembled saddle matrix and against the
oracle's assembly (reference algorithm, mfd.py:545-599), at the BASELINE shapes C2 and C3 scaled to
sizes spsolve finishes in seconds.  north_star gate: |x - x_spsolve| / |x_spsolve| <= 1e-8."""
import numpy as np
import pytest
from scipy import sparse
from scipy.sparse.linalg import spsolve

from oracle import mimpy_oracle as orc

pytestmark = pytest.mark.gpu

BENCH_RTOL = 1e-8   # bench.py --rtol default, measured against ||b_saddle||_2


@pytest.fixture(scope="module")
def ctx():
    import torch
    from mimpy_b200 import _lib

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib.default_context()


def H(ctx, t):
    from mimpy_b200 import device

    return device.to_host(ctx, t)


def _csr(ctx, plan, vals):
    n = plan.ndof
    return sparse.csr_matrix((H(ctx, vals)[:plan.nnz], H(ctx, plan.indices)[:plan.nnz], H(ctx, plan.indptr)[:n + 1]),
                             shape=(n, n))


@pytest.mark.parametrize("n", [16, 24])
def test_bench_path_vs_spsolve(ctx, n):
    """C3 family (bench.build_problem: log-normal diagonal K seed 256, method 1, Dirichlet x-/x+,
    no-flow elsewhere): bench-path solve at the bench's tolerance vs spsolve; n = 16 also checks the
    brick-kernel CSR against the oracle's per-cell assembly."""
    import torch
    import bench
    from mimpy_b200 import device

    pb = bench.build_problem(ctx, n, 0, 1)
    dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
    vals = device.numeric(dm, plan, 1)
    a = _csr(ctx, plan, vals)
    b = H(ctx, rhs)
    if n == 16:
        o = orc.hex_mesh(n, n, n, 1., 1., 1., k_array=bench.lognormal_k(n ** 3, 256))
        zero = lambda p: np.array([0., 0., 0.])
        f = orc.OracleMFD(o, 1)
        f.apply_dirichlet_from_function(0, lambda p: 1.)
        f.apply_dirichlet_from_function(1, lambda p: 0.)
        for m in (2, 3, 4, 5):
            f.apply_neumann_from_function(m, zero)
        f.build_lhs(m_e_list=list(orc.build_m_e_batched(o, 1)))
        f.build_rhs()
        ref_a = f.lhs.tocsr()
        assert abs(a - ref_a).max() <= 1e-12 * abs(ref_a).max()
        assert np.array_equal(b, f.rhs)
    ref = spsolve(a.tocsc(), b)
    hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1)
    sol, info = hyb.solve(rhs, rtol=BENCH_RTOL, maxit=2000, check_every=5, norm="saddle", saddle_vals=vals)
    x = H(ctx, sol)
    err = np.linalg.norm(x - ref) / np.linalg.norm(ref)
    F = plan.flux_dof
    err_p = np.linalg.norm(x[F:] - ref[F:]) / np.linalg.norm(ref[F:])
    res = np.linalg.norm(a @ x - b) / np.linalg.norm(b)
    assert err <= 1e-8 and err_p <= 1e-8, (err, err_p, info.iterations)
    assert res <= 2e-8, res                       # the face residual tracks the saddle residual
    assert info.iterations <= 80
    # the looser stopping test of round 1 (rtol 1e-8 of ||h||) does NOT meet the gate: keep it out
    sol2, info2 = hyb.solve(rhs, rtol=1e-8, maxit=2000, check_every=5, norm="face")
    err2 = np.linalg.norm(H(ctx, sol2) - ref) / np.linalg.norm(ref)
    assert info2.iterations < info.iterations and err2 > 1e-8


def test_bench_path_c2_shape_vs_spsolve(ctx):
    """C2 (SPE10-shaped 60x220x85) scaled to 12x44x17: cell size 6.1 x 3.05 x 0.61 m, k_v = 0.1 k_h,
    sigma = 2 log-normal; fused set-up + auxmg PCG vs spsolve."""
    import torch
    from mimpy_b200 import device

    nx, ny, nz = 12, 44, 17
    dims = (365.76 * nx / 60., 670.56 * ny / 220., 51.816 * nz / 85.)
    nc = nx * ny * nz
    g = np.random.default_rng(10).standard_normal((nc, 3))
    k = np.zeros((nc, 9))
    k[:, 0] = k[:, 4] = np.exp(2.0 * g[:, 0])
    k[:, 8] = 0.1 * np.exp(2.0 * g[:, 1])
    dm = device.DeviceMesh.hex(ctx, nx, ny, nz, *dims, cell_k=k)
    fc, areas = H(ctx, dm.face_centroids), H(ctx, dm.face_areas)
    eps = 1e-6
    mask = ((fc[:, 1] < eps) | (fc[:, 1] > dims[1] - eps) | (fc[:, 2] < eps) | (fc[:, 2] > dims[2] - eps)).astype(np.uint8)
    plan = device.symbolic(dm, mask)
    xm, xp = np.nonzero(fc[:, 0] < eps)[0], np.nonzero(fc[:, 0] > dims[0] - eps)[0]
    rhs = device.build_rhs(dm, plan, np.concatenate([xm, xp]).astype(np.int32),
                           np.concatenate([-areas[xm], np.zeros(len(xp))]), True, 0., None)
    vals = device.numeric(dm, plan, 1)
    a, b = _csr(ctx, plan, vals), H(ctx, rhs)
    # the same matrix from the oracle (reference algorithm) on the same mesh
    o = orc.hex_mesh(nx, ny, nz, *dims, k_array=k)
    zero = lambda p: np.array([0., 0., 0.])
    f = orc.OracleMFD(o, 1)
    f.apply_dirichlet_from_function(0, lambda p: 1.)
    f.apply_dirichlet_from_function(1, lambda p: 0.)
    for m in (2, 3, 4, 5):
        f.apply_neumann_from_function(m, zero)
    f.build_lhs(m_e_list=list(orc.build_m_e_batched(o, 1)))
    ref_a = f.lhs.tocsr()
    assert abs(a - ref_a).max() <= 1e-12 * abs(ref_a).max()
    ref = spsolve(a.tocsc(), b)
    hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1)
    sol, info = hyb.solve(rhs, rtol=BENCH_RTOL, maxit=5000, check_every=5, norm="saddle", saddle_vals=vals)
    x = H(ctx, sol)
    F = plan.flux_dof
    assert np.linalg.norm(a @ x - b) <= 1.05 * BENCH_RTOL * np.linalg.norm(b)     # the verified stopping rule
    err = np.linalg.norm(x - ref) / np.linalg.norm(ref)
    err_p = np.linalg.norm(x[F:] - ref[F:]) / np.linalg.norm(ref[F:])
    assert err <= 1e-8 and err_p <= 1e-8, (err, err_p, info.iterations)


def test_host_pipeline_entry_points(ctx):
    """mimpy_b200/hostpipe.py (what bench.py's e2e figures time): slab-pipelined assembly host -> host equals a
    plain device assembly bit for bit, twice in a row; solve_to_host equals the device solve."""
    import torch
    import bench
    from mimpy_b200 import device, hostpipe

    n = 32
    pb = bench.build_problem(ctx, n, 0, 1)
    dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
    ref_vals = H(ctx, device.numeric(dm, plan, 1))
    k_pin = torch.from_numpy(pb["k_host"]).pin_memory().reshape(dm.nc, 9)
    vals_pin = torch.empty(plan.nnz, dtype=torch.float64).pin_memory()
    pipe = hostpipe.AssemblyPipeline(dm, plan, 1, n_slabs=4)
    assert pipe.piped and len(pipe.bounds) - 1 == 4
    for _ in range(2):
        vals_pin.fill_(float("nan"))
        with ctx.on_stream():
            dm.cell_k.zero_()                      # the pipeline must bring K itself
        pipe.run(k_pin, vals_pin)
        assert np.array_equal(vals_pin.numpy(), ref_vals)
    sol_pin = torch.empty(plan.ndof, dtype=torch.float64).pin_memory()
    info = hostpipe.solve_to_host(dm, plan, 1, k_pin, rhs, sol_pin, rtol=BENCH_RTOL, saddle_vals=device.numeric(dm, plan, 1),
                                  maxit=2000, check_every=5)
    hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1)
    sol, info2 = hyb.solve(rhs, rtol=BENCH_RTOL, maxit=2000, check_every=5, norm="saddle", saddle_vals=device.numeric(dm, plan, 1))
    assert info.iterations == info2.iterations
    assert np.array_equal(sol_pin.numpy(), H(ctx, sol))


def test_block_smoother_vs_jacobi_smoother(ctx):
    """The two face-space terms of the auxiliary-space preconditioner (weighted additive Schwarz over the 6 x 6
    cell blocks / D^-1) solve the same system to the gate; on anisotropic K the blocks need clearly fewer
    iterations, "auto" picks them there and picks Jacobi for a scalar K (csrc/auxmg.cu:k_aux_cellblocks)."""
    import bench
    from mimpy_b200 import device

    n = 24
    pb = bench.build_problem(ctx, n, 0, 1)
    dm, plan, rhs = pb["dm"], pb["plan"], pb["rhs"]
    vals = device.numeric(dm, plan, 1)
    a = _csr(ctx, plan, vals)
    ref = spsolve(a.tocsc(), H(ctx, rhs))
    its = {}
    for sm in ("blocks", "jacobi"):
        for layout in ("sell", "csr"):
            hyb = device.HybridSystem(dm, plan, None, precond="auxmg", method=1, smoother=sm, layout=layout)
            sol, info = hyb.solve(rhs, rtol=BENCH_RTOL, maxit=2000, check_every=1, norm="saddle", saddle_vals=vals)
            err = np.linalg.norm(H(ctx, sol) - ref) / np.linalg.norm(ref)
            assert err <= 1e-8, (sm, layout, err)
            its[sm, layout] = info.iterations
    assert its["blocks", "sell"] == its["blocks", "csr"] and its["jacobi", "sell"] == its["jacobi", "csr"]
    assert its["blocks", "sell"] <= 0.8 * its["jacobi", "sell"], its
    assert device.HybridSystem(dm, plan, None, precond="auxmg", method=1).smoother == "blocks"
    k_iso = np.zeros((n ** 3, 9))
    k_iso[:, 0] = k_iso[:, 4] = k_iso[:, 8] = np.exp(np.random.default_rng(1).standard_normal(n ** 3))
    dm2 = device.DeviceMesh.hex(ctx, n, n, n, 1., 1., 1., cell_k=k_iso)
    plan2 = device.symbolic(dm2, pb["slab"].neumann_mask())
    assert device.HybridSystem(dm2, plan2, None, precond="auxmg", method=1).smoother == "jacobi"
