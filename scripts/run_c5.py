"""BASELINE config 5 (SURVEY 8d "C5"): hex mesh with every other cell (checkerboard) cut by a plane
through its centroid -> mixed cell face counts (7..12 faces), full-tensor K, single-phase solve
through the drop-in API.  usage: run_c5.py [n=32]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mimpy_b200.mesh.hexmesh as hexmesh
import mimpy_b200.mfd.mfd as mfd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nc0 = n ** 3
rng = np.random.default_rng(5)
k = np.zeros((nc0, 3, 3))
for c in range(nc0):
    g = rng.standard_normal(3)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    k[c] = q @ np.diag(np.exp(g)) @ q.T
t0 = time.perf_counter()
mesh = hexmesh.HexMesh()
mesh.build_mesh(n, n, n, 1., 1., 1., k)
nrm = np.array([1.2, .2, 1.3]); nrm /= np.linalg.norm(nrm)
for c in range(nc0):
    i, j, kk = c % n, (c // n) % n, c // (n * n)
    if (i + j + kk) % 2 == 0:
        mesh.divide_cell_by_plane(c, np.array(mesh.get_cell_real_centroid(c)), nrm)
t_mesh = time.perf_counter() - t0
nc, nf = mesh.get_number_of_cells(), mesh.get_number_of_faces()
counts = np.bincount(np.diff(mesh.cells.compact()[0]))
lin = lambda p: 1. + 2. * p[0] - p[1] + .5 * p[2]
print("C5 n=%d: %d cells, %d faces, faces per cell %s; mesh+cuts %.1f s (host Python)" %
      (n, nc, nf, {i: int(v) for i, v in enumerate(counts) if v}, t_mesh))


def run(label):
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    for b in range(6):
        f.apply_dirichlet_from_function(b, lin)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f.build_lhs(); f.build_rhs()
    torch.cuda.synchronize(); t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    sol = f.solve()
    torch.cuda.synchronize(); t_solve = time.perf_counter() - t0
    a = f.lhs.tocsr()
    res = np.linalg.norm(a @ sol - f.rhs) / np.linalg.norm(f.rhs)
    p = f.get_pressure_solution()
    exact = np.array([lin(mesh.get_cell_real_centroid(c)) for c in range(nc)])
    print("   %s: build_lhs+build_rhs (incl. download to scipy) %.3f s, solve %.3f s (%d Jacobi-PCG iterations, %.1f ms on "
          "the device), saddle rel. residual %.1e, max|p - p_linear| = %.2e"
          % (label, t_asm, t_solve, f.solver_info["iterations"], f.solver_info["ms"], res, abs(p - exact).max()))


run("heterogeneous full-tensor K (a linear p is NOT a solution)")
# linear patch test (SURVEY 8c i): with ONE SPD tensor for all cells a linear pressure is reproduced exactly
for c in range(nc):
    mesh.set_cell_k(c, k[0])
run("uniform full-tensor K (linear patch test)")
