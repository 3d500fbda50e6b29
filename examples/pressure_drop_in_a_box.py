"""Single-phase pressure solve on a heterogeneous box through the mimpy-compatible API.

The same calls work with `import mimpy.mesh.hexmesh as hexmesh; import mimpy.mfd.mfd as mfd` (the
reference, CPU) -- here they run on the GPU through libmimpy_b200.so.  Needs a CUDA device.

    python examples/pressure_drop_in_a_box.py [n=48]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mimpy_b200.mesh.hexmesh as hexmesh  # noqa: E402
import mimpy_b200.mfd.mfd as mfd           # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
rng = np.random.default_rng(7)
# permeability: one 3x3 tensor per cell (an array instead of the reference's per-cell callback)
log_k = rng.standard_normal(n ** 3)
K = np.exp(log_k)[:, None, None] * np.diag([1., 1., 0.2])[None]

mesh = hexmesh.HexMesh()
mesh.build_mesh(n, n, n, 1., 1., 0.5, K)

solver = mfd.MFD()
solver.set_m_e_construction_method(1)
solver.set_mesh(mesh)
solver.apply_dirichlet_from_function(0, lambda p: 1.)          # x = 0 : p = 1
solver.apply_dirichlet_from_function(1, lambda p: 0.)          # x = 1 : p = 0
for marker in (2, 3, 4, 5):                                    # no flow through the other sides
    solver.apply_neumann_from_function(marker, lambda p: np.zeros(3))

t0 = time.perf_counter()
solver.build_lhs()
solver.build_rhs()
t1 = time.perf_counter()
solver.solve()
t2 = time.perf_counter()
p = solver.get_pressure_solution()
print("%d cells, %d flux unknowns: assembly (incl. the copy into SciPy) %.3f s, solve %.3f s, %d PCG iterations"
      % (mesh.get_number_of_cells(), solver.flux_dof, t1 - t0, t2 - t1, solver.solver_info["iterations"]))
print("pressure in [%.4f, %.4f], mean %.4f" % (p.min(), p.max(), p.mean()))
res = np.linalg.norm(solver.lhs.tocsr() @ solver.solution - solver.rhs) / np.linalg.norm(solver.rhs)
print("relative residual of the assembled saddle system: %.1e" % res)
mesh.output_vtk_mesh("pressure_drop_in_a_box", [p, log_k], ["pressure", "log_permeability"])
print("wrote pressure_drop_in_a_box.vtk")
