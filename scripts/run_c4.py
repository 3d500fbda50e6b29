"""BASELINE config 4 (SURVEY 8d "C4"): two-phase IMPES (MFD pressure + upwind saturation) on an n^3
hex mesh through the drop-in API.  usage: run_c4.py [n=128] [steps=10]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mimpy_b200.mesh.hexmesh as hexmesh
import mimpy_b200.models.twophase as twophase

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nc = n ** 3
rng = np.random.default_rng(128)
k = np.zeros((nc, 9))
kk = 1e-13 * np.exp(rng.standard_normal(nc))
k[:, 0] = k[:, 4] = k[:, 8] = kk
t0 = time.perf_counter()
mesh = hexmesh.HexMesh()
mesh.build_mesh(n, n, n, float(n), float(n), float(n), k)
t_mesh = time.perf_counter() - t0
zero_flux = lambda p: np.array([0., 0., 0.])
tp = twophase.TwoPhase()
tp.set_mesh(mesh)
for b in (0, 2, 3, 4, 5):
    tp.apply_flux_boundary_from_function(b, zero_flux)
tp.apply_pressure_boundary_from_function(1, lambda p: 0.)
tp.set_initial_p_o(np.zeros(nc)); tp.set_initial_s_w(np.zeros(nc)); tp.set_porosities(np.full(nc, .2))
tp.set_viscosity_water(8.9e-4); tp.set_viscosity_oil(1.78e-3)
tp.set_compressibility_water(0.); tp.set_compressibility_oil(0.)
tp.set_ref_density_water(1000.); tp.set_ref_density_oil(1000.)
tp.set_ref_pressure_oil(0.); tp.set_ref_pressure_water(0.)
tp.set_residual_saturation_water(0.); tp.set_residual_saturation_oil(.2)
tp.set_corey_relperm(2., 2.)
# The reference's Newton loop stops on an ABSOLUTE criterion, norm(rhs)/len(rhs) < 1e-6
# (twophase.py:186, 759, 830): with a total rate of 10 spread over n^2 cells the very first residual
# is already below it and no pressure solve would ever run, so the threshold is lowered here.
tp.newton_threshold = 1.e-11
dt = 1.e4 * (n / 128.) ** 2   # explicit upwind: CFL < 1 in the fastest channels of the log-normal field
tp.set_time_step_size(dt)
for c in range(0, nc, n):     # cells of the i = 0 plane (cell index i + n j + n^2 k)
    tp.add_rate_well(10. / (n * n), 0., c, "w%d" % c)
t0 = time.perf_counter()
tp.initialize_system()
torch.cuda.synchronize()
t_init = time.perf_counter() - t0
times = []
for step in range(1, steps + 1):
    t0 = time.perf_counter()
    tp.update_pressure(step)
    if step == 1 or step % 10 == 0:
        tp.find_upwinding_direction()
    tp.update_saturation(step)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
tp._sync_host()
sw = tp.current_s_w
print("C4 n=%d cells=%d: mesh %.2f s, initialize_system %.2f s, %d IMPES steps: first %.3f s, others mean %.3f s"
      % (n, nc, t_mesh, t_init, steps, times[0], float(np.mean(times[1:])) if steps > 1 else times[0]))
print("   s_w in [%.3e, %.6f], mean %.6e (injected water fraction of pore volume %.6e)"
      % (sw.min(), sw.max(), sw.mean(), steps * dt * 10. / 1000. / (0.2 * nc)))
