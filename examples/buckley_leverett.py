"""Two-phase Buckley-Leverett displacement: the problem of the reference's examples/twophase/buckleyleverett (300 x 1 x 1
cells, 80 m, Corey(2, 2), mu_o = 2 mu_w, S_or = 0.2, one water injector) through the same TwoPhase API -- only the two
import lines differ from a mimpy script -- with a shorter run; the IMPES loop (Newton pressure solve, upwind saturation
update) runs on the GPU through libmimpy_b200.so.  Needs a CUDA device.

At the end the computed water front is compared with the analytic Buckley-Leverett solution (Welge tangent) and the
injected water with the water in place.

    python examples/buckley_leverett.py [steps=200]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mimpy_b200.mesh.hexmesh as hexmesh        # import mimpy.mesh.hexmesh as hexmesh
import mimpy_b200.models.twophase as twophase    # import mimpy.models.twophase as twophase

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nx, length = 300, 80.

no_flow = lambda p: np.zeros(3)
cells = nx

mesh = hexmesh.HexMesh()
mesh.build_mesh(nx, 1, 1, length, 1., 1., lambda point, i, j, k: 1.e-15 * np.eye(3))

model = twophase.TwoPhase()
model.set_mesh(mesh)
for marker in (0, 2, 3, 4, 5):                       # injection end and the four sides: no flow
    model.apply_flux_boundary_from_function(marker, no_flow)
model.apply_pressure_boundary_from_function(1, lambda p: 0.)   # outlet

model.set_initial_p_o(np.full(cells, 100.))
model.set_initial_s_w(np.zeros(cells))
model.set_porosities(np.full(cells, .2))
for setter, value in ((model.set_viscosity_water, 8.90e-4), (model.set_viscosity_oil, 2. * 8.90e-4),
                      (model.set_compressibility_water, 0.), (model.set_compressibility_oil, 0.),
                      (model.set_ref_density_water, 1000.), (model.set_ref_density_oil, 1000.),
                      (model.set_ref_pressure_oil, 0.), (model.set_ref_pressure_water, 0.),
                      (model.set_residual_saturation_water, 0.), (model.set_residual_saturation_oil, .2)):
    setter(value)
model.set_corey_relperm(2., 2.)

delta_t, rate = 200000., 3. / (60. * 60. * 24)
model.set_time_step_size(delta_t)
model.add_rate_well(rate, 0., 0, "WELL1")            # water injector in the first cell
model.set_output_frequency(10 ** 9)                  # no VTK files from an example
model.set_number_of_time_steps(steps)
model.initialize_system()
model.start_solving()

s_w = np.asarray(model.current_s_w)
x = (np.arange(nx) + .5) * length / nx

# analytic solution: f(S) with Corey(2, 2), S_or = 0.2, mu_o = 2 mu_w; the front saturation is Welge's tangent point
s = np.linspace(1e-6, .8, 200001)
se = s / .8
f = se ** 2 / (se ** 2 + .5 * (1. - se) ** 2)
k_front = int(np.argmax(f / s))                  # f(S_f) / S_f = f'(S_f) maximises the chord slope
s_front, speed = s[k_front], f[k_front] / s[k_front]
darcy_over_phi = rate / 1000. / .2               # injected volume per time, area and porosity
x_front = darcy_over_phi * steps * delta_t * speed
cells_behind = int((s_w > .5 * s_front).sum())
x_front_numeric = cells_behind * length / nx
water_in_place = float((s_w * .2 * length / nx).sum())
water_injected = rate / 1000. * steps * delta_t
print("%d IMPES steps on %d cells: front saturation %.3f (analytic %.3f), front at %.2f m (analytic %.2f m)"
      % (steps, nx, float(s_w[max(cells_behind - 3, 0)]), s_front, x_front_numeric, x_front))
print("water in place %.6f m^3, injected %.6f m^3" % (water_in_place, water_injected))
assert abs(water_in_place - water_injected) <= 1e-9 * water_injected, "water is not conserved"
assert abs(x_front_numeric - x_front) <= 0.05 * x_front + 2. * length / nx, "front position off the analytic one"
