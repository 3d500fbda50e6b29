"""Two-phase IMPES cases shared by the golden generator (run on the unmodified reference) and by the
GPU tests (run on mimpy_b200): the same set-up code drives either TwoPhase class."""
import numpy as np

zero_flux = lambda p: np.array([0., 0., 0.])


def krw_user(se):
    """A user water relative permeability that is NOT a Corey power law (passed to set_krw)."""
    se = np.clip(np.asarray(se, dtype=float), 0., 1.)
    return 0.6 * se ** 2.5 / (1. + 0.5 * (1. - se))


def kro_user(se):
    se = np.clip(np.asarray(se, dtype=float), 0., 1.)
    return (1. - se) ** 1.5 * (1. - 0.3 * se)


CASES = {
    # name: (nx, ny, nz, dims, dt, steps)
    "pwell": (6, 3, 3, (6., 3., 3.), 40., 12),     # rate injectors on x-, BHP producers on x+, user relperms, compressible
    "c4_1233": (12, 3, 3, (12., 3., 3.), 40., 12),  # BASELINE config 4 shape, Corey, rate wells + Dirichlet outflow
}


def permeability(name):
    nx, ny, nz = CASES[name][:3]
    rng = np.random.default_rng(128 + nx)
    kiso = 1.e-13 * np.exp(rng.standard_normal(nx * ny * nz))
    return kiso[:, None, None] * np.eye(3)[None]


def configure(tp, mesh, name, well_prefix):
    """Boundary conditions, fluid, wells of case `name` on a TwoPhase model whose mesh is set."""
    nx, ny, nz, dims, dt, steps = CASES[name]
    nc = mesh.get_number_of_cells()
    tp.set_initial_p_o(np.zeros(nc))
    tp.set_initial_s_w(np.zeros(nc))
    tp.set_porosities(np.array([.2] * nc))
    tp.set_viscosity_water(8.9e-4)
    tp.set_viscosity_oil(1.78e-3)
    tp.set_ref_density_water(1000.)
    tp.set_ref_density_oil(1000.)
    tp.set_ref_pressure_oil(0.)
    tp.set_ref_pressure_water(0.)
    tp.set_residual_saturation_water(0.)
    tp.set_residual_saturation_oil(.2)
    tp.set_time_step_size(dt)
    if name == "pwell":
        for b in range(6):
            tp.apply_flux_boundary_from_function(b, zero_flux)
        tp.set_compressibility_water(1.e-10)
        tp.set_compressibility_oil(2.e-10)
        tp.set_krw(krw_user)
        tp.set_kro(kro_user)
        for c in range(nc):
            i = mesh.cell_to_ijk[c][0]
            if i == 0:
                tp.add_rate_well(10. / (ny * nz), 0., c, well_prefix + "inj%d" % c)
            if i == nx - 1:
                tp.add_pressure_well(1.e4 * (1 + c % 3), 2.e-9, c, well_prefix + "prod%d" % c)
    else:
        for b in (0, 2, 3, 4, 5):
            tp.apply_flux_boundary_from_function(b, zero_flux)
        tp.apply_pressure_boundary_from_function(1, lambda p: 0.)
        tp.set_compressibility_water(0.)
        tp.set_compressibility_oil(0.)
        tp.set_corey_relperm(2., 2.)
        for c in range(nc):
            if mesh.cell_to_ijk[c][0] == 0:
                tp.add_rate_well(10. / (ny * nz), 0., c, well_prefix + "w%d" % c)
    return steps
