"""SinglePhase.start_solving_newton (singlephase.py:300-356: 100 passes of the implicit step per time step) against
the unmodified reference on the set-up of fixture singlephase_543.  Run with pytest -m gpu."""
import numpy as np
import pytest

from oracle import mimpy_oracle as orc

pytestmark = pytest.mark.gpu


def test_start_solving_newton_matches_reference(golden):
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.models.singlephase as singlephase

    g, g0 = golden("singlephase_newton_543"), golden("singlephase_543")
    ref = orc.omesh_from_dict(g0)
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(5, 4, 3, 5., 4., 3., ref.cell_k.reshape(-1, 3, 3))
    sp = singlephase.SinglePhase()
    sp.set_mesh(mesh)
    nc = mesh.get_number_of_cells()
    for b in (2, 3, 4, 5):
        sp.apply_flux_boundary_from_function(b, lambda p: np.array([0., 0., 0.]))
    sp.apply_pressure_boundary_from_function(0, lambda p: 2.)
    sp.apply_pressure_boundary_from_function(1, lambda p: 1.)
    sp.set_compressibility(1.e-2)
    sp.set_ref_density(1.3)
    sp.set_ref_pressure(1.)
    sp.set_viscosity(0.7)
    sp.set_porosities(np.linspace(.1, .3, nc))
    sp.set_time_step_size(0.5)
    sp.set_number_of_time_steps(2)
    sp.add_point_rate_well(0.4, 27, "w")
    sp.initialize_system()
    sp.set_initial_pressure(np.full(nc, 1.5))
    calls = []
    sp.time_step_output = lambda t, k: calls.append((t, k, np.array(sp.current_pressure).copy()))
    sp.start_solving_newton()
    assert np.linalg.norm(sp.current_pressure - g["p"]) <= 1e-8 * np.linalg.norm(g["p"])
    assert np.linalg.norm(sp.current_velocity - g["v"]) <= 1e-8 * np.linalg.norm(g["v"])
    assert [c[0] for c in calls] == list(g["output_times"]) and [c[1] for c in calls] == list(g["output_steps"])
    for c, p_ref in zip(calls, g["output_p"]):
        assert np.linalg.norm(c[2] - p_ref) <= 1e-8 * np.linalg.norm(p_ref)
    # the settings of the plain loop are back
    assert sp.number_of_time_steps == 2 and sp.output_frequency == 1 and sp.time_step_output.__name__ == "<lambda>"
