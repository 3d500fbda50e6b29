"""Multiplier faces on the device path (the part round 1/2 refused): periodic boundaries
(Mesh.set_periodic_boundary, mesh.py:1476-1492) and Lagrange pointers between two sub-domains (mesh.py:1420-1467)
-> MFD.build_coupling_terms (mfd.py:806-848) -> solve.  Golden: the unmodified reference on the same set-up
(tests/coupling_cases.py:add_multipliers, tests/golden/gen_golden_r2.py:multiplier_system).  Run with pytest -m gpu;
the host-side numbering is covered on the CPU by tests/test_multipliers.py."""
import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu


def test_periodic_and_lagrange_multipliers_match_reference(golden):
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd
    import coupling_cases as cc

    g = golden("multipliers_433")
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(cc.NX, cc.NY, cc.NZ, cc.DIMS[0], cc.DIMS[1], cc.DIMS[2], cc.permeability())
    pairs, glue = cc.add_multipliers(mesh)
    assert np.array_equal(np.array(pairs), g["periodic_pairs"]) and np.array_equal(np.array(glue), g["glue"])
    f = mfd.MFD()
    cc.configure_multipliers(f, mesh)
    lhs = f.build_lhs()
    rhs = f.build_rhs()
    # the reference's numbering: the 21 multiplier faces are flux DOFs although the device system leaves them out
    assert f.flux_dof == int(g["flux_dof"][0]) and np.array_equal(f.face_to_flux[:, 0], g["face_to_flux"])
    cd, cr, ccol = f.build_coupling_terms()
    assert np.array_equal(cr, g["coupling_row"]) and np.array_equal(ccol, g["coupling_col"])
    assert np.array_equal(cd, g["coupling_data"])
    n = f.get_number_of_dof()
    ref = sparse.csr_matrix((g["csr_data"], g["csr_indices"], g["csr_indptr"]), shape=(n, n))
    assert lhs.shape == (n, n) and abs(lhs - ref).max() <= 1e-12 * abs(ref).max()
    assert np.array_equal(rhs, g["rhs"])
    # the COO list of M carries the reference's row / column numbers as well
    data, row, col = f.build_m()
    F = f.flux_dof
    m_ours = sparse.coo_matrix((data, (row, col)), shape=(n, n)).tocsr()[:F, :F]
    m_ref = (ref - sparse.coo_matrix((cd, (cr, ccol)), shape=(n, n)).tocsr())[:F, :F]
    assert abs(m_ours - m_ref).max() <= 1e-12 * abs(ref).max()
    sol = f.solve()
    err = np.linalg.norm(sol - g["solution"]) / np.linalg.norm(g["solution"])
    assert err <= 1e-8, (err, f.solver_info)
    assert np.linalg.norm(ref @ sol - rhs) <= 1e-9 * np.linalg.norm(rhs)


def test_singlephase_time_loop_with_pointer_couplings(golden):
    """SinglePhase.start_solving on a mesh with Dirichlet / forcing pointers, periodic boundaries and Lagrange pointers
    (singlephase.py:199-223 adds L, L^T to the step matrix): 5 steps against the unmodified reference."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.models.singlephase as singlephase
    import mimpy_b200.models.transport as transport
    import coupling_cases as cc

    g = golden("singlephase_coupled_433")
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(cc.NX, cc.NY, cc.NZ, cc.DIMS[0], cc.DIMS[1], cc.DIMS[2], cc.permeability())
    sp = singlephase.SinglePhase()
    cc.configure_singlephase(sp, mesh)
    sp.initialize_system()
    assert sp.mfd.flux_dof == int(g["flux_dof"][0])
    sp.set_initial_pressure(np.full(mesh.get_number_of_cells(), 1.5))
    for step in range(len(g["p"])):
        sp.start_solving()
        assert np.linalg.norm(sp.current_pressure - g["p"][step]) <= 1e-8 * np.linalg.norm(g["p"][step]), step
        assert np.linalg.norm(sp.current_velocity - g["v"][step]) <= 1e-8 * np.linalg.norm(g["v"][step]), step
    # Transport's loop does not carry the couplings: refused loudly, not ignored
    tr = transport.Transport()
    tr.set_mesh(mesh)
    with pytest.raises(NotImplementedError):
        tr.initialize_system()
