"""Drop-in API parity on the GPU: the scripts a mimpy user writes (HexMesh / MFD / TwoPhase)
produce what the reference produced (golden vectors) -- run with pytest -m gpu."""
import numpy as np
import pytest
from scipy import sparse

from oracle import mimpy_oracle as orc

pytestmark = pytest.mark.gpu

zero_flux = lambda p: np.array([0., 0., 0.])
lin = lambda p: 1. + 2. * p[0] - p[1] + .5 * p[2]


def spd_tensors(nc, seed, full=True):
    rng = np.random.default_rng(seed)
    out = np.zeros((nc, 3, 3))
    for c in range(nc):
        g = rng.standard_normal(3)
        if full:
            q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            out[c] = q @ np.diag(np.exp(g)) @ q.T
        else:
            out[c] = np.diag(np.exp(g))
    return out


def assert_mesh_equals_reference(mesh, ref, geometry_exact=True):
    o = orc.OMesh.from_mesh(mesh)
    for k in ("face_ptr", "face_points", "cell_ptr", "cell_faces", "cell_sigma"):
        assert np.array_equal(getattr(o, k), getattr(ref, k)), k
    assert {m: [(int(a), int(b)) for a, b in v] for m, v in o.boundary_faces.items()} == ref.boundary_faces
    cmp = np.array_equal if geometry_exact else (lambda a, b: np.allclose(a, b, rtol=1e-12, atol=1e-14))
    for k in ("points", "face_areas", "face_normals", "face_centroids"):
        assert cmp(getattr(o, k), getattr(ref, k)), k
    assert np.allclose(o.cell_volume, ref.cell_volume, rtol=1e-12, atol=0)
    assert np.allclose(o.cell_centroid, ref.cell_centroid, rtol=1e-11, atol=1e-14)
    assert np.array_equal(o.cell_k, ref.cell_k)
    return o


def test_example1_script_matches_reference(golden):
    """hexmesh_example_1.py (BASELINE config 1) with only the imports changed."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd

    g = golden("hex_c1_n5")
    res_mfd = mfd.MFD()
    res_mfd.set_compute_diagonality(True)
    res_mfd.set_m_e_construction_method(0)
    K = lambda p, i, j, k: np.eye(3)
    u = lambda p: p[0] ** 2 + p[1] ** 2 + p[2] ** 2
    res_mesh = hexmesh.HexMesh()
    res_mesh.build_mesh(5, 5, 5, 1., 1., 1., K, lambda p, i, j, k: p)
    ref = orc.omesh_from_dict(g)
    assert_mesh_equals_reference(res_mesh, ref)
    assert np.array_equal(res_mesh.face_to_cell[:ref.nf], ref.face_to_cell)
    res_mfd.set_mesh(res_mesh)
    for b in range(6):
        res_mfd.apply_dirichlet_from_function(b, lambda p: u(p))
    res_mfd.apply_forcing_from_function(lambda x: -6.)
    lhs = res_mfd.build_lhs()
    rhs = res_mfd.build_rhs()
    assert np.array_equal(res_mfd.face_to_flux[:, 0], g["face_to_flux"])
    assert res_mfd.flux_dof == int(g["flux_dof"])
    n = lhs.shape[0]
    gref = sparse.csr_matrix((g["csr_data"], g["csr_indices"], g["csr_indptr"]), shape=(n, n))
    assert abs(lhs - gref).max() <= 1e-12 * abs(gref).max()
    assert np.array_equal(rhs, g["rhs"])
    sol = res_mfd.solve()
    assert np.linalg.norm(sol - g["solution"]) <= 1e-8 * np.linalg.norm(g["solution"])
    p = res_mfd.get_pressure_solution()
    exact = np.array(res_mfd.get_analytical_pressure_solution(u))
    assert abs(abs(p - exact).max() - (1. / 5) ** 2 / 4) < 1e-9   # h^2/4 (SURVEY 8c ii)
    assert res_mfd.get_diagonality().max() < 1e-8                    # iso K, method 0: diagonal M_E
    assert abs(res_mfd.compute_l2_error_pressure(u) - np.sqrt(np.sum(ref.cell_volume * (p - exact) ** 2))) < 1e-14


def test_build_m_div_bottom_right_like_reference(golden):
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd

    g = golden("hex_aniso_432")
    ref = orc.omesh_from_dict(g)
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(4, 3, 2, 2., 1.5, 1., ref.cell_k.reshape(-1, 3, 3))
    assert_mesh_equals_reference(mesh, ref)
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    f.apply_dirichlet_from_function(0, lambda p: 1. + 0.3 * p[1])
    f.apply_dirichlet_from_function(1, lambda p: 0.2 * p[2])
    for b in (2, 3, 4, 5):
        f.apply_neumann_from_function(b, zero_flux)
    f.apply_forcing_from_function(lambda x: np.sin(3. * x[0]) + x[1] * x[2])
    data, row, col = f.build_m(save_update_info=True)
    (dd, dr, dc), (td, tr, tc) = f.build_div()
    cd, cr, cc = f.build_bottom_right()
    # full-tensor K + method 1: the reference keeps every structural entry -> identical COO stream
    nm = len(data)
    assert np.array_equal(row, g["m1_coo_row"][:nm]) and np.array_equal(col, g["m1_coo_col"][:nm])
    assert np.allclose(data, g["m1_coo_data"][:nm], rtol=0, atol=1e-12 * abs(g["m1_coo_data"]).max())
    nd = len(dd)
    assert np.array_equal(np.concatenate([dd, td, cd]), g["m1_coo_data"][nm:])
    assert np.array_equal(np.concatenate([dr, tr, cr]), g["m1_coo_row"][nm:])
    assert np.array_equal(np.concatenate([dc, tc, cc]), g["m1_coo_col"][nm:])
    assert f.m_e_locations[-1] == nm
    # per-cell block through the API
    blk = f.build_m_e(7)
    ref_blk = g["m1_m_e_flat"][7 * 36:8 * 36].reshape(6, 6)
    assert np.linalg.norm(blk - ref_blk) <= 1e-12 * np.linalg.norm(ref_blk)
    f.build_lhs()
    f.build_rhs()
    sol = f.solve()
    assert np.linalg.norm(sol - g["m1_solution"]) <= 1e-8 * np.linalg.norm(g["m1_solution"])


def test_reference_unit_tests_divide_cell_by_plane():
    """The reference's own tests (tests/test_divide_cell_by_plane.py:12-58) against our Mesh."""
    import mimpy_b200.mesh.hexmesh as hexmesh

    K = lambda x, i, j, k: np.eye(3)
    for normal in [np.array([1., 0., 0.]), np.array([0., 1., 0.]), np.array([0., 0., 1.]),
                   -np.array([1., 0., 0.]), -np.array([0., 1., 0.]), -np.array([0., 0., 1.])]:
        m = hexmesh.HexMesh()
        m.build_mesh(1, 1, 1, 1., 1., 1., K)
        m.divide_cell_by_plane(0, np.array([.5, .5, .5]), normal)
        assert m.get_number_of_cells() == 2
        assert abs(m.get_cell_volume(0) - .5) < 1.e-12
        assert abs(m.get_cell_volume(1) - .5) < 1.e-12
    for normal in [np.array([1.2, .8, 0.]), np.array([1.2, .2, 1.3]), -np.array([1.2, .8, 0.]), -np.array([1.2, .2, 1.3])]:
        m = hexmesh.HexMesh()
        m.build_mesh(1, 1, 1, 1., 1., 1., K)
        m.divide_cell_by_plane(0, np.array([.5, .5, .5]), normal)
        assert m.get_number_of_cells() == 2
        m = hexmesh.HexMesh()
        m.build_mesh(3, 3, 3, 1., 1., 1., K)
        c = m.find_cell_near_point(np.array([.5, .5, .5]))
        m.divide_cell_by_plane(c, np.array([.5, .5, .5]), normal)
        assert m.get_number_of_cells() == 28
        assert abs(m.cell_volume[:28].sum() - 1.) < 1e-13


def test_cut_mesh_generator_and_solve_match_reference(golden):
    """BASELINE config 5 shape: checkerboard of plane-cut cells built with OUR divide_cell_by_plane
    equals the reference's mesh, and the GPU solve equals the reference's spsolve."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd

    g = golden("cut_checker_4")
    ref = orc.omesh_from_dict(g)
    n = 4
    ktab = spd_tensors(n ** 3, 5, full=True)
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(n, n, n, 1., 1., 1., ktab)
    nrm = np.array([1.2, .2, 1.3])
    nrm /= np.linalg.norm(nrm)
    for c in range(n ** 3):
        i, j, k = mesh.cell_to_ijk[c]
        if (i + j + k) % 2 == 0:
            mesh.divide_cell_by_plane(c, np.array(mesh.get_cell_real_centroid(c)), nrm)
    assert_mesh_equals_reference(mesh, ref, geometry_exact=False)
    # all volumes again on the GPU (Mirtich kernel, general polyhedra)
    vol_cut = mesh.cell_volume[:ref.nc].copy()
    mesh.find_volume_centroid_all()
    assert np.allclose(mesh.cell_volume[:ref.nc], vol_cut, rtol=1e-12, atol=0)
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    for b in range(6):
        f.apply_dirichlet_from_function(b, lin)
    f.build_lhs()
    f.build_rhs()
    sol = f.solve()
    assert np.linalg.norm(sol - g["m1_solution"]) <= 1e-8 * np.linalg.norm(g["m1_solution"])


def test_cut_mesh_6cubed_c2_boundary_conditions(golden):
    """BASELINE config 5 at the parity size BASELINE.md names (6^3 checkerboard-cut polyhedral mesh) with the C2
    boundary conditions (Dirichlet x-/x+, no-flow elsewhere): mesh, saddle matrix and solution vs the reference."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd

    g = golden("cut_checker_6")
    ref = orc.omesh_from_dict(g)
    n = 6
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(n, n, n, 1., 1., 1., spd_tensors(n ** 3, 17, full=True))
    nrm = np.array([1.2, .2, 1.3])
    nrm /= np.linalg.norm(nrm)
    for c in range(n ** 3):
        i, j, k = c % n, (c // n) % n, c // (n * n)
        if (i + j + k) % 2 == 0:
            mesh.divide_cell_by_plane(c, np.array(mesh.get_cell_real_centroid(c)), nrm)
    assert_mesh_equals_reference(mesh, ref, geometry_exact=False)
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    f.apply_dirichlet_from_function(0, lambda p: 1.)
    f.apply_dirichlet_from_function(1, lambda p: 0.)
    for b in (2, 3, 4, 5):
        f.apply_neumann_from_function(b, zero_flux)
    lhs = f.build_lhs()
    rhs = f.build_rhs()
    assert np.array_equal(f.face_to_flux[:, 0], g["m1n_face_to_flux"])
    nd = f.get_number_of_dof()
    ref_a = sparse.csr_matrix((g["m1n_csr_data"], g["m1n_csr_indices"], g["m1n_csr_indptr"]), shape=(nd, nd))
    assert abs(lhs - ref_a).max() <= 1e-12 * abs(ref_a).max()
    assert np.array_equal(rhs, g["m1n_rhs"])
    sol = f.solve()
    assert np.linalg.norm(sol - g["m1n_solution"]) <= 1e-8 * np.linalg.norm(g["m1n_solution"])


@pytest.mark.parametrize("variant", ["wells", "inflow"])
def test_twophase_model_matches_reference_history(golden, variant):
    """BASELINE config 4 shape at 6x3x3: 12 IMPES steps; s_w within 1e-10 per step, p/u within 1e-8."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.models.twophase as twophase

    g = golden("twophase_633_" + variant)
    ref = orc.omesh_from_dict(g)
    nx, ny, nz = 6, 3, 3
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(nx, ny, nz, 6., 3., 3., ref.cell_k.reshape(-1, 3, 3))
    tp = twophase.TwoPhase()
    tp.set_mesh(mesh)
    nc = mesh.get_number_of_cells()
    if variant == "wells":
        for b in (0, 2, 3, 4, 5):
            tp.apply_flux_boundary_from_function(b, zero_flux)
        tp.apply_pressure_boundary_from_function(1, lambda p: 0.)
    else:
        for b in (2, 3, 4, 5):
            tp.apply_flux_boundary_from_function(b, zero_flux)
        tp.apply_pressure_boundary_from_function(0, lambda p: 2.e5)
        tp.apply_pressure_boundary_from_function(1, lambda p: 0.)
        tp.set_saturation_boundaries(0, lambda p: 1.)
    tp.set_initial_p_o(np.zeros(nc))
    tp.set_initial_s_w(np.zeros(nc))
    tp.set_porosities(np.array([.2] * nc))
    tp.set_viscosity_water(8.9e-4)
    tp.set_viscosity_oil(1.78e-3)
    tp.set_compressibility_water(0.)
    tp.set_compressibility_oil(0.)
    tp.set_ref_density_water(1000.)
    tp.set_ref_density_oil(1000.)
    tp.set_ref_pressure_oil(0.)
    tp.set_ref_pressure_water(0.)
    tp.set_residual_saturation_water(0.)
    tp.set_residual_saturation_oil(.2)
    tp.set_corey_relperm(2., 2.)
    tp.set_time_step_size(float(g["delta_t"]))
    if variant == "wells":
        for c in range(nc):
            if mesh.cell_to_ijk[c][0] == 0:
                tp.add_rate_well(10. / (ny * nz), 0., c, "w%d" % c)
    tp.initialize_system()
    # c_start counts M's COO triples: ours is structural, the reference's is value-filtered
    assert tp.c_start >= int(g["c_start"]) and tp.c_end - tp.c_start == nc
    for step in range(1, int(g["steps"]) + 1):
        tp.update_pressure(step)
        if step == 1 or step % 10 == 0:
            tp.find_upwinding_direction()
        tp.update_saturation(step)
        tp._sync_host()
        assert abs(tp.current_s_w - g["s_w"][step - 1]).max() < 1e-10, step
        assert np.linalg.norm(tp.current_p_o - g["p_o"][step - 1]) <= 1e-8 * np.linalg.norm(g["p_o"][step - 1]), step
        assert np.linalg.norm(tp.current_u_t - g["u_t"][step - 1]) <= 1e-8 * np.linalg.norm(g["u_t"][step - 1]), step


@pytest.mark.parametrize("name", ["pwell", "c4_1233"])
def test_twophase_round2_cases_match_reference_history(golden, name, tmp_path, monkeypatch):
    """Two more reference histories (tests/golden/gen_golden_r2.py): `pwell` -- pressure (BHP) wells whose
    saturation term uses the mobilities of the state BEFORE the update (twophase.py:965, 1053-1090), user
    relperm callables (set_krw / set_kro, twophase.py:292-300) and compressible fluids, production logs
    compared with the reference's <well>.prod files; `c4_1233` -- BASELINE config 4 shape at 12x3x3."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.models.twophase as twophase
    import twophase_cases as cases

    g = golden("twophase_" + name)
    nx, ny, nz, dims, dt, steps = cases.CASES[name]
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(nx, ny, nz, dims[0], dims[1], dims[2], cases.permeability(name))
    tp = twophase.TwoPhase()
    tp.set_mesh(mesh)
    tp.write_well_logs = True
    monkeypatch.chdir(tmp_path)
    assert cases.configure(tp, mesh, name, "t_") == int(g["steps"])
    tp.initialize_system()
    for step in range(1, steps + 1):
        tp.update_pressure(step)
        if step == 1 or step % 10 == 0:
            tp.find_upwinding_direction()
        tp.update_saturation(step)
        tp._sync_host()
        assert abs(tp.current_s_w - g["s_w"][step - 1]).max() < 1e-10, step
        assert np.linalg.norm(tp.current_p_o - g["p_o"][step - 1]) <= 1e-8 * np.linalg.norm(g["p_o"][step - 1]), step
        assert np.linalg.norm(tp.current_u_t - g["u_t"][step - 1]) <= 1e-8 * np.linalg.norm(g["u_t"][step - 1]), step
    if name == "pwell":
        for f in tp.production_files:
            f.close()
        for w, nme in enumerate(tp.pressure_wells_name):
            log = np.loadtxt(nme + ".prod").reshape(-1, 3)
            ref = g["prod_log"][w]
            assert log.shape == ref.shape and np.array_equal(log[:, 0], ref[:, 0])
            assert np.allclose(log[:, 1:], ref[:, 1:], rtol=1e-7, atol=1e-12 * abs(ref[:, 1:]).max())


def test_pointer_couplings_match_reference(golden):
    """Dirichlet / forcing pointers (mesh.py:1397-1525) -> MFD.build_coupling_terms
    (mfd.py:776-849).  Coupling COO bit-exact, full saddle matrix 1e-12, rhs exact, and the coupled solve
    (GPU flexible GMRES preconditioned by the hybridised uncoupled solve) within 1e-8 of the reference's spsolve."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd
    import coupling_cases as cc

    g = golden("coupling_433")
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(cc.NX, cc.NY, cc.NZ, cc.DIMS[0], cc.DIMS[1], cc.DIMS[2], cc.permeability())
    faces = cc.add_pointers(mesh)
    assert np.array_equal(faces, g["pointer_faces"])
    f = mfd.MFD()
    cc.configure(f, mesh)
    lhs = f.build_lhs()
    rhs = f.build_rhs()
    cd, cr, ccol = f.build_coupling_terms()
    assert np.array_equal(cr, g["coupling_row"]) and np.array_equal(ccol, g["coupling_col"])
    assert np.array_equal(cd, g["coupling_data"])
    n = f.get_number_of_dof()
    ref = sparse.csr_matrix((g["csr_data"], g["csr_indices"], g["csr_indptr"]), shape=(n, n))
    assert abs(lhs - ref).max() <= 1e-12 * abs(ref).max()
    assert np.array_equal(rhs, g["rhs"])
    sol = f.solve()
    err = np.linalg.norm(sol - g["solution"]) / np.linalg.norm(g["solution"])
    assert err <= 1e-8, (err, f.solver_info)
    assert np.linalg.norm(ref @ sol - rhs) <= 1e-9 * np.linalg.norm(rhs)


def test_add_gravity_matches_reference(golden):
    """MFD.add_gravity (mfd.py:1001-1025): rhs += M_{K=I} (rho g e_g.n) on the device vs the reference."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd

    g = golden("gravity_432")
    ref = orc.omesh_from_dict(g)
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(4, 3, 2, 2., 1.5, 1., ref.cell_k.reshape(-1, 3, 3))
    mesh.set_gravity_vector(np.array([0.1, -0.2, -1.]) / np.linalg.norm([0.1, -0.2, -1.]))
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    for b in range(6):
        f.apply_dirichlet_from_function(b, lambda p: 1. + p[0] - 2. * p[2])
    f.build_lhs()
    assert np.array_equal(f.build_rhs(), g["rhs_before"])
    with pytest.raises(AttributeError):
        f.add_gravity()                      # the reference fails the same way: no fluid density on the mesh
    mesh.set_fluid_density(850.)
    f.add_gravity()
    assert abs(f.rhs - g["rhs_after"]).max() <= 1e-12 * abs(g["rhs_after"]).max()
    sol = f.solve()
    assert np.linalg.norm(sol - g["solution"]) <= 1e-8 * np.linalg.norm(g["solution"])
    # with Neumann faces the reference's face-indexed update does not fit the flux numbering
    f2 = mfd.MFD()
    f2.set_m_e_construction_method(1)
    f2.set_mesh(mesh)
    f2.apply_neumann_from_function(5, zero_flux)
    for b in range(5):
        f2.apply_dirichlet_from_function(b, lambda p: 1.)
    f2.build_lhs()
    f2.build_rhs()
    with pytest.raises(ValueError):
        f2.add_gravity()


def test_voromesh_matches_reference(golden):
    """VoroMesh.build_mesh (voromesh.py:19-201) on a voro++-format file of 48 Voronoi cells with 6..16 faces: the mesh
    equals the reference's array by array (shifted centroids included), and the generic lane-group kernels
    (16 lanes per cell here) reproduce its M_E blocks, saddle matrix, right-hand side and solution."""
    import os
    import mimpy_b200.mesh.voromesh as voromesh
    import mimpy_b200.mfd.mfd as mfd

    g = golden("voro_48")
    ref = orc.omesh_from_dict(g)
    mesh = voromesh.VoroMesh()
    mesh.build_mesh(os.path.join(os.path.dirname(__file__), "golden", "voro_48.vol.gz"),
                    lambda p: np.diag([1. + p[0], 2. - p[1], 1. + .5 * p[2]]))
    assert_mesh_equals_reference(mesh, ref, geometry_exact=False)
    nf, nc = mesh.get_number_of_faces(), mesh.get_number_of_cells()
    assert mesh.is_using_face_shifted_centroid() and mesh.is_using_cell_shifted_centroid()
    assert np.allclose(mesh.face_shifted_centroids[:nf], g["face_shifted_centroids"], rtol=0, atol=1e-14)
    assert np.array_equal(mesh.cell_shifted_centroid[:nc], g["cell_shifted_centroids"])
    n_e = np.diff(ref.cell_ptr)
    assert n_e.min() == 6 and n_e.max() == 16
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    f.apply_dirichlet_from_function(0, lambda p: 1. + p[1])
    f.apply_dirichlet_from_function(1, lambda p: 0.)
    for b in (2, 3, 4, 5):
        f.apply_neumann_from_function(b, zero_flux)
    f.apply_forcing_from_function(lambda p: p[0] * p[2])
    blocks = np.concatenate([f.build_m_e(c).ravel() for c in range(nc)])
    assert abs(blocks - g["m1_m_e_flat"]).max() <= 1e-11 * abs(g["m1_m_e_flat"]).max()
    f.build_lhs()
    assert np.array_equal(f.face_to_flux[:, 0], g["m1_face_to_flux"]) and f.flux_dof == int(g["m1_flux_dof"])
    n = f.flux_dof + nc
    ours = f.lhs.tocsr()
    theirs = sparse.csr_matrix((g["m1_csr_data"], g["m1_csr_indices"], g["m1_csr_indptr"]), shape=(n, n))
    assert abs(ours - theirs).max() <= 1e-11 * abs(theirs).max()
    assert np.allclose(f.build_rhs(), g["m1_rhs"], rtol=1e-12, atol=1e-15)
    sol = f.solve()
    assert np.linalg.norm(sol - g["m1_solution"]) <= 1e-8 * np.linalg.norm(g["m1_solution"])


def test_mfd_helper_methods_match_reference(golden):
    """The small MFD helpers a script may call directly -- build_c_e / build_d_e / build_w_e (mfd.py:314-355),
    proj_vector (:601-616), the three parts of build_rhs (:1027-1041, 1209-1215), apply_forcing_from_grad
    (:1178-1192), compute_x_error_velocity (:1301-1319) -- against the unmodified reference."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd

    g = golden("mfd_helpers_432")
    ref = orc.omesh_from_dict(g)
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(4, 3, 2, 2., 1.5, 1., ref.cell_k.reshape(-1, 3, 3))
    grad = lambda p: np.array([2. * p[0] + .3, -1. + p[2], .5 * p[1]])
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    for b in range(6):
        f.apply_dirichlet_from_function(b, lambda p: 1. + p[0] - 2. * p[2] + p[0] * p[1])
    f.apply_forcing_from_function(lambda p: p[0] - p[1])
    assert np.allclose(f.cell_forcing_function, g["forcing_before"], rtol=1e-13, atol=0)
    f.apply_forcing_from_grad(grad)
    assert np.allclose(f.cell_forcing_function, g["forcing_after"], rtol=1e-12, atol=1e-15)
    for q, c in enumerate((0, 7, 23)):
        w = f.build_w_e(c)
        assert abs(w - g["w_e"][q]).max() <= 1e-12 * abs(g["w_e"][q]).max()
    # method 4 is M_E = W_E^-1 (mfd.py:420-421): the device blocks invert the host W_E
    f4 = mfd.MFD()
    f4.set_m_e_construction_method(4)
    f4.set_mesh(mesh)
    assert abs(f4.build_m_e(7) @ f.build_w_e(7) - np.eye(6)).max() <= 1e-10
    n_e = f.build_n_e(5)
    assert abs(n_e - g["n_e_5"]).max() <= 1e-13 * abs(g["n_e_5"]).max()
    c_e = f.build_c_e(n_e)
    assert c_e.shape == (6, 3) and abs(n_e.T @ c_e).max() <= 1e-12
    assert abs(c_e @ c_e.T - g["c_e_5_projector"]).max() <= 1e-12
    assert f.build_d_e(n_e).shape == (6, 3)
    assert np.allclose(np.array(f.proj_vector(grad)), g["proj"], rtol=1e-12, atol=1e-14)
    f.build_lhs()
    rhs = f.build_rhs().copy()
    assert np.allclose(rhs, g["rhs"], rtol=1e-12, atol=1e-15)
    f.rhs = np.zeros_like(rhs)   # the same vector from its three parts
    f.build_rhs_neumann()
    f.build_rhs_dirichlet()
    f.build_rhs_forcing()
    assert np.array_equal(f.rhs, rhs)
    sol = f.solve()
    assert np.linalg.norm(sol - g["solution"]) <= 1e-8 * np.linalg.norm(g["solution"])
    assert abs(f.compute_x_error_velocity(grad) - float(g["x_error"])) <= 1e-7 * float(g["x_error"])
    assert abs(f.compute_l2_error_velocity(grad) - float(g["l2_error_velocity"])) <= 1e-7 * float(g["l2_error_velocity"])
    per_cell = f.l2_error_velocity_per_cell(grad)
    assert per_cell.shape == (24,) and np.all(np.isfinite(per_cell)) and np.all(per_cell >= 0.)


def test_singlephase_model_matches_reference_history(golden):
    """SinglePhase (implicit compressible stepping, singlephase.py:244-299): 6 steps vs the reference."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.models.singlephase as singlephase

    g = golden("singlephase_543")
    ref = orc.omesh_from_dict(g)
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(5, 4, 3, 5., 4., 3., ref.cell_k.reshape(-1, 3, 3))
    sp = singlephase.SinglePhase()
    sp.set_mesh(mesh)
    nc = mesh.get_number_of_cells()
    for b in (2, 3, 4, 5):
        sp.apply_flux_boundary_from_function(b, zero_flux)
    sp.apply_pressure_boundary_from_function(0, lambda p: 2.)
    sp.apply_pressure_boundary_from_function(1, lambda p: 1.)
    sp.set_compressibility(1.e-2)
    sp.set_ref_density(1.3)
    sp.set_ref_pressure(1.)
    sp.set_viscosity(0.7)
    sp.set_porosities(np.linspace(.1, .3, nc))
    sp.set_time_step_size(0.5)
    sp.set_number_of_time_steps(1)
    sp.add_point_rate_well(0.4, 27, "w")
    sp.initialize_system()
    sp.set_initial_pressure(np.full(nc, 1.5))
    for step in range(len(g["p"])):
        sp.start_solving()
        assert np.linalg.norm(sp.current_pressure - g["p"][step]) <= 1e-8 * np.linalg.norm(g["p"][step]), step
        assert np.linalg.norm(sp.current_velocity - g["v"][step]) <= 1e-8 * np.linalg.norm(g["v"][step]), step


def test_transport_model_matches_reference_history(golden):
    """Transport (transport.py:363-391): SinglePhase pressure step + explicit upwind concentration;
    6 steps against the unmodified reference (c within 1e-10 per step, p/v within 1e-8)."""
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd
    import mimpy_b200.models.transport as transport

    g = golden("transport_543")
    ref = orc.omesh_from_dict(g)
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(5, 4, 3, 2.5, 2., 1.8, ref.cell_k.reshape(-1, 3, 3))
    tr = transport.Transport()
    f = mfd.MFD()
    f.set_m_e_construction_method(1)
    tr.set_mesh_mfd(mesh, f)
    nc = mesh.get_number_of_cells()
    tr.apply_pressure_boundary_from_function(0, lambda p: 2.)
    tr.apply_pressure_boundary_from_function(1, lambda p: 1.)
    for b in (2, 3, 4, 5):
        tr.apply_pressure_boundary_from_function(b, lambda p: 2. - p[0] / 2.5)
    tr.set_concentration_boundaries(0, lambda p: 1. + 0.5 * p[1])
    tr.set_compressibility(1.e-2)
    tr.set_ref_density(1.3)
    tr.set_ref_pressure(1.)
    tr.set_viscosity(0.7)
    tr.set_porosities(np.linspace(.1, .3, nc))
    tr.set_time_step_size(0.004)
    tr.set_number_of_time_steps(1)
    tr.add_point_rate_well(0.4, 27, "w")
    tr.set_initial_pressure(np.full(nc, 1.5))
    tr.set_initial_concentration(np.array(g["c0"]))
    tr.initialize_system()
    for step in range(len(g["p"])):
        tr.start_solving()
        assert np.linalg.norm(tr.current_pressure - g["p"][step]) <= 1e-8 * np.linalg.norm(g["p"][step]), step
        assert np.linalg.norm(tr.current_velocity - g["v"][step]) <= 1e-8 * np.linalg.norm(g["v"][step]), step
        assert abs(tr.current_concentration - g["c"][step]).max() <= 1e-10, step
    assert len(tr.get_upwinded_face_cell()) > 0
    # Neumann boundaries are rejected (the reference would index out of range, transport.py:205-209)
    tr2 = transport.Transport()
    tr2.set_mesh(mesh)
    tr2.apply_flux_boundary_from_function(2, zero_flux)
    with pytest.raises(ValueError):
        tr2.initialize_system()
