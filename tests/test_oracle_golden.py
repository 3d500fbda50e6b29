"""Pin the CPU oracle (oracle/mimpy_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/gen_golden.py) and against analytic known answers.
CPU only."""
import numpy as np
import pytest
from scipy import sparse

from oracle import mimpy_oracle as orc

zero_flux = lambda p: np.array([0., 0., 0.])
lin = lambda p: 1. + 2. * p[0] - p[1] + .5 * p[2]


def blocks_of(o, flat):
    out, pos = [], 0
    for c in range(o.nc):
        n = int(o.cell_ptr[c + 1] - o.cell_ptr[c])
        out.append(flat[pos:pos + n * n].reshape(n, n))
        pos += n * n
    assert pos == len(flat)
    return out


def assert_blocks_close(ours, ref, tol=1e-12):
    for a, b in zip(ours, ref):
        assert np.linalg.norm(a - b) <= tol * np.linalg.norm(b)


def assert_csr_matches(lhs_ours, g, prefix, tol=1e-12):
    """ours (oracle, value-filtered like the reference) vs golden CSR: identical pattern after a
    1e-12 relative threshold, values to tol."""
    ours = lhs_ours.tocsr()
    ours.sort_indices()
    shape = ours.shape
    ref = sparse.csr_matrix((g[prefix + "csr_data"], g[prefix + "csr_indices"], g[prefix + "csr_indptr"]), shape=shape)
    scale = abs(ref).max()
    diff = abs(ours - ref)
    assert diff.max() <= tol * scale
    # patterns agree exactly except for roundoff-level entries (SURVEY 7 hard part 1)
    a = ours.copy()
    a.data[abs(a.data) < 1e-12 * scale] = 0
    a.eliminate_zeros()
    b = ref.copy()
    b.data[abs(b.data) < 1e-12 * scale] = 0
    b.eliminate_zeros()
    assert np.array_equal(a.indptr, b.indptr)
    assert np.array_equal(a.indices, b.indices)


def test_hex_closed_form_topology_and_geometry(golden):
    g = golden("hex_c1_n5")
    ref = orc.omesh_from_dict(g)
    o = orc.hex_mesh(5, 5, 5, 1., 1., 1.)
    for k in ("face_points", "cell_faces", "cell_sigma", "face_to_cell"):
        assert np.array_equal(getattr(o, k), getattr(ref, k)), k
    assert o.boundary_faces == ref.boundary_faces
    for k in ("points", "face_areas", "face_normals", "face_centroids"):
        assert np.array_equal(getattr(o, k), getattr(ref, k)), k
    assert np.allclose(o.cell_volume, ref.cell_volume, rtol=1e-14, atol=0)
    assert np.allclose(o.cell_centroid, ref.cell_centroid, rtol=1e-13, atol=1e-16)


def test_hex_closed_form_rectangular_and_distorted(golden):
    g = golden("hex_aniso_432")
    ref = orc.omesh_from_dict(g)
    o = orc.hex_mesh(4, 3, 2, 2., 1.5, 1.)
    for k in ("face_points", "cell_faces", "cell_sigma", "face_to_cell"):
        assert np.array_equal(getattr(o, k), getattr(ref, k)), k
    assert o.boundary_faces == ref.boundary_faces
    g = golden("hex_distorted_333")
    ref = orc.omesh_from_dict(g)

    def mod(p, i, j, k):
        return p + 0.06 * np.array([np.sin(5. * p[1] + p[2]), np.cos(4. * p[0] * p[2]), np.sin(3. * p[0] + 2. * p[1])])

    o = orc.hex_mesh(3, 3, 3, 1., 1., 1., modification_function=mod)
    assert np.array_equal(o.points, ref.points)
    assert np.allclose(o.face_areas, ref.face_areas, rtol=1e-15, atol=0)
    assert np.allclose(o.face_normals, ref.face_normals, rtol=1e-15, atol=1e-17)
    assert np.allclose(o.cell_volume, ref.cell_volume, rtol=1e-13, atol=0)
    assert np.allclose(o.cell_centroid, ref.cell_centroid, rtol=1e-12, atol=1e-15)


def test_mirtich_on_cut_cells(golden):
    g = golden("cut_checker_4")
    ref = orc.omesh_from_dict(g)
    # the reference's face_to_cell loses the old cell on interior cut faces (mesh.py:2405-2410)
    assert (ref.face_to_cell != orc.rebuild_face_to_cell(ref)).any()
    ref.face_to_cell = orc.rebuild_face_to_cell(ref)
    vol, cen = orc.cell_volumes_centroids(ref)
    assert np.allclose(vol, ref.cell_volume, rtol=1e-12, atol=0)
    assert np.allclose(cen, ref.cell_centroid, rtol=1e-11, atol=1e-14)
    for c in range(0, ref.nc, 5):
        v, x = orc.volume_centroid_from_faces(ref, c)
        assert abs(v - ref.cell_volume[c]) <= 1e-13 * ref.cell_volume[c]
        assert abs(x - ref.cell_centroid[c]).max() < 1e-13
    assert abs(vol.sum() - 1.) < 1e-13
    counts = np.diff(ref.cell_ptr)
    assert counts.min() >= 7 and counts.max() <= 12


@pytest.mark.parametrize("method", range(8))
def test_m_e_all_methods(golden, method):
    g = golden("hex_aniso_432")
    o = orc.omesh_from_dict(g)
    ref = blocks_of(o, g["m%d_m_e_flat" % method])
    ours = orc.build_m_e_all(o, method)
    assert_blocks_close(ours, ref)


@pytest.mark.parametrize("name,methods", [("hex_distorted_333", (0, 1, 6)), ("cut_checker_4", (0, 1))])
def test_m_e_other_meshes(golden, name, methods):
    g = golden(name)
    o = orc.omesh_from_dict(g)
    for method in methods:
        ref = blocks_of(o, g["m%d_m_e_flat" % method])
        assert_blocks_close(orc.build_m_e_all(o, method), ref)


def test_m_e_batched_equals_per_cell(golden):
    g = golden("hex_distorted_333")
    o = orc.omesh_from_dict(g)
    for method in (0, 1, 6):
        b = orc.build_m_e_batched(o, method)
        for c in range(o.nc):
            a = orc.build_m_e(o, c, method)
            assert np.linalg.norm(a - b[c]) <= 1e-13 * np.linalg.norm(a)


def test_projector_identity(golden):
    """C C^T == I - N (N^T N)^-1 N^T  (SURVEY 8c v) -- what the CUDA kernel uses instead of the SVD."""
    for name in ("hex_aniso_432", "cut_checker_4"):
        o = orc.omesh_from_dict(golden(name))
        for c in range(o.nc):
            n_e = orc.build_n_e(o, c)
            c_e = orc.build_c_e(n_e)
            p = np.eye(len(n_e)) - n_e @ np.linalg.inv(n_e.T @ n_e) @ n_e.T
            assert abs(c_e @ c_e.T - p).max() < 1e-12


def _system(o, method, bc, forcing):
    f = orc.OracleMFD(o, method)
    for kind, marker, fn in bc:
        if kind == "d":
            f.apply_dirichlet_from_function(marker, fn)
        else:
            f.apply_neumann_from_function(marker, fn)
    if forcing is not None:
        f.apply_forcing_from_function(forcing)
    f.build_lhs()
    f.build_rhs()
    return f


def test_config1_assembly_rhs_solution(golden):
    g = golden("hex_c1_n5")
    o = orc.omesh_from_dict(g)
    u = lambda p: p[0] ** 2 + p[1] ** 2 + p[2] ** 2
    f = _system(o, 0, [("d", b, u) for b in range(6)], lambda x: -6.)
    assert np.array_equal(f.face_to_flux[:, 0], g["face_to_flux"])
    assert f.flux_dof == int(g["flux_dof"])
    assert_csr_matches(f.lhs, g, "")
    assert np.array_equal(f.rhs, g["rhs"])
    sol = f.solve()
    assert np.linalg.norm(sol - g["solution"]) <= 1e-10 * np.linalg.norm(g["solution"])
    # analytic: max|p - p_exact| = h^2/4 (SURVEY 8c ii)
    ex = np.array([u(c) for c in o.cell_centroid])
    assert abs(abs(f.get_pressure_solution() - ex).max() - (1. / 5) ** 2 / 4) < 1e-12


@pytest.mark.parametrize("method", [0, 1, 2, 3, 6])
def test_aniso_mixed_bc_system(golden, method):
    g = golden("hex_aniso_432")
    o = orc.omesh_from_dict(g)
    bc = [("d", 0, lambda p: 1. + 0.3 * p[1]), ("d", 1, lambda p: 0.2 * p[2])] + [("n", b, zero_flux) for b in (2, 3, 4, 5)]
    f = _system(o, method, bc, lambda x: np.sin(3. * x[0]) + x[1] * x[2])
    pre = "m%d_" % method
    assert np.array_equal(f.face_to_flux[:, 0], g[pre + "face_to_flux"])
    assert_csr_matches(f.lhs, g, pre)
    assert np.allclose(f.rhs, g[pre + "rhs"], rtol=1e-15, atol=0)
    sol = f.solve()
    assert np.linalg.norm(sol - g[pre + "solution"]) <= 1e-9 * np.linalg.norm(g[pre + "solution"])


def test_cut_mesh_patch_test_and_neumann(golden):
    g = golden("cut_checker_4")
    o = orc.omesh_from_dict(g)
    f = _system(o, 1, [("d", b, lin) for b in range(6)], None)
    assert_csr_matches(f.lhs, g, "m1_")
    sol = f.solve()
    assert np.linalg.norm(sol - g["m1_solution"]) <= 1e-9 * np.linalg.norm(g["m1_solution"])
    # linear patch test (SURVEY 8c i): exact for a UNIFORM SPD tensor on any polyhedral mesh
    o2 = orc.omesh_from_dict(g)
    kk = np.array([[2., .5, .3], [.5, 1.5, -.2], [.3, -.2, 1.]])
    o2.cell_k = np.tile(kk.ravel(), (o2.nc, 1))
    for method in (0, 1):
        f2 = _system(o2, method, [("d", b, lin) for b in range(6)], None)
        f2.solve()
        ex = np.array([lin(c) for c in o2.cell_centroid])
        assert abs(f2.get_pressure_solution() - ex).max() < 1e-11
    bc = [("d", 0, lambda p: 1.), ("d", 1, lambda p: 0.)] + [("n", b, zero_flux) for b in (2, 3, 4, 5)]
    f = _system(o, 1, bc, None)
    assert np.array_equal(f.face_to_flux[:, 0], g["m1n_face_to_flux"])
    assert_csr_matches(f.lhs, g, "m1n_")
    assert np.array_equal(f.rhs, g["m1n_rhs"])


def test_flow1d_known_answer(golden):
    g = golden("flow1d_322")
    o = orc.omesh_from_dict(g)
    bc = [("d", 0, lambda p: 1.), ("d", 1, lambda p: 0.)] + [("n", b, zero_flux) for b in (2, 3, 4, 5)]
    f = _system(o, 1, bc, None)
    assert f.flux_dof == 28
    p = f.solve()[f.flux_dof:]
    assert np.allclose(p.reshape(2, 2, 3), np.array([5. / 6, .5, 1. / 6])[None, None, :], atol=1e-13)
    assert np.linalg.norm(f.solution - g["solution"]) <= 1e-11 * np.linalg.norm(g["solution"])


def _twophase(o, variant, dt):
    tp = orc.OracleTwoPhase(o)
    nc = o.nc
    if variant == "wells":
        for b in (0, 2, 3, 4, 5):
            tp.mfd.apply_neumann_from_function(b, zero_flux)
        tp.mfd.apply_dirichlet_from_function(1, lambda p: 0.)
    else:
        for b in (2, 3, 4, 5):
            tp.mfd.apply_neumann_from_function(b, zero_flux)
        tp.mfd.apply_dirichlet_from_function(0, lambda p: 2.e5)
        tp.mfd.apply_dirichlet_from_function(1, lambda p: 0.)
        tp.saturation_boundaries[0] = lambda p: 1.
    tp.current_p_o = np.zeros(nc)
    tp.current_s_w = np.zeros(nc)
    tp.porosities = np.array([.2] * nc)
    tp.viscosity_water, tp.viscosity_oil = 8.9e-4, 1.78e-3
    tp.compressibility_water = tp.compressibility_oil = 0.
    tp.ref_density_water = tp.ref_density_oil = 1000.
    tp.residual_saturation_water, tp.residual_saturation_oil = 0., .2
    tp.set_corey_relperm(2., 2.)
    tp.delta_t = dt
    if variant == "wells":
        for c in range(nc):
            if c % 6 == 0:
                tp.add_rate_well(10. / 9, 0., c)
    tp.initialize_system()
    return tp


@pytest.mark.parametrize("variant", ["wells", "inflow"])
def test_twophase_history(golden, variant):
    g = golden("twophase_633_" + variant)
    o = orc.omesh_from_dict(g)
    tp = _twophase(o, variant, float(g["delta_t"]))
    assert np.array_equal(tp.mfd.m_e_locations, g["m_e_locations"])
    assert np.allclose(tp.mfd.m_data_for_update, g["m_data_for_update"], rtol=1e-11, atol=0)
    assert tp.c_start == int(g["c_start"]) and tp.c_end == int(g["c_end"])
    nup = []
    for step in range(1, int(g["steps"]) + 1):
        tp.update_pressure()
        if step == 1 or step % 10 == 0:
            tp.find_upwinding_direction()
            nup.append(len(tp.upwinded_face_cell))
        tp.update_saturation()
        assert abs(tp.current_s_w - g["s_w"][step - 1]).max() < 1e-10, step
        pn = np.linalg.norm(g["p_o"][step - 1])
        assert np.linalg.norm(tp.current_p_o - g["p_o"][step - 1]) <= 1e-8 * pn
        un = np.linalg.norm(g["u_t"][step - 1])
        assert np.linalg.norm(tp.current_u_t - g["u_t"][step - 1]) <= 1e-8 * un
    assert nup == list(g["n_upwind_pairs"])


def test_singlephase_history(golden):
    g = golden("singlephase_543")
    o = orc.omesh_from_dict(g)
    sp = orc.OracleSinglePhase(o)
    for b in (2, 3, 4, 5):
        sp.mfd.apply_neumann_from_function(b, zero_flux)
    sp.mfd.apply_dirichlet_from_function(0, lambda p: 2.)
    sp.mfd.apply_dirichlet_from_function(1, lambda p: 1.)
    sp.compressibility, sp.ref_density, sp.ref_pressure, sp.viscosity = 1.e-2, 1.3, 1., 0.7
    sp.porosities = g["porosities"]
    sp.delta_t = 0.5
    sp.add_point_rate_well(0.4, 27)
    sp.initialize_system()
    sp.current_pressure = np.full(o.nc, 1.5)
    for step in range(len(g["p"])):
        sp.step()
        assert np.linalg.norm(sp.current_pressure - g["p"][step]) <= 1e-10 * np.linalg.norm(g["p"][step])
        assert np.linalg.norm(sp.current_velocity - g["v"][step]) <= 1e-9 * np.linalg.norm(g["v"][step])


def _setup_transport(t, g, dirichlet, concentration_boundary):
    """The configuration of tests/golden/gen_golden.py section 9, applied through `dirichlet(marker, fn)`."""
    dirichlet(0, lambda p: 2.)
    dirichlet(1, lambda p: 1.)
    for b in (2, 3, 4, 5):
        dirichlet(b, lambda p: 2. - p[0] / 2.5)
    concentration_boundary(0, lambda p: 1. + 0.5 * p[1])
    t.compressibility, t.ref_density, t.ref_pressure, t.viscosity = 1.e-2, 1.3, 1., 0.7
    t.porosities = g["porosities"]
    t.delta_t = 0.004


def test_transport_history(golden):
    """OracleTransport (transport.py:261-361) against the unmodified reference: 6 steps of pressure,
    velocity and concentration, including the missing division by the cell volume (cells of 0.15)."""
    g = golden("transport_543")
    o = orc.omesh_from_dict(g)
    t = orc.OracleTransport(o)
    _setup_transport(t, g, t.mfd.apply_dirichlet_from_function,
                     lambda m, fn: t.concentration_boundaries.__setitem__(m, fn))
    t.add_point_rate_well(0.4, 27)
    t.initialize_system()
    t.current_pressure = np.full(o.nc, 1.5)
    t.current_concentration = np.array(g["c0"])
    for step in range(len(g["p"])):
        t.step()
        assert np.linalg.norm(t.current_pressure - g["p"][step]) <= 1e-10 * np.linalg.norm(g["p"][step])
        assert np.linalg.norm(t.current_velocity - g["v"][step]) <= 1e-9 * np.linalg.norm(g["v"][step])
        assert abs(t.current_concentration - g["c"][step]).max() <= 1e-12, step


# ---- pointer couplings (mfd.py:776-849, singlephase.py:199-223): the oracle's restatement against the reference -----
def _coupled_mesh(multipliers, pointers):
    import coupling_cases as cc
    from test_cpu_abi_and_host import mesh_from_omesh

    m = mesh_from_omesh(orc.hex_mesh(cc.NX, cc.NY, cc.NZ, *cc.DIMS, k_array=cc.permeability()))
    if pointers:
        cc.add_pointers(m)
    if multipliers:
        cc.add_multipliers(m)
    return orc.OMesh.from_mesh(m)


def _assert_coupled_system(f, g):
    cd, cr, ccol = f.build_coupling_terms()
    assert np.array_equal(cd, g["coupling_data"]) and np.array_equal(cr, g["coupling_row"])
    assert np.array_equal(ccol, g["coupling_col"])
    assert_csr_matches(f.lhs, g, "")
    assert np.array_equal(f.rhs, g["rhs"])
    sol = f.solve()
    assert np.linalg.norm(sol - g["solution"]) <= 1e-9 * np.linalg.norm(g["solution"])


def test_dirichlet_and_forcing_pointers(golden):
    g = golden("coupling_433")
    o = _coupled_mesh(multipliers=False, pointers=True)
    bc = [("d", 0, lambda p: 1. + p[1])] + [("n", b, zero_flux) for b in (2, 3, 4, 5)]
    _assert_coupled_system(_system(o, 1, bc, lambda x: 0.3 * x[2]), g)


def test_periodic_boundaries_and_lagrange_pointers(golden):
    g = golden("multipliers_433")
    o = _coupled_mesh(multipliers=True, pointers=False)
    bc = [("d", 0, lambda p: 1. + p[1]), ("d", 1, lambda p: 0.25 * p[2])] + [("n", b, zero_flux) for b in (2, 3)]
    f = _system(o, 1, bc, lambda x: 0.3 * x[2])
    assert f.flux_dof == int(g["flux_dof"][0]) and np.array_equal(f.face_to_flux[:, 0], g["face_to_flux"])
    _assert_coupled_system(f, g)


def test_singlephase_history_with_all_pointer_kinds(golden):
    g = golden("singlephase_coupled_433")
    o = _coupled_mesh(multipliers=True, pointers=True)
    sp = orc.OracleSinglePhase(o)
    for b in (2, 3):
        sp.mfd.apply_neumann_from_function(b, zero_flux)
    sp.mfd.apply_dirichlet_from_function(0, lambda p: 2.)
    sp.compressibility, sp.ref_density, sp.ref_pressure, sp.viscosity = 1.e-2, 1.3, 1., 0.7
    sp.porosities = np.linspace(.1, .3, o.nc)
    sp.delta_t = 0.5
    sp.add_point_rate_well(0.4, 17)
    sp.initialize_system()
    assert sp.mfd.flux_dof == int(g["flux_dof"][0])
    sp.current_pressure = np.full(o.nc, 1.5)
    for step in range(len(g["p"])):
        sp.step()
        assert np.linalg.norm(sp.current_pressure - g["p"][step]) <= 1e-9 * np.linalg.norm(g["p"][step]), step
        assert np.linalg.norm(sp.current_velocity - g["v"][step]) <= 1e-9 * np.linalg.norm(g["v"][step]), step
