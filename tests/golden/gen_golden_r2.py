"""Round-2 golden fixtures from the UNMODIFIED reference (see gen_golden.py for the conventions):

  twophase_pwell    6x3x3, 12 IMPES steps with pressure (BHP) wells, user relperm callables
                    (TwoPhase.set_krw / set_kro, twophase.py:292-300), compressible fluids
  twophase_c4_1233  BASELINE config 4 shape at 12x3x3, Corey relperms, rate wells

Usage (build container only):  python tests/golden/gen_golden_r2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import ref_shim  # noqa: E402
from oracle import mimpy_oracle as orc  # noqa: E402
import gen_golden as g1  # noqa: E402
import twophase_cases as cases  # noqa: E402


def twophase_history(ref, name):
    nx, ny, nz, dims, dt, steps = cases.CASES[name]
    ktab = cases.permeability(name)
    mesh = g1.ref_hex(ref, nx, ny, nz, dims, ktab)
    tp = ref.twophase.TwoPhase()
    tp.set_mesh(mesh)
    cwd = os.getcwd()
    os.chdir("/tmp")          # the reference opens <well>.inj / <well>.prod in the cwd
    try:
        cases.configure(tp, mesh, name, "golden_")
        g1.quiet(tp.initialize_system)
        sw_hist, p_hist, u_hist = [], [], []
        for step in range(1, steps + 1):
            g1.quiet(tp.update_pressure, step)
            if step == 1 or step % 10 == 0:
                tp.find_upwinding_direction()
            for _ in range(tp.saturation_time_steps):
                tp.update_saturation(step)
            sw_hist.append(np.array(tp.current_s_w))
            p_hist.append(np.array(tp.current_p_o))
            u_hist.append(np.array(tp.current_u_t))
        for f in list(tp.production_files) + list(tp.pressure_files):
            f.flush()
        prod = {}
        for w, nme in enumerate(tp.pressure_wells_name):
            prod[nme] = np.loadtxt(nme + ".prod").reshape(-1, 3)
    finally:
        os.chdir(cwd)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["delta_t"] = np.float64(dt)
    d["steps"] = np.int64(steps)
    d["s_w"], d["p_o"], d["u_t"] = np.array(sw_hist), np.array(p_hist), np.array(u_hist)
    if prod:
        d["prod_log"] = np.array([prod[n] for n in tp.pressure_wells_name])
    return d


class _ScalarIndexed(np.ndarray):
    """face_to_flux is an (nf, 1) array and mfd.py:796-846 appends face_to_flux[f] -- a 1-element array -- to an
    array.array('i'): accepted by the NumPy the reference was written for, a TypeError since NumPy 2.  Runtime
    compatibility shim (like oracle/ref_shim.py): integer indexing yields the scalar; no source is edited."""

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return int(np.ndarray.__getitem__(self, (idx, 0)))
        return np.ndarray.__getitem__(self, idx)


def _patch_coupling(ref):
    orig = ref.mfd.MFD.build_coupling_terms
    if getattr(orig, "_shimmed", False):
        return

    def build_coupling_terms(self):
        keep = self.face_to_flux
        self.face_to_flux = np.asarray(keep).view(_ScalarIndexed)
        try:
            return orig(self)
        finally:
            self.face_to_flux = keep

    build_coupling_terms._shimmed = True
    ref.mfd.MFD.build_coupling_terms = build_coupling_terms


def coupling_system(ref):
    """MFD with pointer couplings (mfd.py:776-849): the coupling COO, the full saddle matrix, rhs, spsolve."""
    import coupling_cases as cc

    _patch_coupling(ref)

    mesh = g1.ref_hex(ref, cc.NX, cc.NY, cc.NZ, cc.DIMS, cc.permeability())
    faces = cc.add_pointers(mesh)
    f = ref.mfd.MFD()
    cc.configure(f, mesh)
    lhs = g1.quiet(f.build_lhs)
    rhs = f.build_rhs()
    cpl = f.build_coupling_terms()
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["pointer_faces"] = np.array(faces)
    d["coupling_data"], d["coupling_row"], d["coupling_col"] = np.array(cpl[0]), np.array(cpl[1]), np.array(cpl[2])
    csr = lhs.tocsr()
    csr.sort_indices()
    d["csr_indptr"], d["csr_indices"], d["csr_data"] = csr.indptr.copy(), csr.indices.copy(), csr.data.copy()
    d["rhs"] = rhs.copy()
    d["solution"] = np.array(f.solve())
    return d


def gravity_rhs(ref):
    """MFD.add_gravity (mfd.py:1001-1025) on a full-tensor 4x3x2 box, Dirichlet everywhere (the reference's
    indexing needs flux_dof == number of faces).  Mesh.get_fluid_density does not exist in the reference
    (mfd.py:1008 would raise AttributeError): the getter is supplied on the instance, nothing else changes."""
    ktab = g1.spd_tensors(24, 3, full=True)
    mesh = g1.ref_hex(ref, 4, 3, 2, (2., 1.5, 1.), ktab)
    mesh.set_gravity_vector(np.array([0.1, -0.2, -1.]) / np.linalg.norm([0.1, -0.2, -1.]))
    mesh.get_fluid_density = lambda: 850.
    f = ref.mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    for b in range(6):
        f.apply_dirichlet_from_function(b, lambda p: 1. + p[0] - 2. * p[2])
    g1.quiet(f.build_lhs)
    rhs0 = f.build_rhs().copy()
    g1.quiet(f.add_gravity)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["rhs_before"], d["rhs_after"] = rhs0, f.rhs.copy()
    d["solution"] = np.array(f.solve())
    return d


def cut_checker(ref, n):
    """BASELINE config 5 shape at n^3 (6^3): every other cell cut by a plane through its centroid,
    full-tensor K, the C2 boundary conditions (Dirichlet x-/x+, no-flow elsewhere), method 1."""
    ktab = g1.spd_tensors(n ** 3, 17, full=True)
    mesh = g1.ref_hex(ref, n, n, n, (1., 1., 1.), ktab)
    # the reference's variable_array reallocates under live views while cells are divided: pre-size (as gen_golden.py does)
    nrm = np.array([1.2, .2, 1.3])
    nrm /= np.linalg.norm(nrm)
    for va in (mesh.faces, mesh.cells, mesh.cell_normal_orientation):
        va.set_data_capacity(40 * len(va.data) + 1000)
        va.set_pointer_capacity(8 * len(va.pointers) + 1000)
    for c in range(n ** 3):
        i, j, k = c % n, (c // n) % n, c // (n * n)
        if (i + j + k) % 2 == 0:
            mesh.divide_cell_by_plane(c, np.array(mesh.get_cell_real_centroid(c)), nrm)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    bc = [("d", 0, lambda p: 1.), ("d", 1, lambda p: 0.)] + [("n", b, cases.zero_flux) for b in (2, 3, 4, 5)]
    r = g1.dump_system(ref, mesh, 1, bc, None, with_blocks=False)
    for k_, v in r.items():
        d["m1n_%s" % k_] = v
    return d


def mfd_helpers(ref):
    """The small MFD helpers a script may call directly (mfd.py:314-355 build_c_e / build_d_e / build_w_e, 601-616
    proj_vector, 1027-1041 + 1209-1215 the three parts of build_rhs, 1178-1192 apply_forcing_from_grad, 1301-1319
    compute_x_error_velocity) on a full-tensor 4x3x2 box, Dirichlet everywhere."""
    ktab = g1.spd_tensors(24, 9, full=True)
    mesh = g1.ref_hex(ref, 4, 3, 2, (2., 1.5, 1.), ktab)
    grad = lambda p: np.array([2. * p[0] + .3, -1. + p[2], .5 * p[1]])
    f = ref.mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    for b in range(6):
        f.apply_dirichlet_from_function(b, lambda p: 1. + p[0] - 2. * p[2] + p[0] * p[1])
    f.apply_forcing_from_function(lambda p: p[0] - p[1])
    forcing0 = np.array(f.cell_forcing_function).copy()
    f.get_face_area = mesh.get_face_area   # mfd.py:1190 calls a getter MFD never defines (it is the mesh's)
    f.apply_forcing_from_grad(grad)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["forcing_before"], d["forcing_after"] = forcing0, np.array(f.cell_forcing_function).copy()
    d["w_e"] = np.array([f.build_w_e(c) for c in (0, 7, 23)])
    n_e = f.build_n_e(5)
    c_e = f.build_c_e(n_e)
    d["n_e_5"], d["c_e_5_projector"] = n_e, c_e.dot(c_e.T)
    d["proj"] = np.array(f.proj_vector(grad))
    g1.quiet(f.build_lhs)
    d["rhs"] = f.build_rhs().copy()
    f.solve()
    d["solution"] = np.array(f.solution)
    d["x_error"] = np.float64(f.compute_x_error_velocity(grad))
    d["l2_error_velocity"] = np.float64(f.compute_l2_error_velocity(grad))
    return d


VORO_BOX = [(0., 2.), (0., 1.5), (0., 1.)]


def voro_k(p):
    """Permeability of the Voronoi fixture as a function of the cell centroid (shared with the tests)."""
    return np.diag([1. + p[0], 2. - p[1], 1. + .5 * p[2]])


def voro_mesh(ref):
    """VoroMesh.build_mesh (voromesh.py:19-201) on a bounded Voronoi tessellation of 48 jittered lattice points
    (cells of 6..16 faces), written in voro++'s text format by tests/voro_util.py; the file is committed as
    tests/golden/voro_48.vol.gz.  Dirichlet p on x-/x+, no-flow elsewhere, a source term; method 1 and the
    shifted centroids VoroMesh switches on."""
    import gzip
    import voro_util
    from mimpy.mesh import voromesh
    path = os.path.join(g1.OUT, "voro_48.vol.gz")
    if not os.path.exists(path):
        txt = voro_util.voronoi_vol_text(voro_util.jittered_points(4, 4, 3, VORO_BOX, 7), VORO_BOX)
        with gzip.GzipFile(path, "wb", mtime=0) as fh:
            fh.write(txt.encode())
    cwd = os.getcwd()
    os.chdir("/tmp")   # the reference opens 'shifted_out' and 'ortho_line' for writing in the cwd
    try:
        with open("/tmp/voro_48.vol", "w") as fh:
            fh.write(gzip.open(path, "rt").read())
        mesh = voromesh.VoroMesh()
        g1.quiet(mesh.build_mesh, "/tmp/voro_48.vol", voro_k)
    finally:
        os.chdir(cwd)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["face_shifted_centroids"] = np.array(mesh.face_shifted_centroids[:mesh.get_number_of_faces()])
    d["cell_shifted_centroids"] = np.array(mesh.cell_shifted_centroid[:mesh.get_number_of_cells()])
    bc = [("d", 0, lambda p: 1. + p[1]), ("d", 1, lambda p: 0.)] + [("n", b, cases.zero_flux) for b in (2, 3, 4, 5)]
    r = g1.dump_system(ref, mesh, 1, bc, lambda p: p[0] * p[2], with_blocks=True)
    for k_, v in r.items():
        d["m1_%s" % k_] = v
    return d


def multiplier_system(ref):
    """MFD with multiplier faces (faces of no cell): periodic boundaries + Lagrange pointers between two sub-domains
    (mesh.py:1420-1492, mfd.py:806-848): coupling COO, the full saddle matrix, rhs, spsolve."""
    import coupling_cases as cc

    _patch_coupling(ref)
    mesh = g1.ref_hex(ref, cc.NX, cc.NY, cc.NZ, cc.DIMS, cc.permeability())
    pairs, glue = cc.add_multipliers(mesh)
    # the reference stores zip(...) objects (mesh.py:1454), exhausted after one pass in Python 3: make them lists
    for lag in list(mesh.lagrange_to_face_pointers):
        mesh.lagrange_to_face_pointers[lag] = list(mesh.lagrange_to_face_pointers[lag])
    f = ref.mfd.MFD()
    cc.configure_multipliers(f, mesh)
    lhs = g1.quiet(f.build_lhs)
    rhs = f.build_rhs()
    cpl = f.build_coupling_terms()
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["periodic_pairs"], d["glue"] = np.array(pairs), np.array(glue)
    d["periodic_boundaries"] = np.array(mesh.periodic_boundaries)
    d["face_to_lagrange"] = np.array([[f] + list(mesh.get_face_to_lagrange_pointer(f))
                                      for f in mesh.get_all_face_to_lagrange_pointers()])
    d["flux_dof"] = np.array([f.flux_dof])
    d["face_to_flux"] = np.asarray(f.face_to_flux).reshape(-1).copy()
    d["coupling_data"], d["coupling_row"], d["coupling_col"] = np.array(cpl[0]), np.array(cpl[1]), np.array(cpl[2])
    csr = lhs.tocsr()
    csr.sort_indices()
    d["csr_indptr"], d["csr_indices"], d["csr_data"] = csr.indptr.copy(), csr.indices.copy(), csr.data.copy()
    d["rhs"] = rhs.copy()
    d["solution"] = np.array(f.solve())
    return d


def mesh_helpers(ref):
    """Host-side Mesh helpers next to the path (mesh.py:510-576 TPFA face shift, segment/face intersection; 1604-1637
    projected centroid; 2152-2184 domain hull) and the dense diagnostics of MFD (mfd.py:1481-1508), on a distorted
    4x3x2 box."""
    ktab = g1.spd_tensors(24, 9, full=True)
    mod = lambda p, i, j, k: p + 0.06 * np.array([np.sin(3. * p[1] + k), np.cos(2. * p[0] + i), np.sin(p[0] + 2. * p[1])])
    mesh = g1.ref_hex(ref, 4, 3, 2, (2., 1.5, 1.), ktab, mod)
    nf, nc = mesh.get_number_of_faces(), mesh.get_number_of_cells()
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    mesh.use_face_shifted_centroid()
    mesh.face_shifted_centroids = np.zeros((nf, 3))
    mesh.set_face_shifted_to_tpfa_all()
    d["tpfa_shifted"] = np.array(mesh.face_shifted_centroids[:nf])
    segs = [(np.array([.1, .2, -.5]), np.array([.3, .4, 1.5])), (np.array([-1., .7, .3]), np.array([3., .8, .35])),
            (np.array([.9, -1., .7]), np.array([1.1, 2.5, .2])), (np.array([.2, .2, .2]), np.array([.25, .22, .21]))]
    hit = np.zeros((len(segs), nf), dtype=np.int8)
    for a, (p1, p2) in enumerate(segs):
        for f in range(nf):
            r = mesh.is_line_seg_intersect_face(f, p1.copy(), p2.copy())
            hit[a, f] = -1 if r is None else int(bool(r))
    d["segments"], d["segment_hits"] = np.array(segs), hit
    d["projected"] = np.array([[mesh.find_centroid_for_coordinates(f, co) for co in ((0, 1), (1, 2), (0, 2))] for f in (0, 5, 17, 40)])
    for c in (5, 6, 9, 10, 21):
        mesh.set_cell_domain(c, 1)
    faces, orients = mesh.find_domain_faces(1)
    d["domain_faces"], d["domain_orientations"] = np.array(faces), np.array(orients)
    mesh.has_face_shifted_centroid = False
    f = ref.mfd.MFD()
    f.set_m_e_construction_method(1)
    f.set_mesh(mesh)
    for b in (0, 1):
        f.apply_dirichlet_from_function(b, lambda p: 1. + p[0])
    g1.quiet(f.build_lhs)
    f.lhs = f.lhs.tocsr()
    d["lhs_svd"] = np.array(f.lhs_svd())
    d["lhs_rank"] = np.array([f.lhs_rank()])
    return d


def singlephase_newton(ref):
    """SinglePhase.start_solving_newton (singlephase.py:300-356) on the set-up of fixture singlephase_543: two time
    steps of 100 passes each."""
    nx, ny, nz = 5, 4, 3
    ktab = g1.spd_tensors(nx * ny * nz, 11, full=False)
    mesh = g1.ref_hex(ref, nx, ny, nz, (5., 4., 3.), ktab)
    sp = ref.singlephase.SinglePhase()
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
    g1.quiet(sp.initialize_system)
    sp.set_initial_pressure(np.full(nc, 1.5))
    calls = []
    sp.time_step_output = lambda t, k: calls.append((t, k, np.array(sp.current_pressure).copy()))
    g1.quiet(sp.start_solving_newton)
    return {"p": np.array(sp.current_pressure), "v": np.array(sp.current_velocity),
            "output_times": np.array([c[0] for c in calls]), "output_steps": np.array([c[1] for c in calls]),
            "output_p": np.array([c[2] for c in calls])}


def singlephase_coupled(ref):
    """SinglePhase time loop (singlephase.py:182-299) on a mesh with Dirichlet / forcing pointers, periodic boundaries
    and Lagrange pointers: 5 steps of the unmodified reference."""
    import coupling_cases as cc

    _patch_coupling(ref)
    mesh = g1.ref_hex(ref, cc.NX, cc.NY, cc.NZ, cc.DIMS, cc.permeability())
    sp = ref.singlephase.SinglePhase()
    cc.configure_singlephase(sp, mesh)
    for lag in list(mesh.lagrange_to_face_pointers):   # zip objects (mesh.py:1454) are exhausted after one pass
        mesh.lagrange_to_face_pointers[lag] = list(mesh.lagrange_to_face_pointers[lag])
    g1.quiet(sp.initialize_system)
    nc = mesh.get_number_of_cells()
    sp.set_initial_pressure(np.full(nc, 1.5))
    p_hist, v_hist = [], []
    for step in range(5):
        g1.quiet(sp.start_solving)
        p_hist.append(np.array(sp.current_pressure))
        v_hist.append(np.array(sp.current_velocity))
    return {"p": np.array(p_hist), "v": np.array(v_hist), "flux_dof": np.array([sp.mfd.flux_dof])}


def main():
    ref = ref_shim.load_reference()
    if "--multipliers-only" in sys.argv:
        g1.save("multipliers_433", multiplier_system(ref))
        g1.save("singlephase_coupled_433", singlephase_coupled(ref))
        return
    if "--mesh-helpers-only" in sys.argv:
        g1.save("mesh_helpers_432", mesh_helpers(ref))
        return
    if "--newton-only" in sys.argv:
        g1.save("singlephase_newton_543", singlephase_newton(ref))
        return
    g1.save("voro_48", voro_mesh(ref))
    g1.save("mfd_helpers_432", mfd_helpers(ref))
    if "--helpers-only" in sys.argv:
        return
    g1.save("cut_checker_6", cut_checker(ref, 6))
    for name in cases.CASES:
        g1.save("twophase_" + name, twophase_history(ref, name))
    g1.save("coupling_433", coupling_system(ref))
    g1.save("gravity_432", gravity_rhs(ref))
    g1.save("multipliers_433", multiplier_system(ref))
    g1.save("singlephase_coupled_433", singlephase_coupled(ref))
    g1.save("mesh_helpers_432", mesh_helpers(ref))
    g1.save("singlephase_newton_543", singlephase_newton(ref))


if __name__ == "__main__":
    main()
