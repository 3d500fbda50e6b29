"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; see oracle/ref_shim.py for how the
reference is imported without copying it).  Usage:  python tests/golden/gen_golden.py

Each fixture stores the inputs (the whole mesh as flat arrays, K, BCs) and what the reference
computed from them (M_E blocks, face_to_flux, COO triples, CSR arrays, rhs, spsolve solution,
two-phase states), so that tests can (a) pin the CPU oracle against the reference without the
reference being present and (b) check the CUDA path against the same numbers.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import mimpy_oracle as orc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def spd_tensors(nc, seed, full=True, spread=1.0):
    """Per-cell SPD K: Q diag(exp g) Q^T (full) or diag(exp g)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((nc, 3, 3))
    for c in range(nc):
        g = rng.standard_normal(3) * spread
        if full:
            q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            out[c] = q @ np.diag(np.exp(g)) @ q.T
        else:
            out[c] = np.diag(np.exp(g))
    return out


def ref_hex(ref, nx, ny, nz, dims, ktab, mod=None):
    m = ref.hexmesh.HexMesh()
    m.build_mesh(nx, ny, nz, dims[0], dims[1], dims[2],
                 lambda p, i, j, k: ktab[i + nx * j + nx * ny * k], mod)
    return m


def dump_system(ref, mesh, method, bc, forcing, with_blocks=True, solve=True):
    """bc: list of (kind, marker, fn). Returns dict of reference outputs."""
    f = ref.mfd.MFD()
    f.set_m_e_construction_method(method)
    f.set_mesh(mesh)
    for kind, marker, fn in bc:
        if kind == "d":
            f.apply_dirichlet_from_function(marker, fn)
        else:
            f.apply_neumann_from_function(marker, fn)
    if forcing is not None:
        f.apply_forcing_from_function(forcing)
    out = {}
    if with_blocks:
        blocks = [f.build_m_e(c) for c in range(mesh.get_number_of_cells())]
        out["m_e_flat"] = np.concatenate([b.ravel() for b in blocks])
    lhs = quiet(f.build_lhs)
    rhs = f.build_rhs()
    out["face_to_flux"] = f.face_to_flux[:, 0].copy()
    out["flux_dof"] = np.int64(f.flux_dof)
    out["coo_data"] = np.array(lhs.data)
    out["coo_row"] = np.array(lhs.row)
    out["coo_col"] = np.array(lhs.col)
    csr = lhs.tocsr()
    csr.sort_indices()
    out["csr_indptr"] = csr.indptr.copy()
    out["csr_indices"] = csr.indices.copy()
    out["csr_data"] = csr.data.copy()
    out["rhs"] = rhs.copy()
    if solve:
        out["solution"] = np.array(f.solve())
    return out


def save(name, d):
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only and name not in only:  # python gen_golden.py singlephase_543  -> rewrite only that fixture
        return
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %-28s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024.))


def main():
    ref = ref_shim.load_reference()

    # ---- 1. config-1 shape (hexmesh_example_1.py), n=5 -------------------------------
    n = 5
    ktab = np.tile(np.eye(3), (n ** 3, 1, 1))
    mesh = ref_hex(ref, n, n, n, (1., 1., 1.), ktab, lambda p, i, j, k: p)
    u = lambda p: p[0] ** 2 + p[1] ** 2 + p[2] ** 2
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d.update(dump_system(ref, mesh, 0, [("d", b, u) for b in range(6)], lambda x: -6.))
    d["method"] = np.int64(0)
    save("hex_c1_n5", d)

    # ---- 2. anisotropic full-tensor box, all 8 methods -------------------------------
    nx, ny, nz = 4, 3, 2
    ktab = spd_tensors(nx * ny * nz, 42, full=True)
    mesh = ref_hex(ref, nx, ny, nz, (2., 1.5, 1.), ktab)
    zero_flux = lambda p: np.array([0., 0., 0.])
    bc = [("d", 0, lambda p: 1. + 0.3 * p[1]), ("d", 1, lambda p: 0.2 * p[2]),
          ("n", 2, zero_flux), ("n", 3, zero_flux), ("n", 4, zero_flux), ("n", 5, zero_flux)]
    forcing = lambda x: np.sin(3. * x[0]) + x[1] * x[2]
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    for method in range(8):
        # methods 5 and 7 give diagonal M_E; method 4 is the inverse-of-W_E variant
        r = dump_system(ref, mesh, method, bc, forcing, solve=True)
        for k, v in r.items():
            d["m%d_%s" % (method, k)] = v
    save("hex_aniso_432", d)

    # ---- 3. distorted hexes (non-planar faces), diag K, method 1 ---------------------
    nx = ny = nz = 3
    ktab = spd_tensors(27, 7, full=False)

    def mod(p, i, j, k):
        return p + 0.06 * np.array([np.sin(5. * p[1] + p[2]), np.cos(4. * p[0] * p[2]), np.sin(3. * p[0] + 2. * p[1])])

    mesh = ref_hex(ref, nx, ny, nz, (1., 1., 1.), ktab, mod)
    lin = lambda p: 1. + 2. * p[0] - p[1] + .5 * p[2]
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    for method in (0, 1, 6):
        r = dump_system(ref, mesh, method, [("d", b, lin) for b in range(6)], lambda x: 0.3)
        for k, v in r.items():
            d["m%d_%s" % (method, k)] = v
    save("hex_distorted_333", d)

    # ---- 4. checkerboard-cut polyhedral mesh (config 5 shape), n=4 -------------------
    n = 4
    ktab = spd_tensors(n ** 3, 5, full=True)
    mesh = ref_hex(ref, n, n, n, (1., 1., 1.), ktab)
    nrm = np.array([1.2, .2, 1.3])
    nrm /= np.linalg.norm(nrm)
    # The reference's variable_array reallocates with refcheck=False while divide_cell_by_plane
    # still holds views into the old buffer (mesh.py:2293-2295 vs mesh.py:156-183), so cutting is
    # heap-state dependent.  Pre-size the ragged arrays through the reference's own public API so
    # that no reallocation happens during the cuts (no reference code is modified).
    for va in (mesh.faces, mesh.cells, mesh.cell_normal_orientation):
        va.set_data_capacity(40 * len(va.data) + 1000)
        va.set_pointer_capacity(8 * len(va.pointers) + 1000)
    for c in range(n ** 3):
        i, j, k = mesh.cell_to_ijk[c]
        if (i + j + k) % 2 == 0:
            mesh.divide_cell_by_plane(c, np.array(mesh.get_cell_real_centroid(c)), nrm)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    for method in (0, 1):
        r = dump_system(ref, mesh, method, [("d", b, lin) for b in range(6)], None)
        for k, v in r.items():
            d["m%d_%s" % (method, k)] = v
    # same mesh with the C2 boundary conditions (Dirichlet x-, x+ ; zero Neumann elsewhere)
    bc = [("d", 0, lambda p: 1.), ("d", 1, lambda p: 0.)] + [("n", b, zero_flux) for b in (2, 3, 4, 5)]
    r = dump_system(ref, mesh, 1, bc, None, with_blocks=False)
    for k, v in r.items():
        d["m1n_%s" % k] = v
    save("cut_checker_4", d)

    # ---- 5. 1-D flow known answer (SURVEY 8c iv) -------------------------------------
    ktab = np.tile(np.diag([1., 2., 3.]), (12, 1, 1))
    mesh = ref_hex(ref, 3, 2, 2, (1., 1., 1.), ktab)
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d.update(dump_system(ref, mesh, 1, bc, None))
    save("flow1d_322", d)

    # ---- 6. two-phase IMPES, 6x3x3, 12 steps (upwind recomputed at steps 1 and 10) ----
    for variant in ("wells", "inflow"):
        nx, ny, nz = 6, 3, 3
        rng = np.random.default_rng(128)
        kiso = 1.e-13 * np.exp(rng.standard_normal(nx * ny * nz))
        ktab = kiso[:, None, None] * np.eye(3)[None]
        mesh = ref_hex(ref, nx, ny, nz, (6., 3., 3.), ktab)
        tp = ref.twophase.TwoPhase()
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
        dt = 40. if variant == "wells" else 2000.
        tp.set_time_step_size(dt)
        cwd = os.getcwd()
        os.chdir("/tmp")
        if variant == "wells":
            for c in range(nc):
                if mesh.cell_to_ijk[c][0] == 0:
                    tp.add_rate_well(10. / (ny * nz), 0., c, "golden_w%d" % c)
        os.chdir(cwd)
        quiet(tp.initialize_system)
        steps = 12
        sw_hist, p_hist, u_hist = [], [], []
        nup = []
        for step in range(1, steps + 1):
            quiet(tp.update_pressure, step)
            if step == 1 or step % 10 == 0:
                tp.find_upwinding_direction()
                nup.append(len(tp.upwinded_face_cell))
            for _ in range(tp.saturation_time_steps):
                tp.update_saturation(step)
            sw_hist.append(np.array(tp.current_s_w))
            p_hist.append(np.array(tp.current_p_o))
            u_hist.append(np.array(tp.current_u_t))
        d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
        d["delta_t"] = np.float64(dt)
        d["steps"] = np.int64(steps)
        d["s_w"] = np.array(sw_hist)
        d["p_o"] = np.array(p_hist)
        d["u_t"] = np.array(u_hist)
        d["n_upwind_pairs"] = np.array(nup)
        d["face_to_flux"] = tp.mfd.face_to_flux[:, 0].copy()
        d["m_e_locations"] = np.array(tp.mfd.m_e_locations)
        d["m_data_for_update"] = np.array(tp.mfd.m_data_for_update)
        d["c_start"] = np.int64(tp.c_start)
        d["c_end"] = np.int64(tp.c_end)
        save("twophase_633_" + variant, d)

    # ---- 7. single-phase slightly compressible flow (singlephase.py:244-299), 5x4x3, 6 steps --------
    nx, ny, nz = 5, 4, 3
    ktab = spd_tensors(nx * ny * nz, 11, full=False)
    mesh = ref_hex(ref, nx, ny, nz, (5., 4., 3.), ktab)
    sp = ref.singlephase.SinglePhase()
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
    quiet(sp.initialize_system)
    sp.set_initial_pressure(np.full(nc, 1.5))
    p_hist, v_hist = [], []
    for step in range(6):
        quiet(sp.start_solving)          # one time step per call (number_of_time_steps = 1)
        p_hist.append(np.array(sp.current_pressure))
        v_hist.append(np.array(sp.current_velocity))
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["p"] = np.array(p_hist)
    d["v"] = np.array(v_hist)
    d["porosities"] = np.linspace(.1, .3, nc)
    save("singlephase_543", d)

    # ---- 9. Transport (transport.py): SinglePhase pressure step + explicit upwind concentration ----
    nx, ny, nz = 5, 4, 3
    rng = np.random.default_rng(3)
    kiso = np.exp(rng.standard_normal(nx * ny * nz))
    mesh = ref_hex(ref, nx, ny, nz, (2.5, 2., 1.8), kiso[:, None, None] * np.eye(3)[None])
    tr = ref.transport.Transport()
    f = ref.mfd.MFD()
    f.set_m_e_construction_method(1)
    tr.set_mesh_mfd(mesh, f)
    nc = mesh.get_number_of_cells()
    tr.apply_pressure_boundary_from_function(0, lambda p: 2.)
    tr.apply_pressure_boundary_from_function(1, lambda p: 1.)
    for b in (2, 3, 4, 5):          # Transport indexes with the number of faces: no Neumann faces
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
    tr.set_initial_concentration(np.linspace(0., .2, nc))
    quiet(tr.initialize_system)
    p_hist, v_hist, c_hist = [], [], []
    for step in range(6):
        quiet(tr.start_solving)
        p_hist.append(np.array(tr.current_pressure))
        v_hist.append(np.array(tr.current_velocity))
        c_hist.append(np.array(tr.current_concentration))
    d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
    d["p"], d["v"], d["c"] = np.array(p_hist), np.array(v_hist), np.array(c_hist)
    d["porosities"] = np.linspace(.1, .3, nc)
    d["c0"] = np.linspace(0., .2, nc)
    save("transport_543", d)

    # ---- 8. `.mms` file written by the reference's Mesh.save_mesh (mesh.py:811-949) ---------
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if not only or "mms_cut_232" in only:
        ktab = spd_tensors(12, 9, full=True)
        mesh = ref_hex(ref, 2, 3, 2, (1., 1.5, 1.), ktab)
        for va in (mesh.faces, mesh.cells, mesh.cell_normal_orientation):
            va.set_data_capacity(40 * len(va.data) + 1000)
            va.set_pointer_capacity(8 * len(va.pointers) + 1000)
        mesh.divide_cell_by_plane(4, np.array(mesh.get_cell_real_centroid(4)), nrm)
        mesh.add_internal_no_flow(int(mesh.get_cell(0)[3]), 1)
        mesh.set_forcing_pointer(2, [int(mesh.get_cell(7)[0])], [1])
        path = os.path.join(OUT, "ref_mms_cut_232.mms")
        mesh.save_mesh(open(path, "wb"))
        import gzip
        import shutil
        with open(path, "rb") as src, gzip.GzipFile(path + ".gz", "wb", mtime=0) as dst:
            shutil.copyfileobj(src, dst)  # the pre-sized ragged arrays are written in full: mostly zeros
        os.remove(path)
        path += ".gz"
        print("wrote %-28s %7.1f KiB" % (os.path.basename(path), os.path.getsize(path) / 1024.))
        d = orc.omesh_to_dict(orc.OMesh.from_mesh(mesh))
        d["n_points"] = np.int64(mesh.get_number_of_points())
        save("mms_cut_232", d)


if __name__ == "__main__":
    main()
