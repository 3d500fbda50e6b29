"""Host side of the multiplier faces (periodic boundaries / Lagrange pointers; mesh.py:1420-1492, mfd.py:806-848):
the pointer API of our Mesh reproduces the reference's mesh, the reference's flux numbering is recovered from the
device plan's (which leaves the multiplier faces out), the coupling COO is the reference's bit for bit, the CSR
renaming is exact, and the preconditioner of the coupled solve -- blockdiag(uncoupled matrix, identity on the
multipliers) -- closes the Krylov space after rank + 1 steps.  CPU only; golden from the unmodified reference."""
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import LinearOperator, gmres, splu

import coupling_cases as cc
from oracle import mimpy_oracle as orc
from test_cpu_abi_and_host import mesh_from_omesh


def test_multiplier_numbering_coupling_terms_and_solve_strategy(golden):
    import mimpy_b200.mfd.mfd as mfd

    g = golden("multipliers_433")
    ref = orc.omesh_from_dict(g)
    m = mesh_from_omesh(orc.hex_mesh(cc.NX, cc.NY, cc.NZ, *cc.DIMS, k_array=cc.permeability()))
    pairs, glue = cc.add_multipliers(m)
    o = orc.OMesh.from_mesh(m)
    for key in ("face_ptr", "face_points", "cell_ptr", "cell_faces", "cell_sigma"):
        assert np.array_equal(getattr(o, key), getattr(ref, key)), key
    assert np.array_equal(np.array(pairs), g["periodic_pairs"]) and np.array_equal(np.array(glue), g["glue"])
    assert np.array_equal(np.array(m.periodic_boundaries), g["periodic_boundaries"])
    assert np.array_equal(np.array([[f] + list(m.get_face_to_lagrange_pointer(f)) for f in m.get_all_face_to_lagrange_pointers()]),
                          g["face_to_lagrange"])
    assert m.has_pointer_couplings()

    f = mfd.MFD()
    f.mesh = m
    mult = f._multiplier_faces()
    assert sorted(mult) == sorted([p[4] for p in m.periodic_boundaries] + [t[2] for t in glue])
    nf, nc = m.get_number_of_faces(), m.get_number_of_cells()
    masked = np.zeros(nf, dtype=bool)
    for b in (2, 3):
        masked[[face for face, s in m.get_boundary_faces_by_marker(b)]] = True
    f.neumann_boundary_values = {int(x): 0. for x in np.nonzero(masked)[0]}
    assert np.array_equal(f._neumann_mask().astype(bool), masked | np.isin(np.arange(nf), mult))
    # what the device plan computes (A1, bit-exact on the GPU): unmasked faces in face order
    masked[mult] = True
    f2f_dev = np.where(~masked, np.cumsum(~masked) - 1, -1)
    g_ref, f_ref, dev2ref = mfd.reference_numbering(f2f_dev, mult, nc)
    assert f_ref == int(g["flux_dof"][0]) and np.array_equal(g_ref, g["face_to_flux"])
    assert np.all(np.diff(dev2ref) > 0) and len(dev2ref) == f_ref - len(mult) + nc

    f.face_to_flux, f.flux_dof = g_ref.reshape(-1, 1), f_ref
    f._ensure_plan = lambda: None
    cd, cr, ccol = f.build_coupling_terms()
    assert np.array_equal(cd, g["coupling_data"]) and np.array_equal(cr, g["coupling_row"])
    assert np.array_equal(ccol, g["coupling_col"])

    n = f_ref + nc
    k = sparse.csr_matrix((g["csr_data"], g["csr_indices"], g["csr_indptr"]), shape=(n, n))
    k0 = (k - sparse.coo_matrix((cd, (cr, ccol)), shape=(n, n)).tocsr()).tocsr()
    k0.eliminate_zeros()
    rows = g_ref[mult]
    assert abs(k0[rows]).sum() == 0 and abs(k0[:, rows]).sum() == 0   # multiplier rows hold coupling terms only
    a_dev = k0[dev2ref][:, dev2ref].tocsr()
    a_dev.sort_indices()
    assert abs(mfd.rename_csr(a_dev, dev2ref, n) - k0).max() == 0

    lu = splu(a_dev.tocsc())
    applications = [0]

    def pinv(v):
        applications[0] += 1
        z = np.zeros_like(v)
        z[dev2ref] = lu.solve(v[dev2ref])
        z[rows] = v[rows]
        return z

    x, info = gmres(k, g["rhs"], M=LinearOperator((n, n), matvec=pinv), rtol=1e-13, restart=200, maxiter=3)
    assert info == 0 and applications[0] <= 2 * len(cd) // 4 + 8
    assert np.linalg.norm(x - g["solution"]) <= 1e-12 * np.linalg.norm(g["solution"])


def test_multiplier_mesh_survives_the_mms_round_trip(tmp_path):
    """save_mesh / load_mesh with multiplier faces: faces of no cell, face -> multiplier pointers and multipliers that
    balance TWO faces each (one LAGRANGE_TO_FACE_POINTERS line per pair)."""
    from mimpy_b200.mesh.mesh import Mesh

    m = mesh_from_omesh(orc.hex_mesh(cc.NX, cc.NY, cc.NZ, *cc.DIMS, k_array=cc.permeability()))
    pairs, glue = cc.add_multipliers(m)
    path = str(tmp_path / "glued.mms")
    m.save_mesh(open(path, "w"))
    m2 = Mesh()
    m2.load_mesh(open(path))
    a, b = orc.OMesh.from_mesh(m), orc.OMesh.from_mesh(m2)
    for key in ("face_ptr", "face_points", "cell_ptr", "cell_faces", "cell_sigma", "face_areas", "face_to_cell"):
        assert np.array_equal(getattr(a, key), getattr(b, key)), key
    assert b.face_to_lagrange == a.face_to_lagrange and len(b.face_to_lagrange) == 2 * len(glue)
    assert b.lagrange_to_face == a.lagrange_to_face and all(len(v) == 2 for v in b.lagrange_to_face.values())
    assert m2.has_pointer_couplings()


def test_coupled_gmres_on_the_cpu_with_a_direct_inner_solve(golden):
    """The product's coupled solver (mfd.coupled_gmres: flexible GMRES, blockdiag(K0, I) preconditioner, warm start)
    driven on CPU tensors with SuperLU standing in for the hybridised device solve: the reference's solution of the
    periodic / Lagrange system in about rank + 1 preconditioner applications."""
    import contextlib
    import types

    import torch
    import mimpy_b200.mfd.mfd as mfd

    g = golden("multipliers_433")
    n = len(g["rhs"])
    k = sparse.csr_matrix((g["csr_data"], g["csr_indices"], g["csr_indptr"]), shape=(n, n))
    e = sparse.coo_matrix((g["coupling_data"], (g["coupling_row"], g["coupling_col"])), shape=(n, n)).tocsr()
    k0 = (k - e).tocsr()
    k0.eliminate_zeros()
    mult_rows = np.nonzero(np.diff(k0.indptr) == 0)[0]          # rows that hold coupling terms only
    assert len(mult_rows) == 21
    dev2ref = np.setdiff1d(np.arange(n), mult_rows)
    lu = splu(k0[dev2ref][:, dev2ref].tocsc())
    calls = [0]

    class Inner(object):
        def solve(self, v, rtol, maxit):
            calls[0] += 1
            info = types.SimpleNamespace(iterations=1, ms_total=0., residual=0., rhs_norm=float(torch.linalg.norm(v)))
            return torch.as_tensor(lu.solve(v.numpy())), info

    ctx = types.SimpleNamespace(device=torch.device("cpu"), on_stream=contextlib.nullcontext)
    apply_k = lambda v: torch.as_tensor(k @ v.numpy())
    rhs = torch.as_tensor(g["rhs"])
    x, info = mfd.coupled_gmres(ctx, apply_k, Inner(), rhs, 1e-13, 100, mult=(dev2ref, n, mult_rows))
    assert np.linalg.norm(x.numpy() - g["solution"]) <= 1e-11 * np.linalg.norm(g["solution"])
    assert info.iterations == calls[0] <= 50 and info.residual <= 1e-12 * info.rhs_norm
    # warm start from a nearby state: fewer applications; from the solution itself: none
    first = calls[0]
    x2, _ = mfd.coupled_gmres(ctx, apply_k, Inner(), rhs, 1e-13, 100, mult=(dev2ref, n, mult_rows),
                              x0=x + 1e-6 * torch.as_tensor(np.cos(np.arange(n, dtype=float))))
    assert np.linalg.norm(x2.numpy() - g["solution"]) <= 1e-11 * np.linalg.norm(g["solution"]) and calls[0] - first <= first
    before = calls[0]
    x3, info3 = mfd.coupled_gmres(ctx, apply_k, Inner(), rhs, 1e-10, 100, mult=(dev2ref, n, mult_rows), x0=x)
    assert calls[0] == before and info3.iterations == 0 and torch.equal(x3, x)
