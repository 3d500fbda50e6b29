"""MFD discretisation with the public API of mimpy.mfd.mfd.MFD (reference mfd.py:33-1529), hot
path on the GPU.

What the reference does per cell in Python/NumPy (build_m_e, the n_E^2 COO scatter, build_div,
build_bottom_right, build_rhs) and with SuperLU (solve) runs here through libmimpy_b200:

    build_m / build_lhs  -> mimpy_b200_mfd_symbolic_* + mimpy_b200_mfd_numeric / _local_matrices
    build_rhs            -> mimpy_b200_mfd_rhs
    update_m             -> per-cell multipliers of mimpy_b200_mfd_numeric (mfd_cython.update_m_fast)
    solve                -> mimpy_b200_hybrid_* + mimpy_b200_pcg (exact hybridisation + Jacobi PCG)

Boundary/source callbacks keep the reference's signature (a Python callable of one point); an
array-valued fast path exists for meshes where nc Python calls are not an option.  Documented
deviations from the reference: the M block keeps its full STRUCTURAL pattern (the
reference drops exact zeros, mfd.py:582); get_velocity_solution returns the actual face fluxes
(the reference returns a constant, mfd.py:1477-1478).

Pointer couplings (build_coupling_terms, mfd.py:776-849): the coupling entries L, L^T are
added to the GPU-assembled saddle CSR, and the coupled system K = K0 + E is solved ON THE GPU by flexible GMRES
preconditioned with the hybridised PCG solve of the uncoupled system K0 (E has one entry pair per pointer,
so the Krylov space closes after about rank(E) + 1 steps).  Lagrange-multiplier / periodic pointers add
MULTIPLIER FACES (faces of no cell, Mesh.duplicate_face): the reference numbers them among the flux DOFs, the device
system leaves them out (their rows would be empty), the host renames device rows to reference rows (monotone, so the
CSR order survives) and the preconditioner of the GMRES solve is blockdiag(K0, I on the multipliers).
"""
import ctypes as C

import numpy as np
from scipy import sparse

from .. import _lib, device


def sign(x):
    """Returns the sign of float x (mfd.py:25-31)."""
    return 1. if abs(x) == x else -1.


def reference_numbering(face_to_flux_dev, multiplier_faces, nc):
    """The reference's flux numbering (mfd.py:141-159: every non-Neumann face in face order, multiplier faces
    included) from the device plan's numbering, which leaves the multiplier faces out.
    Returns (face_to_flux of the reference, its flux_dof, device index -> reference index over [flux ; cells])."""
    f2f = np.asarray(face_to_flux_dev)
    is_flux = f2f >= 0
    f_dev = int(is_flux.sum())
    real = is_flux.copy()
    is_flux[np.asarray(multiplier_faces, dtype=np.int64)] = True
    g_ref = np.where(is_flux, np.cumsum(is_flux) - 1, -1)
    f_ref = int(is_flux.sum())
    dev2ref = np.empty(f_dev + nc, dtype=np.int64)
    dev2ref[f2f[real]] = g_ref[real]
    dev2ref[f_dev:] = f_ref + np.arange(nc)
    return g_ref, f_ref, dev2ref


def rename_csr(a, dev2ref, n):
    """Rows and columns of the CSR matrix `a` renamed by the monotone map dev2ref into an n x n matrix (the
    entries keep their order; rows / columns that nothing maps to stay empty)."""
    lens = np.zeros(n, dtype=np.int64)
    lens[dev2ref] = np.diff(a.indptr)
    indptr = np.concatenate([[0], np.cumsum(lens)])
    return sparse.csr_matrix((a.data, dev2ref[a.indices].astype(np.int32), indptr.astype(np.int32)), shape=(n, n))


def coupled_gmres(ctx, apply_k, hybrid, rhs, rtol, maxit, mult=None, x0=None, restart=200):
    """Flexible GMRES for the pointer-coupled saddle system K = K0 + E (device vectors; `apply_k(v)` = K v),
    preconditioned by the hybridised PCG solve of the uncoupled system K0 (`hybrid`, a device.HybridSystem).
    mult: the multiplier faces of Lagrange / periodic pointers (MFD._mult) -- their rows are empty in K0, so the
    preconditioner is blockdiag(K0 on the mesh DOFs, identity on the multipliers): still K up to a low-rank term,
    and the Krylov space closes after about rank + 1 steps.  Returns (x, info of the last inner solve with the
    totals of all of them)."""
    import torch
    total_it, total_ms, info = 0, 0., None
    if mult is not None:
        with ctx.on_stream():
            mesh_rows = torch.as_tensor(mult[0], device=ctx.device)
            mult_rows = torch.as_tensor(np.asarray(mult[2], dtype=np.int64), device=ctx.device)

    def Pinv(v):
        nonlocal total_it, total_ms, info
        if mult is None:
            z, info = hybrid.solve(v, rtol=min(rtol, 1e-12), maxit=maxit)
        else:
            v_mesh = v[mesh_rows].contiguous()
            z = torch.zeros_like(v)
            z[mult_rows] = v[mult_rows]
            if float(torch.linalg.norm(v_mesh)) == 0.:   # a Krylov vector that lives on the multipliers only
                return z
            z_mesh, info = hybrid.solve(v_mesh, rtol=min(rtol, 1e-12), maxit=maxit)
            z[mesh_rows] = z_mesh
        total_it += int(info.iterations)
        total_ms += float(info.ms_total)
        return z

    with ctx.on_stream():
        b = rhs
        bnorm = float(torch.linalg.norm(b))
        x = torch.zeros_like(b) if x0 is None else x0.clone()
        r = b.clone() if x0 is None else b - apply_k(x)
        tol = max(rtol, 1e-13) * max(bnorm, 1e-300)
        for _cycle in range(10):
            beta = float(torch.linalg.norm(r))
            if beta <= tol:
                break
            V, Z = [r / beta], []
            Hm = np.zeros((restart + 1, restart))
            y = None
            for j in range(restart):
                Z.append(Pinv(V[j]))
                w = apply_k(Z[j])
                for i in range(j + 1):            # modified Gram-Schmidt
                    Hm[i, j] = float(torch.dot(w, V[i]))
                    w = w - Hm[i, j] * V[i]
                Hm[j + 1, j] = float(torch.linalg.norm(w))
                e1 = np.zeros(j + 2)
                e1[0] = beta
                y, res, _, _ = np.linalg.lstsq(Hm[:j + 2, :j + 1], e1, rcond=None)
                resid = float(np.linalg.norm(Hm[:j + 2, :j + 1] @ y - e1))
                if resid <= tol or Hm[j + 1, j] <= 1e-300:
                    break
                V.append(w / Hm[j + 1, j])
            for c, z in zip(y, Z):
                x = x + float(c) * z
            r = b - apply_k(x)
    if info is None:   # the right-hand side was met by the initial guess
        import types
        info = types.SimpleNamespace(iterations=0, ms_total=0., residual=0., rhs_norm=bnorm)
    info.iterations, info.ms_total = total_it, total_ms
    info.residual = float(torch.linalg.norm(r))
    info.rhs_norm = bnorm
    return x, info


class MFD(object):
    def __init__(self):
        self.mesh = None
        self.lhs = None
        self.rhs = None
        self.solution = None
        self.flux_dof = 0
        self.pressure_dof = 0
        self.total_dof = 0
        self.face_to_flux = None
        self.neumann_needs_updating = True
        self.check_m_e = False
        self.print_progress = False
        self.verbose = False
        self.compute_diagonality = False
        self.diagonality_index_list = None
        self.m_e_construction_method = 0
        self.m_data_for_update = None
        self.m_e_locations = None
        self.dirichlet_boundary_values = {}
        self.neumann_boundary_values = {}
        self.cell_forcing_function = np.empty(shape=(0), dtype=float)
        self.cell_faces_neumann = {}
        self.non_ortho_count = 0
        self.all_ortho = True
        # device state
        self.ctx = None
        self._plan = None
        self._plan_key = None
        self._m_e = None        # device blocks of the current (mesh, method)
        self._m_e_key = None
        self._hybrid = None
        self._mult = None       # (device index -> reference index, reference dofs, rows of the multiplier faces)
        self.device_lhs = None  # (indptr, indices, vals) tensors of the saddle CSR
        self.device_rhs = None
        self.solver_info = None
        self.pcg_rtol = 1.e-12
        self.pcg_maxit = 50000

    # ---- set-up ----------------------------------------------------------------------------
    def set_mesh(self, mesh, ctx=None):
        """Links the MFD object to a Mesh (mfd.py:91-97)."""
        self.mesh = mesh
        self.ctx = ctx or _lib.default_context()
        self.cell_forcing_function = np.zeros(mesh.get_number_of_cells())
        self.process_internal_boundaries()

    def get_flux_dof(self):
        return self.flux_dof

    def get_pressure_dof(self):
        return self.pressure_dof

    def get_total_dof(self):
        return self.total_dof

    def set_compute_diagonality(self, setting):
        self.compute_diagonality = setting

    def get_diagonality(self):
        return self.diagonality_index_list

    def set_m_e_construction_method(self, method_index):
        """0 => U = [R^T R (R^T N)^-1]_00, 1 => |E|/tr K, 2,3,5,6,7 as in mfd.py:379-468."""
        self.m_e_construction_method = method_index
        self._m_e_key = None

    def process_internal_boundaries(self):
        """mfd.py:1194-1207"""
        for (face_index, _orient) in self.mesh.get_internal_no_flow():
            cell_index = int(self.mesh.face_to_cell[face_index][0])
            self.cell_faces_neumann.setdefault(cell_index, []).append(face_index)
            self.neumann_boundary_values[face_index] = 0.
            self.neumann_needs_updating = True

    # ---- boundary conditions and sources (mfd.py:1060-1176) --------------------------------------
    def apply_dirichlet_from_function(self, boundary_marker, p_function):
        """value = p(x_f) a_f sigma_bdry for every face of the marker (mfd.py:1143-1162)."""
        m = self.mesh
        for (f, orient) in m.get_boundary_faces_by_marker(boundary_marker):
            value = p_function(m.get_face_real_centroid(f))
            value *= m.get_face_area(f)
            self.dirichlet_boundary_values[f] = value * orient

    def apply_neumann_from_function(self, boundary_marker, grad_u):
        """value = grad_u(x_f) . n_f; the face stops being a flux DOF (mfd.py:1060-1090)."""
        m = self.mesh
        for (f, _orient) in m.get_boundary_faces_by_marker(boundary_marker):
            value = np.asarray(grad_u(m.get_face_real_centroid(f))).dot(m.get_face_normal(f))
            self.set_neumann_by_face(f, _orient, value)

    def set_neumann_by_face(self, face_index, face_orientation, value):
        """mfd.py:1092-1109"""
        self.neumann_boundary_values[face_index] = value
        cell_index = int(self.mesh.face_to_cell[face_index][0])
        self.cell_faces_neumann.setdefault(cell_index, []).append(face_index)
        self.neumann_needs_updating = True

    def apply_forcing_from_function(self, forcing_function):
        """F_c = f(x_E) |E| (mfd.py:1164-1176)."""
        m = self.mesh
        nc = m.get_number_of_cells()
        cen = m.get_all_cell_real_centroids()
        vol = m.cell_volume[:nc]
        out = np.empty(nc)
        for c in range(nc):
            out[c] = forcing_function(cen[c]) * vol[c]
        self.cell_forcing_function = out

    def apply_forcing_from_grad(self, grad_p):
        """Source term of a cell from the exact flux through its faces (mfd.py:1178-1192), vectorised over the
        faces: F_c += sum_f sigma_f a_f grad_p(x_f) . n_f"""
        m = self.mesh
        nf = m.get_number_of_faces()
        flux = np.array([np.dot(grad_p(m.get_face_real_centroid(f)), m.get_face_normal(f)) for f in range(nf)])
        flux *= np.asarray(m.face_areas[:nf])
        ptr, data = m.cells.compact()
        sigma = np.asarray(m.cell_normal_orientation.compact()[1], dtype=float)
        contrib = flux[np.asarray(data)] * sigma
        cells = np.repeat(np.arange(m.get_number_of_cells()), np.diff(ptr))
        np.add.at(self.cell_forcing_function, cells, contrib)

    def apply_forcing_from_array(self, values):
        """Extension: f sampled at the cell centroids, as an array (no Python callback per cell)."""
        nc = self.mesh.get_number_of_cells()
        self.cell_forcing_function = np.asarray(values, dtype=float) * self.mesh.cell_volume[:nc]

    def get_dirichlet_boundary_values(self):
        return self.dirichlet_boundary_values.items()

    def get_dirichlet_value(self, face_index):
        return self.dirichlet_boundary_values[face_index]

    def get_number_of_dirichlet_faces(self):
        return len(self.dirichlet_boundary_values)

    def get_neumann_boundary_values(self):
        return self.neumann_boundary_values.items()

    def reset_dirichlet_boundaries(self):
        self.dirichlet_boundary_values = {}

    def reset_neumann_boundaries(self):
        self.neumann_boundary_values = {}
        self.neumann_needs_updating = True

    def get_cell_faces_neumann(self, cell_index):
        return self.cell_faces_neumann.get(cell_index, [])

    def is_neumann_face(self, face_index):
        return face_index in self.neumann_boundary_values

    # ---- device plumbing -----------------------------------------------------------------------
    def _dmesh(self):
        return self.mesh.to_device(self.ctx)

    def _multiplier_faces(self):
        """Faces that belong to no cell: the multipliers of periodic boundaries and of Lagrange pointers
        (Mesh.duplicate_face, mesh.py:1476-1492, 2106-2150).  They are flux DOFs of the reference's numbering whose
        rows hold coupling terms only."""
        m = self.mesh
        if not (m.periodic_boundaries or m.face_to_lagrange_pointers or m.lagrange_to_face_pointers) or len(m.cells) == 0:
            return np.empty(0, dtype=np.int64)
        nf = m.get_number_of_faces()
        used = np.zeros(nf, dtype=bool)
        used[self.mesh.cells.compact()[1]] = True
        return np.nonzero(~used)[0]

    def _neumann_mask(self):
        """Faces that are no flux DOF of the DEVICE plan: Neumann faces, and multiplier faces (no cell assembles
        into their rows; the host keeps them in the reference's numbering, see _ensure_plan)."""
        mask = np.zeros(self.mesh.get_number_of_faces(), dtype=np.uint8)
        if self.neumann_boundary_values:
            mask[np.fromiter(self.neumann_boundary_values.keys(), dtype=np.int64)] = 1
        mask[self._multiplier_faces()] = 1
        return mask

    def _ensure_plan(self):
        dm = self._dmesh()
        key = (id(dm), tuple(sorted(self.neumann_boundary_values.keys())) if len(self.neumann_boundary_values) < 4096
               else hash(self._neumann_mask().tobytes()))
        if self._plan is None or self._plan_key != key or self.neumann_needs_updating:
            self._plan = device.symbolic(dm, self._neumann_mask())
            self._plan_key = key
            self._hybrid = None
            f2f = device.to_host(self.ctx, self._plan.face_to_flux)[:dm.nf]
            self.flux_dof = self._plan.flux_dof
            self._mult = None
            mult = self._multiplier_faces()
            if len(mult):
                # the reference numbers the multiplier faces among the flux DOFs, in face order (mfd.py:141-159); the
                # device system leaves them out.  Both numberings run in face order, so device index -> reference
                # index is monotone: rows and columns of the device CSR keep their order when they are renamed.
                g_ref, f_ref, dev2ref = reference_numbering(f2f, mult, dm.nc)
                self._mult = (dev2ref, f_ref + dm.nc, g_ref[mult])
                f2f, self.flux_dof = g_ref, f_ref
            self.face_to_flux = f2f.reshape(-1, 1).astype(np.dtype("i"))
            self.pressure_dof = dm.nc
            self.total_dof = self.flux_dof + dm.nc
            self.neumann_needs_updating = False
        return dm, self._plan

    def renumber_for_neumann(self):
        """face_to_flux / flux_dof (mfd.py:141-159); computed with the rest of the symbolic plan."""
        self._ensure_plan()

    def _blocks(self, k_unity=False):
        dm, plan = self._ensure_plan()
        key = (id(dm), self.m_e_construction_method, bool(k_unity), self.mesh._version)
        if self._m_e is None or self._m_e_key != key:
            want_d = bool(self.compute_diagonality)
            m_e, diag, cons = device.local_matrices(dm, plan, self.m_e_construction_method, k_unity=k_unity,
                                                    want_diagonality=want_d, want_consistency=bool(self.check_m_e))
            self._m_e, self._m_e_key = m_e, key
            if want_d:
                self.diagonality_index_list = device.to_host(self.ctx, diag)
                self.all_ortho = bool((self.diagonality_index_list <= 1.e-8).all())
            if self.check_m_e:
                bad = device.to_host(self.ctx, cons)
                for c in np.nonzero(bad > 1.e-6)[0]:
                    print("M_E N ne R", bad[c])
        return self._m_e

    # ---- local matrices (host views of the device results; mfd.py:248-483) -----------------------
    def build_r_e(self, cell_index):
        m = self.mesh
        xc = m.get_cell_shifted_centroid(cell_index) if m.is_using_cell_shifted_centroid() else m.get_cell_real_centroid(cell_index)
        f = np.asarray(m.get_cell(cell_index))
        fc = m.face_shifted_centroids[f] if m.is_using_face_shifted_centroid() else m.face_real_centroids[f]
        return m.face_areas[f][:, None] * (fc - xc[None, :])

    def build_n_e(self, cell_index, k_unity=False):
        m = self.mesh
        f = np.asarray(m.get_cell(cell_index))
        s = np.asarray(m.get_cell_normal_orientation(cell_index), dtype=float)
        n = m.face_normals[f]
        if not k_unity:
            n = n @ m.get_cell_k(cell_index).T
        return n * s[:, None]

    def build_c_e(self, n_e):
        """Orthonormal basis of null(N_E^T) from the SVD (mfd.py:314-322); the device kernels only ever need
        C_E C_E^T = I - Q Q^T and build Q by CholeskyQR2 instead."""
        _, _, v = np.linalg.svd(np.transpose(n_e))
        return np.transpose(v)[:, 3:]

    def build_d_e(self, n_e):
        """mfd.py:324-336 (the same null-space basis; two-dimensional meshes keep one more column)."""
        _, _, v = np.linalg.svd(np.transpose(n_e))
        return np.transpose(v)[:, (3 if self.mesh.dim == 3 else 2):]

    def build_w_e(self, cell_index):
        """W_E of Brezzi et al. (mfd.py:338-355); method 4 inverts it (on the device for whole meshes)."""
        r_e = self.build_r_e(cell_index)
        n_e = self.build_n_e(cell_index)
        d_e = self.build_d_e(r_e)
        current_k = self.mesh.get_cell_k(cell_index)
        w_e = n_e.dot(np.linalg.inv(n_e.T.dot(r_e)).dot(n_e.T))
        w_e += d_e.dot(d_e.T)
        w_e *= np.trace(current_k)
        w_e /= self.mesh.get_cell_volume(cell_index)
        return w_e

    def proj_vector(self, vector_f):
        """Per cell, the normal components sigma n_f . F(x_f) of a vector field at the face centroids
        (mfd.py:601-616)."""
        m = self.mesh
        out = []
        for cell_index in range(m.get_number_of_cells()):
            out.append([vector_f(m.get_face_real_centroid(f)).dot(m.get_face_normal(f) * o)
                        for f, o in zip(m.get_cell(cell_index), m.get_cell_normal_orientation(cell_index))])
        return out

    def build_m_e(self, cell_index, k_unity=False):
        """The n_E x n_E block of one cell (computed for all cells on the GPU, then sliced)."""
        blocks = self._blocks(k_unity)
        _, plan = self._ensure_plan()
        ptr = device.to_host(self.ctx, plan.m_ptr)
        n = self.mesh.get_number_of_cell_faces(cell_index)
        a, b = int(ptr[cell_index]), int(ptr[cell_index + 1])
        return device.to_host(self.ctx, blocks[a:b]).reshape(n, n).copy()

    # ---- global blocks ------------------------------------------------------------------------------
    def build_m(self, save_update_info=False, k_unity=False):
        """COO triples of M, cell-major / row-major like the reference (mfd.py:545-599), every
        structural entry kept.  Returns [data, row, col] (NumPy arrays)."""
        dm, plan = self._ensure_plan()
        blocks = self._blocks(k_unity)
        loc, rows, cols, vals = device.coo_view(dm, plan, blocks)
        H = lambda t: device.to_host(self.ctx, t)
        data, row, col = H(vals), H(rows), H(cols)
        if self._mult is not None:
            row, col = self._mult[0][row].astype(row.dtype), self._mult[0][col].astype(col.dtype)
        if save_update_info:
            self.m_data_for_update = data.copy()
            self.m_e_locations = H(loc)
        if self.verbose:
            print("all ortho", self.all_ortho)
        return [data, row, col]

    def build_div(self, shift=1):
        """[[div_data, div_row, div_col], [div_t_data, div_t_row, div_t_col]] (mfd.py:203-246)."""
        dm, plan = self._ensure_plan()
        m = self.mesh
        cptr, cfaces = m.cells.compact()
        _, sig = m.cell_normal_orientation.compact()
        g = self.face_to_flux[:, 0][cfaces]
        keep = g >= 0
        cell_of = np.repeat(np.arange(dm.nc), np.diff(cptr))
        data = (sig * m.face_areas[cfaces])[keep].astype(float)
        row = (cell_of[keep] + shift * self.flux_dof).astype(np.dtype("i"))
        col = g[keep].astype(np.dtype("i"))
        return [[data, row, col], [-data, col.copy(), row.copy()]]

    def build_bottom_right(self, alpha=0., shift=1):
        """C[c,c] = alpha |E| (explicit zeros kept; mfd.py:745-774)."""
        m = self.mesh
        nc = m.get_number_of_cells()
        if self.flux_dof == 0 and self._plan is None:
            self._ensure_plan()
        if m.is_using_alpha_list:
            data = np.array([m.get_alpha_by_cell(c) for c in range(nc)], dtype=float) * m.cell_volume[:nc]
        else:
            data = alpha * m.cell_volume[:nc]
        idx = (np.arange(nc) + shift * self.flux_dof).astype(np.dtype("i"))
        return [np.asarray(data, dtype=float), idx, idx.copy()]

    def build_coupling_terms(self):
        """COO triples [data, row, col] of the matrices L and L^T that couple cells and faces through the
        mesh's forcing / Dirichlet / Lagrange pointers and periodic boundaries (mfd.py:776-849)."""
        self._ensure_plan()
        m, F = self.mesh, self.flux_dof
        g = self.face_to_flux[:, 0]
        data, row, col = [], [], []
        for cell in m.get_forcing_pointer_cells():
            for (f, orient) in m.get_forcing_pointers_for_cell(cell):
                data.append(-m.get_face_area(f) * orient)
                row.append(cell + F)
                col.append(int(g[f]))
        for f in m.get_dirichlet_pointer_faces():
            (cell, orient) = m.get_dirichlet_pointer(f)
            data.append(m.get_face_area(f) * orient)
            row.append(int(g[f]))
            col.append(cell + F)
        for f in m.get_all_face_to_lagrange_pointers():
            (lag, orient) = m.get_face_to_lagrange_pointer(f)
            data.append(m.get_face_area(f) * orient)
            row.append(int(g[f]))
            col.append(int(g[lag]))
        for lag in m.get_all_lagrange_to_face_pointers():
            for (f, orient) in m.get_lagrange_to_face_pointers(lag):
                data.append(-m.get_face_area(f) * orient)
                row.append(int(g[lag]))
                col.append(int(g[f]))
        for (f1, o1, f2, o2, lag) in m.periodic_boundaries:
            data += [m.get_face_area(f1) * o1, m.get_face_area(f2) * o2, o1, o2]     # L^T, L^T, L, L
            row += [int(g[f1]), int(g[f2]), int(g[lag]), int(g[lag])]
            col += [int(g[lag]), int(g[lag]), int(g[f1]), int(g[f2])]
        return [np.asarray(data, dtype=float), np.asarray(row, dtype=np.dtype("i")), np.asarray(col, dtype=np.dtype("i"))]

    def build_lhs(self, alpha=0, beta=None):
        """The saddle-point matrix [M -DIV^T; DIV C] (mfd.py:851-914).  Assembled on the GPU straight
        into sorted CSR (self.device_lhs); self.lhs is its SciPy view on the host."""
        if beta is not None:
            raise NotImplementedError("advection terms (mfd.py:618-668) are not part of this build")
        dm, plan = self._ensure_plan()
        method = self.m_e_construction_method
        c_diag = None
        if self.mesh.is_using_alpha_list:
            import torch
            c_diag = torch.as_tensor(self.build_bottom_right()[0], device=self.ctx.device)
        if self.compute_diagonality or self.check_m_e or method == 4:
            vals = device.numeric_from_blocks(dm, plan, self._blocks(), alpha=float(alpha), c_diag=c_diag)
        else:
            vals = device.numeric(dm, plan, method, alpha=float(alpha), c_diag=c_diag)
        self._alpha, self._c_diag = float(alpha), c_diag
        self.device_lhs = (plan.indptr, plan.indices, vals)
        n = plan.ndof
        H = lambda t: device.to_host(self.ctx, t)
        self.lhs = sparse.csr_matrix((H(vals)[:plan.nnz], H(plan.indices)[:plan.nnz], H(plan.indptr)[:n + 1]), shape=(n, n))
        if self._mult is not None:
            # rename to the reference's numbering: empty rows / columns appear where the multiplier faces sit
            dev2ref, n, _ = self._mult
            self.lhs = rename_csr(self.lhs, dev2ref, n)
        self._coupled = None
        if self.mesh.has_pointer_couplings():
            # the coupling entries sit off the mesh pattern: merged into the assembled CSR (the uncoupled matrix
            # stays on the device as the preconditioner of the coupled solve)
            import torch
            cd, cr, cc = self.build_coupling_terms()
            self.lhs = (self.lhs + sparse.coo_matrix((cd, (cr, cc)), shape=(n, n))).tocsr()
            self.lhs.sort_indices()
            dev = self.ctx.device
            self._coupled = (torch.as_tensor(self.lhs.indptr.astype(np.int32), device=dev),
                             torch.as_tensor(self.lhs.indices.astype(np.int32), device=dev),
                             torch.as_tensor(self.lhs.data, device=dev))
        return self.lhs

    def get_number_of_dof(self):
        return self.mesh.get_number_of_cells() + self.flux_dof

    def build_rhs(self):
        """mfd.py:991-999 (Neumann quirk, Dirichlet, forcing) on the device; returns the host copy."""
        dm, plan = self._ensure_plan()
        dfaces = list(self.dirichlet_boundary_values.keys())
        dvals = [self.dirichlet_boundary_values[f] for f in dfaces]
        nvals = list(self.neumann_boundary_values.values())
        self.device_rhs = device.build_rhs(dm, plan, dfaces, dvals, len(nvals) > 0, nvals[-1] if nvals else 0.,
                                           self.cell_forcing_function)
        if self._mult is not None:
            import torch
            dev2ref, n, _ = self._mult
            with self.ctx.on_stream():
                full = torch.zeros(n, dtype=torch.float64, device=self.ctx.device)
                full[torch.as_tensor(dev2ref, device=self.ctx.device)] = self.device_rhs
            self.device_rhs = full
        self.rhs = device.to_host(self.ctx, self.device_rhs)
        return self.rhs

    # the three parts of build_rhs (mfd.py:1027-1041, 1209-1215) on the host vector self.rhs, for callers that
    # compose a right-hand side themselves; build_rhs itself runs them fused on the device
    def build_rhs_forcing(self):
        self.rhs[self.flux_dof:self.flux_dof + self.mesh.get_number_of_cells()] += \
            np.asarray(self.cell_forcing_function)[:self.mesh.get_number_of_cells()]

    def build_rhs_neumann(self):
        for boundary_index, boundary_value in self.get_neumann_boundary_values():
            self.rhs[self.face_to_flux[boundary_index, 0]] = boundary_value   # (g = -1: the last entry, like the reference)

    def build_rhs_dirichlet(self):
        for boundary_index, boundary_value in self.get_dirichlet_boundary_values():
            self.rhs[self.face_to_flux[boundary_index, 0]] = -boundary_value

    def add_gravity(self):
        """rhs[f] += (M_{K=I} g)_f with g_f = rho * |g| * (e_g . n_f) (mfd.py:1001-1025), computed on the device:
        the k_unity saddle matrix applied to [g; 0].  Like the reference, which indexes the right-hand side by
        FACE and multiplies a flux_dof-sized matrix with a face-sized vector, this needs a problem without
        Neumann faces (flux_dof == number of faces)."""
        import torch
        dm, plan = self._ensure_plan()
        m = self.mesh
        nf = m.get_number_of_faces()
        if plan.flux_dof != nf or self._mult is not None:
            raise ValueError("dimension mismatch: add_gravity needs flux_dof == number of faces (mfd.py:1015-1024)")
        if self.rhs is None:
            self.build_rhs()
        scale = m.get_fluid_density() * m.get_gravity_acceleration()
        gvec = np.asarray(m.get_gravity_vector(), dtype=float)
        with self.ctx.on_stream():
            force = torch.zeros(plan.ndof, dtype=torch.float64, device=self.ctx.device)
            force[:nf] = scale * (dm.face_normals @ torch.as_tensor(gvec, device=self.ctx.device))
            vals = device.numeric(dm, plan, self.m_e_construction_method, k_unity=True)
            mg = device.spmv(self.ctx, plan.indptr, plan.indices, vals, force, n=plan.ndof)
            self.device_rhs = self.device_rhs.clone()
            self.device_rhs[:nf] += mg[:nf]
        self.rhs = device.to_host(self.ctx, self.device_rhs)

    def update_m(self, m_coo, multipliers):
        """m_coo[k] = m_data_for_update[k] * multipliers[cell of k] (mfd.py:729-743,
        mfd_cython.pyx:6-22).  On the device path the same multipliers go to numeric assembly;
        this host version serves code that keeps the reference's COO list."""
        counts = np.diff(self.m_e_locations)
        m_coo[:self.m_e_locations[-1]] = self.m_data_for_update * np.repeat(np.asarray(multipliers), counts)

    # ---- solve -------------------------------------------------------------------------------------------
    def solve(self, initial_guess=0., solver='spsolve', rtol=None, maxit=None):
        """Solves the saddle system for [v ; p] (mfd.py:1349-1368).  Every `solver` value runs the
        same GPU path: exact hybridisation to the SPD face system + PCG (auxiliary-space multigrid
        on structured hex meshes, Jacobi otherwise) + local recovery; there is no sparse direct
        factorisation on the device."""
        import torch

        dm, plan = self._ensure_plan()
        if self.device_rhs is None or self.rhs is None:
            self.build_rhs()
        rhs = self.device_rhs
        if self.rhs is not None and rhs is not None:
            # honour host-side edits of self.rhs made after build_rhs()
            host = device.to_host(self.ctx, rhs)
            if not np.array_equal(host, self.rhs):
                rhs = torch.as_tensor(np.asarray(self.rhs, dtype=float), device=self.ctx.device)
        if self._hybrid is None or self._hybrid_key != (self._m_e_key, getattr(self, "_alpha", 0.)):
            self._hybrid = device.HybridSystem(dm, plan, self._blocks(), alpha=getattr(self, "_alpha", 0.),
                                               c_diag=getattr(self, "_c_diag", None))
            self._hybrid_key = (self._m_e_key, getattr(self, "_alpha", 0.))
        if getattr(self, "_coupled", None) is not None:
            sol, info = self._solve_coupled(rhs, rtol or self.pcg_rtol, maxit or self.pcg_maxit)
        elif self._mult is not None:
            raise ValueError("the mesh has multiplier faces (faces of no cell) but no pointer couples them: singular system")
        else:
            sol, info = self._hybrid.solve(rhs, rtol=rtol or self.pcg_rtol, maxit=maxit or self.pcg_maxit)
        self.solver_info = {"iterations": int(info.iterations), "residual": float(info.residual),
                            "rhs_norm": float(info.rhs_norm), "ms": float(info.ms_total)}
        self.device_solution = sol
        self.solution = device.to_host(self.ctx, sol)
        return self.solution

    def _solve_coupled(self, rhs, rtol, maxit):
        """The coupled saddle CSR (device SpMV) solved by `coupled_gmres` around the hybridised uncoupled solve."""
        ip, ix, vv = self._coupled
        n_full = int(ip.numel() - 1)
        apply_k = lambda v: device.spmv(self.ctx, ip, ix, vv, v, n=n_full)
        return coupled_gmres(self.ctx, apply_k, self._hybrid, rhs, rtol, maxit, mult=self._mult)

    # ---- diagnostics of the assembled saddle matrix (mfd.py:1481-1529; small systems only: dense) -----------------
    def lhs_svd(self):
        return np.linalg.svd(self.lhs.toarray())[1]

    def lhs_rank(self):
        return int((np.abs(self.lhs_svd()) > 1.e-9).sum())

    def print_lhs_to_file(self, filename):
        """The whole matrix in dense text form, `value , value , ...` per row."""
        with open(filename + ".dat", "w") as out:
            for row in self.lhs.toarray():
                out.write(" ".join("%s ," % v for v in row) + " \n")

    def print_lhs_to_ppm(self, filename):
        """The sparsity pattern (|entry| > 1e-10) as a plain-text bitmap."""
        dense = np.abs(self.lhs.toarray()) > 1.e-10
        with open(filename + ".ppm", "w") as out:
            out.write("P1\n%d %d\n" % dense.shape)
            for row in dense:
                out.write(" ".join("1" if x else "0" for x in row) + " \n")

    def get_pressure_solution(self):
        return self.solution[self.flux_dof:]

    def get_velocity_solution_by_index(self, face_index):
        index = self.face_to_flux[face_index, 0]
        return self.solution[index] if index >= 0 else 0.

    def get_velocity_solution(self):
        """Face fluxes in face order, 0 on Neumann faces (the reference's version is broken,
        mfd.py:1468-1479)."""
        g = self.face_to_flux[:, 0]
        out = np.zeros(len(g))
        out[g >= 0] = self.solution[g[g >= 0]]
        return out

    # ---- analytics (host; mfd.py:1217-1347) ----------------------------------------------------------------
    def get_analytical_pressure_solution(self, p_function, quadrature_method=0):
        pts = self.mesh.get_all_cell_real_centroids() if quadrature_method == 0 else self.mesh.get_all_cell_shifted_centroids()
        return [p_function(p) for p in pts]

    def get_analytical_velocity_solution(self, grad_p):
        m = self.mesh
        return [np.dot(grad_p(m.get_face_real_centroid(f)), m.get_face_normal(f)) for f in range(m.get_number_of_faces())]

    def compute_l2_error_pressure(self, p_function, quadrature_method=0):
        m = self.mesh
        nc = m.get_number_of_cells()
        pts = m.get_all_cell_real_centroids() if quadrature_method == 0 else m.get_all_cell_shifted_centroids()
        exact = np.array([p_function(p) for p in pts])
        # the reference indexes the solution with number_of_faces (mfd.py:1342); identical when
        # there are no Neumann faces
        p = self.solution[m.get_number_of_faces():m.get_number_of_faces() + nc] if self.flux_dof == m.get_number_of_faces() \
            else self.get_pressure_solution()
        return np.sqrt(np.sum(m.cell_volume[:nc] * (p - exact) ** 2))

    def l2_error_velocity_per_cell(self, grad_p):
        """Relative velocity error per cell (mfd.py:1245-1271; the reference's last line takes the square root of
        the whole vector inside the loop -- the per-cell value is what it means)."""
        m = self.mesh
        out = np.zeros(m.get_number_of_cells())
        for c in range(m.get_number_of_cells()):
            num = den = 0.
            for f in m.get_cell(c):
                exact = np.dot(grad_p(m.get_face_real_centroid(f)), m.get_face_normal(f))
                num += m.get_cell_volume(c) * (self.get_velocity_solution_by_index(f) - exact) ** 2
                den += m.get_cell_volume(c) * exact ** 2
            out[c] = np.sqrt(num / den) if den > 0. else 0.
        return out

    def compute_x_error_velocity(self, grad_p):
        """||| grad_p^I - v_h |||_X with the assembled M block (mfd.py:1301-1319); like the reference it addresses
        the solution by face index, i.e. it is meant for systems without Neumann faces."""
        m = self.mesh
        nf = m.get_number_of_faces()
        diff = np.zeros(nf + m.get_number_of_cells())
        for f in range(nf):
            diff[f] = np.dot(grad_p(m.get_face_real_centroid(f)), m.get_face_normal(f)) - self.solution[f]
        lhs = self.lhs.tocsr() if hasattr(self.lhs, "tocsr") else self.lhs
        err = lhs.dot(diff)[:nf].dot(diff[:nf])
        return np.sqrt(err)

    def compute_l2_error_velocity(self, grad_p, quadrature_method=0):
        m = self.mesh
        total = 0.
        for c in range(m.get_number_of_cells()):
            for f in m.get_cell(c):
                pt = m.get_face_real_centroid(f) if quadrature_method == 0 else m.get_face_shifted_centroid(f)
                exact = np.dot(grad_p(pt), m.get_face_normal(f))
                total += m.get_cell_volume(c) * (self.get_velocity_solution_by_index(f) - exact) ** 2
        return np.sqrt(total)
