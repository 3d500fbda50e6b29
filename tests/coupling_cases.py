"""Pointer-coupling set-up shared by the golden generator (unmodified reference) and the GPU tests.

A 4x3x3 hex box: Dirichlet p = 1 + y on x-, no-flow on y/z, and every x+ boundary face is a Dirichlet
POINTER to the cell at the x- end of its row, which in turn takes the flux through that face as its source
(forcing pointer) -- the fracture/reservoir coupling pattern of mesh.py:1397-1525, mfd.py:776-849."""
import numpy as np

NX, NY, NZ = 4, 3, 3
DIMS = (2., 1.5, 1.)
zero_flux = lambda p: np.array([0., 0., 0.])


def permeability():
    rng = np.random.default_rng(77)
    g = np.exp(0.7 * rng.standard_normal((NX * NY * NZ, 3)))
    k = np.zeros((NX * NY * NZ, 3, 3))
    k[:, 0, 0], k[:, 1, 1], k[:, 2, 2] = g[:, 0], g[:, 1], g[:, 2]
    k[:, 0, 1] = k[:, 1, 0] = 0.2 * np.sqrt(g[:, 0] * g[:, 1])
    return k


def add_pointers(mesh):
    """x+ faces (marker 1) point to the x- cell of the same (j, k) row; returns the pointer faces."""
    faces = []
    for (f, orient) in mesh.get_boundary_faces_by_marker(1):
        cell = int(mesh.face_to_cell[f][0])
        j, k = (cell // NX) % NY, cell // (NX * NY)
        target = 0 + NX * j + NX * NY * k
        mesh.set_dirichlet_face_pointer(f, orient, target)
        mesh.set_forcing_pointer(target, [f], [orient])
        faces.append(f)
    return faces


def configure(mfd_obj, mesh):
    mfd_obj.set_m_e_construction_method(1)
    mfd_obj.set_mesh(mesh)
    mfd_obj.apply_dirichlet_from_function(0, lambda p: 1. + p[1])
    for b in (2, 3, 4, 5):
        mfd_obj.apply_neumann_from_function(b, zero_flux)
    mfd_obj.apply_forcing_from_function(lambda x: 0.3 * x[2])


# ---- multiplier faces: periodic boundaries and Lagrange pointers (mesh.py:1420-1492, mfd.py:806-848) ----------------
def add_multipliers(mesh):
    """The same box, (i) periodic in z: every z- face is tied to the z+ face above it (set_periodic_boundary creates
    the shared multiplier face), (ii) cut into two sub-domains between i = 1 and i = 2: the interior x-face of every
    (j, k) row is doubled -- one copy per side -- and a third copy is the Lagrange multiplier that glues them
    (pressure continuity through face_to_lagrange, flux continuity through lagrange_to_face).  Works on the reference's
    Mesh and on ours (public API only).  Returns (periodic pairs, [(face, new_face, lagrange face)])."""
    lo = {}
    for (f, orient) in mesh.get_boundary_faces_by_marker(4):
        cell = int([c for c in mesh.face_to_cell[f] if c >= 0][0])
        lo[cell % (NX * NY)] = (f, orient)
    pairs = []
    for (f, orient) in mesh.get_boundary_faces_by_marker(5):
        cell = int([c for c in mesh.face_to_cell[f] if c >= 0][0])
        f_lo, o_lo = lo[cell % (NX * NY)]
        mesh.set_periodic_boundary(f_lo, o_lo, f, orient)
        pairs.append((f_lo, f))
    glue = []
    for k in range(NZ):
        for j in range(NY):
            c1, c2 = 1 + NX * j + NX * NY * k, 2 + NX * j + NX * NY * k
            f = [int(x) for x in mesh.get_cell(c1) if x in set(int(y) for y in mesh.get_cell(c2))][0]
            o1 = int(mesh.get_cell_normal_orientation(c1)[list(mesh.get_cell(c1)).index(f)])
            new_face, lag = mesh.duplicate_face(f), mesh.duplicate_face(f)
            for g in (new_face, lag):
                mesh.set_face_normal(g, mesh.get_face_normal(f))
                mesh.set_face_real_centroid(g, mesh.get_face_real_centroid(f))
            faces2 = [int(x) for x in mesh.get_cell(c2)]
            faces2[faces2.index(f)] = new_face
            mesh.remove_from_face_to_cell(f, c2)
            mesh.set_cell_faces(c2, faces2)
            mesh.set_face_to_lagrange_pointer(f, o1, lag)
            mesh.set_face_to_lagrange_pointer(new_face, -o1, lag)
            mesh.set_lagrange_to_face_pointers(lag, [f, new_face], [o1, -o1])
            glue.append((f, new_face, lag))
    return pairs, glue


def configure_multipliers(mfd_obj, mesh):
    mfd_obj.set_m_e_construction_method(1)
    mfd_obj.set_mesh(mesh)
    mfd_obj.apply_dirichlet_from_function(0, lambda p: 1. + p[1])
    mfd_obj.apply_dirichlet_from_function(1, lambda p: 0.25 * p[2])
    for b in (2, 3):
        mfd_obj.apply_neumann_from_function(b, zero_flux)
    mfd_obj.apply_forcing_from_function(lambda x: 0.3 * x[2])


# ---- SinglePhase time loop on a mesh with every kind of pointer (singlephase.py:199-223) ----------------------------
def configure_singlephase(sp, mesh):
    """x+ faces are Dirichlet pointers feeding forcing pointers (add_pointers), z is periodic and the box is glued
    from two sub-domains (add_multipliers); pressure 2 on x-, no-flow on y; slightly compressible fluid, one well."""
    sp.set_mesh(mesh)
    nc = mesh.get_number_of_cells()
    add_pointers(mesh)
    add_multipliers(mesh)
    for b in (2, 3):
        sp.apply_flux_boundary_from_function(b, zero_flux)
    sp.apply_pressure_boundary_from_function(0, lambda p: 2.)
    sp.set_compressibility(1.e-2)
    sp.set_ref_density(1.3)
    sp.set_ref_pressure(1.)
    sp.set_viscosity(0.7)
    sp.set_porosities(np.linspace(.1, .3, nc))
    sp.set_time_step_size(0.5)
    sp.set_number_of_time_steps(1)
    sp.add_point_rate_well(0.4, 17, "w")
