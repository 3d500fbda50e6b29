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
