"""The mesh-wide plane cut (Mesh.divide_cells_by_plane, tests/test_cutting.py) through the device path: a HexMesh
with a slanted plane through a whole layer of face-adjacent cells -- tetrahedra to 7-face cells -- assembled and
solved through the drop-in MFD API against the CPU oracle (reference algorithm + SuperLU) on the same mesh.
Run with pytest -m gpu."""
import numpy as np
import pytest

from oracle import mimpy_oracle as orc
from test_gpu_api import lin, spd_tensors, zero_flux

pytestmark = pytest.mark.gpu


def test_many_cell_cut_mesh_on_device():
    import mimpy_b200.mesh.hexmesh as hexmesh
    import mimpy_b200.mfd.mfd as mfd

    n = 4
    mesh = hexmesh.HexMesh()
    mesh.build_mesh(n, n, n, 1., 1., 1., spd_tensors(n ** 3, 23, full=True))
    new = mesh.divide_cells_by_plane(None, np.array([.5, .5, .5]), np.array([.3, .2, 1.]))
    assert len(new) == 22 and mesh.get_number_of_cells() == 86
    # Mirtich kernel on the cut polyhedra == the host integrals made while cutting
    nc = mesh.get_number_of_cells()
    vol, cen = mesh.cell_volume[:nc].copy(), mesh.cell_real_centroid[:nc].copy()
    mesh.find_volume_centroid_all()
    assert np.allclose(mesh.cell_volume[:nc], vol, rtol=1e-12, atol=0)
    assert np.allclose(mesh.cell_real_centroid[:nc], cen, rtol=1e-11, atol=1e-14)
    assert abs(vol.sum() - 1.) < 1e-13
    o = orc.OMesh.from_mesh(mesh)
    for bcs in ("dirichlet", "mixed"):
        f, r = mfd.MFD(), orc.OracleMFD(o, 1)
        f.set_m_e_construction_method(1)
        f.set_mesh(mesh)
        for g in (f, r):
            if bcs == "dirichlet":
                for b in range(6):
                    g.apply_dirichlet_from_function(b, lin)
            else:
                g.apply_dirichlet_from_function(0, lambda p: 1.)
                g.apply_dirichlet_from_function(1, lambda p: 0.)
                for b in (2, 3, 4, 5):
                    g.apply_neumann_from_function(b, zero_flux)
        lhs, rhs = f.build_lhs(), f.build_rhs()
        r.build_lhs()
        r.build_rhs()
        ref_a = r.lhs.tocsr()
        assert abs(lhs - ref_a).max() <= 1e-12 * abs(ref_a).max()
        assert np.array_equal(rhs, r.rhs)
        ref_sol = r.solve()
        sol = f.solve()
        assert np.linalg.norm(sol - ref_sol) <= 1e-8 * np.linalg.norm(ref_sol)
