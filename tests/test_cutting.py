"""Mesh.divide_cells_by_plane -- cutting MANY (face-adjacent) cells by one plane.

The reference's divide_cell_by_plane (mesh.py:2227-2419) works on isolated cells only; its fracture code re-implements
the cut with a `done_faces` list (hexmeshwmsfracs.py:725-737, Python 2 only).  There is no reference output to compare
a mesh-wide cut with, so these tests pin it (a) to the reference-equal single-cell cut where both apply (golden
`cut_checker_4`, produced by the unmodified reference), and (b) by what any conforming polyhedral mesh must satisfy:
volumes add up, every cell is closed, every interior face has two cells with opposite orientations, boundary markers
keep their area, and the MFD linear patch test (SURVEY 8c i) is exact on it -- checked with the CPU oracle.  CPU only
(the GPU run of the same mesh is tests/test_gpu_api.py::test_many_cell_cut_mesh_patch_test_on_device)."""
import numpy as np
import pytest

from oracle import mimpy_oracle as orc
from test_cpu_abi_and_host import mesh_from_omesh

lin = lambda p: 1. + 2. * p[0] - p[1] + .5 * p[2]


def conforming(m, total_volume, marker_area):
    """Structural and geometric invariants of a conforming polyhedral mesh; returns its oracle view."""
    o = orc.OMesh.from_mesh(m)
    nc, nf = o.nc, o.nf
    vol = np.array([m.get_cell_volume(c) for c in range(nc)])
    assert vol.min() > 0.
    assert abs(vol.sum() - total_volume) <= 1e-13 * total_volume
    count, osum = np.zeros(nf, int), np.zeros(nf, int)
    for c in range(nc):
        closed = np.zeros(3)
        for f, s in zip(m.get_cell(c), m.get_cell_normal_orientation(c)):
            closed += s * m.get_face_area(f) * m.get_face_normal(f)
            count[f] += 1
            osum[f] += s
        assert abs(closed).max() < 1e-14, "cell %d is not closed" % c
    boundary = {}
    for marker in m.get_boundary_markers():
        area = 0.
        for f, s in m.get_boundary_faces_by_marker(marker):
            assert f not in boundary
            boundary[f] = marker
            area += m.get_face_area(f)
        assert abs(area - marker_area[marker]) < 1e-13
    for f in range(nf):
        if f in boundary:
            assert count[f] == 1
        else:
            assert count[f] == 2 and osum[f] == 0, "interior face %d: %d cells, orientations sum to %d" % (f, count[f], osum[f])
    assert np.array_equal(np.sort(m.face_to_cell[:nf], axis=1), np.sort(orc.rebuild_face_to_cell(o), axis=1))
    # the host volumes are the Mirtich integrals of the faces as stored
    v2, c2 = orc.cell_volumes_centroids(o)
    assert np.allclose(v2, vol, rtol=1e-12, atol=1e-16)
    # no two cells share more than one face (what the symbolic pass of the library requires)
    pairs = set()
    for f in range(nf):
        if count[f] == 2:
            key = tuple(sorted(int(x) for x in m.face_to_cell[f]))
            assert key not in pairs
            pairs.add(key)
    return o


def patch_test(o, methods=(0, 1)):
    kk = np.array([[2., .5, .3], [.5, 1.5, -.2], [.3, -.2, 1.]])
    o.cell_k = np.tile(kk.ravel(), (o.nc, 1))
    for method in methods:
        f = orc.OracleMFD(o, method)
        for b in sorted(o.boundary_faces):
            f.apply_dirichlet_from_function(b, lin)
        f.build_lhs()
        f.build_rhs()
        f.solve()
        exact = np.array([lin(c) for c in o.cell_centroid])
        assert abs(f.get_pressure_solution() - exact).max() < 1e-10


UNIT = {b: 1. for b in range(6)}


def test_slanted_plane_through_a_block_of_cells():
    n = 4
    m = mesh_from_omesh(orc.hex_mesh(n, n, n, 1., 1., 1.))
    new = m.divide_cells_by_plane(None, np.array([.5, .5, .5]), np.array([1.2, .2, 1.3]))
    assert len(new) == 30 and m.get_number_of_cells() == 64 + 30
    assert sorted(new.values()) == list(range(64, 94))
    o = conforming(m, 1., UNIT)
    # every cut cell lies below the plane, its new twin above
    nrm = np.array([1.2, .2, 1.3]) / np.linalg.norm([1.2, .2, 1.3])
    for c, c2 in new.items():
        assert (m.get_cell_real_centroid(c) - .5).dot(nrm) < 0. < (m.get_cell_real_centroid(c2) - .5).dot(nrm)
        assert np.array_equal(m.get_cell_k(c), m.get_cell_k(c2))
    # one new point per crossed EDGE: shared by the faces and cells around it
    pts = m.points[:m.get_number_of_points()]
    assert len(np.unique(np.round(pts, 12), axis=0)) == len(pts)
    patch_test(o)


def test_planes_through_vertices_edges_and_faces():
    """Planes that contain mesh vertices, edges or whole faces: no point is created next to an existing vertex,
    a cell that is only touched stays whole, repeated cuts of an already cut mesh stay conforming."""
    n = 4
    m = mesh_from_omesh(orc.hex_mesh(n, n, n, 1., 1., 1.))
    npts = m.get_number_of_points()
    assert m.divide_cells_by_plane(None, np.array([.5, .5, .5]), np.array([1., 0., 0.])) == {}   # a grid plane
    assert m.get_number_of_cells() == 64 and m.get_number_of_faces() == 240 and m.get_number_of_points() == npts
    assert len(m.divide_cells_by_plane(None, np.array([.375, .5, .5]), np.array([1., 0., 0.]))) == 16
    conforming(m, 1., UNIT)
    assert len(m.divide_cells_by_plane(None, np.array([.5, .5, .5]), np.array([1., 1., 0.]))) == 20   # through edges
    conforming(m, 1., UNIT)
    assert len(m.divide_cells_by_plane(None, np.array([.5, .5, .5]), np.array([1., 1., 1.]))) == 31   # through vertices
    o = conforming(m, 1., UNIT)
    assert m.get_number_of_cells() == 64 + 16 + 20 + 31
    patch_test(o, methods=(1,))


def test_subset_of_cells_keeps_the_neighbours_conforming():
    """Only some of the crossed cells are divided: their uncut neighbours receive the split faces."""
    n = 3
    m = mesh_from_omesh(orc.hex_mesh(n, n, n, 3., 3., 3.))
    cells = [13, 14, 16]   # the centre cell and two face neighbours
    new = m.divide_cells_by_plane(cells, np.array([1.5, 1.5, 1.4]), np.array([.1, .2, 1.]))
    assert sorted(new) == cells and m.get_number_of_cells() == 30
    o = conforming(m, 27., {b: 9. for b in range(6)})
    assert len(m.get_cell(12)) == 7 and len(m.get_cell(17)) == 8   # uncut neighbours of one / of two divided cells
    patch_test(o, methods=(1,))
    with pytest.raises(ValueError):
        m.divide_cells_by_plane([1, 1], np.array([1.5, 1.5, 1.4]), np.array([.1, .2, 1.]))


def test_isolated_cells_equal_the_reference_cut(golden):
    """On isolated cells both routines apply: cell by cell the two halves are the reference's (golden cut_checker_4:
    every other cell of 4^3 cut through its centroid by the unmodified reference)."""
    g = golden("cut_checker_4")
    ref = orc.omesh_from_dict(g)
    n = 4
    m = mesh_from_omesh(orc.hex_mesh(n, n, n, 1., 1., 1.))
    nrm = np.array([1.2, .2, 1.3])
    nrm /= np.linalg.norm(nrm)
    for c in range(n ** 3):
        i, j, k = c % n, (c // n) % n, c // (n * n)
        if (i + j + k) % 2 == 0:
            assert list(m.divide_cells_by_plane([c], np.array(m.get_cell_real_centroid(c)), nrm)) == [c]
    o = conforming(m, 1., UNIT)
    assert o.nc == ref.nc and o.nf == ref.nf
    assert np.allclose(o.cell_volume, ref.cell_volume, rtol=1e-12, atol=0)
    assert np.allclose(o.cell_centroid, ref.cell_centroid, rtol=1e-11, atol=1e-14)
    assert np.array_equal(np.diff(o.cell_ptr), np.diff(ref.cell_ptr))   # same number of faces, cell by cell
    # the reference creates one point per (face, crossed edge), this routine one per crossed edge
    assert m.get_number_of_points() < len(ref.points)


def test_reference_routine_fails_on_adjacent_cells():
    """What divide_cells_by_plane is for: the second of two face-adjacent single-cell cuts breaks down in the
    reference's algorithm (which divide_cell_by_plane restates faithfully), the mesh-wide cut does not."""
    m = mesh_from_omesh(orc.hex_mesh(2, 1, 1, 2., 1., 1.))
    plane = (np.array([1., .5, .5]), np.array([.1, .2, 1.]))
    m.divide_cell_by_plane(0, *plane)
    with pytest.raises(Exception):
        m.divide_cell_by_plane(1, *plane)
    m = mesh_from_omesh(orc.hex_mesh(2, 1, 1, 2., 1., 1.))
    assert m.divide_cells_by_plane([0, 1], *plane) == {0: 2, 1: 3}
    conforming(m, 2., {0: 1., 1: 1., 2: 2., 3: 2., 4: 2., 5: 2.})


def test_one_plane_per_cell_is_the_checkerboard_generator(golden):
    """point_on_plane with one row per cell: parallel planes, every listed cell cut through its own point -- BASELINE
    config 5's generator in one call (bench.py:run_c5).  Same halves as the reference's mesh; face-sharing cells refused."""
    g = golden("cut_checker_4")
    ref = orc.omesh_from_dict(g)
    n = 4
    m = mesh_from_omesh(orc.hex_mesh(n, n, n, 1., 1., 1.))
    nrm = np.array([1.2, .2, 1.3])
    cells = [c for c in range(n ** 3) if (c % n + (c // n) % n + c // (n * n)) % 2 == 0]
    points = np.array([m.get_cell_real_centroid(c) for c in cells])
    new = m.divide_cells_by_plane(cells, points, nrm)
    assert list(new) == cells and list(new.values()) == list(range(64, 64 + len(cells)))
    o = conforming(m, 1., UNIT)
    assert o.nc == ref.nc and o.nf == ref.nf
    assert np.allclose(o.cell_volume, ref.cell_volume, rtol=1e-12, atol=0)
    assert np.allclose(o.cell_centroid, ref.cell_centroid, rtol=1e-11, atol=1e-14)
    assert np.array_equal(np.diff(o.cell_ptr), np.diff(ref.cell_ptr))
    patch_test(o, methods=(1,))
    m = mesh_from_omesh(orc.hex_mesh(2, 1, 1, 2., 1., 1.))
    with pytest.raises(ValueError):
        m.divide_cells_by_plane([0, 1], np.array([[.5, .5, .5], [1.5, .5, .4]]), nrm)
    with pytest.raises(ValueError):
        m.divide_cells_by_plane([0], np.array([[.5, .5, .5], [1.5, .5, .4]]), nrm)


def test_random_and_degenerate_planes_keep_the_mesh_conforming():
    """Seeded stress: up to three successive cuts of a 3^3 box by planes in general position, through grid vertices,
    along grid planes and through grid edges -- every result is a conforming mesh of unchanged volume."""
    rng = np.random.default_rng(1)
    o0 = orc.hex_mesh(3, 3, 3, 1., 1., 1.)
    for trial in range(60):
        m = mesh_from_omesh(o0)
        for _ in range(int(rng.integers(1, 4))):
            kind = int(rng.integers(0, 4))
            if kind == 0:
                p, n = rng.random(3), rng.standard_normal(3)
            elif kind == 1:
                p, n = rng.integers(0, 4, 3) / 3., rng.standard_normal(3)
            elif kind == 2:
                p, n = rng.integers(0, 7, 3) / 6., np.eye(3)[int(rng.integers(0, 3))]
            else:
                p, n = rng.integers(0, 4, 3) / 3., rng.standard_normal(3)
                n[int(rng.integers(0, 3))] = 0.
            m.divide_cells_by_plane(None, p, n)
        conforming(m, 1., UNIT)


def test_cells_crossed_by_a_bounded_piece_of_the_plane():
    """find_cells_crossed_by_plane with an `inside` predicate (a disc-shaped fracture): the same cells as the
    reference's edge-by-edge selection loop (hexmeshwmsfracs.py:683-713) restated per cell, and cutting exactly those
    cells leaves a conforming mesh whose uncut neighbours carry the split faces."""
    n = 5
    m = mesh_from_omesh(orc.hex_mesh(n, n, n, 1., 1., 1.))
    p0, nrm = np.array([.5, .5, .45]), np.array([.3, .2, 1.])
    nrm /= np.linalg.norm(nrm)
    disc = lambda x: np.linalg.norm(x - p0, axis=1) < .3
    cells = m.find_cells_crossed_by_plane(p0, nrm, inside=disc)
    expected = []
    for c in range(n ** 3):
        hit = False
        for f in m.get_cell(c):
            loop = [int(q) for q in m.get_face(f)]
            for a, b in zip(loop, loop[1:] + loop[:1]):
                da, db = (m.get_point(a) - p0).dot(nrm), (m.get_point(b) - p0).dot(nrm)
                if da * db < 0. and min(abs(da), abs(db)) > 1e-8:
                    x = m.get_point(a) + da / (da - db) * (m.get_point(b) - m.get_point(a))
                    hit = hit or bool(disc(x[None, :])[0])
        if hit:
            expected.append(c)
    assert cells == expected and 0 < len(cells) < len(m.find_cells_crossed_by_plane(p0, nrm)) <= n * n * 2
    new = m.divide_cells_by_plane(cells, p0, nrm)
    assert sorted(new) == cells
    conforming(m, 1., UNIT)
    assert m.find_cells_crossed_by_plane(p0, nrm, inside=disc) == []      # nothing left to cut inside the disc
