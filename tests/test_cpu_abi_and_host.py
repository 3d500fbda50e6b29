"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the
ctypes structures match the C layouts, and the host-side mesh logic (ragged arrays, plane cutting,
planar polygon geometry) reproduces the reference.  No compute call is made without a GPU."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import mimpy_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mimpy_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from mimpy_b200 import build

    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    hdr = open(HEADER).read()
    names = sorted(set(re.findall(r"\b(mimpy_b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.mimpy_b200_version() >= 100


def test_ctypes_signatures_cover_the_header(lib_path):
    from mimpy_b200 import _lib

    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    names = set(re.findall(r"\b(mimpy_b200_[a-z0-9_]+)\s*\(", hdr)) - {"mimpy_b200_last_error"}
    undeclared = [n for n in sorted(names) if n not in _lib.SIGNATURES]
    assert not undeclared, undeclared
    # argument counts agree with the prototypes
    for name in names:
        m = re.search(r"int\s+" + name + r"\s*\(([^;]*?)\)\s*;", hdr, re.S)
        assert m, name
        args = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(args) == len(_lib.SIGNATURES[name]), name


def test_struct_layouts_match_c(lib_path):
    """sizeof/offsetof of every struct that crosses the boundary, from a C program built with gcc."""
    from mimpy_b200 import _lib

    structs = {"mimpy_mesh_view": _lib.MeshView, "mimpy_symbolic": _lib.Symbolic,
               "mimpy_pcg_info": _lib.PcgInfo, "mimpy_fluid": _lib.Fluid}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mimpy_b200.h"', 'int main(void){']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append('return 0;}')
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "layout.c")
        open(src, "w").write("\n".join(lines))
        exe = os.path.join(tmp, "layout")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_product_has_no_cpu_fallback_and_never_imports_the_oracle():
    """The product path must fail loudly without CUDA and must not route through oracle/."""
    import torch
    from mimpy_b200 import _lib

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _lib.Context()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mimpy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "scipy.sparse.linalg" not in text and "spsolve" not in text.replace("'spsolve'", "").replace("('spsolve')", "").replace("solve('spsolve')", ""), f


# ---- host mesh logic -----------------------------------------------------------------------------
def test_variable_array_semantics():
    from mimpy_b200.mesh.mesh import variable_array

    a = variable_array(dtype=np.dtype('i'))
    idx = [a.add_entry(list(range(k, k + 3 + k % 4))) for k in range(2500)]
    assert idx == list(range(2500)) and len(a) == 2500
    assert list(a[7]) == list(range(7, 7 + 3 + 3))
    a[7] = [1, 2]                      # shorter: in place
    assert list(a[7]) == [1, 2]
    a[8] = list(range(40))             # longer: re-appended
    assert list(a[8]) == list(range(40)) and list(a[9]) == list(range(9, 9 + 3 + 1))
    ptr, data = a.compact()
    assert ptr[-1] == len(data) and list(data[ptr[8]:ptr[9]]) == list(range(40))
    with pytest.raises(IndexError):
        a[2500]


def mesh_from_omesh(o):
    """Host-only construction through the public Mesh API (no GPU)."""
    from mimpy_b200.mesh.mesh import Mesh

    m = Mesh()
    for p in o.points:
        m.add_point(p)
    for f in range(o.nf):
        m.add_face(list(o.face(f)))
        m.set_face_area(f, o.face_areas[f])
        m.set_face_normal(f, o.face_normals[f])
        m.set_face_real_centroid(f, o.face_centroids[f])
    m.set_boundary_markers(sorted(o.boundary_faces), ["b%d" % b for b in sorted(o.boundary_faces)])
    for b, lst in o.boundary_faces.items():
        for f, s in lst:
            m.add_boundary_face(b, f, s)
    for c in range(o.nc):
        m.add_cell(list(o.cell(c)), list(o.sigma(c)))
        m.set_cell_volume(c, o.cell_volume[c])
        m.set_cell_real_centroid(c, o.cell_centroid[c])
        m.set_cell_k(c, o.cell_k[c].reshape(3, 3))
    return m


def test_mesh_roundtrip_and_cutting_reproduce_reference(golden):
    g = golden("cut_checker_4")
    ref = orc.omesh_from_dict(g)
    rng = np.random.default_rng(5)
    n = 4
    ktab = np.zeros((n ** 3, 3, 3))
    for c in range(n ** 3):  # same generator as tests/golden/gen_golden.py:spd_tensors(64, 5)
        gg = rng.standard_normal(3)
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        ktab[c] = q @ np.diag(np.exp(gg)) @ q.T
    m = mesh_from_omesh(orc.hex_mesh(n, n, n, 1., 1., 1., k_array=ktab))
    assert m.get_number_of_cells() == 64 and m.get_number_of_faces() == 240
    nrm = np.array([1.2, .2, 1.3])
    nrm /= np.linalg.norm(nrm)
    for c in range(n ** 3):
        i, j, k = c % n, (c // n) % n, c // (n * n)
        if (i + j + k) % 2 == 0:
            m.divide_cell_by_plane(c, np.array(m.get_cell_real_centroid(c)), nrm)
    o = orc.OMesh.from_mesh(m)
    for key in ("face_ptr", "face_points", "cell_ptr", "cell_faces", "cell_sigma"):
        assert np.array_equal(getattr(o, key), getattr(ref, key)), key
    assert {b: [(int(a), int(s)) for a, s in v] for b, v in o.boundary_faces.items()} == ref.boundary_faces
    assert np.allclose(o.points, ref.points, rtol=1e-14, atol=1e-16)
    assert np.allclose(o.face_areas, ref.face_areas, rtol=1e-12, atol=0)
    assert np.allclose(o.face_centroids, ref.face_centroids, rtol=1e-12, atol=1e-15)
    assert np.allclose(o.face_normals, ref.face_normals, rtol=1e-12, atol=1e-15)
    assert np.allclose(o.cell_volume, ref.cell_volume, rtol=1e-12, atol=0)
    assert np.allclose(o.cell_centroid, ref.cell_centroid, rtol=1e-11, atol=1e-14)
    assert np.array_equal(o.cell_k, ref.cell_k)
    # unlike the reference (mesh.py:2405-2410) the cut face keeps BOTH of its cells
    assert np.array_equal(np.sort(m.face_to_cell[:o.nf], axis=1), np.sort(orc.rebuild_face_to_cell(o), axis=1))


def test_hexmesh_closed_form_tables_match_reference(golden):
    """HexMesh's closed-form face numbering / boundary lists (host part, no GPU needed)."""
    from mimpy_b200.mesh.hexmesh import HexMesh

    for name, dims in (("hex_c1_n5", (5, 5, 5)), ("hex_aniso_432", (4, 3, 2)), ("flow1d_322", (3, 2, 2))):
        ref = orc.omesh_from_dict(golden(name))
        h = HexMesh()
        h.ni, h.nj, h.nk = dims[0] + 1, dims[1] + 1, dims[2] + 1
        got = {m: [(f, s) for f, s in v] for m, v in h._boundary_lists().items()}
        assert got == ref.boundary_faces
        zf, yf, xf = h._face_tables()
        cx, cy, cz = dims
        c = np.arange(cx * cy * cz)
        i, j, k = c % cx, (c // cx) % cy, c // (cx * cy)
        cells = np.stack([zf(i, j, k), yf(i, j, k), xf(i, j, k), xf(i + 1, j, k), yf(i, j + 1, k), zf(i, j, k + 1)], 1)
        assert np.array_equal(cells.ravel(), ref.cell_faces)
        assert h.ijk_to_cell_index(1, 1, 1) == 1 + cx + cx * cy


# ---- on-disk formats -------------------------------------------------------------
def _assert_same_mesh(o, ref, exact=True):
    for key in ("face_ptr", "face_points", "cell_ptr", "cell_faces", "cell_sigma"):
        assert np.array_equal(getattr(o, key), getattr(ref, key)), key
    assert {b: [(int(a), int(s)) for a, s in v] for b, v in o.boundary_faces.items()} == \
        {b: [(int(a), int(s)) for a, s in v] for b, v in ref.boundary_faces.items()}
    assert [tuple(map(int, x)) for x in o.internal_no_flow] == [tuple(map(int, x)) for x in ref.internal_no_flow]
    for key in ("points", "face_areas", "face_centroids", "face_normals", "cell_volume", "cell_centroid", "cell_k"):
        a, b = getattr(o, key), getattr(ref, key)
        assert np.array_equal(a, b) if exact else np.allclose(a, b, rtol=1e-15, atol=0), key


def test_mms_file_written_by_the_reference_loads(golden):
    """Mesh.load_mesh reads a `.mms` file produced by the UNMODIFIED reference's save_mesh
    (tests/golden/ref_mms_cut_232.mms.gz: a 2x3x2 hex mesh with one cell cut by a plane, an internal
    no-flow face and a forcing pointer) into exactly the arrays the reference held."""
    import gzip
    from mimpy_b200.mesh.mesh import Mesh

    g = golden("mms_cut_232")
    ref = orc.omesh_from_dict(g)
    m = Mesh()
    with gzip.open(os.path.join(os.path.dirname(__file__), "golden", "ref_mms_cut_232.mms.gz"), "rt") as fh:
        m.load_mesh(fh)
    assert m.get_number_of_cells() == ref.nc == 13 and m.get_number_of_faces() == ref.nf
    # the reference writes its whole over-allocated point table (mesh.py:826-828) and sizes the mesh
    # from that count on load (mesh.py:601-603); the live points come first
    assert m.get_number_of_points() >= int(g["n_points"])
    m.number_of_points = int(g["n_points"])
    _assert_same_mesh(orc.OMesh.from_mesh(m), ref, exact=False)   # text round trip: %.18e is exact, str(area) too
    assert m.get_forcing_pointers_for_cell(2) == [(int(m.get_cell(7)[0]), 1)] or \
        [tuple(x) for x in m.get_forcing_pointers_for_cell(2)] == [(int(m.get_cell(7)[0]), 1)]


@pytest.mark.parametrize("mode", ["wb", "w"])
def test_mms_save_load_round_trip(golden, mode, tmp_path):
    """save_mesh -> load_mesh is the identity on a cut polyhedral mesh (binary handle like the
    reference's own test, tests/test_read_write.py:18-26, or a text handle), counts included."""
    from mimpy_b200.mesh.mesh import Mesh

    ref = orc.omesh_from_dict(golden("cut_checker_4"))
    m = mesh_from_omesh(ref)
    m.add_internal_no_flow(int(m.get_cell(3)[2]), -1)
    m.set_dirichlet_face_pointer(5, 1, 2)
    path = str(tmp_path / "mesh.mms")
    m.save_mesh(open(path, mode))
    m2 = Mesh()
    m2.load_mesh(open(path))
    assert m2.get_number_of_faces() == m.get_number_of_faces()      # the reference's own assertions
    assert m2.get_number_of_cells() == m.get_number_of_cells()
    assert m2.get_number_of_points() == m.get_number_of_points()
    _assert_same_mesh(orc.OMesh.from_mesh(m2), orc.OMesh.from_mesh(m), exact=True)
    assert m2.get_dirichlet_pointer(5) == m.get_dirichlet_pointer(5)
    assert np.array_equal(m2.face_to_cell[:m.get_number_of_faces()], m.face_to_cell[:m.get_number_of_faces()])


@pytest.mark.skipif(not os.path.isdir("/root/reference/mimpy"), reason="needs the reference tree (build container only)")
def test_mms_file_written_here_loads_in_the_reference(golden, tmp_path):
    from oracle import ref_shim

    ref = ref_shim.load_reference()
    o = orc.omesh_from_dict(golden("cut_checker_4"))
    m = mesh_from_omesh(o)
    path = str(tmp_path / "ours.mms")
    m.save_mesh(open(path, "wb"))
    r = ref.mesh.Mesh()
    r.load_mesh(open(path))
    assert r.get_number_of_cells() == o.nc and r.get_number_of_faces() == o.nf
    _assert_same_mesh(orc.OMesh.from_mesh(r), o, exact=True)


def test_vtk_writers(golden, tmp_path):
    """Legacy-VTK output (mesh.py:1985-2087): polyhedral cells (type 42) and face polygons (type 7),
    connectivity counts consistent, scalars attached."""
    o = orc.omesh_from_dict(golden("cut_checker_4"))
    m = mesh_from_omesh(o)
    base = str(tmp_path / "out")
    p = np.arange(o.nc, dtype=float)
    m.output_vtk_mesh(base, [p, 2 * p], ["pressure", "twice"])
    txt = open(base + ".vtk").read().split("\n")
    assert txt[0] == "# vtk DataFile Version 2.0" and txt[3] == "DATASET UNSTRUCTURED_GRID"
    i = next(k for k, l in enumerate(txt) if l.startswith("CELLS "))
    ncell, total = map(int, txt[i].split()[1:])
    rows = [list(map(int, l.split())) for l in txt[i + 1:i + 1 + ncell]]
    assert ncell == o.nc and sum(len(r) for r in rows) == total and all(r[0] == len(r) - 1 for r in rows)
    assert all(r[1] == o.cell_ptr[c + 1] - o.cell_ptr[c] for c, r in enumerate(rows))
    assert txt[i + 1 + ncell] == "CELL_TYPES %d" % o.nc and txt[i + 2 + ncell] == "42"
    j = txt.index("SCALARS twice double 1")
    assert [float(x) for x in txt[j + 2:j + 2 + o.nc]] == list(2 * p)
    faces = [0, 5, 17]
    m.output_vtk_faces(base + "_f", faces, [np.array([1., 2., 3.])], ["v"])
    t2 = open(base + "_f.vtk").read().split("\n")
    i = next(k for k, l in enumerate(t2) if l.startswith("CELLS "))
    assert t2[i] == "CELLS 3 %d" % sum(len(o.face(f)) + 1 for f in faces)
    assert t2[i + 1].split() == [str(len(o.face(0)))] + [str(int(q)) for q in o.face(0)]
    assert t2[i + 4:i + 8] == ["CELL_TYPES 3", "7", "7", "7"]
    # glyph fields (mesh.py:1892-1983): face fluxes as magnitude x normal at the face centroids; a cell's outward normals
    flux = np.linspace(-1., 2., o.nf)
    m.output_vector_field(base + "_v", [flux], ["flux"])
    t3 = open(base + "_v.vtk").read().split("\n")
    assert t3[4] == "POINTS %d float" % o.nf and [float(x) for x in t3[5 + 7].split()] == list(o.face_centroids[7])
    assert t3[5 + o.nf] == "CELLS %d %d" % (o.nf, 2 * o.nf) and t3[6 + o.nf] == "1 0"
    j = t3.index("VECTORS flux double")
    assert t3[j - 1] == "POINT_DATA %d" % o.nf
    assert np.allclose([float(x) for x in t3[j + 1 + 11].split()], flux[11] * o.face_normals[11], rtol=0, atol=0)
    m.output_cell_normals(base + "_n", 3)
    t4 = open(base + "_n.vtk").read().split("\n")
    n3 = len(o.cell(3))
    j = t4.index("VECTORS OUT_NORMAL double")
    got = np.array([[float(x) for x in l.split()] for l in t4[j + 1:j + 1 + n3]])
    assert np.array_equal(got, o.sigma(3)[:, None] * o.face_normals[o.cell(3)])
    area_vec = (got * o.face_areas[o.cell(3)][:, None]).sum(axis=0)
    assert abs(area_vec).max() < 1e-14   # outward area vectors of a closed cell cancel


def test_voromesh_reads_voro_file_like_reference(golden):
    """VoroMesh.build_mesh is host code: the mesh read from the voro++-format fixture equals the reference's
    (voromesh.py:19-201) point by point, face by face, cell by cell -- no GPU involved."""
    import os
    from oracle import mimpy_oracle as orc
    import mimpy_b200.mesh.voromesh as voromesh

    g = golden("voro_48")
    ref = orc.omesh_from_dict(g)
    mesh = voromesh.VoroMesh()
    mesh.build_mesh(os.path.join(os.path.dirname(__file__), "golden", "voro_48.vol.gz"),
                    lambda p: np.diag([1. + p[0], 2. - p[1], 1. + .5 * p[2]]))
    o = orc.OMesh.from_mesh(mesh)
    for k in ("points", "face_ptr", "face_points", "cell_ptr", "cell_faces", "cell_sigma", "face_normals", "face_areas",
              "face_to_cell", "cell_volume", "cell_centroid", "cell_k"):
        assert np.array_equal(getattr(o, k), getattr(ref, k)), k
    assert np.allclose(o.face_centroids, ref.face_centroids, rtol=0, atol=1e-14)
    assert np.allclose(o.face_shifted_centroids, g["face_shifted_centroids"], rtol=0, atol=1e-14)   # (boundary faces: find_face_centroid)
    assert np.array_equal(o.cell_shifted_centroids, g["cell_shifted_centroids"])
    assert {m: [(int(a), int(b)) for a, b in v] for m, v in o.boundary_faces.items()} == ref.boundary_faces
    assert abs(o.cell_volume.sum() - 3.0) < 1e-12


def test_mesh_helpers_and_mfd_diagnostics_match_reference(golden, tmp_path):
    """Host-side helpers next to the path -- TPFA face shift (mesh.py:510-527), segment / face intersection
    (:529-576), projected centroid (:1604-1637), domain hull (:2152-2184) -- against the unmodified reference on a
    distorted 4x3x2 box; the dense diagnostics of MFD (mfd.py:1481-1529) on a matrix taken from the golden mesh."""
    import mimpy_b200.mfd.mfd as mfd
    from scipy import sparse

    g = golden("mesh_helpers_432")
    o = orc.omesh_from_dict(g)
    m = mesh_from_omesh(o)
    nf = m.get_number_of_faces()
    m.set_face_shifted_to_tpfa_all()
    assert m.is_using_face_shifted_centroid()
    assert np.allclose(m.face_shifted_centroids[:nf], g["tpfa_shifted"], rtol=1e-13, atol=1e-15)
    interior = (m.face_to_cell[:nf] >= 0).all(axis=1)
    assert np.array_equal(m.face_shifted_centroids[:nf][~interior], m.face_real_centroids[:nf][~interior])
    for a, (p1, p2) in enumerate(g["segments"]):
        for f in range(nf):
            r = m.is_line_seg_intersect_face(f, p1, p2)
            assert (-1 if r is None else int(r)) == g["segment_hits"][a, f], (a, f)
    assert (g["segment_hits"] == 1).sum() >= 4 and (g["segment_hits"] == 0).sum() >= 4
    for row, f in zip(g["projected"], (0, 5, 17, 40)):
        for ref3, co in zip(row, ((0, 1), (1, 2), (0, 2))):
            assert np.allclose(m.find_centroid_for_coordinates(f, co), ref3, rtol=1e-12, atol=1e-15)
    for c in (5, 6, 9, 10, 21):
        m.set_cell_domain(c, 1)
    faces, orients = m.find_domain_faces(1)
    assert np.array_equal(faces, g["domain_faces"]) and np.array_equal(orients, g["domain_orientations"])
    # MFD diagnostics: any small matrix will do -- the oracle's saddle matrix of the same problem
    r = orc.OracleMFD(o, 1)
    for b in (0, 1):
        r.apply_dirichlet_from_function(b, lambda p: 1. + p[0])
    r.build_lhs()
    f = mfd.MFD()
    f.lhs = r.lhs.tocsr()
    assert np.allclose(f.lhs_svd(), g["lhs_svd"], rtol=1e-9, atol=1e-12) and f.lhs_rank() == int(g["lhs_rank"][0])
    f.lhs = sparse.csr_matrix(np.array([[1.5, 0.], [1e-12, -2.]]))
    f.print_lhs_to_file(str(tmp_path / "a"))
    f.print_lhs_to_ppm(str(tmp_path / "a"))
    assert open(str(tmp_path / "a.dat")).read().split() == ["1.5", ",", "0.0", ",", "1e-12", ",", "-2.0", ","]
    assert open(str(tmp_path / "a.ppm")).read().split() == ["P1", "2", "2", "1", "0", "0", "1"]


def test_brick_kernels_stage_with_bulk_copies(lib_path):
    """SASS census of the built library (scripts/sass_census.py, profiles/r2_sass_bulk_copy.txt): the headline assembly
    kernel and the fused set-up kernel move their inputs with cp.async.bulk (UBLKCP) through mbarriers (SYNCS) and use
    setmaxnreg (USETMAXREG); the assembly kernel has no Ampere-style cp.async (LDGSTS) left."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sass_census

    counts = sass_census.census(lib_path)
    brick3 = [k for k in counts if "k_numeric_brick3" in k]
    setup = [k for k in counts if "k_hybrid_setup_brick" in k]
    assert len(brick3) == 3 and len(setup) == 6      # methods 0 / 1 / 6 (x CSR / SELL row stride for the set-up)
    for k in brick3 + setup:
        assert counts[k]["UBLKCP"] >= 6 and counts[k]["SYNCS"] >= 6 and counts[k]["USETMAXREG"] == 2, (k, counts[k])
    assert all(counts[k]["LDGSTS"] == 0 for k in brick3)


def test_save_cell_round_trip_and_reference_interchange(golden, tmp_path):
    """Mesh.save_cell (mesh.py:766-809): one cut cell written as a mesh of its own and read back -- faces in cell
    order, points renumbered by first use, geometry and K carried over; with the reference tree present, the same
    cell saved by the UNMODIFIED reference loads into the same arrays."""
    from mimpy_b200.mesh.mesh import Mesh

    o = orc.omesh_from_dict(golden("cut_checker_4"))
    m = mesh_from_omesh(o)
    c = int(np.argmax(np.diff(o.cell_ptr)))          # a cell with the most faces
    path = str(tmp_path / "cell.mms")
    m.save_cell(c, open(path, "w"))
    one = Mesh()
    one.load_mesh(open(path))
    faces = o.cell(c)
    assert one.get_number_of_cells() == 1 and one.get_number_of_faces() == len(faces)
    assert one.get_number_of_points() == len(np.unique(np.concatenate([o.face(f) for f in faces])))
    assert list(one.get_cell(0)) == list(range(len(faces)))
    assert np.array_equal(one.get_cell_normal_orientation(0), o.sigma(c))
    for g, f in enumerate(faces):
        assert np.array_equal(one.points[one.get_face(g)], o.points[o.face(f)])
        assert one.get_face_area(g) == o.face_areas[f] and np.array_equal(one.get_face_normal(g), o.face_normals[f])
    assert one.get_cell_volume(0) == o.cell_volume[c] and np.array_equal(one.get_cell_k(0).ravel(), o.cell_k[c])
    volume, centroid = one.find_volume_centroid(0)
    assert abs(volume - o.cell_volume[c]) <= 1e-14 and np.allclose(centroid, o.cell_centroid[c], rtol=0, atol=1e-13)
    if os.path.isdir("/root/reference/mimpy"):
        from oracle import ref_shim

        ref = ref_shim.load_reference()
        r = ref.mesh.Mesh()
        r.load_mesh(open(str(_save_full(m, tmp_path))))
        ref_path = str(tmp_path / "cell_ref.mms")
        r.save_cell(c, open(ref_path, "wb"))
        theirs = Mesh()
        theirs.load_mesh(open(ref_path))
        theirs.number_of_points = one.get_number_of_points()   # the reference writes its over-allocated point table
        _assert_same_mesh(orc.OMesh.from_mesh(theirs), orc.OMesh.from_mesh(one), exact=True)


def _save_full(m, tmp_path):
    path = tmp_path / "full.mms"
    m.save_mesh(open(str(path), "w"))
    return path
