"""On-disk formats either side of the hot path: the reference's `.mms` text
format (Mesh.load_mesh / save_mesh, mesh.py:584-949) and its legacy-VTK writers
(output_vtk_mesh / output_vtk_faces, mesh.py:1985-2087).

Files written here load in the reference and vice versa (tests/test_cpu_abi_and_host.py).  The
reference writes whole over-allocated arrays (`len(self.points)`, mesh.py:826-828) and sizes the
mesh from those counts on load; this writer emits exactly the live entries, which the reference's
loader reads the same way.  Its VTK writer opens the file in binary mode and then print()s text
(mesh.py:2039, a Python-3 defect); the format produced here is what it intends.
"""
import io

import numpy as np

_HEADER = ("this is just a test", "mimpy_b200", "date", "name", "comments", "#", "#")


def _w(fh, text):
    if isinstance(fh, io.TextIOBase):
        fh.write(text)
    else:
        fh.write(text.encode("ascii"))


def _table(fh, arr, fmt):
    arr = np.asarray(arr)
    if arr.size:
        np.savetxt(fh, arr, fmt=fmt)


def _ragged(fh, va):
    """data count, data, pointer count, pointers (offset, length) of the live entries, compacted."""
    ptr, data = va.compact()
    n = va.number_of_entries
    _w(fh, "%d\n" % len(data))
    _table(fh, data, "%i")
    _w(fh, "%d\n" % n)
    _table(fh, np.stack([ptr[:-1], np.diff(ptr)], axis=1) if n else np.empty((0, 2), int), "%i %i")


def save_mesh(mesh, output_file):
    """mesh.py:811-949.  output_file: a file opened for writing (binary like the reference, or text);
    closed on return like the reference does."""
    fh = output_file
    nf, nc, npnt = mesh.get_number_of_faces(), mesh.get_number_of_cells(), mesh.get_number_of_points()
    for line in _HEADER:
        _w(fh, line + "\n")
    _w(fh, "POINTS %d\n" % npnt)
    _table(fh, mesh.points[:npnt], "%.18e")
    _w(fh, "FACES %d\n" % nf)
    _ragged(fh, mesh.faces)
    _w(fh, "FACE_NORMALS %d\n" % nf)
    _table(fh, mesh.face_normals[:nf], "%.18e")
    _w(fh, "FACE_AREAS %d\n" % nf)
    _table(fh, mesh.face_areas[:nf], "%.18e")
    _w(fh, "FACE_REAL_CENTROIDS %d\n" % nf)
    _table(fh, mesh.face_real_centroids[:nf], "%.18e")
    if mesh.has_face_shifted_centroid:
        _w(fh, "FACE_SHIFTED_CENTROIDS %d\n" % nf)
        _table(fh, mesh.face_shifted_centroids[:nf], "%.18e")
    _w(fh, "FACE_TO_CELL %d\n" % nf)
    _table(fh, mesh.face_to_cell[:nf], "%i %i")
    _w(fh, "CELLS %d\n" % nc)
    _ragged(fh, mesh.cells)
    _w(fh, "CELL_NORMAL_ORIENTATION %d\n" % nc)
    _ragged(fh, mesh.cell_normal_orientation)
    _w(fh, "CELL_VOLUMES %d\n" % nc)
    _table(fh, mesh.cell_volume[:nc], "%.18e")
    _w(fh, "CELL_REAL_CENTROIDS %d\n" % nc)
    _table(fh, mesh.cell_real_centroid[:nc], "%.18e")
    if mesh.has_cell_shifted_centroid:
        _w(fh, "CELL_SHIFTED_CENTROIDS %d\n" % nc)
        _table(fh, mesh.cell_shifted_centroid[:nc], "%.18e")
    _w(fh, "CELL_K %d\n" % nc)
    _table(fh, np.asarray(mesh.cell_k[:nc]).reshape(nc, 9), "%.18e")
    _w(fh, "BOUNDARY_MARKERS %d\n" % len(mesh.boundary_markers))
    for marker in mesh.boundary_markers:
        flat = [int(marker)]
        for face_index, orientation in mesh.get_boundary_faces_by_marker(marker):
            flat += [int(face_index), int(orientation)]
        _w(fh, " ".join(map(str, flat)) + " \n")
    _w(fh, "DIRICHLET_BOUNDARY_POINTERS %d\n" % len(mesh.dirichlet_boundary_pointers))
    for key, (cell_index, orientation) in mesh.dirichlet_boundary_pointers.items():
        _w(fh, "%d %d %d\n" % (key, cell_index, orientation))
    _w(fh, "INTERNAL_NO_FLOW %d\n" % len(mesh.internal_no_flow))
    for face_index, orientation in mesh.internal_no_flow:
        _w(fh, "%d %d\n" % (face_index, orientation))
    _w(fh, "FORCING_FUNCTION_POINTERS %d\n" % len(mesh.forcing_function_pointers))
    for cell_index, pairs in mesh.forcing_function_pointers.items():
        flat = [int(cell_index)]
        for face_index, orientation in pairs:
            flat += [int(face_index), int(orientation)]
        _w(fh, " ".join(map(str, flat)) + " \n")
    _w(fh, "FACE_TO_LAGRANGE_POINTERS %d\n" % len(mesh.face_to_lagrange_pointers))
    for key, (lagrange_index, orientation) in mesh.face_to_lagrange_pointers.items():
        _w(fh, "%d %d %d\n" % (key, lagrange_index, orientation))
    # the reference writes one (face, orientation) per key here (mesh.py:941-947); keys whose
    # list is longer are written one line per pair, which its loader accepts as well
    lines = []
    for key, value in mesh.lagrange_to_face_pointers.items():
        pairs = value if (len(value) and np.ndim(value[0])) else [value]
        lines += ["%d %d %d\n" % (key, f, o) for f, o in pairs]
    _w(fh, "LAGRANGE_TO_FACE_POINTERS %d\n" % len(lines))
    for line in lines:
        _w(fh, line)
    fh.close()


class _Lines(object):
    """Line reader that accepts text or binary files."""

    def __init__(self, fh):
        self.it = iter(fh)

    def __iter__(self):
        return self

    def __next__(self):
        line = next(self.it)
        return line.decode("ascii") if isinstance(line, bytes) else line

    def block(self, n, dtype, width=None):
        rows = [next(self).split() for _ in range(n)]
        if n == 0:
            return np.empty((0,) if not width or width == 1 else (0, width), dtype=dtype)
        arr = np.array(rows, dtype=np.float64).astype(dtype)
        return arr[:, 0] if arr.shape[1] == 1 else arr


def _read_ragged(lines, va, n_entries):
    n_data = int(next(lines))
    data = lines.block(n_data, np.dtype("i"))
    n_ptr = int(next(lines))
    ptrs = lines.block(n_ptr, np.dtype("i"), 2).reshape(n_ptr, 2)
    va.data = np.ascontiguousarray(data, dtype=np.dtype("i")).reshape(-1)
    va.pointers = np.ascontiguousarray(ptrs, dtype=np.dtype("i"))
    va.number_of_entries = n_entries
    va.next_data_pos = n_data
    va.pointer_capacity, va.data_capacity = len(va.pointers), len(va.data)


def load_mesh(mesh, input_file):
    """mesh.py:584-764: same keywords, same semantics (counts in the file size the mesh)."""
    lines = _Lines(input_file)
    for line in lines:  # header lines carry no keyword and are skipped like any unknown line
        t = line.split()
        if not t:
            continue
        key = t[0]
        try:
            n = int(t[1]) if len(t) > 1 else 0
        except ValueError:
            continue  # a header line
        if key == "POINTS":
            mesh.number_of_points = n
            mesh.points = lines.block(n, np.float64, 3).reshape(n, 3)
        elif key == "FACES":
            _read_ragged(lines, mesh.faces, n)
        elif key == "FACE_NORMALS":
            mesh.face_normals = lines.block(n, np.float64, 3).reshape(n, 3)
        elif key == "FACE_AREAS":
            mesh.face_areas = lines.block(n, np.float64).reshape(n)
        elif key == "FACE_REAL_CENTROIDS":
            mesh.face_real_centroids = lines.block(n, np.float64, 3).reshape(n, 3)
        elif key == "FACE_SHIFTED_CENTROIDS":
            mesh.has_face_shifted_centroid = True
            mesh.face_shifted_centroids = lines.block(n, np.float64, 3).reshape(n, 3)
        elif key == "FACE_TO_CELL":
            mesh.face_to_cell = lines.block(n, np.dtype("i"), 2).reshape(n, 2)
        elif key == "CELLS":
            _read_ragged(lines, mesh.cells, n)
        elif key == "CELL_NORMAL_ORIENTATION":
            _read_ragged(lines, mesh.cell_normal_orientation, n)
        elif key == "CELL_VOLUMES":
            mesh.cell_volume = lines.block(n, np.float64).reshape(n)
        elif key == "CELL_REAL_CENTROIDS":
            mesh.cell_real_centroid = lines.block(n, np.float64, 3).reshape(n, 3)
        elif key == "CELL_SHIFTED_CENTROIDS":
            mesh.has_cell_shifted_centroid = True
            mesh.cell_shifted_centroid = lines.block(n, np.float64, 3).reshape(n, 3)
        elif key == "CELL_K":
            mesh.cell_k = lines.block(n, np.float64, 9).reshape(n, 9)
        elif key == "BOUNDARY_MARKERS":
            for _ in range(n):
                e = [int(x) for x in next(lines).split()]
                marker = e[0]
                mesh.add_boundary_marker(marker, "FROMFILE")
                for face_index, orientation in zip(e[1::2], e[2::2]):
                    mesh.add_boundary_face(marker, face_index, orientation)
        elif key == "DIRICHLET_BOUNDARY_POINTERS":
            for _ in range(n):
                k, cell_index, orientation = [int(x) for x in next(lines).split()[:3]]
                mesh.set_dirichlet_face_pointer(k, orientation, cell_index)
        elif key == "INTERNAL_NO_FLOW":
            for _ in range(n):
                face_index, orientation = [int(x) for x in next(lines).split()[:2]]
                mesh.internal_no_flow.append([face_index, orientation])
        elif key == "FORCING_FUNCTION_POINTERS":
            for _ in range(n):
                e = [int(x) for x in next(lines).split()]
                mesh.set_forcing_pointer(e[0], e[1::2], e[2::2])
        elif key == "FACE_TO_LAGRANGE_POINTERS":
            for _ in range(n):
                face_index, lagrange_index, orientation = [int(x) for x in next(lines).split()[:3]]
                mesh.set_face_to_lagrange_pointer(face_index, orientation, lagrange_index)
        elif key == "LAGRANGE_TO_FACE_POINTERS":
            # one (face, orientation) per line; a multiplier that balances several faces has several lines (the
            # reference's loader passes the two integers to a setter that zips them, mesh.py:759-764, 1447-1455)
            pairs = {}
            for _ in range(n):
                lagrange_index, face_index, orientation = [int(x) for x in next(lines).split()[:3]]
                pairs.setdefault(lagrange_index, []).append((face_index, orientation))
            for lagrange_index, entries in pairs.items():
                mesh.set_lagrange_to_face_pointers(lagrange_index, [e[0] for e in entries], [e[1] for e in entries])
    nc = mesh.get_number_of_cells()
    if len(mesh.cell_domain) < nc:
        mesh.cell_domain = np.zeros(nc, dtype=int)
    mesh._touch()


def _vtk_head(out, mesh):
    npnt = mesh.get_number_of_points()
    out.write("# vtk DataFile Version 2.0\n# unstructured mesh\nASCII\nDATASET UNSTRUCTURED_GRID\n")
    out.write("POINTS %d float\n" % npnt)
    for p in mesh.points[:npnt]:
        out.write("%s %s %s\n" % (repr(float(p[0])), repr(float(p[1])), repr(float(p[2]))))


def _vtk_data(out, n, values, labels):
    if len(values):
        out.write("CELL_DATA %d\n" % n)
        for entry, name in zip(values, labels):
            out.write("SCALARS %s double 1\nLOOKUP_TABLE default\n" % name)
            for v in np.asarray(entry)[:n]:
                out.write("%s\n" % repr(float(v)))


def output_vtk_mesh(mesh, file_name, cell_values=(), cell_value_labels=()):
    """Polyhedral cells (VTK type 42) with optional per-cell scalars (mesh.py:2032-2087)."""
    nc = mesh.get_number_of_cells()
    with open(file_name + ".vtk", "w") as out:
        _vtk_head(out, mesh)
        rows, total = [], 0
        for c in range(nc):
            faces = mesh.get_cell(c)
            body = [len(faces)]
            for f in faces:
                pts = mesh.get_face(f)
                body.append(len(pts))
                body.extend(int(p) for p in pts)
            rows.append([len(body)] + body)
            total += len(body) + 1
        out.write("CELLS %d %d\n" % (nc, total))
        for r in rows:
            out.write(" ".join(map(str, r)) + "\n")
        out.write("CELL_TYPES %d\n" % nc)
        out.write("42\n" * nc)
        _vtk_data(out, nc, cell_values, cell_value_labels)


def output_vtk_faces(mesh, file_name, face_indices, face_values=(), face_value_labels=()):
    """The listed faces as polygons (VTK type 7) (mesh.py:1985-2030)."""
    face_indices = list(face_indices)
    with open(file_name + ".vtk", "w") as out:
        _vtk_head(out, mesh)
        total = sum(mesh.get_number_of_face_points(f) + 1 for f in face_indices)
        out.write("CELLS %d %d\n" % (len(face_indices), total))
        for f in face_indices:
            pts = mesh.get_face(f)
            out.write(" ".join(map(str, [len(pts)] + [int(p) for p in pts])) + "\n")
        out.write("CELL_TYPES %d\n" % len(face_indices))
        out.write("7\n" * len(face_indices))
        _vtk_data(out, len(face_indices), face_values, face_value_labels)


def _vtk_vertices(out, points):
    """n points, each a VTK vertex cell (type 1) -- the carrier of the glyph fields below."""
    n = len(points)
    out.write("# vtk DataFile Version 2.0\n# face centroids\nASCII\nDATASET UNSTRUCTURED_GRID\n")
    out.write("POINTS %d float\n" % n)
    for p in points:
        out.write("%s %s %s\n" % (repr(float(p[0])), repr(float(p[1])), repr(float(p[2]))))
    out.write("CELLS %d %d\n" % (n, 2 * n))
    for i in range(n):
        out.write("1 %d\n" % i)
    out.write("CELL_TYPES %d\n" % n)
    out.write("1\n" * n)
    out.write("POINT_DATA %d\n" % n)


def _vtk_vectors(out, name, vectors):
    out.write("VECTORS %s double\n" % name)
    for v in vectors:
        out.write("%s %s %s\n" % (repr(float(v[0])), repr(float(v[1])), repr(float(v[2]))))


def output_vector_field(mesh, file_name, vector_magnitudes=(), vector_labels=()):
    """Face fluxes as vectors for ParaView's glyph filter: one vertex per face centroid carrying
    magnitude x face normal (mesh.py:1892-1942; the reference numbers its vertex cells from 1, which legacy VTK
    reads as an off-by-one connectivity -- 0-based here)."""
    nf = mesh.get_number_of_faces()
    with open(file_name + ".vtk", "w") as out:
        _vtk_vertices(out, mesh.face_real_centroids[:nf])
        for values, name in zip(vector_magnitudes, vector_labels):
            values = np.asarray(values, dtype=float)
            field = np.zeros((nf, 3))
            field[:len(values)] = values[:, None] * mesh.face_normals[:len(values)]
            _vtk_vectors(out, name, field)


def output_cell_normals(mesh, file_name, cell_index):
    """The outward normals (orientation x face normal) of one cell at its face centroids, to check a cell's
    orientation list by eye (mesh.py:1944-1983)."""
    faces = [int(f) for f in mesh.get_cell(cell_index)]
    orientation = np.asarray(mesh.get_cell_normal_orientation(cell_index), dtype=float)
    with open(file_name + ".vtk", "w") as out:
        _vtk_vertices(out, mesh.face_real_centroids[faces])
        _vtk_vectors(out, "OUT_NORMAL", orientation[:, None] * mesh.face_normals[faces])


def save_cell(mesh, cell_index, output_file):
    """One cell as a mesh of its own in the .mms format (mesh.py:766-809): its faces in cell order, its points
    renumbered in order of first use, geometry and K carried over."""
    single = mesh.__class__()
    local = {}
    faces = []
    for f in mesh.get_cell(cell_index):
        points = []
        for p in mesh.get_face(f):
            p = int(p)
            if p not in local:
                local[p] = single.add_point(mesh.get_point(p))
            points.append(local[p])
        g = single.add_face(points)
        single.set_face_area(g, mesh.get_face_area(f))
        single.set_face_normal(g, mesh.get_face_normal(f))
        single.set_face_real_centroid(g, mesh.get_face_real_centroid(f))
        faces.append(g)
    single.add_cell(faces, [int(o) for o in mesh.get_cell_normal_orientation(cell_index)])
    single.set_cell_k(0, mesh.get_cell_k(cell_index))
    single.set_cell_volume(0, mesh.get_cell_volume(cell_index))
    single.set_cell_real_centroid(0, mesh.get_cell_real_centroid(cell_index))
    save_mesh(single, output_file)
