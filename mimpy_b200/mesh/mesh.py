"""Host-side mesh data model with the public API of mimpy.mesh.mesh (reference mesh.py:20-503,
951-1573) for the part of it that the MFD hot path reads, plus the exporter that gathers it into
SoA device arrays (mimpy_b200.device.DeviceMesh).

The mesh stays a mutable host object (points / ragged faces / ragged cells, exactly the arrays the
reference exposes); everything numeric that the reference does in native code on it runs on the
GPU through libmimpy_b200:

    find_volume_centroid_all   -> mimpy_b200_geom_cells        (mesh_cython.pyx:5-207)
    HexMesh face geometry      -> mimpy_b200_geom_hex_faces    (hexmesh_cython.pyx:30-118)

.mms I/O and VTK writers: meshio.py; VoroMesh: voromesh.py.  Out of scope: fracture extrusion.
"""
import numpy as np


class variable_array(object):
    """Ragged 2-D array: flat `data` + `pointers[n, 2] = [offset, length]` (mesh.py:20-183).

    Same public attributes and methods as the reference class; growth is geometric so that
    appends are amortised O(1)."""

    def __init__(self, dtype=np.dtype('d'), size=(2, 2), dim=1):
        self.pointer_capacity = int(size[0])
        self.data_capacity = int(size[1])
        self.dim = dim
        self.dtype = dtype
        self.pointers = np.zeros((self.pointer_capacity, 2), dtype=np.dtype("i"))
        shape = (self.data_capacity,) if dim == 1 else (self.data_capacity, dim)
        self.data = np.zeros(shape, dtype=dtype)
        self.number_of_entries = 0
        self.next_data_pos = 0

    # -- capacity -------------------------------------------------------------------------
    def set_pointer_capacity(self, capacity):
        self.pointer_capacity = int(capacity)
        new = np.zeros((self.pointer_capacity, 2), dtype=np.dtype("i"))
        n = min(len(self.pointers), self.pointer_capacity)
        new[:n] = self.pointers[:n]
        self.pointers = new

    def set_data_capacity(self, capacity):
        self.data_capacity = int(capacity)
        shape = (self.data_capacity,) if self.dim == 1 else (self.data_capacity, self.dim)
        new = np.zeros(shape, dtype=self.dtype)
        n = min(len(self.data), self.data_capacity)
        new[:n] = self.data[:n]
        self.data = new

    def new_size(self, size, minimum=1000):
        return max(size + size // 2 + 2, minimum)

    def _reserve(self, n_new):
        if self.number_of_entries >= len(self.pointers):
            self.set_pointer_capacity(self.new_size(len(self.pointers)))
        need = self.next_data_pos + n_new
        if need > len(self.data):
            self.set_data_capacity(max(self.new_size(len(self.data)), need))

    # -- entries -------------------------------------------------------------------------
    def add_entry(self, data):
        n = len(data)
        self._reserve(n)
        self.pointers[self.number_of_entries] = (self.next_data_pos, n)
        self.data[self.next_data_pos:self.next_data_pos + n] = data
        self.next_data_pos += n
        self.number_of_entries += 1
        return self.number_of_entries - 1

    def get_entry(self, index):
        if index >= self.number_of_entries or index < -self.number_of_entries:
            raise IndexError("No entry with index " + str(index))
        pos, length = self.pointers[index]
        return self.data[pos:pos + length]

    def set_entry(self, index, data):
        """New content for an existing entry; a longer one is re-appended at the end of `data`
        (the old slot is abandoned, like the reference)."""
        pos, length = self.pointers[index]
        n = len(data)
        if length >= n:
            self.data[pos:pos + n] = data
            self.pointers[index, 1] = n
        else:
            need = self.next_data_pos + n
            if need > len(self.data):
                self.set_data_capacity(max(self.new_size(len(self.data)), need))
            self.data[self.next_data_pos:self.next_data_pos + n] = data
            self.pointers[index] = (self.next_data_pos, n)
            self.next_data_pos += n

    def __getitem__(self, index):
        return self.get_entry(index)

    def __setitem__(self, index, value):
        self.set_entry(index, value)

    def __len__(self):
        return self.number_of_entries

    # -- bulk (used by the structured builders; not in the reference) --------------------------
    def set_uniform(self, table):
        """All entries at once from an (n, width) table."""
        table = np.ascontiguousarray(table, dtype=self.dtype)
        n, w = table.shape
        self.data = table.reshape(-1).copy()
        self.pointers = np.empty((n, 2), dtype=np.dtype("i"))
        self.pointers[:, 0] = np.arange(n, dtype=np.int64) * w
        self.pointers[:, 1] = w
        self.number_of_entries = n
        self.next_data_pos = n * w
        self.pointer_capacity, self.data_capacity = n, n * w

    def compact(self):
        """(ptr[n+1], data) in entry order, without the abandoned slots."""
        n = self.number_of_entries
        lens = self.pointers[:n, 1].astype(np.int64)
        ptr = np.zeros(n + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(lens)
        offs = self.pointers[:n, 0].astype(np.int64)
        if n and np.array_equal(offs, ptr[:-1]):
            return ptr, self.data[:ptr[-1]]
        idx = np.repeat(offs - ptr[:-1], lens) + np.arange(ptr[-1])
        return ptr, self.data[idx]


def _cross3(a, b):
    """np.cross for two 3-vectors without its axis bookkeeping (same products, same order: bit-identical)."""
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _grow(arr, n, fill=None):
    """Grow the first axis geometrically until index n-1 exists (the reference's _memory_extension)."""
    if len(arr) >= n:
        return arr
    new_len = max(n, len(arr) + len(arr) // 2 + 1)
    new = np.zeros((new_len,) + arr.shape[1:], dtype=arr.dtype)
    if fill is not None:
        new[...] = fill
    new[:len(arr)] = arr
    return new


class _PlaneCut(object):
    """State of one plane cut (Mesh.divide_cells_by_plane): vertex classification, the new point of every crossed
    edge, the side of every face met so far."""

    def __init__(self, mesh, point, normal, tol, boundary_of, local):
        self.mesh, self.p0, self.nrm, self.tol, self.boundary_of = mesh, point, normal, tol, boundary_of
        self.edge_point = {}   # (low point, high point) -> the point where the plane crosses the edge
        self.face_side = {}    # face -> -1 below, +1 above, 0 in the plane
        if local:              # a cut of one cell: distances on demand (the mesh may have millions of points)
            self.distance = {}
        else:                  # mesh-wide: every existing vertex classified at once; points created here lie on the plane
            n_old = mesh.get_number_of_points()
            self.distance = (mesh.points[:n_old] - point).dot(normal)
        self.n_old = mesh.get_number_of_points()
        self.local = local

    def dist(self, p):
        if not self.local:
            return float(self.distance[p]) if p < self.n_old else 0.
        d = self.distance.get(p)
        if d is None:
            d = self.distance[p] = float((self.mesh.get_point(p) - self.p0).dot(self.nrm))
        return d

    def side(self, p):
        d = self.dist(p)
        return 0 if abs(d) <= self.tol else (1 if d > 0. else -1)

    def crossing(self, a, b):
        key = (a, b) if a < b else (b, a)
        p = self.edge_point.get(key)
        if p is None:
            da, db = self.dist(key[0]), self.dist(key[1])
            pa, pb = self.mesh.get_point(key[0]), self.mesh.get_point(key[1])
            p = self.mesh.add_point(pa + (da / (da - db)) * (pb - pa))
            if self.local:
                self.distance[p] = 0.
            self.edge_point[key] = p
        return p

    def split_faces_of(self, c):
        m = self.mesh
        for f in [int(x) for x in m.get_cell(c)]:
            if f in self.face_side:
                continue
            loop = [int(p) for p in m.get_face(f)]
            full = []
            for a, b in zip(loop, loop[1:] + loop[:1]):
                full.append(a)
                if self.side(a) * self.side(b) < 0:
                    full.append(self.crossing(a, b))
            signs = [self.side(p) for p in full]
            below, above = -1 in signs, 1 in signs
            if not (below and above):
                self.face_side[f] = -1 if below else (1 if above else 0)
                continue
            on = [i for i, s in enumerate(signs) if s == 0]
            if len(on) != 2:
                raise Exception("divide_cells_by_plane: face %d meets the plane in %d points (not convex, "
                                "or it carries collinear vertices in the plane)" % (f, len(on)))
            i0, i1 = on
            part_a = full[i0:i1 + 1]
            part_b = full[i1:] + full[:i0 + 1]
            if len(part_a) < 3 or len(part_b) < 3:
                raise Exception("divide_cells_by_plane: degenerate split of face %d" % f)
            if signs[i0 + 1] > 0:
                part_a, part_b = part_b, part_a   # part_a: below the plane, keeps the index
            m.set_face(f, part_a)
            new_face = m.add_face(part_b)
            m.set_face_normal(new_face, m.get_face_normal(f))
            for g in (f, new_face):
                area, centroid = m.find_face_centroid(g)
                m.set_face_area(g, area)
                m.set_face_real_centroid(g, centroid)
                if m.has_face_shifted_centroid:
                    m.set_face_shifted_centroid(g, centroid)
            self.face_side[f], self.face_side[new_face] = -1, 1
            for cell in m.get_face_to_cell(f):
                cf = list(m.get_cell(cell))
                co = list(m.get_cell_normal_orientation(cell))
                local = cf.index(f)
                m.set_cell_faces(cell, cf + [new_face])
                m.set_cell_orientation(cell, np.array(co + [co[local]]))
                if f in self.boundary_of:
                    m.add_boundary_face(self.boundary_of[f], new_face, co[local])
                    self.boundary_of[new_face] = self.boundary_of[f]

    def divide(self, c, new_cells):
        m = self.mesh
        cf = [int(x) for x in m.get_cell(c)]
        co = [int(x) for x in m.get_cell_normal_orientation(c)]
        sides = [self.face_side[f] for f in cf]
        if not (-1 in sides and 1 in sides):
            return   # the plane touches the cell at most in a face, an edge or a vertex
        if 0 in sides:
            raise Exception("divide_cells_by_plane: cell %d has a face in the plane and faces on both sides" % c)
        segments = set()
        for f in cf:
            loop = [int(p) for p in m.get_face(f)]
            for a, b in zip(loop, loop[1:] + loop[:1]):
                if self.side(a) == 0 and self.side(b) == 0:
                    segments.add((a, b) if a < b else (b, a))
        polygon = m._chain_segments(segments, c)
        cut_face = m.add_face(polygon)
        area, centroid = m.find_face_centroid(cut_face)
        m.set_face_area(cut_face, area)
        m.set_face_real_centroid(cut_face, centroid)
        if m.has_face_shifted_centroid:
            m.set_face_shifted_centroid(cut_face, centroid)
        cut_normal = m.find_face_normal(cut_face)
        m.set_face_normal(cut_face, cut_normal)
        self.face_side[cut_face] = 0
        same = cut_normal.dot(self.nrm) > 0.
        faces_1 = [f for f, s in zip(cf, sides) if s < 0] + [cut_face]
        orient_1 = [o for o, s in zip(co, sides) if s < 0] + [1 if same else -1]
        faces_2 = [f for f, s in zip(cf, sides) if s > 0] + [cut_face]
        orient_2 = [o for o, s in zip(co, sides) if s > 0] + [-1 if same else 1]
        for f in faces_2[:-1]:
            m.remove_from_face_to_cell(f, c)
        m.set_cell_faces(c, faces_1)
        m.set_cell_orientation(c, orient_1)
        new_cell = m.add_cell(faces_2, orient_2)
        m.set_cell_k(new_cell, m.get_cell_k(c))
        m.cell_domain[new_cell] = m.cell_domain[c]
        for cell in (c, new_cell):
            volume, centroid = m.find_volume_centroid(cell)
            m.set_cell_volume(cell, volume)
            m.set_cell_real_centroid(cell, centroid)
            if m.has_cell_shifted_centroid:
                m.set_cell_shifted_centroid(cell, centroid)
        new_cells[c] = new_cell


class Mesh(object):
    """Polyhedral mesh container (mesh.py:185-302) -- same attribute names as the reference."""

    def __init__(self):
        self.points = np.empty((0, 3), dtype=np.dtype('d'))
        self.number_of_points = 0
        self.faces = variable_array(dtype=np.dtype('i'))
        self.face_normals = np.empty((0, 3), dtype=np.dtype('d'))
        self.face_areas = np.empty((0,), dtype=np.dtype('d'))
        self.face_real_centroids = np.empty((0, 3))
        self.face_to_cell = np.empty((0, 2), dtype=np.dtype('i'))
        self.face_shifted_centroids = np.empty((0, 3))
        self.has_face_shifted_centroid = False
        self.has_cell_shifted_centroid = False
        self.has_alpha = False
        self.boundary_markers = []
        self.boundary_descriptions = []
        self.boundary_faces = {}
        self.cells = variable_array(dtype=np.dtype('i'))
        self.cell_normal_orientation = variable_array(dtype=np.dtype('i'))
        self.cell_volume = np.empty((0,), dtype=np.dtype('d'))
        self.cell_real_centroid = np.empty((0, 3), dtype=np.dtype('d'))
        self.cell_shifted_centroid = np.empty((0, 3), dtype=np.dtype('d'))
        self.cell_k = np.empty((0, 9))
        self.cell_domain = np.empty((0,), dtype=int)
        self.cell_domain_tags = set([0])
        self.dim = 3
        self.dirichlet_boundary_pointers = {}
        self.periodic_boundaries = []
        self.internal_no_flow = []
        self.face_to_lagrange_pointers = {}
        self.lagrange_to_face_pointers = {}
        self.forcing_function_pointers = {}
        self.cell_alpha = []
        self.is_using_alpha_list = False
        self.gravity_vector = None
        self.gravity_acceleration = 9.8
        # device snapshot cache (invalidated by any mutation through the API below)
        self._version = 0
        self._device = None
        self._device_version = -1

    def _touch(self):
        self._version += 1

    # ---- points (mesh.py:304-341) --------------------------------------------------------
    def add_point(self, new_point):
        self.points = _grow(self.points, self.number_of_points + 1)
        self.points[self.number_of_points] = new_point
        self.number_of_points += 1
        self._touch()
        return self.number_of_points - 1

    def get_point(self, point_index):
        return self.points[point_index]

    def get_number_of_points(self):
        return self.number_of_points

    # ---- faces (mesh.py:349-503) -----------------------------------------------------------
    def add_face(self, list_of_points):
        f = self.faces.add_entry(list_of_points)
        self.face_normals = _grow(self.face_normals, f + 1)
        self.face_areas = _grow(self.face_areas, f + 1)
        self.face_real_centroids = _grow(self.face_real_centroids, f + 1)
        self.face_to_cell = _grow(self.face_to_cell, f + 1, fill=-1)
        if self.has_face_shifted_centroid:
            self.face_shifted_centroids = _grow(self.face_shifted_centroids, f + 1)
        self._touch()
        return f

    def set_face(self, face_index, points):
        self.faces[face_index] = points
        self._touch()

    def get_face(self, face_index):
        return self.faces[face_index]

    def get_number_of_face_points(self, face_index):
        return len(self.faces[face_index])

    def get_number_of_faces(self):
        return self.faces.number_of_entries

    def duplicate_face(self, face_index):
        new_index = self.add_face(self.get_face(face_index))
        self.set_face_area(new_index, self.get_face_area(face_index))
        return new_index

    def remove_from_face_to_cell(self, face_index, cell_index):
        row = self.face_to_cell[face_index]
        if row[0] == cell_index:
            row[0] = -1
        elif row[1] == cell_index:
            row[1] = -1
        else:
            raise Exception("cell_index %s not found in face_to_cell for %s" % (cell_index, face_index))
        self._touch()

    def add_to_face_to_cell(self, face_index, cell_index):
        row = self.face_to_cell[face_index]
        if row[0] == -1:
            row[0] = cell_index
        elif row[1] == -1:
            row[1] = cell_index
        else:
            raise Exception("cell_index %s could not be added to %s" % (cell_index, face_index))
        self._touch()

    def get_face_to_cell(self, face_index):
        return [int(x) for x in self.face_to_cell[face_index] if x >= 0]

    def set_face_real_centroid(self, face_index, centroid):
        self.face_real_centroids[face_index] = centroid
        self._touch()

    def get_face_real_centroid(self, face_index):
        return self.face_real_centroids[face_index]

    def use_face_shifted_centroid(self):
        self.has_face_shifted_centroid = True
        self.face_shifted_centroids = _grow(self.face_shifted_centroids, self.get_number_of_faces())
        self._touch()

    def is_using_face_shifted_centroid(self):
        return self.has_face_shifted_centroid

    def set_face_shifted_centroid(self, face_index, centroid):
        self.face_shifted_centroids[face_index] = centroid
        self._touch()

    def get_face_shifted_centroid(self, face_index):
        return self.face_shifted_centroids[face_index]

    def set_face_shifted_to_tpfa_all(self):
        """Shifted centroid of every interior face = the point where the line between the centroids of its two
        cells pierces the face's plane (boundary faces keep their real centroid): with these points the MFD
        inner product of a K-orthogonal cell is the two-point flux one (mesh.py:510-527).  All faces at once."""
        nf = self.get_number_of_faces()
        if not self.has_face_shifted_centroid:
            self.use_face_shifted_centroid()
        self.face_shifted_centroids = _grow(self.face_shifted_centroids, nf)
        f2c = self.face_to_cell[:nf]
        shifted = self.face_real_centroids[:nf].copy()
        two = np.nonzero((f2c >= 0).all(axis=1))[0]
        if len(two):
            c1 = self.cell_real_centroid[f2c[two, 0]]
            direction = self.cell_real_centroid[f2c[two, 1]] - c1
            direction = direction / np.linalg.norm(direction, axis=1)[:, None]
            normal = self.face_normals[two]
            d = ((self.face_real_centroids[two] - c1) * normal).sum(axis=1) / (direction * normal).sum(axis=1)
            shifted[two] = d[:, None] * direction + c1
        self.face_shifted_centroids[:nf] = shifted
        self._touch()

    def is_line_seg_intersect_face(self, face_index, p1, p2):
        """True when the segment p1-p2 pierces the (convex, planar) face: the point where it meets the face's plane
        lies on the same side of every edge (mesh.py:529-576)."""
        p1, p2 = np.asarray(p1, dtype=float), np.asarray(p2, dtype=float)
        length = np.linalg.norm(p2 - p1)
        vector = (p2 - p1) / length
        normal = self.get_face_normal(face_index)
        denom = vector.dot(normal)
        if abs(denom) < 1e-10:
            return None   # parallel to the face: the reference falls through without a verdict
        d = (self.get_face_real_centroid(face_index) - p1).dot(normal) / denom
        if not (1.e-8 < d <= length + 1.e-8):
            return None
        hit = d * vector + p1
        pts = self.points[np.asarray(self.get_face(face_index))]
        previous = np.roll(pts, 1, axis=0)
        direction = np.cross(pts - previous, previous - hit).dot(normal)
        return bool((direction > 0.).all() or (direction < 0.).all())

    def find_centroid_for_coordinates(self, face_index, coordinates):
        """(signed area, centroid coordinate 1, centroid coordinate 2) of the face projected on the two chosen
        coordinate axes (mesh.py:1604-1637)."""
        pts = self.points[np.asarray(self.get_face(face_index))]
        a, b = pts[:, coordinates[0]], pts[:, coordinates[1]]
        an, bn = np.roll(a, -1), np.roll(b, -1)
        cross = a * bn - an * b
        area = .5 * cross.sum()
        return (area, ((a + an) * cross).sum() / (6. * area), ((b + bn) * cross).sum() / (6. * area))

    def find_domain_faces(self, domain):
        """(faces, orientations) of the hull of `domain`, seen from its cells, in cell order: every face whose other
        side is not a cell of the domain -- another domain's cell or, like in the reference, nothing at all (a face
        on the mesh boundary has -1 there, which is in no domain; mesh.py:2152-2184)."""
        nc = self.get_number_of_cells()
        inside = np.zeros(nc, dtype=bool)
        inside[self.get_cells_in_domain(domain)] = True
        faces, orientations = [], []
        for c in np.nonzero(inside)[0]:
            for f, s in zip(self.get_cell(c), self.get_cell_normal_orientation(c)):
                row = self.face_to_cell[f]
                other = row[1] if row[0] == c else row[0]
                if other < 0 or not inside[other]:
                    faces.append(int(f))
                    orientations.append(int(s))
        return (faces, orientations)

    def set_face_area(self, face_index, area):
        self.face_areas[face_index] = area
        self._touch()

    def get_face_area(self, face_index):
        return self.face_areas[face_index]

    def set_face_normal(self, face_index, normal):
        self.face_normals[face_index] = normal
        self._touch()

    def get_face_normal(self, face_index):
        return self.face_normals[face_index]

    # ---- cells (mesh.py:951-1240) ---------------------------------------------------------
    def add_cell(self, list_of_faces, list_of_orientations):
        c = self.cells.add_entry(list_of_faces)
        self.cell_normal_orientation.add_entry(list_of_orientations)
        self.cell_volume = _grow(self.cell_volume, c + 1)
        self.cell_k = _grow(self.cell_k, c + 1)
        self.cell_domain = _grow(self.cell_domain, c + 1)
        self.cell_real_centroid = _grow(self.cell_real_centroid, c + 1)
        if self.has_cell_shifted_centroid:
            self.cell_shifted_centroid = _grow(self.cell_shifted_centroid, c + 1)
        if self.has_alpha:
            self.cell_alpha.append(None)
        for f in list_of_faces:
            row = self.face_to_cell[f]
            if row[0] == -1:
                row[0] = c
            elif row[1] == -1:
                row[1] = c
            else:
                raise Exception("setting face %s to cell %s already set to two cells %s" % (f, c, row))
        self._touch()
        return c

    def set_cell_faces(self, cell_index, faces):
        self.cells[cell_index] = faces
        for f in faces:
            if cell_index not in self.face_to_cell[f]:
                self.add_to_face_to_cell(f, cell_index)
        self._touch()

    def set_cell_orientation(self, cell_index, orientation):
        self.cell_normal_orientation[cell_index] = orientation
        self._touch()

    def get_cell(self, cell_index):
        return self.cells[cell_index]

    def get_cell_normal_orientation(self, cell_index):
        return self.cell_normal_orientation[cell_index]

    def get_number_of_cells(self):
        return len(self.cells)

    def get_number_of_cell_faces(self, cell_index):
        return len(self.cells[cell_index])

    def set_cell_real_centroid(self, cell_index, centroid):
        self.cell_real_centroid[cell_index] = centroid
        self._touch()

    def get_cell_real_centroid(self, cell_index):
        return self.cell_real_centroid[cell_index]

    def get_all_cell_real_centroids(self):
        return self.cell_real_centroid[:self.get_number_of_cells()]

    def get_all_cell_shifted_centroids(self):
        return self.cell_shifted_centroid[:self.get_number_of_cells()]

    def use_cell_shifted_centroid(self):
        self.has_cell_shifted_centroid = True
        self.cell_shifted_centroid = _grow(self.cell_shifted_centroid, self.get_number_of_cells())
        self._touch()

    def is_using_cell_shifted_centroid(self):
        return self.has_cell_shifted_centroid

    def set_cell_shifted_centroid(self, cell_index, centroid):
        self.cell_shifted_centroid[cell_index] = centroid
        self._touch()

    def get_cell_shifted_centroid(self, cell_index):
        return self.cell_shifted_centroid[cell_index]

    def set_cell_volume(self, cell_index, volume):
        self.cell_volume[cell_index] = volume
        self._touch()

    def get_cell_volume(self, cell_index):
        return self.cell_volume[cell_index]

    def set_cell_k(self, cell_index, k):
        self.cell_k[cell_index] = np.asarray(k).reshape((1, 9))
        self._touch()

    def get_cell_k(self, cell_index):
        return self.cell_k[cell_index].reshape((3, 3))

    def get_all_k_entry(self, i, j):
        return self.cell_k[:self.get_number_of_cells(), i * 3 + j]

    def get_all_k(self):
        return self.cell_k

    def use_alpha(self):
        self.has_alpha = True

    def set_alpha_by_cell(self, alpha, cell_index):
        self.cell_alpha[cell_index] = alpha

    def get_alpha_by_cell(self, cell_index):
        return self.cell_alpha[cell_index]

    def set_cell_domain(self, cell_index, domain):
        self.cell_domain[cell_index] = domain
        self.cell_domain_tags.add(domain)

    def get_domain_tags(self):
        return list(self.cell_domain_tags)

    def get_cell_domain(self, cell_index):
        return self.cell_domain[cell_index]

    def get_cell_domain_all(self):
        return self.cell_domain[:self.get_number_of_cells()]

    def get_cells_in_domain(self, domain):
        return [int(c) for c in np.nonzero(self.cell_domain[:self.get_number_of_cells()] == domain)[0]]

    # ---- boundary bookkeeping (mesh.py:1242-1395) ------------------------------------------
    def set_boundary_markers(self, boundary_markers, boundary_descriptions):
        self.boundary_markers = boundary_markers
        self.boundary_descriptions = boundary_descriptions
        for marker in boundary_markers:
            self.boundary_faces[marker] = []

    def add_boundary_marker(self, boundary_marker, boundary_description):
        self.boundary_markers.append(boundary_marker)
        self.boundary_descriptions.append(boundary_description)
        self.boundary_faces[boundary_marker] = []

    def create_new_boundary_marker(self, boundary_description):
        new_index = len(self.boundary_markers)
        self.add_boundary_marker(new_index, boundary_description)
        return new_index

    def has_boundary_marker(self, boundary_marker):
        return boundary_marker in self.boundary_markers

    def get_boundary_markers(self):
        return self.boundary_markers

    def get_boundary_description(self, boundary_marker):
        return self.boundary_descriptions[boundary_marker]

    def add_boundary_face(self, boundary_marker, face_index, face_orientation):
        self.boundary_faces[boundary_marker].append([face_index, face_orientation])

    def set_boundary_faces(self, boundary_marker, face_orientation_list):
        self.boundary_faces[boundary_marker] = face_orientation_list

    def get_boundary_faces_by_marker(self, boundary_marker):
        return self.boundary_faces[boundary_marker]

    def is_boundary_face(self, face_index, markers):
        return self.find_boundary_marker(face_index, markers) is not None

    def find_boundary_marker(self, face_index, markers):
        for marker in markers:
            for face in self.boundary_faces[marker]:
                if face_index == face[0]:
                    return marker
        return None

    def set_boundary_face_orientation(self, face_index, new_orientation):
        for marker in self.boundary_markers:
            for face in self.boundary_faces[marker]:
                if face_index == face[0]:
                    face[1] = new_orientation

    def get_number_of_boundary_faces(self):
        return sum(len(self.boundary_faces[m]) for m in self.boundary_markers)

    def add_internal_no_flow(self, face_index, face_orientation):
        self.internal_no_flow.append([face_index, face_orientation])

    def get_internal_no_flow(self):
        return self.internal_no_flow

    # ---- pointer couplings (mesh.py:1397-1525): stored, consumed by "next" components ------------
    def set_dirichlet_face_pointer(self, face_index, face_orientation, cell_index):
        self.dirichlet_boundary_pointers[face_index] = (cell_index, face_orientation)

    def get_dirichlet_pointer_faces(self):
        return list(self.dirichlet_boundary_pointers.keys())

    def get_dirichlet_pointer(self, face_index):
        return self.dirichlet_boundary_pointers[face_index]

    def set_face_to_lagrange_pointer(self, face_index, face_orientation, lagrange_index):
        self.face_to_lagrange_pointers[face_index] = (lagrange_index, face_orientation)

    def get_all_face_to_lagrange_pointers(self):
        return list(self.face_to_lagrange_pointers.keys())

    def get_face_to_lagrange_pointer(self, face_index):
        return self.face_to_lagrange_pointers[face_index]

    def set_lagrange_to_face_pointers(self, lagrange_index, face_index, orientation):
        self.lagrange_to_face_pointers[lagrange_index] = list(zip(face_index, orientation))

    def get_all_lagrange_to_face_pointers(self):
        return list(self.lagrange_to_face_pointers.keys())

    def get_lagrange_to_face_pointers(self, lagrange_index):
        return self.lagrange_to_face_pointers[lagrange_index]

    def set_periodic_boundary(self, face_index_1, face_orientation_1, face_index_2, face_orientation_2):
        """Ties two boundary faces through one shared multiplier face, a duplicate of the first (mesh.py:1476-1492):
        MFD imposes continuity of pressure and flux across the pair."""
        lagrange_index_1 = self.duplicate_face(face_index_1)
        self.periodic_boundaries.append((face_index_1, face_orientation_1, face_index_2, face_orientation_2,
                                         lagrange_index_1))

    def set_forcing_pointer(self, cell_index, face_indices, face_orientations):
        self.forcing_function_pointers[cell_index] = list(zip(face_indices, face_orientations))

    def get_forcing_pointer_cells(self):
        return list(self.forcing_function_pointers.keys())

    def get_forcing_pointers_for_cell(self, cell_index):
        return self.forcing_function_pointers[cell_index]

    def has_pointer_couplings(self):
        return bool(self.dirichlet_boundary_pointers or self.face_to_lagrange_pointers or
                    self.lagrange_to_face_pointers or self.forcing_function_pointers or
                    self.periodic_boundaries)

    def set_gravity_vector(self, gravity_vector):
        self.gravity_vector = gravity_vector

    def get_gravity_vector(self):
        return self.gravity_vector

    def get_gravity_acceleration(self):
        return self.gravity_acceleration

    def set_fluid_density(self, density):
        """MFD.add_gravity reads mesh.get_fluid_density() (mfd.py:1008); the reference's Mesh never defines it.
        Supplied here so that add_gravity is usable."""
        self.fluid_density = density

    def get_fluid_density(self):
        if getattr(self, "fluid_density", None) is None:
            raise AttributeError("fluid density not set (Mesh.set_fluid_density)")
        return self.fluid_density

    # ---- planar polygon geometry, host side (mesh.py:1575-1756): used when cells are cut --------
    def find_basis_for_face(self, face_index):
        face = self.get_face(face_index)
        n = len(face)
        for i in range(n):
            v1 = self.get_point(face[(i + 1) % n]) - self.get_point(face[i])
            v2 = self.get_point(face[i]) - self.get_point(face[i - 1])
            v1 = v1 / np.linalg.norm(v1)
            v2 = v2 / np.linalg.norm(v2)
            if 1. - abs(v1.dot(v2)) > 1.e-6:
                return (v1, v2, face[i])
        raise Exception("Couldn't compute basis for face " + str(face_index))

    def find_face_normal(self, face_index):
        """First non-degenerate corner, cross(previous edge, next edge), normalised (mesh.py:1589-1602)."""
        face = self.get_face(face_index)
        n = len(face)
        for i in range(n):
            v1 = self.get_point(face[(i + 1) % n]) - self.get_point(face[i])
            v2 = self.get_point(face[i]) - self.get_point(face[i - 1])
            nrm = _cross3(v2, v1)
            length = np.linalg.norm(nrm)
            if length > 1.e-10:
                return nrm / length
        raise Exception("Couldn't compute normal for face " + str(face_index))

    def compute_polygon_area(self, polygon, dims=[0, 1]):
        """Signed shoelace area in the two chosen coordinates (mesh.py:1733-1756)."""
        p = np.asarray(polygon, dtype=float)
        x, y = p[:, dims[0]], p[:, dims[1]]
        return .5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))

    def find_face_centroid(self, face_index):
        """(area, centroid) of a planar polygon through an in-plane orthonormal basis (mesh.py:1639-1731)."""
        v1, v2, origin_index = self.find_basis_for_face(face_index)
        polygon = np.array([self.get_point(x) for x in self.get_face(face_index)], dtype=float)
        v1 = v1 / np.linalg.norm(v1)
        v2 = _cross3(_cross3(v1, v2), v1)
        if np.linalg.norm(v2) < 1.e-10:
            v2 = _cross3(_cross3(v1, polygon[-2] - polygon[-1]), v1)
        v2 = v2 / np.linalg.norm(v2)
        rel = polygon - self.get_point(origin_index)
        a, b = rel.dot(v1), rel.dot(v2)
        an, bn = np.roll(a, -1), np.roll(b, -1)
        cross = a * bn - an * b
        area = .5 * cross.sum()
        cx = ((a + an) * cross).sum() / (6. * area)
        cy = ((b + bn) * cross).sum() / (6. * area)
        # the reference anchors the in-plane coordinates at polygon[0] (mesh.py:1717-1727)
        centroid = polygon[0] + cx * v1 + cy * v2
        return (abs(area), centroid)

    # ---- cell volumes / centroids ---------------------------------------------------------------
    def find_volume_centroid_all(self, ctx=None):
        """All cell volumes and centroids on the GPU (replaces mesh_cython.all_cell_volumes_centroids,
        mesh.py:1758-1781).  Adjacency comes from the cell lists themselves."""
        from .. import _lib, device
        import torch

        ctx = ctx or _lib.default_context()
        nc, nf = self.get_number_of_cells(), self.get_number_of_faces()
        cptr, cdata = self.cells.compact()
        _, sdata = self.cell_normal_orientation.compact()
        fptr, fdata = self.faces.compact()
        dev = ctx.device
        with ctx.on_stream():
            t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=dev)
            vol = torch.empty(nc, dtype=torch.float64, device=dev)
            cen = torch.empty((nc, 3), dtype=torch.float64, device=dev)
            args = [t(cptr, np.int32), t(cdata, np.int32), t(sdata, np.int8), t(fptr, np.int32), t(fdata, np.int32),
                    t(self.points[:self.number_of_points], np.float64), t(self.face_normals[:nf], np.float64)]
            max_faces = int(np.diff(cptr).max()) if nc else 1
            ctx.call("mimpy_b200_geom_cells", nc, max_faces, *[_lib.ptr(a) for a in args], _lib.ptr(vol), _lib.ptr(cen))
            self.cell_volume = _grow(self.cell_volume, nc)
            self.cell_real_centroid = _grow(self.cell_real_centroid, nc)
            self.cell_volume[:nc] = vol.cpu().numpy()
            self.cell_real_centroid[:nc] = cen.cpu().numpy()
        self._touch()

    def find_volume_from_faces(self, face_list, orientation_list):
        """Mirtich volume/centroid of ONE cell on the host (mesh.py:1783-1882); used while cutting."""
        volume = 0.
        centroid = np.zeros(3)
        for f, s in zip(face_list, orientation_list):
            nrm = self.get_face_normal(f) * s
            an = np.abs(nrm)
            C = 0 if (an[0] > an[1] and an[0] > an[2]) else (1 if an[1] > an[2] else 2)
            A, B = (C + 1) % 3, (C + 2) % 3
            pts = np.array([self.get_point(p) for p in self.get_face(f)], dtype=float)
            if s > 0:
                cur, nxt = pts, np.roll(pts, -1, axis=0)
            else:
                nxt, cur = pts, np.roll(pts, -1, axis=0)
            a0, b0, a1, b1 = cur[:, A], cur[:, B], nxt[:, A], nxt[:, B]
            da, db = a1 - a0, b1 - b0
            C1 = a1 + a0
            Ca = a1 * C1 + a0 * a0
            Caa = a1 * Ca + a0 ** 3
            Cb = b1 * (b1 + b0) + b0 * b0
            Cbb = b1 * Cb + b0 ** 3
            Cab = 3. * a1 * a1 + 2. * a1 * a0 + a0 * a0
            Kab = a1 * a1 + 2. * a1 * a0 + 3. * a0 * a0
            P1 = (db * C1).sum() / 2.
            Pa = (db * Ca).sum() / 6.
            Paa = (db * Caa).sum() / 12.
            Pb = (da * Cb).sum() / -6.
            Pbb = (da * Cbb).sum() / -12.
            Pab = (db * (b1 * Cab + b0 * Kab)).sum() / 24.
            w = -nrm.dot(pts[0])
            k1 = 1. / nrm[C]
            Fa, Fb = k1 * Pa, k1 * Pb
            Fc = -k1 * k1 * (nrm[A] * Pa + nrm[B] * Pb + w * P1)
            Faa, Fbb = k1 * Paa, k1 * Pbb
            Fcc = k1 ** 3 * (nrm[A] ** 2 * Paa + 2 * nrm[A] * nrm[B] * Pab + nrm[B] ** 2 * Pbb +
                             w * (2. * (nrm[A] * Pa + nrm[B] * Pb) + w * P1))
            volume += nrm[0] * (Fa if A == 0 else (Fb if B == 0 else Fc))
            centroid[A] += nrm[A] * Faa
            centroid[B] += nrm[B] * Fbb
            centroid[C] += nrm[C] * Fcc
        centroid /= volume * 2.
        return (volume, centroid)

    def find_volume_centroid(self, cell_index):
        return self.find_volume_from_faces(self.get_cell(cell_index), self.get_cell_normal_orientation(cell_index))

    def find_cell_near_point(self, point):
        c = self.cell_real_centroid[:self.get_number_of_cells()] - np.asarray(point)[None, :]
        return int(np.argmin((c * c).sum(axis=1)))

    # ---- config-5 mesh generator: split one cell by a plane (mesh.py:2186-2419) ------------------
    def construct_polygon_from_segs(self, segments):
        """Chains point pairs (matched by position, 1e-7) into one closed polygon."""
        segs = [list(s) for s in segments]
        poly = [segs[0][0]]
        target = segs[0][1]
        segs.pop(0)
        while segs:
            hit = None
            for k, (p, q) in enumerate(segs):
                if np.linalg.norm(self.get_point(target) - self.get_point(p)) < 1.e-7:
                    hit = (k, p, q)
                elif np.linalg.norm(self.get_point(target) - self.get_point(q)) < 1.e-7:
                    hit = (k, q, p)
            if hit is None:
                raise Exception("Faild at constructing polygon from segments")
            k, here, there = hit
            poly.append(here)
            target = there
            segs.pop(k)
        return poly

    def divide_cell_by_plane(self, cell_index, point_on_plane, plane_normal):
        """Splits `cell_index` by a plane: every cut face is split in two (the neighbour cell gains
        the coplanar sub-face), the cut polygon becomes a new interior face, faces are distributed
        to the two halves by the side of their centroid; returns the new cell's index."""
        point_on_plane = np.asarray(point_on_plane, dtype=float)
        plane_normal = np.asarray(plane_normal, dtype=float)
        interior_segments = []
        for face_index in list(self.get_cell(cell_index)):
            face = list(self.get_face(face_index))
            side_a, side_b = [], []
            on_a = True
            seg = None
            for p_idx, q_idx in zip(face, face[1:] + face[:1]):
                (side_a if on_a else side_b).append(p_idx)
                p1, p2 = self.get_point(p_idx), self.get_point(q_idx)
                length = np.linalg.norm(p2 - p1)
                direction = (p2 - p1) / length
                denom = direction.dot(plane_normal)
                if abs(denom) < 1e-10:
                    continue
                d = (point_on_plane - p1).dot(plane_normal) / denom
                if -1.e-8 < d <= length + 1.e-8:
                    new_point = self.add_point(d * direction + p1)
                    side_a.append(new_point)
                    side_b.append(new_point)
                    if on_a:
                        seg = [new_point]
                    else:
                        seg.append(new_point)
                        interior_segments.append(seg)
                    on_a = not on_a
            if not side_b:
                continue
            assert len(side_a) > 2
            self.set_face(face_index, side_a)
            area, centroid = self.find_face_centroid(face_index)
            self.set_face_real_centroid(face_index, centroid)
            self.set_face_area(face_index, area)
            new_face = self.add_face(side_b)
            area, centroid = self.find_face_centroid(new_face)
            self.set_face_real_centroid(new_face, centroid)
            self.set_face_area(new_face, area)
            self.set_face_normal(new_face, self.get_face_normal(face_index))
            for c in [cell_index] + [x for x in self.get_face_to_cell(face_index) if x != cell_index]:
                cf = list(self.get_cell(c))
                co = list(self.get_cell_normal_orientation(c))
                local = cf.index(face_index)
                if c == cell_index:
                    marker = self.find_boundary_marker(face_index, self.get_boundary_markers())
                    if marker is not None:
                        self.add_boundary_face(marker, new_face, co[local])
                self.set_cell_faces(c, cf + [new_face])
                self.set_cell_orientation(c, np.array(co + [co[local]]))
        new_poly = self.construct_polygon_from_segs(interior_segments)
        cut_face = self.add_face(new_poly)
        area, centroid = self.find_face_centroid(cut_face)
        self.set_face_real_centroid(cut_face, centroid)
        self.set_face_area(cut_face, area)
        cut_normal = self.find_face_normal(cut_face)
        self.set_face_normal(cut_face, cut_normal)
        faces_1, faces_2, orient_1, orient_2 = [], [], [], []
        cf = list(self.get_cell(cell_index))
        co = list(self.get_cell_normal_orientation(cell_index))
        for f, s in zip(cf, co):
            if (point_on_plane - self.get_face_real_centroid(f)).dot(plane_normal) > 0.:
                faces_1.append(f)
                orient_1.append(s)
            else:
                faces_2.append(f)
                orient_2.append(s)
        faces_1.append(cut_face)
        faces_2.append(cut_face)
        same = cut_normal.dot(plane_normal) > 0.
        orient_1.append(1 if same else -1)
        orient_2.append(-1 if same else 1)
        self.set_cell_faces(cell_index, faces_1)
        self.set_cell_orientation(cell_index, orient_1)
        volume, centroid = self.find_volume_centroid(cell_index)
        self.set_cell_real_centroid(cell_index, centroid)
        self.set_cell_volume(cell_index, volume)
        for f in faces_2:
            if f != cut_face and cell_index in self.face_to_cell[f]:
                self.remove_from_face_to_cell(f, cell_index)
        new_cell = self.add_cell(faces_2, orient_2)
        volume, centroid = self.find_volume_centroid(new_cell)
        self.set_cell_volume(new_cell, volume)
        self.set_cell_real_centroid(new_cell, centroid)
        self.set_cell_k(new_cell, self.get_cell_k(cell_index))
        return new_cell

    # ---- many cells, one plane ----------------------------------------------------
    def find_cells_crossed_by_plane(self, point_on_plane, plane_normal, inside=None, tol=1.e-8):
        """Cells with at least one edge that crosses the plane -- at a point x with inside(x) true, when a bounded
        piece of the plane is meant: `inside` maps an (m, 3) array of points in the plane to m booleans (a disc, an
        ellipse, a polygon: what Fracture.is_point_on_frac decides in the reference's selection loop,
        hexmeshwmsfracs.py:683-713).  All edges of the mesh at once; the result is what divide_cells_by_plane takes."""
        p0 = np.asarray(point_on_plane, dtype=float)
        nrm = np.asarray(plane_normal, dtype=float)
        nrm = nrm / np.linalg.norm(nrm)
        nf, nc = self.get_number_of_faces(), self.get_number_of_cells()
        fptr, fdata = self.faces.compact()
        fdata = np.asarray(fdata, dtype=np.int64)
        lens = np.diff(fptr)
        successor = np.arange(len(fdata)) + 1
        successor[fptr[1:][lens > 0] - 1] = fptr[:-1][lens > 0]      # the loop of a face closes on its first point
        a, b = fdata, fdata[successor]
        distance = (self.points[:self.get_number_of_points()] - p0).dot(nrm)
        da, db = distance[a], distance[b]
        crossed = (da * db < 0.) & (np.abs(da) > tol) & (np.abs(db) > tol)
        if inside is not None and crossed.any():
            pa, pb = self.points[a[crossed]], self.points[b[crossed]]
            x = pa + (da[crossed] / (da[crossed] - db[crossed]))[:, None] * (pb - pa)
            keep = np.asarray(inside(x), dtype=bool)
            crossed[np.nonzero(crossed)[0][~keep]] = False
        face_hit = np.zeros(nf, dtype=bool)
        face_hit[np.repeat(np.arange(nf), lens)[crossed]] = True
        cptr, cdata = self.cells.compact()
        cell_of_entry = np.repeat(np.arange(nc), np.diff(cptr))
        return [int(c) for c in np.unique(cell_of_entry[face_hit[np.asarray(cdata, dtype=np.int64)]])]

    def divide_cells_by_plane(self, cell_indices, point_on_plane, plane_normal, tol=1.e-8):
        """Splits every cell of `cell_indices` (None: all cells) that the plane really crosses; returns
        {cell_index: new_cell_index}.  Cells that the plane only touches are left alone.

        `divide_cell_by_plane` (mesh.py:2227-2419) cannot be applied to two face-adjacent cells: the second call
        meets the already-split shared face, whose new vertices lie ON the plane, and splits it again into
        degenerate pieces (the reference raises; HexMeshWMSFracs.add_fractures works around it with its
        `done_faces` list, hexmeshwmsfracs.py:725-737).  Here the cut is done mesh-wide and conforming:
          * every vertex is classified once by its signed distance (|s| <= tol: ON the plane, it becomes a
            vertex of the cut polygon and no point is created next to it);
          * every crossed EDGE gets one new point, shared by all faces and cells around it (the reference
            creates one per face and matches them by position afterwards);
          * every crossed FACE is split once, into its part below (keeps the face index) and above the plane
            (new index, same normal, same boundary marker); both cells of the face receive the new part;
          * every cell then sorts its faces by side, chains the in-plane edges of its faces into the cut
            polygon (by point id) and is divided like the reference does: `cell_index` keeps the part
            below the plane (`(point_on_plane - x) . plane_normal > 0`), the new cell gets the part above
            and the same K and domain.
        `point_on_plane` may also hold ONE POINT PER CELL (shape (len(cell_indices), 3)): parallel planes, each cell
        cut by its own -- the checkerboard generator of BASELINE config 5 in one call.  The listed cells must then
        not share faces (a shared face cannot be split by two planes).
        Faces are assumed convex (two crossings per face), as in the reference."""
        p0 = np.asarray(point_on_plane, dtype=float)
        nrm = np.asarray(plane_normal, dtype=float)
        nrm = nrm / np.linalg.norm(nrm)
        if cell_indices is None:
            cell_indices = range(self.get_number_of_cells())
        cell_indices = [int(c) for c in cell_indices]
        if len(set(cell_indices)) != len(cell_indices):
            raise ValueError("divide_cells_by_plane: a cell is listed twice")
        boundary_of = {}
        for marker in self.get_boundary_markers():
            for face_index, orientation in self.boundary_faces[marker]:
                boundary_of[int(face_index)] = marker

        if p0.ndim == 2:
            if p0.shape != (len(cell_indices), 3):
                raise ValueError("divide_cells_by_plane: one point per listed cell expected")
            seen = set()
            for c in cell_indices:
                faces = set(int(f) for f in self.get_cell(c))
                if seen & faces:
                    raise ValueError("divide_cells_by_plane: cells that share a face cannot be cut by different planes")
                seen |= faces
            new_cells = {}
            for c, point in zip(cell_indices, p0):
                cut = _PlaneCut(self, point, nrm, tol, boundary_of, local=True)
                cut.split_faces_of(c)
                cut.divide(c, new_cells)
            return new_cells

        cut = _PlaneCut(self, p0, nrm, tol, boundary_of, local=False)
        for c in cell_indices:      # 1. split the crossed faces, once each
            cut.split_faces_of(c)
        new_cells = {}
        for c in cell_indices:      # 2. divide the cells
            cut.divide(c, new_cells)
        return new_cells

    @staticmethod
    def _chain_segments(segments, cell_index):
        """Closed polygon (point ids) from the in-plane edges of one cell: every point must end two edges."""
        ends = {}
        for a, b in segments:
            ends.setdefault(a, []).append(b)
            ends.setdefault(b, []).append(a)
        if len(segments) < 3 or any(len(v) != 2 for v in ends.values()):
            raise Exception("divide_cells_by_plane: the plane does not cut cell %d in one simple polygon "
                            "(%d in-plane edges)" % (cell_index, len(segments)))
        start = min(ends)
        polygon, previous, here = [start], None, start
        while True:
            a, b = ends[here]
            nxt = b if a == previous else a
            if nxt == start:
                break
            polygon.append(nxt)
            previous, here = here, nxt
            if len(polygon) > len(ends):
                raise Exception("divide_cells_by_plane: open chain of in-plane edges in cell %d" % cell_index)
        if len(polygon) != len(ends):
            raise Exception("divide_cells_by_plane: the in-plane edges of cell %d form more than one loop" % cell_index)
        return polygon

    # ---- SoA export ---------------------------------------------------------------------------------
    # ---- on-disk formats (mesh.py:584-949, 1985-2087) -> meshio.py --------------------------------
    def load_mesh(self, input_file):
        from . import meshio
        meshio.load_mesh(self, input_file)

    def save_mesh(self, output_file):
        from . import meshio
        meshio.save_mesh(self, output_file)

    def save_cell(self, cell_index, output_file):
        from . import meshio
        meshio.save_cell(self, cell_index, output_file)

    def output_vtk_mesh(self, file_name, cell_values=[], cell_value_labels=[]):
        from . import meshio
        meshio.output_vtk_mesh(self, file_name, cell_values, cell_value_labels)

    def output_vector_field(self, file_name, vector_magnitudes=[], vector_labels=[]):
        from . import meshio
        meshio.output_vector_field(self, file_name, vector_magnitudes, vector_labels)

    def output_cell_normals(self, file_name, cell_index):
        from . import meshio
        meshio.output_cell_normals(self, file_name, cell_index)

    def output_vtk_faces(self, file_name, face_indices, face_values=[], face_value_labels=[]):
        from . import meshio
        meshio.output_vtk_faces(self, file_name, face_indices, face_values, face_value_labels)

    def to_device(self, ctx=None):
        """Gathers the mesh into SoA device arrays (cached until the mesh is modified)."""
        from .. import _lib, device

        if self._device is not None and self._device_version == self._version:
            return self._device
        ctx = ctx or _lib.default_context()
        nc, nf = self.get_number_of_cells(), self.get_number_of_faces()
        cptr, cdata = self.cells.compact()
        _, sdata = self.cell_normal_orientation.compact()
        fptr, fdata = self.faces.compact()
        fcen = self.face_shifted_centroids[:nf] if self.has_face_shifted_centroid else self.face_real_centroids[:nf]
        ccen = self.cell_shifted_centroid[:nc] if self.has_cell_shifted_centroid else self.cell_real_centroid[:nc]
        dm = device.DeviceMesh.from_arrays(ctx, cptr, cdata, sdata, self.face_areas[:nf], self.face_normals[:nf],
                                           fcen, self.cell_k[:nc], ccen, self.cell_volume[:nc],
                                           self.points[:self.number_of_points], fptr, fdata)
        self._device, self._device_version = dm, self._version
        return dm
