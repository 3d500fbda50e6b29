"""Voronoi meshes from voro++ output, with the API of mimpy.mesh.voromesh.VoroMesh (voromesh.py:12-201).

The file holds one line per cell, written by
    voro++ -c "%q$%v$%C$%P$%t$%f$%l$%n" x_min x_max y_min y_max z_min z_max [points_file]
i.e. id $ generator point $ volume $ centroid $ vertices $ faces (vertex tuples) $ face areas $ face normals $
neighbours (negative: wall -1 .. -6 = x_min, x_max, y_min, y_max, z_min, z_max).

build_mesh() numbers points, faces and cells exactly as the reference does, so that a script gets the same
arrays: every cell adds its own copy of its vertices; a face between two cells is created by the cell that
comes first in the file (orientation +1, the later cell lists it with -1); walls become boundary markers 0..5;
areas, normals, volumes and centroids are voro++'s; the SHIFTED centroids (generator points for cells, the
midpoint of the two generators for interior faces) are switched on, which makes the MFD inner product the
two-point flux one on the orthogonal Voronoi pairs.  The resulting general polyhedral mesh (cells of 6 .. 30
faces) goes through the generic lane-group kernels of the device library like any other Mesh.
"""
import gzip

import numpy as np

from . import mesh as _mesh


def _tuples(field, conv):
    """'(a,b,c) (d,e)' -> [[a, b, c], [d, e]]"""
    return [[conv(x) for x in item.split(",")] for item in field.replace("(", " ").replace(")", " ").split()]


class VoroMesh(_mesh.Mesh):
    """Loads a mesh from the voro++ Voronoi generator by Chris H. Rycroft."""

    MERGE_INTERIOR = 1.e-8   # vertices of a face closer than this are one point (voromesh.py:117, 163)
    MERGE_BOUNDARY = 1.e-12

    def _distinct(self, point_ids, tol):
        kept = []
        for p in point_ids:
            if not any(np.linalg.norm(self.get_point(q) - self.get_point(p)) < tol for q in kept):
                kept.append(p)
        return kept

    def build_mesh(self, voro_file_name, K):
        """Takes the voro++ output file (plain or .gz) and the permeability tensor function K(centroid)."""
        opener = gzip.open if str(voro_file_name).endswith(".gz") else open
        with opener(voro_file_name, "rt") as fh:
            lines = [ln for ln in fh.read().split("\n") if ln.strip()]
        self.set_boundary_markers([0, 1, 2, 3, 4, 5],
                                  ["BottomX", "TopX", "BottomY", "TopY", "BottomZ,", "TopZ"])
        self.use_face_shifted_centroid()
        self.use_cell_shifted_centroid()
        for _ in lines:
            self.add_cell([], [])
        done = set()
        face_between = {}   # (first cell, second cell) -> face
        for ln in lines:
            fld = ln.split("$")
            cell = int(fld[0])
            generator = np.array([float(x) for x in fld[1].split()])
            volume = float(fld[2])
            centroid = np.array([float(x) for x in fld[3].split()])
            vertices = [np.array(v) for v in _tuples(fld[4], float)]
            faces = [list(reversed(f)) for f in _tuples(fld[5], int)]
            areas = [float(a) for a in fld[6].split()]
            normals = [np.array(v) for v in _tuples(fld[7], float)]
            neighbours = [int(x) for x in fld[8].split()]
            point_of = [self.add_point(v) for v in vertices]
            cell_faces, cell_sigma = [], []
            for lf, nb in enumerate(neighbours):
                if nb >= 0 and nb in done:
                    # the neighbour made the face; its shifted centroid is the midpoint of the two generators
                    f = face_between.get((nb, cell))
                    if f is not None:
                        cell_faces.append(f)
                        cell_sigma.append(-1)
                        self.set_face_shifted_centroid(f, (generator + self.get_cell_shifted_centroid(nb)) * .5)
                    continue
                pts = self._distinct([point_of[v] for v in faces[lf]],
                                     self.MERGE_INTERIOR if nb >= 0 else self.MERGE_BOUNDARY)
                if len(pts) <= 2:
                    continue             # a face that collapsed to an edge
                f = self.add_face(pts)
                self.set_face_normal(f, normals[lf])
                self.set_face_area(f, areas[lf])
                self.set_face_real_centroid(f, self.find_face_centroid(f)[1])
                cell_faces.append(f)
                cell_sigma.append(1)
                if nb >= 0:
                    face_between[(cell, nb)] = f
                else:
                    self.add_boundary_face(abs(nb) - 1, f, 1)
                    self.set_face_shifted_centroid(f, self.find_face_centroid(f)[1])
            self.set_cell_faces(cell, cell_faces)
            self.set_cell_domain(cell, 0)
            self.set_cell_orientation(cell, cell_sigma)
            self.set_cell_volume(cell, volume)
            self.set_cell_real_centroid(cell, centroid)
            self.set_cell_shifted_centroid(cell, generator)
            done.add(cell)
        for cell in range(self.get_number_of_cells()):
            self.set_cell_k(cell, K(np.array(self.get_cell_real_centroid(cell))))
