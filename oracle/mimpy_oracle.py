"""CPU oracle: a NumPy/SciPy restatement of the reference's MFD hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module, and only as the checker / the CPU arm.  Nothing
under mimpy_b200/ imports it; the product path fails loudly when the CUDA library is missing.

Every function cites the reference file:line (relative to /root/reference) it restates.  The
restatement is pinned (tests/test_oracle_golden.py) against golden vectors produced by the
unmodified reference running in the build container (tests/golden/gen_golden.py) and against
the analytic known answers of SURVEY.md 8c.

Third-party arithmetic on this path that is NOT under /root/reference (unpinned by the
reference: requirements.txt lists only wheel/numpy/pytest): numpy.linalg.svd / inv / norm
(LAPACK), scipy.sparse coo->csr, scipy SuperLU spsolve.  The oracle calls the same routines.
"""
import numpy as np
from scipy import sparse
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------------------
# mesh container (what the reference Mesh exposes through its getters, as flat arrays)
# --------------------------------------------------------------------------------------
class OMesh(object):
    """Flat-array view of a reference-style Mesh (mesh.py:194-302).

    ragged faces/cells are (ptr, data) pairs; `cell_sigma` is cell_normal_orientation.
    """

    def __init__(self):
        self.points = None
        self.face_ptr = None
        self.face_points = None
        self.cell_ptr = None
        self.cell_faces = None
        self.cell_sigma = None
        self.face_normals = None
        self.face_areas = None
        self.face_centroids = None
        self.face_shifted_centroids = None
        self.cell_shifted_centroids = None
        self.face_to_cell = None
        self.cell_volume = None
        self.cell_centroid = None
        self.cell_k = None
        self.boundary_faces = {}
        self.internal_no_flow = []
        # pointer couplings (mesh.py:1397-1525)
        self.dirichlet_pointers = {}   # face -> (cell, orientation)
        self.forcing_pointers = {}     # cell -> [(face, orientation), ...]
        self.face_to_lagrange = {}     # face -> (multiplier face, orientation)
        self.lagrange_to_face = {}     # multiplier face -> [(face, orientation), ...]
        self.periodic = []             # (face 1, orientation 1, face 2, orientation 2, multiplier face)

    @property
    def nc(self):
        return len(self.cell_ptr) - 1

    @property
    def nf(self):
        return len(self.face_ptr) - 1

    def cell(self, c):
        return self.cell_faces[self.cell_ptr[c]:self.cell_ptr[c + 1]]

    def sigma(self, c):
        return self.cell_sigma[self.cell_ptr[c]:self.cell_ptr[c + 1]]

    def face(self, f):
        return self.face_points[self.face_ptr[f]:self.face_ptr[f + 1]]

    @classmethod
    def from_mesh(cls, mesh):
        """Snapshot any object that follows the reference Mesh API (mesh.py:453-503, 1034-1240)."""
        o = cls()
        nf = mesh.get_number_of_faces()
        nc = mesh.get_number_of_cells()
        npnt = mesh.get_number_of_points()
        o.points = np.array(mesh.points[:npnt], dtype=np.float64)
        fl = [np.asarray(mesh.get_face(f), dtype=np.int32) for f in range(nf)]
        o.face_ptr = np.zeros(nf + 1, dtype=np.int64)
        o.face_ptr[1:] = np.cumsum([len(x) for x in fl])
        o.face_points = np.concatenate(fl).astype(np.int32) if nf else np.zeros(0, np.int32)
        cl = [np.asarray(mesh.get_cell(c), dtype=np.int32) for c in range(nc)]
        sl = [np.asarray(mesh.get_cell_normal_orientation(c), dtype=np.int32) for c in range(nc)]
        o.cell_ptr = np.zeros(nc + 1, dtype=np.int64)
        o.cell_ptr[1:] = np.cumsum([len(x) for x in cl])
        o.cell_faces = np.concatenate(cl).astype(np.int32)
        o.cell_sigma = np.concatenate(sl).astype(np.int32)
        o.face_normals = np.array(mesh.face_normals[:nf], dtype=np.float64)
        o.face_areas = np.array(mesh.face_areas[:nf], dtype=np.float64)
        o.face_centroids = np.array(mesh.face_real_centroids[:nf], dtype=np.float64)
        o.face_to_cell = np.array(mesh.face_to_cell[:nf], dtype=np.int32)
        o.cell_volume = np.array(mesh.cell_volume[:nc], dtype=np.float64)
        o.cell_centroid = np.array(mesh.cell_real_centroid[:nc], dtype=np.float64)
        o.cell_k = np.array(mesh.cell_k[:nc], dtype=np.float64)
        if mesh.is_using_face_shifted_centroid():
            o.face_shifted_centroids = np.array(mesh.face_shifted_centroids[:nf])
        if mesh.is_using_cell_shifted_centroid():
            o.cell_shifted_centroids = np.array(mesh.cell_shifted_centroid[:nc])
        o.boundary_faces = {m: [(int(a), int(b)) for a, b in mesh.get_boundary_faces_by_marker(m)]
                            for m in mesh.get_boundary_markers()}
        o.internal_no_flow = [(int(a), int(b)) for a, b in mesh.get_internal_no_flow()]
        # pointer dictionaries carry the reference's attribute names (mesh.py:271-292)
        o.dirichlet_pointers = {int(f): (int(v[0]), int(v[1])) for f, v in getattr(mesh, "dirichlet_boundary_pointers", {}).items()}
        o.forcing_pointers = {int(c): [(int(f), int(s)) for f, s in v]
                              for c, v in getattr(mesh, "forcing_function_pointers", {}).items()}
        o.face_to_lagrange = {int(f): (int(v[0]), int(v[1])) for f, v in getattr(mesh, "face_to_lagrange_pointers", {}).items()}
        o.lagrange_to_face = {int(l): [(int(f), int(s)) for f, s in v]
                              for l, v in getattr(mesh, "lagrange_to_face_pointers", {}).items()}
        o.periodic = [tuple(int(x) for x in p) for p in getattr(mesh, "periodic_boundaries", [])]
        return o


# --------------------------------------------------------------------------------------
# G2: structured hex topology in closed form (hexmesh.py:161-234, 269-350)
# --------------------------------------------------------------------------------------
def hex_face_index_tables(cx, cy, cz):
    """Face numbering of HexMesh._build_faces (hexmesh.py:172-228).

    One k->j->i sweep over the (cx+1)(cy+1)(cz+1) nodes emits, at node (i,j,k): a z-face if
    i<cx and j<cy, then a y-face if k<cz and i<cx, then an x-face if j<cy and k<cz.
    Returns three integer arrays zf[k,j,i], yf[k,j,i], xf[k,j,i] (node-indexed).
    """
    ni, nj, nk = cx + 1, cy + 1, cz + 1
    k, j, i = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
    hz = (i < cx) & (j < cy)
    hy = (k < cz) & (i < cx)
    hx = (j < cy) & (k < cz)
    per_node = hz.astype(np.int64) + hy + hx
    start = np.concatenate(([0], np.cumsum(per_node.ravel())[:-1])).reshape(per_node.shape)
    zf = np.where(hz, start, -1)
    yf = np.where(hy, start + hz, -1)
    xf = np.where(hx, start + hz + hy, -1)
    return zf, yf, xf


def hex_mesh(cx, cy, cz, dim_x, dim_y, dim_z, k_array=None, modification_function=None):
    """OMesh equal to what HexMesh.build_mesh produces (hexmesh.py:269-350).

    k_array: (nc,3,3) or (nc,9) permeabilities in cell order i + cx*j + cx*cy*k, or None (identity).
    """
    ni, nj, nk = cx + 1, cy + 1, cz + 1
    # hexmesh.py:295-297 -- dx = dim/float(n-1.)
    dx = dim_x / float(ni - 1.)
    dy = dim_y / float(nj - 1.)
    dz = dim_z / float(nk - 1.)
    kk, jj, ii = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
    pts = np.stack([ii.ravel() * dx, jj.ravel() * dy, kk.ravel() * dz], axis=1).astype(np.float64)
    if modification_function is not None:  # hexmesh.py:314-322
        pts = np.array([modification_function(p, int(a), int(b), int(c))
                        for p, a, b, c in zip(pts, ii.ravel(), jj.ravel(), kk.ravel())])

    def pid(i, j, k):  # hexmesh.py:236-243
        return i + ni * j + ni * nj * k

    zf, yf, xf = hex_face_index_tables(cx, cy, cz)
    nf = int(max(zf.max(), yf.max(), xf.max())) + 1
    faces = np.zeros((nf, 4), dtype=np.int32)
    K, J, I = kk, jj, ii
    m = zf >= 0  # hexmesh.py:178-182
    faces[zf[m]] = np.stack([pid(I, J, K)[m], pid(I + 1, J, K)[m], pid(I + 1, J + 1, K)[m], pid(I, J + 1, K)[m]], 1)
    m = yf >= 0  # hexmesh.py:195-199
    faces[yf[m]] = np.stack([pid(I, J, K)[m], pid(I, J, K + 1)[m], pid(I + 1, J, K + 1)[m], pid(I + 1, J, K)[m]], 1)
    m = xf >= 0  # hexmesh.py:211-215
    faces[xf[m]] = np.stack([pid(I, J, K)[m], pid(I, J + 1, K)[m], pid(I, J + 1, K + 1)[m], pid(I, J, K + 1)[m]], 1)

    # boundary markers, in emission order (hexmesh.py:186-226)
    bnd = {b: [] for b in range(6)}
    order = []
    for arr, lo_marker, hi_marker, axis_idx, n_axis in ((zf, 4, 5, K, nk), (yf, 2, 3, J, nj), (xf, 0, 1, I, ni)):
        msk = arr >= 0
        fid = arr[msk]
        ax = axis_idx[msk]
        order.append((fid[ax == 0], lo_marker, -1))
        order.append((fid[ax == n_axis - 1], hi_marker, 1))
    for fid, marker, orient in order:
        for f in np.sort(fid):
            bnd[marker].append((int(f), orient))

    # cells (hexmesh.py:329-342): [z-, y-, x-, x+, y+, z+], sigma [-1,-1,-1,1,1,1]
    ck, cj, ci = np.meshgrid(np.arange(cz), np.arange(cy), np.arange(cx), indexing="ij")
    ck, cj, ci = ck.ravel(), cj.ravel(), ci.ravel()
    cells = np.stack([zf[ck, cj, ci], yf[ck, cj, ci], xf[ck, cj, ci],
                      xf[ck, cj, ci + 1], yf[ck, cj + 1, ci], zf[ck + 1, cj, ci]], 1).astype(np.int32)
    nc = cells.shape[0]
    sig = np.tile(np.array([-1, -1, -1, 1, 1, 1], dtype=np.int32), nc)

    o = OMesh()
    o.points = pts
    o.face_ptr = np.arange(nf + 1, dtype=np.int64) * 4
    o.face_points = faces.ravel()
    o.cell_ptr = np.arange(nc + 1, dtype=np.int64) * 6
    o.cell_faces = cells.ravel()
    o.cell_sigma = sig
    # face_to_cell filled in add_cell order (mesh.py:1003-1012)
    f2c = -np.ones((nf, 2), dtype=np.int32)
    for c in range(nc):
        for f in cells[c]:
            if f2c[f, 0] == -1:
                f2c[f, 0] = c
            else:
                f2c[f, 1] = c
    o.face_to_cell = f2c
    o.boundary_faces = bnd
    o.face_areas = hex_face_areas(pts, faces)
    o.face_centroids = hex_face_centroids(pts, faces)
    o.face_normals = hex_face_normals(pts, faces)
    o.cell_volume, o.cell_centroid = cell_volumes_centroids(o)
    if k_array is None:
        k_array = np.tile(np.eye(3).ravel(), (nc, 1))
    o.cell_k = np.asarray(k_array, dtype=np.float64).reshape(nc, 9)
    return o


# --------------------------------------------------------------------------------------
# G3-G5: quad face geometry (hexmesh_cython.pyx:30-118, hexmesh.py:26-42)
# --------------------------------------------------------------------------------------
def hex_face_areas(points, faces):
    """Four Heron triangles about the 4-point mean (hexmesh_cython.pyx:47-88)."""
    p = points[faces]  # nf,4,3
    c = .25 * (p[:, 0] + p[:, 1] + p[:, 2] + p[:, 3])

    def dist(a, b):
        d = a - b
        return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])

    area = np.zeros(len(faces))
    for e in range(4):
        p0, p1 = p[:, e], p[:, (e + 1) % 4]
        a = dist(p0, p1)
        b = dist(p1, c)
        cc = dist(c, p0)
        s = (a + b + cc) / 2.
        area = area + np.sqrt(s * (s - a) * (s - b) * (s - cc))
    return area


def hex_face_normals(points, faces):
    """normalize((p1-p2) x (p2-p3)) (hexmesh_cython.pyx:103-117)."""
    p = points[faces]
    v1 = p[:, 0] - p[:, 1]
    v2 = p[:, 1] - p[:, 2]
    n = np.stack([v1[:, 1] * v2[:, 2] - v1[:, 2] * v2[:, 1],
                  v1[:, 2] * v2[:, 0] - v1[:, 0] * v2[:, 2],
                  v1[:, 0] * v2[:, 1] - v1[:, 1] * v2[:, 0]], 1)
    nn = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2])
    return n / nn[:, None]


def hex_face_centroids(points, faces):
    """.25*(p1+p2+p3+p4) (hexmesh.py:34-42)."""
    p = points[faces]
    return .25 * (p[:, 0] + p[:, 1] + p[:, 2] + p[:, 3])


# --------------------------------------------------------------------------------------
# G6: Mirtich cell volumes / centroids (mesh_cython.pyx:5-207)
# --------------------------------------------------------------------------------------
def face_mirtich_integrals(points, face_pts, normal):
    """Per-face projection integrals and face integrals (mesh_cython.pyx:62-150).

    Returns (A, B, C, Fa, Fb, Fc, Faa, Fbb, Fcc)."""
    an = np.abs(normal)
    if an[0] > an[1] and an[0] > an[2]:
        C = 0
    elif an[1] > an[2]:
        C = 1
    else:
        C = 2
    A = (C + 1) % 3
    B = (A + 1) % 3
    P1 = Pa = Pb = Paa = Pab = Pbb = 0.
    n = len(face_pts)
    for l in range(n):
        cur = points[face_pts[l]]
        nxt = points[face_pts[(l + 1) % n]]
        a0, b0, a1, b1 = cur[A], cur[B], nxt[A], nxt[B]
        da, db = a1 - a0, b1 - b0
        a0_2 = a0 * a0
        a0_3 = a0_2 * a0
        b0_2 = b0 * b0
        b0_3 = b0_2 * b0
        a1_2 = a1 * a1
        C1 = a1 + a0
        Ca = a1 * C1 + a0_2
        Caa = a1 * Ca + a0_3
        Cb = b1 * (b1 + b0) + b0_2
        Cbb = b1 * Cb + b0_3
        Cab = 3. * a1_2 + 2. * a1 * a0 + a0_2
        Kab = a1_2 + 2 * a1 * a0 + 3 * a0_2
        P1 += db * C1
        Pa += db * Ca
        Paa += db * Caa
        Pb += da * Cb
        Pbb += da * Cbb
        Pab += db * (b1 * Cab + b0 * Kab)
    P1 /= 2.0
    Pa /= 6.0
    Paa /= 12.0
    Pb /= -6.0
    Pbb /= -12.0
    Pab /= 24.0
    fp = points[face_pts[0]]
    w = 0.
    w -= normal[0] * fp[0]
    w -= normal[1] * fp[1]
    w -= normal[2] * fp[2]
    k1 = 1. / normal[C]
    k2 = k1 * k1
    k3 = k2 * k1
    Fa = k1 * Pa
    Fb = k1 * Pb
    Fc = -k2 * (normal[A] * Pa + normal[B] * Pb + w * P1)
    Faa = k1 * Paa
    Fbb = k1 * Pbb
    Fcc = k3 * ((normal[A] * normal[A]) * Paa + 2 * normal[A] * normal[B] * Pab +
                (normal[B] * normal[B]) * Pbb + w * (2. * (normal[A] * Pa + normal[B] * Pb) + w * P1))
    return A, B, C, Fa, Fb, Fc, Faa, Fbb, Fcc


def cell_volumes_centroids(o):
    """Face loop with scatter-add into the <=2 owners, then /(2 vol) (mesh_cython.pyx:62-207)."""
    nc, nf = o.nc, o.nf
    vol = np.zeros(nc)
    cen = np.zeros((nc, 3))
    for f in range(nf):
        nrm = o.face_normals[f]
        A, B, C, Fa, Fb, Fc, Faa, Fbb, Fcc = face_mirtich_integrals(o.points, o.face(f), nrm)
        for slot in range(2):
            c = o.face_to_cell[f, slot]
            if c < 0:
                continue
            loc = -1
            cf = o.cell(c)
            for idx in range(len(cf)):  # last match wins (mesh_cython.pyx:160-162)
                if cf[idx] == f:
                    loc = idx
            s = o.sigma(c)[loc]
            if A == 0:
                vol[c] += nrm[0] * Fa * s
            elif B == 0:
                vol[c] += nrm[0] * Fb * s
            else:
                vol[c] += nrm[0] * Fc * s
            cen[c, A] += nrm[A] * Faa * s
            cen[c, B] += nrm[B] * Fbb * s
            cen[c, C] += nrm[C] * Fcc * s
    cen = cen / (vol * 2.)[:, None]
    return vol, cen


def rebuild_face_to_cell(o):
    """face -> (<=2) cells derived from the cell lists, in cell order.

    The reference's own face_to_cell is NOT reliable after divide_cell_by_plane: the interior
    cut face ends up registered to the new cell only (mesh.py:2405-2410 removes the old cell from
    every face of the second half, including the shared cut face)."""
    f2c = -np.ones((o.nf, 2), dtype=np.int32)
    for c in range(o.nc):
        for f in o.cell(c):
            if f2c[f, 0] == -1:
                f2c[f, 0] = c
            elif f2c[f, 1] == -1 and f2c[f, 0] != c:
                f2c[f, 1] = c
    return f2c


def volume_centroid_from_faces(o, c):
    """Per-cell Mirtich with the normal flipped / traversal reversed by the orientation
    (mesh.py:1783-1882, the Python twin used by divide_cell_by_plane, mesh.py:2401-2415)."""
    volume = 0.
    centroid = np.zeros(3)
    for f, s in zip(o.cell(c), o.sigma(c)):
        nrm = o.face_normals[f] * s
        pts = o.face(f)
        if s > 0:
            order = list(pts)
        else:  # mesh.py:1814-1816: edges run next->current, i.e. the reversed polygon
            order = list(pts[1:]) + list(pts[:1])
            order = order[::-1]
            order = order[-1:] + order[:-1]
        A, B, C, Fa, Fb, Fc, Faa, Fbb, Fcc = face_mirtich_integrals(o.points, np.array(order), nrm)
        # w uses the face's first stored point (mesh.py:1852-1853); same plane => same w
        if A == 0:
            volume += nrm[0] * Fa
        elif B == 0:
            volume += nrm[0] * Fb
        else:
            volume += nrm[0] * Fc
        centroid[A] += nrm[A] * Faa
        centroid[B] += nrm[B] * Fbb
        centroid[C] += nrm[C] * Fcc
    centroid /= volume * 2.
    return volume, centroid


# --------------------------------------------------------------------------------------
# L1-L4: local matrices (mfd.py:248-483)
# --------------------------------------------------------------------------------------
def build_r_e(o, c):
    """R[f,:] = a_f (x_f - x_E) (mfd.py:248-285)."""
    xc = o.cell_shifted_centroids[c] if o.cell_shifted_centroids is not None else o.cell_centroid[c]
    fc = o.face_shifted_centroids if o.face_shifted_centroids is not None else o.face_centroids
    f = o.cell(c)
    return o.face_areas[f][:, None] * (fc[f] - xc[None, :])


def build_n_e(o, c, k_unity=False):
    """N[f,:] = sigma_f K n_f (mfd.py:287-312)."""
    f = o.cell(c)
    s = o.sigma(c).astype(np.float64)
    nrm = o.face_normals[f]
    if k_unity:
        return nrm * s[:, None]
    k = o.cell_k[c].reshape(3, 3)
    return (nrm @ k.T) * s[:, None]


def build_c_e(n_e):
    """Orthonormal basis of null(N^T) from the SVD (mfd.py:314-322)."""
    u, s, v = np.linalg.svd(n_e.T)
    return v.T[:, 3:]


def build_m_e(o, c, method=0, k_unity=False):
    """Local inner-product matrix M_E, all 8 construction methods (mfd.py:357-483)."""
    r_e = build_r_e(o, c)
    n_e = build_n_e(o, c, k_unity)
    k = np.eye(3) if k_unity else o.cell_k[c].reshape(3, 3)
    f = o.cell(c)
    nfc = len(f)
    c_e = build_c_e(n_e)
    t_inv = np.linalg.inv(r_e.T @ n_e)
    m_0 = r_e @ (t_inv @ r_e.T)  # mfd.py:377
    if method == 0:  # mfd.py:379-384
        u = ((r_e.T @ r_e) @ t_inv)[0, 0]
        m_e = m_0 + c_e @ (u * c_e.T)
    elif method == 1:  # mfd.py:386-389
        u = o.cell_volume[c] / np.trace(k)
        m_e = m_0 + c_e @ (u * c_e.T)
    elif method == 2:  # mfd.py:391-404
        k_inv = np.linalg.inv(k)
        u = 0.
        for fi in f:
            d = o.face_centroids[fi] - o.cell_centroid[c]
            u += d @ (k_inv @ d)
        u /= 3 ** 2
        u /= nfc
        m_e = m_0 + c_e @ (u * c_e.T)
    elif method == 3:  # mfd.py:406-418
        lam = .5 * np.trace(r_e @ (t_inv @ r_e.T))
        m_e = np.eye(nfc) - (n_e @ np.linalg.inv(n_e.T @ n_e)) @ n_e.T
        m_e = lam * m_e + m_0
    elif method == 4:  # mfd.py:420-421, 338-355
        u_, s_, v_ = np.linalg.svd(r_e.T)
        d_e = v_.T[:, 3:]
        w_e = n_e @ (np.linalg.inv(n_e.T @ r_e) @ n_e.T)
        w_e = w_e + d_e @ d_e.T
        w_e = w_e * np.trace(o.cell_k[c].reshape(3, 3))
        w_e = w_e / o.cell_volume[c]
        m_e = np.linalg.inv(w_e)
    elif method == 5:  # mfd.py:423-438
        diag = []
        for i in range(nfc):
            mi = 0
            cur = abs(r_e[i, 0])
            for a in range(1, 3):
                if abs(r_e[i, a]) > cur:
                    mi = a
                    cur = abs(r_e[i, a])
            diag.append(r_e[i, mi] / n_e[i, mi])
        m_e = np.diag(diag)
    elif method == 6:  # mfd.py:440-456
        k_inv = np.linalg.inv(k)
        sig = o.sigma(c)
        dvec = np.zeros(nfc)
        for i, fi in enumerate(f):
            nrm = sig[i] * o.face_normals[fi]
            e = nrm @ (k_inv @ nrm)
            e *= o.face_areas[fi]
            e *= np.linalg.norm(o.face_centroids[fi] - o.cell_centroid[c])
            dvec[i] = e
        m_e = m_0 + np.diag(dvec) @ (c_e @ c_e.T)
    elif method == 7:  # mfd.py:460-468
        m_e = np.zeros((nfc, nfc))
        for i, fi in enumerate(f):
            m_e[i, i] = np.linalg.norm(r_e[i, :]) * o.face_areas[fi] / k[0, 0]
    else:
        raise ValueError("unknown M_E construction method %r" % (method,))
    return m_e


def build_m_e_all(o, method=0, k_unity=False):
    """Per-cell list of M_E (the loop of mfd.py:567-568)."""
    return [build_m_e(o, c, method, k_unity) for c in range(o.nc)]


def build_m_e_batched(o, method=1):
    """Vectorised restatement for UNIFORM face count (hex), methods 0/1/6 -- same math as
    build_m_e (batched np.linalg.svd/inv); validated against it in tests. Returns (nc,n,n)."""
    nc = o.nc
    n = int(o.cell_ptr[1] - o.cell_ptr[0])
    assert np.all(np.diff(o.cell_ptr) == n)
    f = o.cell_faces.reshape(nc, n)
    s = o.cell_sigma.reshape(nc, n).astype(np.float64)
    k = o.cell_k.reshape(nc, 3, 3)
    r = o.face_areas[f][:, :, None] * (o.face_centroids[f] - o.cell_centroid[:, None, :])
    nn = np.einsum("cab,cfb->cfa", k, o.face_normals[f]) * s[:, :, None]
    t_inv = np.linalg.inv(np.einsum("cfa,cfb->cab", r, nn))
    m0 = np.einsum("cia,cab,cjb->cij", r, t_inv, r)
    u_, s_, vt = np.linalg.svd(np.transpose(nn, (0, 2, 1)))
    ce = np.transpose(vt, (0, 2, 1))[:, :, 3:]
    proj = np.einsum("cia,cja->cij", ce, ce)
    if method == 0:
        u = np.einsum("cfa,cfb,cbd->cad", r, r, t_inv)[:, 0, 0]
        return m0 + u[:, None, None] * proj
    if method == 1:
        u = o.cell_volume / np.trace(k, axis1=1, axis2=2)
        return m0 + u[:, None, None] * proj
    if method == 6:
        kinv = np.linalg.inv(k)
        nrm = o.face_normals[f] * s[:, :, None]
        e = np.einsum("cfa,cab,cfb->cf", nrm, kinv, nrm) * o.face_areas[f]
        e = e * np.linalg.norm(o.face_centroids[f] - o.cell_centroid[:, None, :], axis=2)
        return m0 + e[:, :, None] * proj
    raise ValueError("batched oracle covers methods 0, 1, 6")


# --------------------------------------------------------------------------------------
# A1-A7: boundary data, DOF numbering, global assembly, RHS (mfd.py:141-246, 545-599, 745-914, 991-1215)
# --------------------------------------------------------------------------------------
class OracleMFD(object):
    """Restatement of the reference MFD object for the hot path (mfd.py:33-1529)."""

    def __init__(self, o, method=0):
        self.o = o
        self.method = method
        self.dirichlet_boundary_values = {}
        self.neumann_boundary_values = {}
        self.cell_faces_neumann = {}
        self.cell_forcing_function = np.zeros(o.nc)
        self.face_to_flux = None
        self.flux_dof = 0
        self.m_e_locations = None
        self.m_data_for_update = None
        self.lhs = None
        self.rhs = None
        self.solution = None
        # process_internal_boundaries (mfd.py:1194-1207)
        for f, _orient in o.internal_no_flow:
            c = int(o.face_to_cell[f][0])
            self.cell_faces_neumann.setdefault(c, []).append(f)
            self.neumann_boundary_values[f] = 0.

    # --- boundary conditions / source -------------------------------------------------
    def apply_dirichlet_from_function(self, marker, p_function):
        """value = p(x_f) a_f sigma_bdry (mfd.py:1143-1162)."""
        o = self.o
        for f, orient in o.boundary_faces[marker]:
            v = p_function(o.face_centroids[f])
            v *= o.face_areas[f]
            self.dirichlet_boundary_values[f] = v * orient

    def apply_neumann_from_function(self, marker, grad_u):
        """value = grad_u(x_f).n_f; registers the face as Neumann (mfd.py:1060-1090)."""
        o = self.o
        for f, _orient in o.boundary_faces[marker]:
            self.neumann_boundary_values[f] = np.asarray(grad_u(o.face_centroids[f])).dot(o.face_normals[f])
            c = int(o.face_to_cell[f][0])
            self.cell_faces_neumann.setdefault(c, []).append(f)

    def apply_forcing_from_function(self, forcing_function):
        """F_c = f(x_E) vol(E) (mfd.py:1164-1176)."""
        o = self.o
        for c in range(o.nc):
            self.cell_forcing_function[c] = forcing_function(o.cell_centroid[c]) * o.cell_volume[c]

    # --- numbering ----------------------------------------------------------------------
    def renumber_for_neumann(self):
        """face_to_flux[f] = f - #Neumann before f, -1 on Neumann faces (mfd.py:141-159)."""
        nf = self.o.nf
        is_n = np.zeros(nf, dtype=bool)
        if self.neumann_boundary_values:
            is_n[np.fromiter(self.neumann_boundary_values.keys(), dtype=np.int64)] = True
        off = np.cumsum(is_n)
        f2f = (np.arange(nf) - off).astype(np.int32)
        f2f[is_n] = -1
        self.face_to_flux = f2f.reshape(nf, 1)
        self.flux_dof = int(nf - is_n.sum())

    # --- blocks -------------------------------------------------------------------------
    def build_m(self, save_update_info=False, k_unity=False, m_e_list=None):
        """Cell-major COO triples of M with the 1e-20 relative filter (mfd.py:545-599)."""
        o = self.o
        self.renumber_for_neumann()
        data, row, col = [], [], []
        locs = [0]
        g = self.face_to_flux[:, 0]
        for c in range(o.nc):
            m_e = m_e_list[c] if m_e_list is not None else build_m_e(o, c, self.method, k_unity)
            nrm = np.linalg.norm(m_e)
            neu = self.cell_faces_neumann.get(c, [])
            f = o.cell(c)
            s = o.sigma(c)
            for i in range(len(m_e)):
                if f[i] in neu:
                    continue
                for j in range(len(m_e)):
                    if f[j] in neu:
                        continue
                    if abs(m_e[i][j]) / nrm > 1.e-20:
                        data.append(m_e[i, j] * s[i] * s[j])
                        row.append(g[f[i]])
                        col.append(g[f[j]])
            locs.append(len(data))
        if save_update_info:
            self.m_data_for_update = np.array(data)
            self.m_e_locations = np.array(locs)
        return [np.array(data, dtype=np.float64), np.array(row, dtype=np.int32), np.array(col, dtype=np.int32)]

    def build_div(self, shift=1):
        """DIV and -DIV^T COO triples, skipping Neumann faces (mfd.py:203-246)."""
        o = self.o
        if self.face_to_flux is None:
            self.renumber_for_neumann()
        g = self.face_to_flux[:, 0]
        d, r, cc = [], [], []
        for c in range(o.nc):
            neu = self.cell_faces_neumann.get(c, [])
            for f, s in zip(o.cell(c), o.sigma(c)):
                if f in neu:
                    continue
                d.append(s * o.face_areas[f])
                r.append(c + shift * self.flux_dof)
                cc.append(g[f])
        d = np.array(d, dtype=np.float64)
        r = np.array(r, dtype=np.int32)
        cc = np.array(cc, dtype=np.int32)
        return [[d, r, cc], [-d, cc.copy(), r.copy()]]

    def build_bottom_right(self, alpha=0., shift=1):
        """C[c,c] = alpha vol(E), explicit even when zero (mfd.py:745-774)."""
        o = self.o
        idx = (np.arange(o.nc) + shift * self.flux_dof).astype(np.int32)
        return [alpha * o.cell_volume, idx, idx.copy()]

    def build_coupling_terms(self):
        """COO triples of the pointer couplings L, L^T in the reference's order: forcing pointers, Dirichlet
        pointers, face -> multiplier, multiplier -> faces, periodic pairs (mfd.py:776-849)."""
        o, F = self.o, self.flux_dof
        g = self.face_to_flux[:, 0]
        data, row, col = [], [], []
        for cell, pointers in o.forcing_pointers.items():          # the cell's source = flux through the faces
            for f, orient in pointers:
                data.append(-o.face_areas[f] * orient)
                row.append(cell + F)
                col.append(g[f])
        for f, (cell, orient) in o.dirichlet_pointers.items():      # the face's pressure = the cell's
            data.append(o.face_areas[f] * orient)
            row.append(g[f])
            col.append(cell + F)
        for f, (lag, orient) in o.face_to_lagrange.items():         # the face's pressure = the multiplier
            data.append(o.face_areas[f] * orient)
            row.append(g[f])
            col.append(g[lag])
        for lag, pointers in o.lagrange_to_face.items():            # the multiplier's equation: fluxes balance
            for f, orient in pointers:
                data.append(-o.face_areas[f] * orient)
                row.append(g[lag])
                col.append(g[f])
        for f1, o1, f2, o2, lag in o.periodic:
            data += [o.face_areas[f1] * o1, o.face_areas[f2] * o2, o1, o2]
            row += [g[f1], g[f2], g[lag], g[lag]]
            col += [g[lag], g[lag], g[f1], g[f2]]
        return [np.asarray(data, dtype=np.float64), np.asarray(row, dtype=np.int32), np.asarray(col, dtype=np.int32)]

    def build_lhs(self, alpha=0., m_e_list=None):
        """M | DIV | -DIV^T | C | couplings concatenated into one COO matrix (mfd.py:851-914)."""
        m = self.build_m(m_e_list=m_e_list)
        dv, dvt = self.build_div()
        c = self.build_bottom_right(alpha)
        cpl = self.build_coupling_terms()
        data = np.concatenate([m[0], dv[0], dvt[0], c[0], cpl[0]])
        row = np.concatenate([m[1], dv[1], dvt[1], c[1], cpl[1]])
        col = np.concatenate([m[2], dv[2], dvt[2], c[2], cpl[2]])
        self.lhs = sparse.coo_matrix((data, (row, col)))
        return self.lhs

    def build_rhs(self):
        """Neumann (rhs[g]=value, g=-1 writes rhs[-1]), Dirichlet (rhs[g]=-value), forcing (+=)
        (mfd.py:991-999, 1035-1041, 1209-1215, 1027-1033)."""
        o = self.o
        rhs = np.zeros(o.nc + self.flux_dof)
        g = self.face_to_flux[:, 0]
        for f, v in self.neumann_boundary_values.items():
            rhs[g[f]] = v
        for f, v in self.dirichlet_boundary_values.items():
            rhs[g[f]] = -v
        rhs[self.flux_dof:] += self.cell_forcing_function
        self.rhs = rhs
        return rhs

    def solve(self):
        """SuperLU on the CSC matrix (mfd.py:1349-1358)."""
        self.solution = spla.spsolve(self.lhs.tocsc(), self.rhs)
        return self.solution

    def get_pressure_solution(self):
        return self.solution[self.flux_dof:]

    def update_m(self, m_coo, multipliers):
        """m_coo[k] = m_data_for_update[k]*multipliers[cell] (mfd_cython.pyx:6-22)."""
        loc = self.m_e_locations
        counts = np.diff(loc)
        m_coo[:loc[-1]] = self.m_data_for_update * np.repeat(multipliers, counts)


# --------------------------------------------------------------------------------------
# T1-T6: two-phase IMPES (twophase.py:268-344, 476-560, 703-1092)
# --------------------------------------------------------------------------------------
class OracleTwoPhase(object):
    """Restatement of the TwoPhase hot path (no pointer couplings, no PETSc)."""

    def __init__(self, o):
        self.o = o
        self.mfd = OracleMFD(o, method=6)  # twophase.py:231
        self.saturation_boundaries = {}
        self.rate_wells = []
        self.rate_wells_rate_water = []
        self.rate_wells_rate_oil = []
        self.pressure_wells = []
        self.pressure_wells_bhp = []
        self.pressure_wells_pi = []
        self.delta_t = 1.
        self.saturation_time_steps = 1
        self.newton_threshold = 1.e-6
        self.newton_step_max = 10
        self.k_rw_o, self.n_w, self.n_o = 1., 2., 2.
        self.newton_history = []

    # Corey relperm + mobilities (twophase.py:268-344)
    def set_corey_relperm(self, n_o, n_w, k_rw_o=1.):
        self.n_o, self.n_w, self.k_rw_o = n_o, n_w, k_rw_o

    def krw(self, se):
        return self.k_rw_o * (np.clip(se, 0., 1.) ** self.n_w)

    def kro(self, se):
        return (1.0 - np.clip(se, 0., 1.)) ** self.n_o

    def sefromsw(self, sw):
        return (sw - self.residual_saturation_water) / (1. - self.residual_saturation_water - self.residual_saturation_oil)

    def water_mobility(self, sw, p):
        return self.krw(self.sefromsw(sw)) / self.viscosity_water * self.ref_density_water * (1. + self.compressibility_water * p)

    def oil_mobility(self, sw, p):
        return self.kro(self.sefromsw(sw)) / self.viscosity_oil * self.ref_density_oil * (1. + self.compressibility_oil * p)

    def add_rate_well(self, qw, qo, cell):
        self.rate_wells.append(cell)
        self.rate_wells_rate_water.append(qw)
        self.rate_wells_rate_oil.append(qo)

    def initialize_system(self):
        """twophase.py:476-560."""
        mfd = self.mfd
        mfd.renumber_for_neumann()
        dv, dvt = mfd.build_div()
        m = mfd.build_m(save_update_info=True)
        self.current_u_t = np.zeros(mfd.flux_dof)
        self.div = sparse.coo_matrix((dv[0], (dv[1] - mfd.flux_dof, dv[2])),
                                     shape=(self.o.nc, mfd.flux_dof)).tocsr()
        self.m_x_coo_length = len(m[0])
        c = mfd.build_bottom_right()
        data = np.concatenate([m[0], dv[0], dvt[0]])
        self.c_start = len(data)
        data = np.concatenate([data, c[0]])
        self.c_end = self.c_start + len(c[0])
        row = np.concatenate([m[1], dv[1], dvt[1], c[1]])
        col = np.concatenate([m[2], dv[2], dvt[2], c[2]])
        self.lhs_coo = sparse.coo_matrix((data, (row, col)))
        self.rhs_mfd = mfd.build_rhs()

    def update_pressure(self):
        """Newton loop (twophase.py:744-926), scipy solver branch."""
        o, mfd = self.o, self.mfd
        po_k = np.array(self.current_p_o)
        ut_k = np.array(self.current_u_t)
        res = 100.
        step = 0
        vol = o.cell_volume
        while abs(res > self.newton_threshold):
            mob = self.water_mobility(self.current_s_w, po_k) + self.oil_mobility(self.current_s_w, po_k)
            mob = 1. / mob
            cmat = self.ref_density_water * self.current_s_w * self.compressibility_water
            cmat = cmat + self.ref_density_oil * (self.compressibility_oil * (1. - self.current_s_w))
            cmat = cmat * self.porosities * vol / self.delta_t
            mfd.update_m(self.lhs_coo.data[:self.m_x_coo_length], mob)
            for cell, pi in zip(self.pressure_wells, self.pressure_wells_pi):
                cmat[cell] += pi * 1. / mob[cell]
            self.lhs_coo.data[self.c_start:self.c_end] = cmat
            lhs = self.lhs_coo.tocsr()
            rhs = -mfd.build_rhs()
            rhs += lhs.dot(np.concatenate((ut_k, po_k)))
            f1 = self.ref_density_water * self.current_s_w * self.porosities / self.delta_t * vol
            f2 = self.ref_density_oil * (1. - self.current_s_w) * self.porosities / self.delta_t * vol
            f3 = self.ref_density_water * (1. + self.compressibility_water * self.current_p_o) * self.current_s_w
            f3 = f3 + self.ref_density_oil * (1 + self.compressibility_oil * self.current_p_o) * (1. - self.current_s_w)
            f3 = f3 * self.porosities / self.delta_t * vol
            rhs[mfd.flux_dof:] += f1
            rhs[mfd.flux_dof:] += f2
            rhs[mfd.flux_dof:] -= f3
            for w, cell in enumerate(self.rate_wells):
                rhs[cell + mfd.flux_dof] += -self.rate_wells_rate_water[w]
                rhs[cell + mfd.flux_dof] += -self.rate_wells_rate_oil[w]
            for cell, bhp, pi in zip(self.pressure_wells, self.pressure_wells_bhp, self.pressure_wells_pi):
                rhs[cell + mfd.flux_dof] -= pi * bhp * 1. / mob[cell]
            res = np.linalg.norm(rhs) / float(len(rhs))
            self.newton_history.append(res)
            if res > self.newton_threshold:
                sol = spla.spsolve(lhs.tocsc(), -rhs)
                po_k += sol[mfd.flux_dof:]
                ut_k += sol[:mfd.flux_dof]
            step += 1
            if step > self.newton_step_max:
                raise ZeroDivisionError("newton did not converge (twophase.py:919-920)")
        self.previous_p_o = np.array(self.current_p_o)
        self.previous_u_t = np.array(self.current_u_t)
        self.current_p_o = po_k
        self.current_u_t = ut_k

    def find_upwinding_direction(self):
        """(flux_index, cell) for sigma*u_t[g] > 0, INCLUDING the g=-1 quirk (twophase.py:928-958)."""
        o, mfd = self.o, self.mfd
        g = mfd.face_to_flux[:, 0]
        pairs = []
        for c in range(o.nc):
            for f, s in zip(o.cell(c), o.sigma(c)):
                gi = g[f]
                if s * self.current_u_t[gi] > 0:
                    pairs.append((int(gi), c))
        self.upwinded_face_cell = pairs
        self.current_f_w = np.zeros(mfd.flux_dof)
        self.current_u_w = np.zeros(mfd.flux_dof)

    def update_saturation(self):
        """Explicit upwind saturation update (twophase.py:960-1092), wells of both kinds."""
        o, mfd = self.o, self.mfd
        dt = self.delta_t / float(self.saturation_time_steps)
        wm = self.water_mobility(self.current_s_w, self.current_p_o)
        om = self.oil_mobility(self.current_s_w, self.current_p_o)
        for gi, c in self.upwinded_face_cell:
            self.current_f_w[gi] = wm[c] / (wm[c] + om[c])
        g = mfd.face_to_flux[:, 0]
        for marker, sfun in self.saturation_boundaries.items():
            for f, orient in o.boundary_faces[marker]:
                gi = g[f]
                if orient * self.current_u_t[gi] < 0:
                    s = np.array([sfun(o.face_centroids[f])])
                    p = np.array([mfd.dirichlet_boundary_values[f]])
                    w = self.water_mobility(s, p)
                    self.current_f_w[gi] = (w / (w + self.oil_mobility(s, p)))[0]
        self.current_u_w[:] = self.current_f_w * self.current_u_t
        div_uw = self.div.dot(self.current_u_w)
        pw, ppw = self.current_p_o, self.previous_p_o
        cw = self.compressibility_water
        nxt = (ppw * cw + 1.) * self.ref_density_water * self.current_s_w / self.ref_density_water / (1. + cw * pw)
        t = 1. / self.porosities / self.ref_density_water / (1. + cw * pw) / o.cell_volume * div_uw * (-dt)
        nxt = nxt + t
        for w, cell in enumerate(self.rate_wells):
            wi = self.rate_wells.index(cell)
            v = dt / self.porosities[cell] / (1. + cw * pw[cell]) / self.ref_density_water
            v *= self.rate_wells_rate_water[wi]
            v /= o.cell_volume[cell]
            nxt[cell] += v
        for w in range(len(self.pressure_wells)):
            cell = self.pressure_wells[w]
            v = self.pressure_wells_pi[w] * (self.pressure_wells_bhp[w] - self.current_p_o[cell]) * dt * wm[cell]
            v = v / self.porosities[cell] / self.ref_density_water / (1. + cw * pw[cell]) / o.cell_volume[cell]
            nxt[cell] += v
        self.current_s_w = nxt

    def start_solving(self, number_of_time_steps):
        """IMPES driver (twophase.py:703-742)."""
        for step in range(1, number_of_time_steps + 1):
            self.update_pressure()
            if step == 1 or step % 10 == 0:
                self.find_upwinding_direction()
            for _ in range(self.saturation_time_steps):
                self.update_saturation()


# --------------------------------------------------------------------------------------
# P1: single-phase slightly compressible flow (singlephase.py:88-96, 182-299)
# --------------------------------------------------------------------------------------
class OracleSinglePhase(object):
    """Restatement of SinglePhase.initialize_system / start_solving (implicit time stepping)."""

    def __init__(self, o):
        self.o = o
        self.mfd = OracleMFD(o, method=1)  # singlephase.py:95
        self.rate_wells = []
        self.rate_wells_rate = []
        self.delta_t = 1.

    def add_point_rate_well(self, rate, cell):
        self.rate_wells.append(cell)
        self.rate_wells_rate.append(rate)

    def initialize_system(self):
        """singlephase.py:182-237"""
        mfd = self.mfd
        mfd.renumber_for_neumann()
        dv, dvt = mfd.build_div()
        m = mfd.build_m(save_update_info=True)
        self.m_x_coo_length = len(m[0])
        c = mfd.build_bottom_right()
        data = np.concatenate([m[0], dv[0], dvt[0]])
        self.c_start = len(data)
        cpl = mfd.build_coupling_terms()   # singlephase.py:199-223: appended after the C block
        data = np.concatenate([data, c[0], cpl[0]])
        row = np.concatenate([m[1], dv[1], dvt[1], c[1], cpl[1]])
        col = np.concatenate([m[2], dv[2], dvt[2], c[2], cpl[2]])
        self.lhs_coo = sparse.coo_matrix((data, (row, col)))
        self.rhs_mfd = mfd.build_rhs()

    def step(self):
        """One pass of the loop body of start_solving (singlephase.py:256-291)."""
        o, mfd = self.o, self.mfd
        rhs = np.zeros(o.nc + mfd.flux_dof)
        rhs += self.rhs_mfd
        density = (-self.ref_pressure + self.current_pressure) * self.compressibility
        density = (density + 1.) * self.ref_density
        mult = self.viscosity / density
        c_entry = self.compressibility * self.porosities / self.delta_t * o.cell_volume
        rhs[mfd.flux_dof:] += c_entry * self.current_pressure
        self.lhs_coo.data[self.c_start:self.c_start + o.nc] = c_entry
        for w, cell in enumerate(self.rate_wells):
            rhs[mfd.flux_dof + cell] += self.rate_wells_rate[w]
        mfd.update_m(self.lhs_coo.data[:self.m_x_coo_length], mult)
        sol = spla.spsolve(self.lhs_coo.tocsr().tocsc(), rhs)
        self.current_pressure = sol[mfd.flux_dof:]
        self.current_velocity = sol[:mfd.flux_dof]


class OracleTransport(OracleSinglePhase):
    """Restatement of Transport (transport.py:186-391): the pressure step is SinglePhase's
    (transport.py:261-301 == singlephase.py:256-291), followed by an explicit upwind update of a
    concentration.  Like the reference it only makes sense without Neumann faces (transport.py:205-209
    and 276-277 index the system with the number of faces, not flux_dof)."""

    def __init__(self, o):
        OracleSinglePhase.__init__(self, o)
        self.concentration_boundaries = {}

    def step(self):
        self.prev_pressure = self.current_pressure
        OracleSinglePhase.step(self)
        self.find_upwinding_direction()
        self.update_concentration()

    def find_upwinding_direction(self):
        """transport.py:346-361 (no Dirichlet pointers)"""
        o = self.o
        self.upwinded_face_cell = []
        for c in range(o.nc):
            for f, s in zip(o.cell(c), o.sigma(c)):
                if s * self.current_velocity[f] > 0:
                    self.upwinded_face_cell.append((f, c))

    def update_concentration(self):
        """transport.py:303-344; the result is NOT divided by the cell volume (the line is commented
        out in the reference, transport.py:344)."""
        o = self.o
        c_uw = np.zeros(o.nf)
        for f, c in self.upwinded_face_cell:
            c_uw[f] = self.current_concentration[c]
        for marker, fn in self.concentration_boundaries.items():
            for f, s in o.boundary_faces[marker]:
                if s * self.current_velocity[f] < 0:
                    c_uw[f] = fn(o.face_centroids[f])
        c_uw *= self.current_velocity
        div_c = np.zeros(o.nc)
        for c in range(o.nc):
            for f, s in zip(o.cell(c), o.sigma(c)):
                div_c[c] += s * o.face_areas[f] * c_uw[f]          # DIV of mfd.py:203-246
        prev_density = ((self.prev_pressure - self.ref_pressure) * self.compressibility + 1.) * self.ref_density
        cur_density = ((self.current_pressure - self.ref_pressure) * self.compressibility + 1.) * self.ref_density
        conc = self.current_concentration * (prev_density * self.porosities)
        conc = conc / self.delta_t
        conc = conc - div_c
        conc = conc * (self.delta_t / self.porosities)
        self.current_concentration = conc / cur_density


# --------------------------------------------------------------------------------------
# (de)serialisation of OMesh for the golden fixtures
# --------------------------------------------------------------------------------------
_OMESH_ARRAYS = ("points", "face_ptr", "face_points", "cell_ptr", "cell_faces", "cell_sigma",
                 "face_normals", "face_areas", "face_centroids", "face_to_cell", "cell_volume",
                 "cell_centroid", "cell_k")


def omesh_to_dict(o, prefix="mesh_"):
    d = {prefix + k: getattr(o, k) for k in _OMESH_ARRAYS}
    markers = sorted(o.boundary_faces.keys())
    d[prefix + "bnd_markers"] = np.array(markers, dtype=np.int32)
    for m in markers:
        d[prefix + "bnd_%d" % m] = np.array(o.boundary_faces[m], dtype=np.int32).reshape(-1, 2)
    d[prefix + "internal_no_flow"] = np.array(o.internal_no_flow, dtype=np.int32).reshape(-1, 2)
    return d


def omesh_from_dict(d, prefix="mesh_"):
    o = OMesh()
    for k in _OMESH_ARRAYS:
        setattr(o, k, np.array(d[prefix + k]))
    o.boundary_faces = {int(m): [(int(a), int(b)) for a, b in d[prefix + "bnd_%d" % int(m)]]
                        for m in d[prefix + "bnd_markers"]}
    o.internal_no_flow = [(int(a), int(b)) for a, b in d[prefix + "internal_no_flow"]]
    return o
