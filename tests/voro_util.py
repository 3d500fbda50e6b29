"""Writes a bounded Voronoi tessellation in the text format mimpy's VoroMesh reads (voromesh.py:13-17:
voro++ -c "%q$%v$%C$%P$%t$%f$%l$%n" ...), without voro++: scipy's Qhull wrapper on the generator points plus
their mirror images across the six walls of the box (the cell of a point then ends exactly at the walls).

Test infrastructure only: tests/golden/gen_golden_r2.py uses it to make the committed fixture
tests/golden/voro_48.vol.gz."""
import numpy as np


def jittered_points(nx, ny, nz, box, seed, jitter=0.3):
    rng = np.random.default_rng(seed)
    h = np.array([(box[d][1] - box[d][0]) / n for d, n in enumerate((nx, ny, nz))])
    pts = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = np.array([box[0][0] + (i + .5) * h[0], box[1][0] + (j + .5) * h[1], box[2][0] + (k + .5) * h[2]])
                pts.append(c + jitter * h * rng.uniform(-1., 1., 3))
    return np.array(pts)


def _ordered(verts, normal):
    """Indices that walk the convex polygon `verts` counter-clockwise seen against `normal`."""
    c = verts.mean(axis=0)
    a = verts[0] - c
    a /= np.linalg.norm(a)
    b = np.cross(normal, a)
    ang = np.arctan2((verts - c) @ b, (verts - c) @ a)
    return np.argsort(ang)


def voronoi_vol_text(points, box):
    from scipy.spatial import Voronoi

    n = len(points)
    blocks = [points]
    for d in range(3):
        for side in range(2):
            m = points.copy()
            m[:, d] = 2. * box[d][side] - m[:, d]
            blocks.append(m)            # block 1 + 2 d + side  <->  voro++ wall id -(1 + 2 d + side)
    vor = Voronoi(np.vstack(blocks))
    faces_of = [[] for _ in range(n)]
    for (p, q), rv in zip(vor.ridge_points, vor.ridge_vertices):
        if -1 in rv:
            continue
        for a, b in ((p, q), (q, p)):
            if a >= n:
                continue
            nb = b if b < n else (-(1 + (b // n - 1)) if b % n == a else None)
            if nb is None:
                continue                 # (a degenerate ridge with the mirror image of another point)
            faces_of[a].append((nb, list(rv), vor.points[b] - vor.points[a]))
    lines = []
    for c in range(n):
        vid = {}
        verts = []
        flist, areas, normals, nbrs = [], [], [], []
        vol, cen = 0., np.zeros(3)
        x0 = points[c]
        for nb, rv, direction in faces_of[c]:
            nrm = direction / np.linalg.norm(direction)
            pv = vor.vertices[rv]
            order = _ordered(pv, nrm)
            pv = pv[order]
            # area and centroid of the polygon (fan from its first vertex)
            area, fcen = 0., np.zeros(3)
            for t in range(1, len(pv) - 1):
                tri = .5 * np.linalg.norm(np.cross(pv[t] - pv[0], pv[t + 1] - pv[0]))
                area += tri
                fcen += tri * (pv[0] + pv[t] + pv[t + 1]) / 3.
            if area < 1e-12:
                continue
            fcen /= area
            hgt = float(np.dot(fcen - x0, nrm))
            pyr = area * hgt / 3.
            vol += pyr
            cen += pyr * (x0 + .75 * (fcen - x0))
            idx = []
            for v, gi in zip(pv, np.array(rv)[order]):
                if gi not in vid:
                    vid[gi] = len(verts)
                    verts.append(v)
                idx.append(vid[gi])
            flist.append(idx)
            areas.append(area)
            normals.append(nrm)
            nbrs.append(nb)
        cen /= vol
        fmt = lambda v: ",".join("%.17g" % x for x in v)
        lines.append("$".join([
            str(c), " ".join("%.17g" % x for x in x0), "%.17g" % vol, " ".join("%.17g" % x for x in cen),
            " ".join("(%s)" % fmt(v) for v in verts),
            " ".join("(%s)" % ",".join(str(i) for i in f) for f in flist),
            " ".join("%.17g" % a for a in areas),
            " ".join("(%s)" % fmt(v) for v in normals),
            " ".join(str(b) for b in nbrs)]))
    return "\n".join(lines) + "\n"
