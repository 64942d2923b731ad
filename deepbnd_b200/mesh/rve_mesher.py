"""Circles-in-a-box triangle mesher that follows the reference's RVE mesh spec.

The reference drives pygmsh -> gmsh (front2d) -> meshio (core/multiscale/mesh_RVE.py:38-64,
core/mesh/ellipse2_mesh.py:5-42, core/mesh/ellipse_two_domains_mesh.py:5-33,
core/mesh/wrapper_gmsh.py:38-55).  Neither gmsh nor meshio exists in this image, so this
module produces meshes to the same *spec* with scipy's Delaunay:

  * geometry constants = paramRVE_default (mesh_RVE.py:12-35): H=1, internal box [-1,1]^2,
    total box [-3,3]^2, lcar = 2/30, NpLxL = 31 / NpLxt = 91 transfinite nodes per side;
  * rows are permuted in place by orderedIndexesTotal (generation_inclusions.py:82-92) exactly
    like buildRVEmesh (mesh_RVE.py:45-50);
  * 'reduced': the 4 internal inclusions in [-1,1]^2, tags inclusions=0, matrix=1
    (ellipse2_mesh.py:22-25);
  * 'full': tags internal inclusions=0, internal matrix=1, external inclusions=2,
    external matrix=3; the internal interface is conforming with 31 nodes per side
    (ellipse_two_domains_mesh.py:22-30, mesh_RVE.py:57-62);
  * circles are inscribed polygons with ~lcar segments (what gmsh does with 4 ellipse arcs).

Every constrained segment (box sides, interface, circle chords) is checked to be a
triangulation edge; the mesher raises if not.
"""
import numpy as np
from scipy.spatial import Delaunay


class paramRVE_default:
    """Mirror of core/multiscale/mesh_RVE.py:12-35 (same attribute names)."""

    def __init__(self, NxL=2, NyL=2, maxOffset=2, Vfrac=0.282743):
        self.Vfrac = Vfrac
        self.maxOffset = maxOffset
        self.NxL = NxL
        self.NyL = NyL
        self.H = 1.0
        self.NL = self.NxL * self.NyL
        self.LxL = self.NxL * self.H
        self.LyL = self.NyL * self.H
        self.x0L = -0.5 * self.LxL
        self.y0L = -0.5 * self.LyL
        self.lcar = (2 / 30) * self.H
        self.Nx = self.NxL + 2 * self.maxOffset
        self.Ny = self.NyL + 2 * self.maxOffset
        self.Lxt = self.Nx * self.H
        self.Lyt = self.Ny * self.H
        self.NpLxt = int(self.Lxt / self.lcar) + 1
        self.NpLxL = int(self.LxL / self.lcar) + 1
        self.x0 = -0.5 * self.Lxt
        self.y0 = -0.5 * self.Lyt
        self.r0 = 0.2 * self.H
        self.r1 = 0.4 * self.H
        self.rm = self.H * np.sqrt(self.Vfrac / np.pi)


def orderedIndexesBox(Nx, Ny, offset):
    """generation_inclusions.py:68-72 (the set difference is sorted: ints hash to themselves)."""
    indexes = np.arange(0, Nx * Ny).reshape((Nx, Ny))
    inner = indexes[offset:Nx - offset, offset:Ny - offset].flatten()
    rest = np.setdiff1d(indexes.flatten(), inner)
    return np.concatenate([inner, rest]).astype(int)


def orderedIndexesTotal(Nx, Ny, minNx):
    """generation_inclusions.py:82-92: standard row/column numbering -> internal-to-external layers."""
    L = np.arange(0, Nx * Ny).astype(int)
    maxOffset = int((Nx - minNx) / 2)
    for i in range(maxOffset):
        ni = (Nx - 2 * i) * (Ny - 2 * i)
        n1 = int(np.sqrt(float(ni)))
        L = np.concatenate([L[:ni][orderedIndexesBox(n1, n1, 1)], L[ni:]])
    return L


def _side_points(x0, y0, Lx, Ly, n_per_side):
    """Box boundary nodes, counter-clockwise from (x0,y0); n_per_side includes both corners."""
    n = n_per_side - 1
    s = np.arange(n) / n
    return np.vstack([np.stack([x0 + Lx * s, np.full(n, y0)], 1),
                      np.stack([np.full(n, x0 + Lx), y0 + Ly * s], 1),
                      np.stack([x0 + Lx * (1 - s), np.full(n, y0 + Ly)], 1),
                      np.stack([np.full(n, x0), y0 + Ly * (1 - s)], 1)])


def _circle_points(cx, cy, r, lcar):
    n = int(np.ceil(2 * np.pi * r / lcar / 4.0)) * 4      # multiple of 4: the 4 arc end points are nodes
    n = max(n, 8)
    t = 2 * np.pi * np.arange(n) / n
    return np.stack([cx + r * np.cos(t), cy + r * np.sin(t)], 1)


def _ellipse_points(cx, cy, l, e, theta, lcar):
    """Boundary nodes of the ellipse of ellipse2_mesh.py:27-39 (semi-axis l along theta, e*l across): the four axis
    end points (gmsh's arc end points) and, on every quarter arc, nodes at equal arc-length spacing of about lcar."""
    a, b = float(l), float(e) * float(l)
    tt = np.linspace(0.0, 0.5 * np.pi, 513)
    seg = np.hypot(np.diff(a * np.cos(tt)), np.diff(b * np.sin(tt)))
    s = np.concatenate([[0.0], np.cumsum(seg)])                    # arc length along the first quadrant
    nq = max(2, int(np.ceil(s[-1] / lcar)))
    tq = np.interp(np.arange(nq) * s[-1] / nq, s, tt)              # nq nodes, the quadrant's last point belongs to the next
    # the other quadrants are mirror images
    t = np.concatenate([[0.0], tq[1:], [0.5 * np.pi], (np.pi - tq[1:])[::-1], [np.pi], np.pi + tq[1:], [1.5 * np.pi],
                        (2 * np.pi - tq[1:])[::-1]])
    c, sn = np.cos(theta), np.sin(theta)
    x, y = a * np.cos(t), b * np.sin(t)
    return np.stack([cx + c * x - sn * y, cy + sn * x + c * y], 1)


def _dist_to_ellipse(p, cx, cy, l, e, theta):
    """first-order distance of points to the ellipse curve (exact for a circle): (F - 1) / |grad F|, F = |(x'/a, y'/b)|"""
    a, b = float(l), float(e) * float(l)
    c, sn = np.cos(theta), np.sin(theta)
    dx, dy = p[:, 0] - cx, p[:, 1] - cy
    xl, yl = c * dx + sn * dy, -sn * dx + c * dy
    F = np.sqrt((xl / a) ** 2 + (yl / b) ** 2)
    with np.errstate(invalid='ignore', divide='ignore'):
        g = np.hypot(xl / a ** 2, yl / b ** 2) / F
        d = np.abs(F - 1.0) / g
    return np.where(F > 1e-12, d, min(a, b))


def _in_polygon(pts, poly):
    """crossing-number test of points against a closed polygon (vertices in order)"""
    x, y = pts[:, 0][:, None], pts[:, 1][:, None]
    x0, y0 = poly[:, 0][None, :], poly[:, 1][None, :]
    x1, y1 = np.roll(poly[:, 0], -1)[None, :], np.roll(poly[:, 1], -1)[None, :]
    cond = (y0 > y) != (y1 > y)
    with np.errstate(invalid='ignore', divide='ignore'):
        xint = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
    return (np.sum(cond & (x < xint), axis=1) % 2) == 1


def _dist_to_box(p, x0, y0, Lx, Ly):
    """Unsigned distance of points to the boundary lines of a box (inside or outside)."""
    dx = np.maximum(np.maximum(x0 - p[:, 0], p[:, 0] - (x0 + Lx)), 0)
    dy = np.maximum(np.maximum(y0 - p[:, 1], p[:, 1] - (y0 + Ly)), 0)
    outside = np.hypot(dx, dy)
    inside = np.minimum(np.minimum(p[:, 0] - x0, x0 + Lx - p[:, 0]),
                        np.minimum(p[:, 1] - y0, y0 + Ly - p[:, 1]))
    return np.where((dx > 0) | (dy > 0), outside, inside)


def mesh_circles_in_box(circles, box, lcar, n_side, inner_box=None, n_side_inner=None,
                        n_inner_incl=0, smooth_iters=3):
    """Triangulate box minus/with circular inclusions.

    circles: (nc,3) [cx,cy,r] or (nc,5) [cx,cy,l,e,theta] (elliptical inclusions, ellipse2_mesh.py:27-39);
    box=(x0,y0,Lx,Ly); inner_box likewise (conforming interface).
    Returns points (nv,2) f8, triangles (nt,3) i4 (CCW), region (nt,) i4 with the reference's
    tags (0/1 and, with an inner box, 2/3)."""
    x0, y0, Lx, Ly = box
    fixed = [_side_points(x0, y0, Lx, Ly, n_side)]
    segs = []

    def add_loop(start, n):
        idx = start + np.arange(n)
        segs.append(np.stack([idx, np.roll(idx, -1)], 1))

    nfix = 0
    add_loop(0, fixed[0].shape[0])
    nfix += fixed[0].shape[0]
    if inner_box is not None:
        pin = _side_points(*inner_box, n_side_inner)
        fixed.append(pin)
        add_loop(nfix, pin.shape[0])
        nfix += pin.shape[0]
    circles = np.asarray(circles, dtype=float)
    circles = circles.reshape(-1, circles.shape[-1] if circles.ndim == 2 else 3)
    if circles.shape[1] == 3:
        circles = np.hstack([circles, np.ones((len(circles), 1)), np.zeros((len(circles), 1))])
    polys = []
    for cx, cy, r, ecc, th in circles:
        pc = _circle_points(cx, cy, r, lcar) if ecc == 1.0 else _ellipse_points(cx, cy, r, ecc, th, lcar)
        polys.append(pc)
        fixed.append(pc)
        add_loop(nfix, pc.shape[0])
        nfix += pc.shape[0]
    fixed = np.vstack(fixed)
    segs = np.vstack(segs)

    # interior candidates: triangular lattice of pitch lcar, clipped away from every constrained curve
    hy = lcar * np.sqrt(3.0) / 2.0
    ny = int(np.floor(Ly / hy))
    nx = int(np.floor(Lx / lcar)) + 1
    jj, ii = np.meshgrid(np.arange(1, ny + 1), np.arange(-1, nx + 1), indexing='ij')
    gy = y0 + (Ly - ny * hy) / 2.0 + (jj - 0.5) * hy
    gx = x0 + (ii + 0.5 * (jj % 2)) * lcar
    cand = np.stack([gx.ravel(), gy.ravel()], 1)
    dmin = 0.62 * lcar
    keep = _dist_to_box(cand, x0, y0, Lx, Ly) > dmin
    keep &= (cand[:, 0] > x0) & (cand[:, 0] < x0 + Lx) & (cand[:, 1] > y0) & (cand[:, 1] < y0 + Ly)
    if inner_box is not None:
        keep &= _dist_to_box(cand, *inner_box) > dmin
    for cx, cy, r, ecc, th in circles:
        if ecc == 1.0:
            d = np.abs(np.hypot(cand[:, 0] - cx, cand[:, 1] - cy) - r)
        else:
            d = _dist_to_ellipse(cand, cx, cy, r, ecc, th)
        keep &= d > dmin
    free = cand[keep]

    pts = np.vstack([fixed, free])
    tri = Delaunay(pts)
    for _ in range(smooth_iters):
        # Laplacian smoothing of the free points (fixed ones stay), then re-triangulate
        indptr, indices = tri.vertex_neighbor_vertices
        cnt = np.diff(indptr)
        acc = np.zeros_like(pts)
        np.add.at(acc, np.repeat(np.arange(pts.shape[0]), cnt), pts[indices])
        new = acc / np.maximum(cnt, 1)[:, None]
        pts[nfix:] = 0.5 * pts[nfix:] + 0.5 * new[nfix:]
        tri = Delaunay(pts)
    if tri.coplanar.size:
        raise RuntimeError("mesher: qhull dropped %d points" % tri.coplanar.shape[0])
    T = tri.simplices.astype(np.int64)
    p0, p1, p2 = pts[T[:, 0]], pts[T[:, 1]], pts[T[:, 2]]
    area2 = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p1[:, 1] - p0[:, 1]) * (p2[:, 0] - p0[:, 0])
    T = T[np.abs(area2) > 1e-12 * lcar * lcar]            # slivers on collinear hull points
    area2 = area2[np.abs(area2) > 1e-12 * lcar * lcar]
    neg = area2 < 0
    T[neg, 1], T[neg, 2] = T[neg, 2].copy(), T[neg, 1].copy()

    # conformity: every constrained segment must be an edge
    nv = pts.shape[0]
    e = np.vstack([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]])
    e.sort(axis=1)
    ekeys = np.unique(e[:, 0] * nv + e[:, 1])
    s = np.sort(segs, axis=1)
    skeys = s[:, 0] * nv + s[:, 1]
    if not np.all(np.isin(skeys, ekeys)):
        raise RuntimeError("mesher: %d constrained segments missing" % int((~np.isin(skeys, ekeys)).sum()))

    # region tags by centroid-in-polygon (exact for the inscribed regular polygons)
    cen = (pts[T[:, 0]] + pts[T[:, 1]] + pts[T[:, 2]]) / 3.0
    incl = -np.ones(T.shape[0], dtype=np.int64)
    for k, (cx, cy, r, ecc, tht) in enumerate(circles):
        if ecc != 1.0:                                      # polygon of the ellipse nodes: crossing-number test nearby
            rmax = r * max(1.0, ecc)
            near = np.flatnonzero(np.hypot(cen[:, 0] - cx, cen[:, 1] - cy) < rmax * 1.01)
            incl[near[_in_polygon(cen[near], polys[k])]] = k
            continue
        n = polys[k].shape[0]
        d = np.hypot(cen[:, 0] - cx, cen[:, 1] - cy)
        th = np.arctan2(cen[:, 1] - cy, cen[:, 0] - cx)
        rel = np.mod(th, 2 * np.pi / n) - np.pi / n
        inside = d * np.cos(rel) < r * np.cos(np.pi / n)
        incl[inside] = k
    if inner_box is None:
        region = np.where(incl >= 0, 0, 1)
    else:
        xi, yi, Lxi, Lyi = inner_box
        in_inner = (cen[:, 0] > xi) & (cen[:, 0] < xi + Lxi) & (cen[:, 1] > yi) & (cen[:, 1] < yi + Lyi)
        region = np.where(in_inner, np.where(incl >= 0, 0, 1), np.where(incl >= 0, 2, 3))
    return np.ascontiguousarray(pts), T.astype(np.int32), region.astype(np.int32)


def buildRVEmesh_arrays(paramRVEdata, isOrdered=False, size='reduced', NxL=2, NyL=2, maxOffset=2,
                        Vfrac=0.282743, lcar=None):
    """Array-returning twin of buildRVEmesh (mesh_RVE.py:38-64).  Permutes paramRVEdata IN PLACE
    exactly like the reference (:45-50) and returns (points, triangles, region)."""
    p = paramRVE_default(NxL, NyL, maxOffset, Vfrac)
    p.lcar = lcar * p.H if lcar is not None else p.lcar
    if isOrdered:
        perm = np.arange(0, p.Nx * p.Ny).astype(int)
    else:
        perm = orderedIndexesTotal(p.Nx, p.Ny, p.NxL)
    paramRVEdata[:, :] = paramRVEdata[perm, :]
    npl = int(round(p.LxL / p.lcar)) + 1 if lcar is not None else p.NpLxL
    npt = int(round(p.Lxt / p.lcar)) + 1 if lcar is not None else p.NpLxt
    # columns: cx, cy, l (radius / semi-axis along theta), e (excentricity: second semi-axis e*l), theta
    # (ellipse2_mesh.py:27-39; the reference datasets use e = 1, theta = 0: generation_inclusions.py:29-30,42)
    if size == 'reduced':
        return mesh_circles_in_box(paramRVEdata[:p.NL, :5], (p.x0L, p.y0L, p.LxL, p.LyL), p.lcar, npl)
    elif size == 'full':
        return mesh_circles_in_box(paramRVEdata[:, :5], (p.x0, p.y0, p.Lxt, p.Lyt), p.lcar, npt,
                                   inner_box=(p.x0L, p.y0L, p.LxL, p.LyL), n_side_inner=npl,
                                   n_inner_incl=p.NL)
    raise ValueError(size)
