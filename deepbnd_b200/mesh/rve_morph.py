"""Fixed-topology RVE meshes by radial morphing (mesh supply must not become the bottleneck).

The reference re-meshes every snapshot with gmsh (mesh_RVE.py:38-64); its datasets vary only the 36 inclusion radii
at fixed centres (build_snapshots_param.py:21-41).  Here ONE template mesh is generated for a reference radius r0
(same geometric spec, deepbnd_b200/mesh/rve_mesher.py) and every RVE of the campaign is obtained by moving the
nodes radially around each inclusion centre c_k:

    d <= r0          :  d -> d r_k / r0                                   (the inclusion scales as a whole)
    r0 < d < R       :  d -> r_k + (d - r0) (R - r_k) / (R - r0)          (the annulus up to R absorbs the change)
    d >= R           :  unchanged                                        (R = half the centre spacing)

so that the inclusion boundary stays a regular polygon of radius r_k, the outer boundary, the internal interface
(the Vref box) and therefore every constraint / sample location of the solver stay where they are, and the
topology, region tags and P2 numbering are shared by all RVEs.  A morphed mesh costs a few vectorised numpy
operations (≈0.1 ms) instead of four Delaunay passes (≈0.2 s for a `full` mesh).

The solutions differ from those on a re-meshed RVE at the level of the discretisation error (tests/test_morph.py
compares the homogenised tangents); use `buildRVEmesh_arrays` when the reference's per-snapshot meshing must be
mirrored, this when throughput matters."""
import numpy as np

from .rve_mesher import buildRVEmesh_arrays, orderedIndexesTotal, paramRVE_default


class MorphTemplate:
    def __init__(self, paramRVEdata, size='reduced', lcar=None, r0=None, isOrdered=False, NxL=2, NyL=2, maxOffset=2,
                 Vfrac=0.282743):
        """paramRVEdata: one (Nx*Ny, 5) table [cx, cy, r, e, theta] in the reference's (unpermuted) order; only its
        centres are used.  r0: template radius (default: the midpoint of the radii range of the table)."""
        p = paramRVE_default(NxL, NyL, maxOffset, Vfrac)
        self.size, self.lcar, self.isOrdered = size, lcar, isOrdered
        self.perm = np.arange(p.Nx * p.Ny) if isOrdered else orderedIndexesTotal(p.Nx, p.Ny, p.NxL)
        row = np.array(paramRVEdata, dtype=float, copy=True)
        self.r0 = float(r0) if r0 is not None else 0.5 * (row[:, 2].min() + row[:, 2].max())
        row[:, 2] = self.r0
        row[:, 3] = 1.0
        self.points, self.tri, self.region = buildRVEmesh_arrays(row, isOrdered=isOrdered, size=size, NxL=NxL, NyL=NyL,
                                                                 maxOffset=maxOffset, Vfrac=Vfrac, lcar=lcar)
        self.n_incl = p.NL if size == 'reduced' else p.Nx * p.Ny
        self.centres = row[:self.n_incl, :2].copy()                    # permuted order = the mesher's circle order
        cc = self.centres
        if len(cc) > 1:
            dmin = np.sqrt(((cc[:, None, :] - cc[None, :, :]) ** 2).sum(-1) + np.eye(len(cc)) * 1e300).min()
        else:
            dmin = 2.0 * p.H
        self.R = 0.5 * dmin
        if not self.r0 < self.R:
            raise ValueError("template radius must be smaller than half the centre spacing")
        # nearest centre, distance and direction of every node
        diff = self.points[:, None, :] - cc[None, :, :]
        dist = np.hypot(diff[..., 0], diff[..., 1])
        k = dist.argmin(1)
        d = dist[np.arange(len(k)), k]
        self.moving = np.flatnonzero(d < self.R * (1.0 - 1e-12))
        self.k = k[self.moving]
        self.d = d[self.moving]
        with np.errstate(invalid='ignore', divide='ignore'):
            u = diff[self.moving, self.k] / self.d[:, None]
        self.dir = np.where(self.d[:, None] > 0, u, 0.0)
        # the box and the internal interface must not move
        x0, y0 = self.points.min(0)
        x1, y1 = self.points.max(0)
        tol = 1e-9 * max(x1 - x0, y1 - y0)
        lines = [x0, x1] + ([p.x0L, p.x0L + p.LxL] if size == 'full' else [])
        on = np.zeros(len(self.points), bool)
        for v in lines:
            on |= np.abs(self.points[:, 0] - v) < tol
        for v in [y0, y1] + ([p.y0L, p.y0L + p.LyL] if size == 'full' else []):
            on |= np.abs(self.points[:, 1] - v) < tol
        if np.intersect1d(np.flatnonzero(on), self.moving).size:
            raise ValueError("an inclusion's zone of influence reaches the box or the internal interface")

    def mesh(self, radii, e=None, theta=None):
        """points of the RVE with inclusion radii `radii` (mesher order); triangles and regions are the template's.
        With excentricities `e` and angles `theta` the inclusions become the ellipses of ellipse2_mesh.py:27-39
        (semi-axis radii along theta, e*radii across): the target radius of a node is the ellipse radius in its direction."""
        radii = np.asarray(radii, dtype=float)
        if radii.shape != (self.n_incl,):
            raise ValueError("expected %d radii" % self.n_incl)
        r = radii[self.k]
        if e is not None and np.any(np.abs(np.asarray(e, dtype=float) - 1.0) > 1e-12):
            e = np.asarray(e, dtype=float)
            theta = np.zeros(self.n_incl) if theta is None else np.asarray(theta, dtype=float)
            a, b, th = radii[self.k], (e * radii)[self.k], theta[self.k]
            phi = np.arctan2(self.dir[:, 1], self.dir[:, 0]) - th
            r = a * b / np.sqrt((b * np.cos(phi)) ** 2 + (a * np.sin(phi)) ** 2)
            radii = np.maximum(radii, e * radii)
        if np.any(radii <= 0) or np.any(radii >= self.R):
            raise ValueError("radii must lie in (0, %g)" % self.R)
        dn = np.where(self.d <= self.r0, self.d * r / self.r0, r + (self.d - self.r0) * (self.R - r) / (self.R - self.r0))
        pts = self.points.copy()
        pts[self.moving] = self.centres[self.k] + self.dir * dn[:, None]
        return pts, self.tri, self.region


def buildRVEmesh_morphed(paramRVEdata, template):
    """Twin of buildRVEmesh_arrays on a MorphTemplate: permutes paramRVEdata IN PLACE like the reference
    (mesh_RVE.py:45-50), checks that the centres are the template's and returns (points, triangles, region)."""
    paramRVEdata[:, :] = paramRVEdata[template.perm, :]
    n = template.n_incl
    if np.abs(paramRVEdata[:n, :2] - template.centres).max() > 1e-9:
        raise ValueError("the inclusion centres differ from the template's: re-mesh (buildRVEmesh_arrays) instead")
    return template.mesh(paramRVEdata[:n, 2], paramRVEdata[:n, 3], paramRVEdata[:n, 4])



def morph_parameters(paramRVEdata, template):
    """(l, e, theta) of every inclusion of a batch in the mesher's order, for RVEBatch.set_morphed (the morphing then runs
    on the GPU): paramRVEdata (B, Nx*Ny, 5) is permuted IN PLACE, every row like the reference (mesh_RVE.py:45-50)."""
    paramRVEdata[:, :, :] = paramRVEdata[:, template.perm, :]
    n = template.n_incl
    if np.abs(paramRVEdata[:, :n, :2] - template.centres[None]).max() > 1e-9:
        raise ValueError("the inclusion centres differ from the template's: re-mesh (buildRVEmesh_arrays) instead")
    return np.ascontiguousarray(paramRVEdata[:, :n, 2:5])
