"""The internal-boundary reference space Vref in the library's DOF order (node k = 0..4(Nb-1)-1 counter-clockwise
from (x0,y0), centre node last, DOF = 2*node + comp): node coordinates and the boundary (ds) mass matrix of its
vector P1 space (core/mesh/degenerated_rectangle_mesh.py:3-35; getMassMatrix with dotProduct = inner(u,v)*ds,
creation_model/RB/RB_utils.py:125-132 and build_inout.py:34)."""
import numpy as np


def vref_points(x0=-1.0, y0=-1.0, Lx=2.0, Ly=2.0, Nb=21):
    n = Nb - 1
    s = np.arange(n) / n
    return np.vstack([np.stack([x0 + Lx * s, np.full(n, y0)], 1),
                      np.stack([np.full(n, x0 + Lx), y0 + Ly * s], 1),
                      np.stack([x0 + Lx * (1 - s), np.full(n, y0 + Ly)], 1),
                      np.stack([np.full(n, x0), y0 + Ly * (1 - s)], 1),
                      [[x0 + 0.5 * Lx, y0 + 0.5 * Ly]]])


def vref_mass(pts):
    """cyclic P1 boundary mass, 2 components interleaved; the rows/columns of the centre node are zero."""
    nb = pts.shape[0] - 1
    k = np.arange(nb)
    k2 = (k + 1) % nb
    L = np.hypot(*(pts[k2] - pts[k]).T)
    M = np.zeros((2 * (nb + 1), 2 * (nb + 1)))
    for c in range(2):
        np.add.at(M, (2 * k + c, 2 * k + c), L / 3)
        np.add.at(M, (2 * k2 + c, 2 * k2 + c), L / 3)
        np.add.at(M, (2 * k + c, 2 * k2 + c), L / 6)
        np.add.at(M, (2 * k2 + c, 2 * k + c), L / 6)
    return M
