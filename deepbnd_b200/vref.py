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


# ------------------------------------------------------------------------------------------------------------------
# DOF order of Vref vectors in files (reference-produced bcs.hd5 / Wbasis.hd5 are in dolfin's DOF order)
# ------------------------------------------------------------------------------------------------------------------
# Every file this package writes with Vref vectors in it (snapshots: solutions_*; bcs.hd5: u0,u1,u2; Wbasis.hd5)
# carries two small datasets that make the order self-describing, for consumers on either side:
#     vref_dof_coordinates (nh, 2)   coordinates of the node each DOF sits on
#     vref_dof_component   (nh,)     its vector component
# A file WITHOUT them is a reference-produced file in dolfin's order.  It is accepted only together with the sidecar
# the FEniCS exporter writes (tools/export_goldens_fenics.py: `vref_dof_order.hd5` with the same two datasets taken
# from Vref.tabulate_dof_coordinates() / Vref.sub(c).dofmap().dofs()), placed in the same directory; the vectors are
# then permuted into the library order.  Without either, loading fails loudly instead of feeding wrongly ordered
# Dirichlet data to the solver (DEEPBND_ASSUME_LIBRARY_VREF_ORDER=1 restores the old silent behaviour).
COORD_LABEL, COMP_LABEL, SIDECAR = "vref_dof_coordinates", "vref_dof_component", "vref_dof_order.hd5"


class VrefOrderError(RuntimeError):
    pass


def vref_order_datasets(box=(-1.0, -1.0, 2.0, 2.0), Nb=21):
    """the two self-description datasets of the library's own order"""
    pts = vref_points(box[0], box[1], box[2], box[3], Nb)
    return np.repeat(pts, 2, axis=0), np.tile(np.arange(2.0), len(pts))


def permutation_to_library(coords, comps, box=(-1.0, -1.0, 2.0, 2.0), Nb=21, tol=1e-8):
    """perm such that v_library = v_file[..., perm]: for every library DOF the index of the file DOF on the same
    node and component"""
    pts = vref_points(box[0], box[1], box[2], box[3], Nb)
    coords = np.asarray(coords, dtype=float).reshape(-1, 2)
    comps = np.asarray(comps).astype(int).ravel()
    nh = 2 * len(pts)
    if len(coords) != nh or len(comps) != nh:
        raise VrefOrderError("Vref order description has %d DOFs, expected %d" % (len(coords), nh))
    perm = np.empty(nh, dtype=np.int64)
    scale = max(box[2], box[3])
    for k, p in enumerate(pts):
        d = np.abs(coords - p).max(1)
        for c in range(2):
            hit = np.flatnonzero((d < tol * scale) & (comps == c))
            if len(hit) != 1:
                raise VrefOrderError("Vref order description: node (%g, %g) component %d matched %d DOFs" % (p[0], p[1], c, len(hit)))
            perm[2 * k + c] = hit[0]
    return perm


def load_vref_vectors(filename, labels, box=(-1.0, -1.0, 2.0, 2.0), Nb=21):
    """datasets `labels` of `filename` whose last axis is a Vref vector, returned in the library's DOF order"""
    import os
    from . import wrapper_h5py as myhd
    single = isinstance(labels, str)
    arrs = [myhd.loadhd5(filename, l) for l in ([labels] if single else labels)]
    desc = None
    if myhd.checkExistenceDataset(filename, COORD_LABEL):
        desc = filename
    else:
        side = os.path.join(os.path.dirname(os.path.abspath(filename)), SIDECAR)
        if os.path.exists(side):
            desc = side
    if desc is None:
        if os.environ.get("DEEPBND_ASSUME_LIBRARY_VREF_ORDER") == "1":
            return arrs[0] if single else arrs
        raise VrefOrderError(
            "%s carries Vref vectors but no DOF-order description (%s / %s) and there is no %s next to it: a file written "
            "by the reference is in dolfin's DOF order and would give silently wrong boundary data.  Export the order "
            "with tools/export_goldens_fenics.py, or set DEEPBND_ASSUME_LIBRARY_VREF_ORDER=1 if the file is known to be "
            "in the library's order (CCW nodes from (x0,y0), centre last, 2*node+comp)." % (filename, COORD_LABEL, COMP_LABEL, SIDECAR))
    perm = permutation_to_library(myhd.loadhd5(desc, COORD_LABEL), myhd.loadhd5(desc, COMP_LABEL), box, Nb)
    out = [a[..., perm] for a in arrs]
    return out[0] if single else out
