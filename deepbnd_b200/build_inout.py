"""Drop-in for creation_model/RB/build_inout.py (reference :36-87), the consumer on the far side of
snapshots.hd5: (0) translateSolution, (1) computingBasis -> Wbasis.hd5, (2) extractAlpha, (3) createXY -> XY.hd5.
Same file/dataset names; the dolfin pieces (Vref functions, the ds mass matrix) are replaced by their closed
forms on the library's Vref order (deepbnd_b200/vref.py).  The dense algebra (SVD of the ns x 162 snapshot
matrix, the projections) runs in fp64 through torch on the GPU when one is present, on the CPU otherwise —
library calls, this stage is not part of the hand-written hot path.  The device is "cuda"
unless the caller asks for "cpu" explicitly (module attribute DEVICE or the `device` arguments): there is no
silent fallback."""
import os

import numpy as np

from . import wrapper_h5py as myhd
from .simulation_snapshots import vref_box_from_mesh
from .vref import COMP_LABEL, COORD_LABEL, load_vref_vectors, vref_mass, vref_order_datasets, vref_points


DEVICE = "cuda"


def _dev(device=None):
    import torch
    d = torch.device(device or DEVICE)
    if d.type == "cuda" and not torch.cuda.is_available():
        raise RuntimeError("deepbnd_b200.build_inout: no CUDA device (pass device='cpu' explicitly to run the RB algebra on the host)")
    return d


def getMassMatrix(nameMeshRefBnd=None):
    box, nb = vref_box_from_mesh(nameMeshRefBnd)
    return vref_mass(vref_points(*box, nb))


def translateSolution(nameSnaps, nameMeshRefBnd=None):
    """solutions_trans_partial_X = solutions_X + a_X at every Vref node (reference :36-51, B = 0)."""
    for load_flag in ['A', 'S']:
        Isol = load_vref_vectors(nameSnaps, 'solutions_%s' % load_flag, *vref_box_from_mesh(nameMeshRefBnd))
        a = myhd.loadhd5(nameSnaps, 'a_%s' % load_flag)
        trans = Isol + np.tile(a, (1, Isol.shape[1] // 2))
        myhd.addDataset(nameSnaps, trans, 'solutions_trans_partial_%s' % load_flag)


def computingBasis_svd(Isol, M, Nmax, device=None):
    """RB_utils.computingBasis_svd (:134-157): Msqrt from the SVD of M, thin SVD of Isol Msqrt,
    Wbasis[i] = (U[:, i] / sig[i])^T Isol; returns (Wbasis (Nmax, Nh), sig**2)."""
    import torch
    dev = _dev(device)
    UM, SM, VMT = np.linalg.svd(M)
    Msqrt = (UM[:, :len(SM)] * np.sqrt(SM)) @ VMT
    I = torch.from_numpy(np.ascontiguousarray(Isol)).to(dev)
    U, sig, _ = torch.linalg.svd(I @ torch.from_numpy(Msqrt).to(dev), full_matrices=False)
    W = (U[:, :Nmax] / sig[:Nmax]).T @ I
    return W.cpu().numpy(), (sig ** 2).cpu().numpy()


def computingBasis(nameSnaps, nameWbasis, Nmax, Nh=None, nameMeshRefBnd=None, device=None):
    if os.path.exists(nameWbasis):
        os.remove(nameWbasis)
    M = getMassMatrix(nameMeshRefBnd)
    out = {'massMat': M}
    for load_flag in ['A', 'S']:
        Isol = myhd.loadhd5(nameSnaps, 'solutions_trans_partial_%s' % load_flag)
        W, sig2 = computingBasis_svd(Isol, M, Nmax, device)
        out['Wbasis_%s' % load_flag] = W
        out['sig_%s' % load_flag] = sig2[:Nmax]
    out[COORD_LABEL], out[COMP_LABEL] = vref_order_datasets(*vref_box_from_mesh(nameMeshRefBnd))
    labels = ['Wbasis_A', 'Wbasis_S', 'sig_A', 'sig_S', 'massMat', COORD_LABEL, COMP_LABEL]
    myhd.savehd5(nameWbasis, [out[l] for l in labels], labels, 'w')


def getAlphas_fast(Wbasis_M, Isol, Nmax, device=None):
    """RB_utils.getAlphas_fast (:211-213): Isol M Wbasis[:Nmax]^T."""
    import torch
    dev = _dev(device)
    W, M = Wbasis_M
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return (t(Isol) @ t(M) @ t(W[:Nmax]).T).cpu().numpy()


def extractAlpha(nameSnaps, nameWbasis, Nmax, nameYlist, device=None):
    if os.path.exists(nameYlist):
        os.remove(nameYlist)
    for load_flag in ['A', 'S']:
        Wbasis_M = myhd.loadhd5(nameWbasis, ['Wbasis_%s' % load_flag, 'massMat'])
        Isol = myhd.loadhd5(nameSnaps, 'solutions_trans_partial_%s' % load_flag)
        Ylist = getAlphas_fast(Wbasis_M, Isol, Nmax, device)
        if os.path.exists(nameYlist):
            myhd.addDataset(nameYlist, Ylist, 'Ylist_%s' % load_flag)
        else:
            myhd.savehd5(nameYlist, Ylist, 'Ylist_%s' % load_flag, mode='w')


def createXY(nameParamRVEdataset, nameYlist, nameXYlist):
    if os.path.exists(nameXYlist):
        os.remove(nameXYlist)
    Y = myhd.loadhd5(nameYlist, ['Ylist_%s' % s for s in ['A', 'S']])
    X = myhd.loadhd5(nameParamRVEdataset, 'param')[:, :, 2]
    myhd.savehd5(nameXYlist, Y + [X], ['Y_%s' % s for s in ['A', 'S']] + ['X'], mode='w')
    os.remove(nameYlist)
