"""CPU restatement of the reference's RB stage (creation_model/RB/RB_utils.py:134-157, 211-213 and
creation_model/RB/build_inout.py:36-87) in plain numpy.  TEST INFRASTRUCTURE ONLY (checker of
deepbnd_b200/build_inout.py).  The reference's own code here IS numpy (np.linalg.svd, matmul) around a dolfin
mass matrix; the mass matrix is restated in oracle/rve_oracle.py:vref_mass."""
import numpy as np


def computingBasis_svd(Isol, M, Nmax):
    UM, SM, VMT = np.linalg.svd(M)
    Msqrt = (UM[:, :len(SM)] * np.sqrt(SM)) @ VMT
    U, sig, VT = np.linalg.svd(Isol @ Msqrt, full_matrices=False)
    W = np.zeros((Nmax, Isol.shape[1]))
    for i in range(Nmax):
        W[i, :] = (U[:, i] / sig[i]).reshape((1, len(U))) @ Isol
    return W, sig ** 2


def getAlphas_fast(W, M, Isol, Nmax):
    return Isol @ M @ W[:Nmax, :].T


def translate(Isol, a):
    out = Isol.copy()
    out[:, 0::2] += a[:, [0]]
    out[:, 1::2] += a[:, [1]]
    return out
