"""Material helpers, mirror of the reference's core/elasticity/misc.py:25-33 and the region table of
getLameInclusions (core/elasticity/fenics_utils.py:41-53)."""
import numpy as np

eng2lamb = lambda nu, E: nu * E / ((1. - 2. * nu) * (1. + nu))
eng2mu = lambda nu, E: E / (2. * (1. + nu))
lame2lambPlane = lambda lamb, mu: (2.0 * mu * lamb) / (lamb + 2.0 * mu)
eng2lambPlane = lambda nu, E: lame2lambPlane(eng2lamb(nu, E), eng2mu(nu, E))


def getLameTable(nu1, E1, nu2, E2):
    """[lambda*, mu] for region tags 0..3 (inclusion, matrix, inclusion, matrix), plane-stress lambda."""
    mu1, lamb1 = eng2mu(nu1, E1), eng2lambPlane(nu1, E1)
    mu2, lamb2 = eng2mu(nu2, E2), eng2lambPlane(nu2, E2)
    return np.array([[lamb1, mu1], [lamb2, mu2], [lamb1, mu1], [lamb2, mu2]])


def macro_strain_voigt(i):
    """Unit Voigt macro strain i with engineering shear ([EXT] mscm.macro_strain)."""
    e = np.zeros(3)
    e[i] = 1.0
    return e
