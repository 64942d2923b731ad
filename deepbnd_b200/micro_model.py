"""Object-level mirrors of the reference's micro models on top of the batched C-ABI.

  MicroModel                  core/multiscale/micro_model.py:24-96
  MicroConstitutiveModelDNN   core/multiscale/micro_model_dnn.py:26-113

One object = one RVE = a batch of one on the GPU; the drivers (simulation_snapshots.py,
tangents_prediction.py) use RVEBatch directly with thousands of systems instead.  `nameMesh` may be the
reference's XDMF name (read from the cache / the XDMF triplet) or an array triple
(points, triangles, region).  Vref DOF order: the library's own — node k = 0..79
counter-clockwise from (x0,y0), centre node last, DOF = 2*node + comp.
"""
import numpy as np

from .batch import RVEBatch
from .elasticity import macro_strain_voigt
from .mesh.mesh_io import load_mesh


def _mesh_of(nameMesh):
    return load_mesh(nameMesh) if isinstance(nameMesh, str) else tuple(nameMesh)


class _Sol:
    """`sol[j][0]` of the reference is a dolfin Function; here it is the nodal array (n_nodes, 2)."""

    def __init__(self, u):
        self.u = u

    def __getitem__(self, k):
        if k != 0:
            raise IndexError("only the displacement block is exposed (the multipliers are eliminated)")
        return self.u


class MicroModel:
    def __init__(self, nameMesh, param, model, device=0, vref_nb=21, vref_box=None, rtol=1e-10):
        self.nameMesh = nameMesh
        self.param = param
        self.model = model
        self.rtol = rtol
        self.mesh = _mesh_of(nameMesh)
        xy = self.mesh[0]
        self.others = {'polyorder': 2, 'x0': xy[:, 0].min(), 'x1': xy[:, 0].max(), 'y0': xy[:, 1].min(), 'y1': xy[:, 1].max()}
        self.batch = RVEBatch(device).set_meshes([self.mesh], tuple(param), model=model, vref_nb=vref_nb, vref_box=vref_box)
        self.sol = None
        self._snap = None

    def compute(self):
        """K once, all three unit macro strains as right-hand sides of one solve (micro_model.py:52-82)."""
        self.batch.assemble()
        eps = np.stack([macro_strain_voigt(i) for i in range(3)])[None]
        self.batch.solve(eps, rtol=self.rtol)
        U = self.batch.solution(0)
        self.sol = [_Sol(U[:, i, :]) for i in range(3)]
        self._snap = self.batch.snapshot()

    def homogenise(self, regions, i_voigt):
        """2x2 average stress of load i over `regions` (micro_model.py:85-96); regions [0,1] or [0,1,2,3]."""
        regions = sorted(int(r) for r in regions)
        if regions == [0, 1]:
            s = self._snap["sigma"][0, i_voigt]
        elif regions == [0, 1, 2, 3]:
            s = self._snap["sigmaT"][0, i_voigt]
        else:
            raise NotImplementedError("homogenise over regions %s" % regions)
        return np.array([[s[0], s[2]], [s[2], s[1]]])

    def getAffineTransformationLocal(self, i_voigt):
        """(a, B) of core/multiscale/misc.py:49-61 over regions [0,1]."""
        return self._snap["a"][0, i_voigt], self._snap["B"][0, i_voigt]

    def interpolate_on_Vref(self, i_voigt):
        """usol.interpolate(sol[i][0]); usol.vector().get_local() (simulation_snapshots.py:56-58)."""
        return self._snap["sol"][0, i_voigt]


class MicroConstitutiveModelDNN:
    def __init__(self, nameMesh, param, model, device=0, rtol=1e-10):
        self.nameMesh = nameMesh
        self.param = param
        self.model = model
        self.rtol = rtol
        self.device = device
        self.mesh = _mesh_of(nameMesh)
        xy = self.mesh[0]
        self.others = {'polyorder': 2, 'x0': xy[:, 0].min(), 'x1': xy[:, 0].max(), 'y0': xy[:, 1].min(), 'y1': xy[:, 1].max()}
        self.isTangentComputed = False
        self.Chom_ = None

    def computeTangent(self):
        model = self.model
        uD = None
        if model in ('dnn', 'lin') and 'uD0_' in self.others:
            uD = np.stack([np.asarray(self.others['uD%d_' % i], dtype=np.float64) for i in range(3)])[None]
            if not np.any(uD):
                uD, model = None, 'lin'
            else:
                model = 'dnn'
        elif model == 'dnn':
            raise ValueError("model 'dnn' needs others['uD0_'], ['uD1_'], ['uD2_'] (micro_model_dnn.py:83)")
        b = RVEBatch(self.device).set_meshes([self.mesh], tuple(self.param), model=model).assemble()
        eps = np.stack([macro_strain_voigt(i) for i in range(3)])[None]
        b.solve(eps, uD=uD, rtol=self.rtol)
        self.Chom_ = b.tangent()[0]
        b.close()
        self.isTangentComputed = True

    def getTangent(self):
        if not self.isTangentComputed:
            self.computeTangent()
        return self.Chom_


class MicroConstitutiveModelGen:
    """Mirror of core/multiscale/micro_model_gen.py:31-160: average stress and strain over the whole RVE and
    over the internal regions `domainL` = [0, 1], and the tangents sigma eps^-1 built from them.  Voigt
    conventions of core/elasticity/misc.py:58-62 (stress [00,11,01], strain [00,11,2*01]).  The dnn variant
    symmetrises the B correction (:122), which the library does under RVE_OPT_SYM_B."""

    def __init__(self, nameMesh, param, model, device=0, rtol=1e-10):
        self.nameMesh = nameMesh
        self.param = param
        self.model = model
        self.device = device
        self.rtol = rtol
        self.others = {'polyorder': 2}
        self.ndim = 2
        self.nvoigt = 3
        self.Hom = None

    def computeHomogenisation(self, domainL=(0, 1)):
        if sorted(domainL) != [0, 1]:
            raise NotImplementedError("domainL must be the internal regions [0, 1]")
        mesh = _mesh_of(self.nameMesh)
        xy = mesh[0]
        self.others.update(x0=xy[:, 0].min(), x1=xy[:, 0].max(), y0=xy[:, 1].min(), y1=xy[:, 1].max())
        model, uD = self.model, None
        if model == 'dnn':
            uD = np.stack([np.asarray(self.others['uD%d_' % i], dtype=np.float64) for i in range(3)])[None]
        b = RVEBatch(self.device).set_meshes([mesh], tuple(self.param), model=model)
        b.set_option(1, 1)                                              # RVE_OPT_SYM_B
        b.assemble()
        b.solve(np.stack([macro_strain_voigt(i) for i in range(3)])[None], uD=uD, rtol=self.rtol)
        h = b.homogenisation()
        b.close()
        Hom = {k: np.zeros((3, 3)) for k in ("sigma", "sigmaL", "eps", "epsL", "tangent", "tangentL")}
        for i in range(3):
            e = h["eps_eff"][0, i]
            for key, g, s in (("L", h["gradL"][0, i], h["sigmaL"][0, i]), ("", h["gradT"][0, i], h["sigmaT"][0, i])):
                Hom["sigma" + key][:, i] = s                             # stress2voigt: [00, 11, 01]
                Hom["eps" + key][:, i] = [e[0] + g[0, 0], e[1] + g[1, 1], e[2] + g[0, 1] + g[1, 0]]   # strain2voigt
        Hom["tangent"] = Hom["sigma"] @ np.linalg.inv(Hom["eps"])
        Hom["tangentL"] = Hom["sigmaL"] @ np.linalg.inv(Hom["epsL"])
        self.Hom = Hom
        return Hom

    def getHomogenisation(self):
        return self.Hom if self.Hom is not None else self.computeHomogenisation()
