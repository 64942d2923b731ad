"""Matrix-free operator (rve_matfree.cuh) against the assembled SELL-32 operator and the CPU oracle: the
element-wise application must be the same linear map as the assembled matrix (to rounding), give the
same PCG iteration counts and the same outputs, for every boundary model, ragged batches and both
mesh sizes."""
import numpy as np
import pytest

from conftest import smooth_uD

pytestmark = pytest.mark.gpu

PARAM = (0.3, 10.0, 0.3, 1.0)


def _rows(b):
    return [b.sizes(s)["n_rows"] for s in range(b.n_sys)]


@pytest.mark.parametrize("model", ["lin", "per", "MR"])
@pytest.mark.parametrize("K", [1, 2, 3])
def test_matfree_apply_equals_assembled_matrix(reduced_meshes, coarse_reduced_meshes, model, K):
    """q = A p of both kernels against the CSR matrix of rve_get_csr times p in numpy (ragged batch)."""
    from deepbnd_b200.batch import RVEBatch
    meshes = [reduced_meshes[0], coarse_reduced_meshes[1], reduced_meshes[2], coarse_reduced_meshes[0]]
    b = RVEBatch().set_meshes(meshes, PARAM, model=model)
    b.set_operator("matfree").assemble()
    nr = _rows(b)
    rng = np.random.RandomState(3)
    p = rng.randn(sum(nr), K, 2)
    q_mf = b.apply_operator(p)
    b.set_operator("assembled").assemble()
    q_as = b.apply_operator(p)
    r0 = 0
    for s, n in enumerate(nr):
        A = b.csr(s)
        for k in range(K):
            ref = (A @ p[r0:r0 + n, k].ravel()).reshape(n, 2)
            scale = np.abs(ref).max()
            assert np.abs(q_as[r0:r0 + n, k] - ref).max() <= 1e-13 * scale, ("assembled", model, s, k)
            assert np.abs(q_mf[r0:r0 + n, k] - ref).max() <= 1e-13 * scale, ("matfree", model, s, k)
        r0 += n
    b.close()


@pytest.mark.parametrize("model", ["lin", "dnn", "per", "MR"])
def test_matfree_solve_matches_assembled_and_oracle(reduced_meshes, model):
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE, vref_points
    n = len(reduced_meshes)
    eps = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    uD = None
    if model == "dnn":
        uD = np.stack([np.stack([smooth_uD(81, 100 * s + i) for i in range(3)]) for s in range(n)])
    out = {}
    for op in ("matfree", "assembled"):
        b = RVEBatch().set_meshes(reduced_meshes, PARAM, model=model)
        b.set_operator(op).assemble()
        st, it = b.solve(eps, uD=uD, rtol=1e-11)
        out[op] = (b.tangent(), it.copy(), b.solution(0))
        b.close()
    Cm, itm, um = out["matfree"]
    Ca, ita, ua = out["assembled"]
    assert np.abs(itm - ita).max() <= 1                      # same operator, same Krylov sequence
    assert np.abs(Cm - Ca).max() <= 1e-9 * np.abs(Ca).max()
    assert np.abs(um - ua).max() <= 1e-8 * np.abs(ua).max()
    pts = vref_points(-1, -1, 2, 2, 21)
    for s in range(n):
        o = OracleRVE(*reduced_meshes[s], param=PARAM, model=model)
        Co = o.compute_tangent(uD_list=None if uD is None else list(uD[s]), pts=pts)
        assert np.abs(Cm[s] - Co).max() <= 1e-8 * np.abs(Co).max(), (model, s)


def test_matfree_full_mesh_snapshot_matches_oracle(full_mesh):
    """the shipped configuration (matrix-free, rtol 1e-10) on a `full` RVE: snapshot outputs against sparse LU"""
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.elasticity import macro_strain_voigt
    from oracle.rve_oracle import OracleRVE, vref_points
    eps = np.stack([macro_strain_voigt(i) for i in (2, 0)])[None]
    b = RVEBatch().set_meshes([full_mesh], PARAM, model="per", vref_box=(-1.0, -1.0, 2.0, 2.0)).assemble()     # default operator
    b.solve(eps, rtol=1e-10)
    out = b.snapshot()
    o = OracleRVE(*full_mesh, param=PARAM, model="per")
    ref = o.snapshot(vref_points(-1, -1, 2, 2, 21))
    for j, lab in enumerate(("S", "A")):
        for key, mine in (("solutions_" + lab, out["sol"][0, j]), ("sigma_" + lab, out["sigma"][0, j]),
                          ("sigmaTotal_" + lab, out["sigmaT"][0, j]), ("a_" + lab, out["a"][0, j]),
                          ("B_" + lab, out["B"][0, j])):
            r = np.asarray(ref[key])
            assert np.abs(mine.reshape(r.shape) - r).max() <= 1e-8 * max(np.abs(r).max(), 1e-3), key
    b.close()
