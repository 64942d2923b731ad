"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same
meshes.  Tolerances: 1e-8 relative (BASELINE.json north_star) on nodal displacement, Vref samples,
RB coefficients and tangent entries; assembled matrices/vectors to 1e-12."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import smooth_uD

pytestmark = pytest.mark.gpu

PARAM = (0.3, 10.0, 0.3, 1.0)
EPS3 = np.eye(3)


def _oracle_rows(o, model):
    from oracle.rve_oracle import periodic_map
    m = o.mesh
    if o._lu is None:
        o._factor()
    if model in ("lin", "dnn"):
        return o.free[0::2] // 2
    if model == "per":
        return np.unique(periodic_map(m))
    return np.arange(m.nn)


def _perm(batch, sys, o, model):
    """position in the oracle's active ordering of every library row"""
    m = o.mesh
    va, vb = batch.rowmap(sys)
    ed = {(int(a), int(b)): m.nv + i for i, (a, b) in enumerate(m.edges)}
    rown = np.array([a if a == b else ed[(int(a), int(b))] for a, b in zip(va, vb)])
    pos = {int(n): i for i, n in enumerate(_oracle_rows(o, model))}
    return np.array([pos[int(n)] for n in rown]), rown


def _oracle_matrix(o, model):
    o._factor()
    if model in ("lin", "dnn"):
        return o.K[o.free][:, o.free].tocsr()
    if model == "per":
        return (o.P.T @ o.K @ o.P).tocsr()
    return o.K.tocsr()


@pytest.mark.parametrize("model", ["lin", "per", "MR"])
def test_assembly_matches_oracle(reduced_meshes, model):
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE
    b = RVEBatch().set_meshes(reduced_meshes, PARAM, model=model).assemble()
    for s in (0, 2):
        o = OracleRVE(*reduced_meshes[s], param=PARAM, model=model)
        A = _oracle_matrix(o, model)
        perm, _ = _perm(b, s, o, model)
        dperm = np.stack([2 * perm, 2 * perm + 1], 1).ravel()
        Al = b.csr(s)
        D = (Al - A[dperm][:, dperm]).tocsr()
        err = np.abs(D.data).max() if D.nnz else 0.0
        assert err <= 1e-12 * np.abs(A.data).max(), (model, s, err)
    b.close()


def _level1_reference(o, model, rown, nc):
    """A1 = sum_e Z_e^T K_e Z_e with the bilinear transfer Z and the library's clamping rule for elements
    that reach a third coarse cell (k_galerkin1): identical to Z^T A Z when no element does."""
    from oracle.rve_oracle import periodic_map
    m = o.mesh
    nx = nc + 1
    master = np.arange(m.nn)
    if model == "lin":
        master[m.bnd_nodes] = -1
    elif model == "per":
        master = periodic_map(m)
    pos = -np.ones(m.nn, dtype=np.int64)
    pos[rown] = np.arange(len(rown))
    node_row = np.where(master >= 0, pos[np.maximum(master, 0)], -1)
    X = m.xy[m.cells]
    H = (m.x1 - m.x0) / nc
    fx, fy = (X[:, :, 0] - m.x0) / H, (X[:, :, 1] - m.y0) / H
    ci, cj = np.clip(fx.astype(int), 0, nc - 1), np.clip(fy.astype(int), 0, nc - 1)
    s, t = np.clip(fx - ci, 0, 1), np.clip(fy - cj, 0, 1)
    imin, jmin = ci.min(1, keepdims=True), cj.min(1, keepdims=True)
    far_i, far_j = (ci - imin) >= 2, (cj - jmin) >= 2
    ci, s = np.where(far_i, imin + 1, ci), np.where(far_i, 1.0, s)
    cj, t = np.where(far_j, jmin + 1, cj), np.where(far_j, 1.0, t)
    R = node_row[m.cells]
    W, N = [], []
    for di, dj, w in ((0, 0, (1 - s) * (1 - t)), (1, 0, s * (1 - t)), (0, 1, (1 - s) * t), (1, 1, s * t)):
        i, j = ci + di, cj + dj
        if model == "per":
            nd, ok = (j % nc) * nx + (i % nc), np.ones_like(i, dtype=bool)
        elif model == "lin":
            ok, nd = (i >= 1) & (i <= nc - 1) & (j >= 1) & (j <= nc - 1), j * nx + i
        else:
            ok, nd = np.ones_like(i, dtype=bool), j * nx + i
        ok = ok & (R >= 0)
        W.append(np.where(ok, w, 0.0))
        N.append(np.where(ok, nd, 0))
    W, N = np.stack(W, 2), np.stack(N, 2)
    nt = m.nt
    Ke4 = o.ed['Ke'].reshape(nt, 6, 2, 6, 2)
    I, J, V = [], [], []
    for ca in range(4):
        for cb in range(4):
            ww = W[:, :, ca][:, :, None] * W[:, :, cb][:, None, :]
            for c in range(2):
                for d in range(2):
                    I.append((2 * N[:, :, ca][:, :, None] + c + 0 * N[:, :, cb][:, None, :]).ravel())
                    J.append((2 * N[:, :, cb][:, None, :] + d + 0 * N[:, :, ca][:, :, None]).ravel())
                    V.append((ww * Ke4[:, :, c, :, d]).ravel())
    A1 = sp.coo_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(2 * nx * nx, 2 * nx * nx))
    return A1.toarray(), int((far_i | far_j).any(1).sum())


@pytest.mark.parametrize("model", ["lin", "per", "MR"])
def test_coarse_operators_are_galerkin(reduced_meshes, model):
    """level-1 operator == sum_e Z_e^T K_e Z_e (Z^T A Z up to the clamping of over-long elements);
    deeper levels == P^T A P."""
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE
    meshes = reduced_meshes[:2]
    b = RVEBatch().set_meshes(meshes, PARAM, model=model).assemble()
    dims = b.coarse_levels()
    s = 1
    o = OracleRVE(*meshes[s], param=PARAM, model=model)
    m = o.mesh
    _, rown = _perm(b, s, o, model)
    nx = int(round(np.sqrt(dims[0])))
    nc = nx - 1

    def node(i, j, ncl):
        nxl = ncl + 1
        if model == "per":
            return (j % ncl) * nxl + (i % ncl), np.ones_like(i, dtype=bool)
        if model == "lin":
            ok = (i >= 1) & (i <= ncl - 1) & (j >= 1) & (j <= ncl - 1)
        else:
            ok = np.ones_like(i, dtype=bool)
        return j * nxl + i, ok

    A1, nclamped = _level1_reference(o, model, rown, nc)
    print("level-1 grid", nc, "clamped elements", nclamped)
    A1l = b.coarse_dense(s, 0)
    assert np.abs(A1 - A1l).max() <= 1e-11 * np.abs(A1).max(), np.abs(A1 - A1l).max()
    # next level
    if len(dims) > 1:
        ncc = nc // 2
        rows, cols, vals = [], [], []
        for j in range(nc + 1):
            for i in range(nc + 1):
                nf, okf = node(np.array([i]), np.array([j]), nc)
                if not okf[0] or (model == "per" and (i >= nc or j >= nc)):
                    continue
                for I, wi in ([(i // 2, 1.0)] if i % 2 == 0 else [(i // 2, 0.5), (i // 2 + 1, 0.5)]):
                    for J, wj in ([(j // 2, 1.0)] if j % 2 == 0 else [(j // 2, 0.5), (j // 2 + 1, 0.5)]):
                        ncd, okc = node(np.array([I]), np.array([J]), ncc)
                        if okc[0]:
                            for c in range(2):
                                rows.append(2 * nf[0] + c); cols.append(2 * ncd[0] + c); vals.append(wi * wj)
        P = sp.coo_matrix((vals, (rows, cols)), shape=(2 * dims[0], 2 * dims[1])).tocsr()
        A2 = (P.T @ sp.csr_matrix(A1l) @ P).toarray()
        A2l = b.coarse_dense(s, 1)
        assert np.abs(A2 - A2l).max() <= 1e-11 * np.abs(A2).max(), np.abs(A2 - A2l).max()
    b.close()


@pytest.mark.parametrize("precond", ["jacobi", "twolevel"])
@pytest.mark.parametrize("model", ["lin", "dnn", "per", "MR"])
def test_solution_tangent_snapshot_match_oracle(reduced_meshes, model, precond):
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE, vref_points, macro_strain
    meshes = reduced_meshes
    n = len(meshes)
    pts = vref_points(-1, -1, 2, 2, 21)
    b = RVEBatch().set_meshes(meshes, PARAM, model=model, vref_nb=21).assemble()
    eps = np.broadcast_to(EPS3, (n, 3, 3)).copy()
    uD = None
    if model == "dnn":
        uD = np.stack([np.stack([smooth_uD(81, 100 * s + i) for i in range(3)]) for s in range(n)])
    status, iters = b.solve(eps, uD=uD, rtol=1e-11, maxit=6000, precond=precond)
    assert (status == 0).all()
    print(model, precond, "iters", iters)
    C = b.tangent()
    snap = b.snapshot()
    for s in range(n):
        o = OracleRVE(*meshes[s], param=PARAM, model=model)
        Co = o.compute_tangent(uD_list=None if uD is None else uD[s], pts=pts)
        U = b.solution(s)
        for i in range(3):
            uo = o.sol[i].reshape(-1, 2)
            err = np.linalg.norm(U[:, i, :] - uo) / max(np.linalg.norm(uo), 1e-300)
            assert err < 1e-8, (model, precond, s, i, err)
        assert np.abs(C[s] - Co).max() < 1e-8 * np.abs(Co).max(), (model, s, np.abs(C[s] - Co).max())
        for i in range(3):
            a_o, B_o = o.affine_local(i, (0, 1))
            sig_o = o.homogenise([0, 1], i).flatten()[[0, 3, 2]]
            sigT_o = o.homogenise([0, 1, 2, 3], i).flatten()[[0, 3, 2]]
            sol_o = o.eval_at(o.sol[i], pts).ravel()
            # scale = the displacement field itself: under 'lin' the samples on the boundary vanish
            sc = max(np.abs(sol_o).max(), np.abs(o.sol[i]).max())
            assert np.abs(snap["sol"][s, i] - sol_o).max() < 1e-8 * sc
            assert np.abs(snap["sigma"][s, i] - sig_o).max() < 1e-8 * np.abs(sig_o).max()
            assert np.abs(snap["sigmaT"][s, i] - sigT_o).max() < 1e-8 * np.abs(sigT_o).max()
            assert np.abs(snap["a"][s, i] - a_o).max() < 1e-8 * max(np.abs(a_o).max(), sc)
            assert np.abs(snap["B"][s, i] - B_o).max() < 1e-8 * max(np.abs(B_o).max(), sc)
    b.close()


def test_rhs_matches_oracle(reduced_meshes):
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE, load_vector, vref_points, macro_strain, strain2voigt
    meshes = reduced_meshes[:2]
    pts = vref_points(-1, -1, 2, 2, 21)
    b = RVEBatch().set_meshes(meshes, PARAM, model="dnn", vref_nb=21).assemble()
    uD = np.stack([np.stack([smooth_uD(81, 7 * s + i) for i in range(3)]) for s in range(2)])
    eps = np.broadcast_to(EPS3, (2, 3, 3)).copy()
    b.solve(eps, uD=uD, rtol=1e-6, maxit=500)
    for s in range(2):
        o = OracleRVE(*meshes[s], param=PARAM, model="dnn")
        o._factor()
        m = o.mesh
        perm, _ = _perm(b, s, o, "dnn")
        rl = b.rhs(s)
        for i in range(3):
            g, B = o.dirichlet_from_vref(uD[s, i], pts)
            F = load_vector(m, o.ed, strain2voigt(macro_strain(i) - B))
            u = np.zeros(2 * m.nn)
            u[2 * m.bnd_nodes] = g[:, 0]
            u[2 * m.bnd_nodes + 1] = g[:, 1]
            ro = (F - o.K @ u)[o.free].reshape(-1, 2)[perm]
            assert np.abs(rl[:, i, :] - ro).max() < 1e-12 * np.abs(ro).max()
    b.close()


def test_full_mesh_snapshot_per_and_projection(full_mesh):
    """config-1 style: one 'full' RVE, periodic model, loads S (2) and A (0), Vref = internal boundary,
    plus the RB projection Y = U M W^T against numpy."""
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE, vref_points, vref_mass, rb_project
    pts = vref_points(-1, -1, 2, 2, 21)
    for model in ("per", "MR"):
        b = RVEBatch().set_meshes([full_mesh], PARAM, model=model, vref_nb=21, vref_box=(-1, -1, 2, 2)).assemble()
        eps = np.array([[[0, 0, 1.0], [1.0, 0, 0]]])
        status, iters = b.solve(eps, rtol=1e-11, maxit=8000)
        snap = b.snapshot()
        o = OracleRVE(*full_mesh, param=PARAM, model=model)
        so = o.snapshot(pts)
        for k, lab in enumerate(("S", "A")):
            for key, okey in (("sol", "solutions_"), ("sigma", "sigma_"), ("sigmaT", "sigmaTotal_"), ("a", "a_"), ("B", "B_")):
                ref = np.asarray(so[okey + lab]).ravel()
                got = snap[key][0, k].ravel()
                assert np.abs(got - ref).max() < 1e-8 * max(np.abs(ref).max(), 1e-3), (model, key, lab)
        rng = np.random.RandomState(0)
        W = np.linalg.qr(rng.randn(162, 160))[0].T
        M = vref_mass(pts)
        Y = b.project(np.stack([W, W]), M, translate=snap["a"])
        for k in range(2):
            Ut = snap["sol"][0, k].reshape(-1, 2) + snap["a"][0, k]
            Yo = rb_project(Ut.ravel()[None], W, M)[0]
            assert np.abs(Y[0, k] - Yo).max() < 1e-10 * np.abs(Yo).max()
        b.close()


def test_homogeneous_material_known_answer(coarse_reduced_meshes):
    """analytic known answer: homogeneous material => u~ = 0, C = plane-stress D, every model."""
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.elasticity import getLameTable
    lam, mu = getLameTable(0.3, 1.0, 0.3, 1.0)[0]
    D = np.array([[lam + 2 * mu, lam, 0], [lam, lam + 2 * mu, 0], [0, 0, mu]])
    n = len(coarse_reduced_meshes)
    eps = np.broadcast_to(EPS3, (n, 3, 3)).copy()
    for model in ("lin", "per", "MR"):
        b = RVEBatch().set_meshes(coarse_reduced_meshes, (0.3, 1.0, 0.3, 1.0), model=model).assemble()
        b.solve(eps, rtol=1e-11, maxit=4000)
        C = b.tangent()
        assert np.abs(C - D[None]).max() < 1e-9, (model, np.abs(C - D[None]).max())
        b.close()


def test_energy_ordering_and_symmetry(coarse_reduced_meshes):
    """MR <= per <= lin in the energy sense; tangents symmetric positive definite."""
    from deepbnd_b200.batch import RVEBatch
    n = len(coarse_reduced_meshes)
    eps = np.broadcast_to(EPS3, (n, 3, 3)).copy()
    C = {}
    for model in ("lin", "per", "MR"):
        b = RVEBatch().set_meshes(coarse_reduced_meshes, PARAM, model=model).assemble()
        b.solve(eps, rtol=1e-11, maxit=4000)
        C[model] = b.tangent()
        b.close()
    for s in range(n):
        for model in C:
            assert np.abs(C[model][s] - C[model][s].T).max() < 1e-8
            assert np.linalg.eigvalsh(0.5 * (C[model][s] + C[model][s].T)).min() > 0
        assert np.linalg.eigvalsh(C["per"][s] - C["MR"][s]).min() > -1e-9
        assert np.linalg.eigvalsh(C["lin"][s] - C["per"][s]).min() > -1e-9


def test_config1_against_committed_goldens():
    """BASELINE config 1: the single RVE of the reference's parameter table (tests/test_meshes.py:63-98) on the
    GPU against tests/golden/c1_oracle.npz (oracle sparse LU; generator: tests/golden/make_goldens.py)."""
    import os
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = np.load(os.path.join(gdir, "c1_oracle.npz"))
    table = np.loadtxt(os.path.join(gdir, "paramRVE_test_meshes.txt"))
    reduced = buildRVEmesh_arrays(table.copy(), size="reduced")
    for model in ("lin", "per", "MR"):
        b = RVEBatch().set_meshes([reduced], PARAM, model=model).assemble()
        b.solve(EPS3[None].copy(), rtol=1e-11, maxit=6000)
        C = b.tangent()[0]
        ref = gold["tangent_reduced_" + model]
        assert np.abs(C - ref).max() < 1e-8 * np.abs(ref).max(), model
        b.close()
    full = buildRVEmesh_arrays(table.copy(), size="full")
    for model in ("per", "MR"):
        b = RVEBatch().set_meshes([full], PARAM, model=model, vref_nb=21, vref_box=(-1, -1, 2, 2)).assemble()
        b.solve(np.array([[[0, 0, 1.0], [1.0, 0, 0]]]), rtol=1e-11, maxit=8000)
        snap = b.snapshot()
        for k, lab in enumerate(("S", "A")):
            for key, okey in (("sol", "solutions_"), ("sigma", "sigma_"), ("sigmaT", "sigmaTotal_"), ("a", "a_"), ("B", "B_")):
                ref = gold["full_%s_%s%s" % (model, okey, lab)].ravel()
                got = snap[key][0, k].ravel()
                assert np.abs(got - ref).max() < 1e-8 * max(np.abs(ref).max(), 1e-3), (model, key, lab)
        b.close()


def test_ragged_batch_and_refined_mesh(coarse_reduced_meshes):
    """one batch mixing mesh resolutions (lcar 0.1, 2/30 and the 'refined' 1/30 of BASELINE config 5): every
    system must match its own oracle solve."""
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    from oracle.rve_oracle import OracleRVE
    from conftest import synthetic_param
    par = synthetic_param(3, seed=21)
    meshes = [coarse_reduced_meshes[0], buildRVEmesh_arrays(par[1].copy(), size="reduced"),
              buildRVEmesh_arrays(par[2].copy(), size="reduced", lcar=1.0 / 30.0)]
    sizes = [len(m[1]) for m in meshes]
    assert sizes[2] > 3 * sizes[1] > 3 * sizes[0]
    b = RVEBatch().set_meshes(meshes, PARAM, model="per").assemble()
    status, iters = b.solve(np.broadcast_to(EPS3, (3, 3, 3)).copy(), rtol=1e-11, maxit=6000)
    assert (status == 0).all()
    C = b.tangent()
    for s in range(3):
        Co = OracleRVE(*meshes[s], param=PARAM, model="per").compute_tangent()
        assert np.abs(C[s] - Co).max() < 1e-8 * np.abs(Co).max(), (s, np.abs(C[s] - Co).max())
    b.close()


def test_error_codes_and_state_machine(coarse_reduced_meshes):
    """the C-ABI reports misuse through status codes + rve_last_error, never by crashing or falling back."""
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200 import _lib as L
    b = RVEBatch()
    with pytest.raises(L.RVEError) as ei:
        b.assemble()                                                   # before set_batch
    assert ei.value.code == -3
    m = coarse_reduced_meshes[0]
    with pytest.raises(L.RVEError) as ei:
        b.set_meshes([m], PARAM, model="lin").solve(EPS3[None].copy())  # solve before assemble
    assert ei.value.code == -3
    b.assemble()
    with pytest.raises(L.RVEError) as ei:
        b.solve(EPS3[None].copy(), uD=np.zeros((1, 3, 162)))           # uD on a non-dnn batch
    assert ei.value.code == -1
    with pytest.raises(L.RVEError) as ei:
        b.tangent()                                                    # epilogue before solve
    assert ei.value.code == -3
    bad = (m[0], np.vstack([m[1], [[0, 1, 10 ** 6]]]).astype(np.int32), np.append(m[2], 0).astype(np.int32))
    with pytest.raises(L.RVEError) as ei:
        b.set_meshes([bad], PARAM, model="lin")                        # vertex id out of range
    assert ei.value.code == -4
    shifted = (m[0] * 1.5, m[1], m[2])
    with pytest.raises(L.RVEError) as ei:
        b.set_meshes([m, shifted], PARAM, model="lin")                 # different box sizes in one batch
    assert ei.value.code == -4
    # maxit too small: RVE_E_NOCONV with per-system status, partial result still retrievable
    b.set_meshes([m], PARAM, model="lin").assemble()
    status, iters = b.solve(EPS3[None].copy(), maxit=3, raise_on_noconv=False)
    assert status[0] == 1 and iters[0] == 3
    with pytest.raises(L.RVEError) as ei:
        b.solve(EPS3[None].copy(), maxit=3)
    assert ei.value.code == -5
    b.close()


@pytest.mark.parametrize("model", ["lin", "per", "MR"])
def test_single_load_and_no_vref(coarse_reduced_meshes, model):
    """n_rhs = 1 (the K = 1 instantiation of the PCG kernels; MR still solves its 3 traction problems) and a batch
    without a Vref space (vref_nb = 0): nodal field and sigma of the shear load against the oracle."""
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import OracleRVE
    meshes = coarse_reduced_meshes[:3]
    b = RVEBatch().set_meshes(meshes, PARAM, model=model, vref_nb=0).assemble()
    eps = np.broadcast_to(np.array([[0.0, 0.0, 1.0]]), (3, 1, 3)).copy()
    status, iters = b.solve(eps, rtol=1e-11, maxit=6000)
    assert (status == 0).all()
    snap = b.snapshot()
    assert snap["sol"] is None
    for s in range(3):
        o = OracleRVE(*meshes[s], param=PARAM, model=model)
        o.solve_load(2)
        uo = o.sol[2].reshape(-1, 2)
        U = b.solution(s)[:, 0, :]
        assert np.linalg.norm(U - uo) < 1e-8 * np.linalg.norm(uo), (model, s)
        sig = o.homogenise([0, 1], 2).flatten()[[0, 3, 2]]
        assert np.abs(snap["sigma"][s, 0] - sig).max() < 1e-8 * np.abs(sig).max()
    b.close()


def test_spmv_microbenchmark_entry(coarse_reduced_meshes):
    """rve_bench_spmv runs the production SpMV kernel for 1..3 right-hand sides and reports the CSR-accounted bytes."""
    from deepbnd_b200.batch import RVEBatch
    b = RVEBatch().set_meshes(coarse_reduced_meshes, PARAM, model="per").assemble()
    nrow = sum(b.sizes(s)["n_rows"] for s in range(len(coarse_reduced_meshes)))
    nblk = sum(b.sizes(s)["n_blocks"] for s in range(len(coarse_reduced_meshes)))
    for k in (1, 2, 3):
        ms, by = b.bench_spmv(n_rhs=k, reps=5)
        assert ms > 0
        n, nnz = 2 * nrow, 4 * nblk
        assert by == pytest.approx(12.0 * nnz + 4.0 * (n + len(coarse_reduced_meshes)) + 16.0 * n * k)
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["lin", "per", "MR", "dnn"])
def test_device_symbolic_matches_host_hook(reduced_meshes, full_mesh, model):
    """the GPU symbolic phase (rve_devsym.cuh: P2 numbering, constraint map, Morton row order, window sort) yields the
    numbering of the CPU test hook rve_host_symbolic entry by entry, on a reduced and on a full mesh."""
    from deepbnd_b200.batch import RVEBatch, host_symbolic
    for mesh in (reduced_meshes[0], full_mesh):
        ref = host_symbolic(*mesh, model=model, vref_nb=21)
        b = RVEBatch().set_meshes([mesh], PARAM, model=model, vref_nb=21)
        sz = b.sizes(0)
        assert sz["n_nodes"] == ref["n_nodes"] and sz["n_rows"] == ref["n_rows"]
        va, vb = b.rowmap(0)
        assert np.array_equal(va, ref["row_va"]) and np.array_equal(vb, ref["row_vb"])
        nva, nvb = b.nodemap(0)
        tri = np.asarray(mesh[1])
        nv = len(mesh[0])
        assert np.array_equal(nva[:nv], np.arange(nv)) and np.array_equal(nvb[:nv], np.arange(nv))
        keys = nva[nv:].astype(np.int64) * (1 << 32) + nvb[nv:]
        assert np.all(np.diff(keys) > 0)                                 # mid-edge nodes sorted by (va, vb), unique
        e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
        assert np.array_equal(np.unique(e[:, 0].astype(np.int64) * (1 << 32) + e[:, 1]), keys)
        b.assemble()
        assert b.sizes(0)["n_blocks"] == ref["n_blocks"]                 # true block count found from the element fans
        b.close()


def test_two_handles_solve_concurrently(reduced_meshes, full_mesh):
    """two handles driven from two host threads at once (what DoubleBuffer and bench.py's e2e loop do), with batches
    whose shared-memory tiles differ: same results as the sequential runs, no launch failure."""
    import threading
    from deepbnd_b200.batch import RVEBatch
    eps2 = np.stack([EPS3[2], EPS3[0]])
    jobs = [([full_mesh] * 3, "per"), (list(reduced_meshes[:4]), "lin")]

    def run(meshes, model, reps, out):
        b = RVEBatch()
        for _ in range(reps):
            b.set_meshes(meshes, PARAM, model=model, vref_nb=21).assemble()
            b.solve(np.broadcast_to(eps2, (len(meshes), 2, 3)).copy())
            out.append(b.snapshot()["sol"].copy())
        b.close()

    ref = [[], []]
    for k, (meshes, model) in enumerate(jobs):
        run(meshes, model, 1, ref[k])
    got = [[], []]
    errs = []

    def guarded(k):
        try:
            run(jobs[k][0], jobs[k][1], 4, got[k])
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=guarded, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for k in range(2):
        scale = np.abs(ref[k][0]).max()
        for sol in got[k]:
            assert np.abs(sol - ref[k][0]).max() <= 1e-8 * scale


def test_device_symbolic_rejects_bad_meshes(coarse_reduced_meshes):
    """the GPU symbolic phase validates the meshes itself (region tags, degenerate / non-manifold triangles, periodic
    boundary matching, Vref box) and reports RVE_E_MESH with the offending system; a good batch still works after."""
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200 import _lib as L
    xy, tri, reg = [np.array(a) for a in coarse_reduced_meshes[0]]
    good = (xy, tri, reg)
    b = RVEBatch()

    def expect_mesh_error(meshes, model, text, **kw):
        with pytest.raises(L.RVEError) as ei:
            b.set_meshes(meshes, PARAM, model=model, **kw)
        assert ei.value.code == -4 and text in str(ei.value), str(ei.value)

    bad_reg = reg.copy(); bad_reg[0] = reg.min() + 7
    expect_mesh_error([good, (xy, tri, bad_reg)], "lin", "system 1: region tag")
    flat = tri.copy(); flat[5] = [tri[5, 0], tri[5, 1], tri[5, 0]]                   # zero-area triangle
    expect_mesh_error([(xy, flat, reg)], "lin", "degenerate triangle")
    dup = np.vstack([tri, tri[:1]]); dup_reg = np.append(reg, reg[0])                 # an element twice: edges with 3-4 owners
    expect_mesh_error([(xy, dup.astype(np.int32), dup_reg.astype(np.int32))], "lin", "non-manifold")
    # periodic model on a mesh whose right side is shifted: no partner on the left side
    onr = np.abs(xy[:, 0] - xy[:, 0].max()) < 1e-12
    inner = onr & (np.abs(xy[:, 1] - xy[:, 1].min()) > 1e-9) & (np.abs(xy[:, 1] - xy[:, 1].max()) > 1e-9)
    moved = xy.copy(); moved[np.flatnonzero(inner)[0], 1] += 1e-3
    expect_mesh_error([(moved, tri, reg)], "per", "periodic")
    # dnn: the mesh boundary must coincide with the Vref box
    expect_mesh_error([good], "dnn", "Vref box", vref_box=(xy[:, 0].min() - 0.1, xy[:, 1].min(), np.ptp(xy[:, 0]), np.ptp(xy[:, 1])))
    # Vref sample points outside the mesh
    expect_mesh_error([good], "lin", "outside the mesh", vref_box=(xy[:, 0].min() - 1.0, xy[:, 1].min(), np.ptp(xy[:, 0]), np.ptp(xy[:, 1])))
    b.set_meshes([good], PARAM, model="lin").assemble()
    b.solve(EPS3[None].copy())
    C = b.tangent()[0]
    assert np.all(np.linalg.eigvalsh(0.5 * (C + C.T)) > 0)
    b.close()
