"""GPU property tests at BASELINE.json's full sizes ('full' 6x6 RVEs at lcar = 2/30, ~70k DOFs each), independent of
any LU solve: linearity in the macro strain, equilibrium residual against an independently assembled stiffness,
and the constraints each boundary model imposes (zero boundary / periodicity + zero mean / zero mean +
zero int u (x) n ds).  The stiffness and constraint operators come from the oracle's assembly routines (numpy)."""
import numpy as np
import pytest

from conftest import synthetic_param

pytestmark = pytest.mark.gpu
PARAM = (0.3, 10.0, 0.3, 1.0)


@pytest.fixture(scope="module")
def full_meshes_default():
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    par = synthetic_param(3, seed=31)
    return [buildRVEmesh_arrays(par[i].copy(), size="full") for i in range(3)]


def _node_perm(b, s, m):
    """library node order -> oracle node order"""
    va, vb = b.nodemap(s)
    ed = {(int(a), int(c)): m.nv + i for i, (a, c) in enumerate(m.edges)}
    return np.array([a if a == c else ed[(int(a), int(c))] for a, c in zip(va, vb)])


@pytest.mark.parametrize("model", ["lin", "per", "MR"])
def test_full_size_properties(full_meshes_default, model):
    from deepbnd_b200.batch import RVEBatch
    from oracle.rve_oracle import (P2Mesh, assemble, lame_table, load_vector, mean_constraint, outer_constraint,
                                   periodic_map)
    meshes = full_meshes_default
    n = len(meshes)
    b = RVEBatch().set_meshes(meshes, PARAM, model=model, vref_nb=21, vref_box=(-1, -1, 2, 2)).assemble()
    assert b.sizes(0)["n_rows"] > 30000                                   # BASELINE 'full' size
    e0, e2 = np.array([1.0, 0, 0]), np.array([0, 0, 1.0])
    eps = np.broadcast_to(np.stack([e0, e2, 0.3 * e0 - 1.7 * e2]), (n, 3, 3)).copy()
    status, iters = b.solve(eps, rtol=1e-11, maxit=8000)
    assert (status == 0).all()
    C = None
    for s in range(n):
        m = P2Mesh(*meshes[s])
        K, ed = assemble(m, lame_table(*PARAM))
        U = np.zeros((m.nn, 3, 2))
        U[_node_perm(b, s, m)] = b.solution(s)
        u = U.transpose(1, 0, 2).reshape(3, -1)                            # (load, 2*node+comp)
        scale = np.abs(u[:2]).max()
        # linearity in the macro strain
        assert np.abs(u[2] - (0.3 * u[0] - 1.7 * u[1])).max() < 1e-8 * scale
        F = np.stack([load_vector(m, ed, eps[s, k]) for k in range(3)])
        R = F - (K @ u.T).T                                                # residual on every dof
        if model == "lin":
            bd = np.zeros(2 * m.nn, dtype=bool)
            bd[2 * m.bnd_nodes] = bd[2 * m.bnd_nodes + 1] = True
            assert np.abs(u[:, bd]).max() == 0.0                           # zero fluctuation on the boundary
            assert np.abs(R[:, ~bd]).max() < 1e-8 * np.abs(F).max()        # equilibrium on the free dofs
        elif model == "per":
            master = periodic_map(m)
            un = u.reshape(3, m.nn, 2)
            assert np.abs(un - un[:, master]).max() < 1e-12 * scale        # periodic fluctuation
            Cm = mean_constraint(m)
            assert np.abs(Cm @ u.T).max() < 1e-9 * scale * 36.0            # zero mean
            # equilibrium: residual summed onto the master dofs vanishes up to the constant multiplier of the mean
            Rm = np.zeros((3, m.nn, 2))
            np.add.at(Rm, (slice(None), master), R.reshape(3, m.nn, 2))
            w = np.zeros(m.nn)
            np.add.at(w, master, Cm[0, 0::2])
            act = np.unique(master)
            for k in range(3):
                for c in range(2):
                    lam = (Rm[k, act, c] @ w[act]) / (w[act] @ w[act])
                    assert np.abs(Rm[k, act, c] - lam * w[act]).max() < 1e-8 * np.abs(F).max()
        else:
            Cm, Co = mean_constraint(m), outer_constraint(m)
            assert np.abs(Cm @ u.T).max() < 1e-9 * scale * 36.0
            assert np.abs(Co @ u.T).max() < 1e-9 * scale * 24.0            # int u (x) n ds = 0
            # equilibrium: the residual lies in the span of the 6 constraint rows
            Call = np.vstack([Cm, Co])
            for k in range(3):
                lam, *_ = np.linalg.lstsq(Call.T, R[k], rcond=None)
                assert np.abs(R[k] - Call.T @ lam).max() < 1e-8 * np.abs(F).max()
    b.close()


@pytest.mark.gpu
def test_piola_rotation_matches_a_rotated_rve(reduced_meshes):
    """physical check of bc_prediction.PiolaTransform_rotation_matricial and of the Vref node order: the boundary
    fluctuation of the RVE rotated by B under the load B eps B^T is B u(B^T x), i.e. P times the original Vref vector
    (the identity NN_elast.py:48-55 uses to get the eps22 prediction from the eps11 network)."""
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.bc_prediction import PiolaTransform_rotation_matricial
    xy, tri, reg = [np.asarray(a) for a in reduced_meshes[0]]
    x0, y0 = xy.min(0)
    L = np.ptp(xy[:, 0])
    assert abs(x0 + 0.5 * L) < 1e-12 and abs(y0 + 0.5 * L) < 1e-12 and abs(np.ptp(xy[:, 1]) - L) < 1e-12
    theta = 3 * np.pi / 2
    s, c = np.sin(theta), np.cos(theta)
    B = np.array([[c, -s], [s, c]])
    eps22 = np.array([[[0.0, 1.0, 0.0]]])                               # Voigt (e11, e22, g12)
    eps11 = np.array([[[1.0, 0.0, 0.0]]])                               # B e2 = e1: B eps22 B^T = eps11
    b = RVEBatch()
    b.set_meshes([(xy, tri, reg)], PARAM, model="per", vref_nb=21).assemble()
    b.solve(eps22.copy())
    S = b.snapshot()["sol"][0, 0]
    b.set_meshes([(np.ascontiguousarray(xy @ B.T), tri, reg)], PARAM, model="per", vref_nb=21).assemble()
    b.solve(eps11.copy())
    Sr = b.snapshot()["sol"][0, 0]
    b.close()
    P = PiolaTransform_rotation_matricial(theta, box=(x0, y0, L, L), nb=21)
    assert np.abs(Sr - P @ S).max() <= 1e-8 * np.abs(S).max()
