"""a16: elliptical inclusions (ellipse2_mesh.py:27-39) in the mesher and the morphing, the boundary reference mesh
writer (degenerated_rectangle_mesh.py:3-35) and the macro-scale tangent lookup (multiscale/misc.py:9-44)."""
import os

import numpy as np
import pytest

from conftest import synthetic_param

PARAM = (0.3, 10.0, 0.3, 1.0)


def _ellipse_param(seed=3):
    par = synthetic_param(1, seed=seed)[0]
    par[:, 3] = 0.7
    par[:, 4] = np.linspace(0.0, np.pi, 36)
    par[:, 2] *= 1.1
    return par


def _areas(P, T):
    p0, p1, p2 = P[T[:, 0]], P[T[:, 1]], P[T[:, 2]]
    return 0.5 * ((p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p1[:, 1] - p0[:, 1]) * (p2[:, 0] - p0[:, 0]))


@pytest.mark.parametrize("size,n,total", [("reduced", 4, 4.0), ("full", 36, 36.0)])
def test_elliptical_inclusions_are_meshed(size, n, total):
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    par = _ellipse_param()
    P, T, R = buildRVEmesh_arrays(par, size=size)                       # permutes par in place (mesher order)
    A = _areas(P, T)
    assert A.min() > 0 and abs(A.sum() - total) < 1e-12
    exact = (np.pi * par[:n, 2] ** 2 * par[:n, 3]).sum()                # sum of pi a b
    incl = A[(R == 0) | (R == 2)].sum()
    assert -0.015 < incl / exact - 1.0 < 0.0                             # inscribed polygons: slightly smaller
    assert sorted(np.unique(R)) == ([0, 1] if size == "reduced" else [0, 1, 2, 3])


def test_circles_are_unchanged_by_the_ellipse_code_path():
    """e = 1 rows take the circle path: the reference's volume goldens (tests/test_oracle.py) stay bit-identical"""
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays, mesh_circles_in_box
    par = synthetic_param(1, seed=5)[0]
    P, T, R = buildRVEmesh_arrays(par.copy(), size="reduced")
    q = par.copy()
    buildRVEmesh_arrays(q, size="reduced")
    P2, T2, R2 = mesh_circles_in_box(q[:4, :3], (-1.0, -1.0, 2.0, 2.0), 2.0 / 30.0, 31)
    assert np.array_equal(P, P2) and np.array_equal(T, T2) and np.array_equal(R, R2)


def test_morphed_ellipses_keep_topology_and_area():
    from deepbnd_b200.mesh.rve_morph import MorphTemplate, buildRVEmesh_morphed
    par = _ellipse_param(seed=4)
    par[:, 2] *= 0.85
    tpl = MorphTemplate(par, size="reduced")
    P, T, R = buildRVEmesh_morphed(par, tpl)
    A = _areas(P, T)
    assert T is tpl.tri and A.min() > 0 and abs(A.sum() - 4.0) < 1e-12
    exact = (np.pi * par[:4, 2] ** 2 * par[:4, 3]).sum()
    assert abs(A[R == 0].sum() / exact - 1.0) < 0.015


def test_oracle_on_elliptical_inclusions_is_anisotropic_but_sane():
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    from oracle.rve_oracle import OracleRVE
    par = _ellipse_param()
    par[:, 4] = 0.0                                                      # all major axes along x: stiffer in x
    mesh = buildRVEmesh_arrays(par, size="reduced", lcar=0.1)
    C = OracleRVE(*mesh, param=PARAM, model="per").compute_tangent()
    assert np.allclose(C, C.T, atol=1e-9 * np.abs(C).max()) and C[0, 0] > C[1, 1] * 1.01
    Ch = OracleRVE(*mesh, param=(0.3, 1.0, 0.3, 1.0), model="per").compute_tangent()
    lam, mu = 0.3 / (1 - 0.09), 0.5 / 1.3                                # plane stress: lam* = nu E / (1 - nu^2)
    D = np.array([[lam + 2 * mu, lam, 0], [lam, lam + 2 * mu, 0], [0, 0, mu]])
    assert np.abs(Ch - D).max() < 1e-10


def test_boundary_reference_mesh_writer_roundtrip(tmp_path):
    from deepbnd_b200.mesh.degenerated_rectangle_mesh import degeneratedBoundaryRectangleMesh
    from deepbnd_b200.mesh.mesh_io import read_xdmf_mesh
    from deepbnd_b200.simulation_snapshots import vref_box_from_mesh
    from deepbnd_b200.vref import vref_points
    fn = str(tmp_path / "boundaryMesh.xdmf")
    m = degeneratedBoundaryRectangleMesh(x0=-1.0, y0=-1.0, Lx=2.0, Ly=2.0, Nb=21)      # simulation_snapshots.py:132-136
    m.generate()
    m.write(fn, 'fenics')
    for suffix in ("", "_faces", "_regions"):
        assert os.path.exists(fn[:-5] + suffix + ".xdmf") and os.path.exists(fn[:-5] + suffix + ".h5")
    assert vref_box_from_mesh(fn) == ((-1.0, -1.0, 2.0, 2.0), 21)
    P, T, R = read_xdmf_mesh(fn)
    assert np.array_equal(P, vref_points()) and T.shape == (80, 3) and abs(_areas(P, T).sum() - 4.0) < 1e-14
    with pytest.raises(Exception):
        open(fn[:-5] + ".h5", "wb").write(b"not hdf5")                  # an unreadable file must not fall back silently
        vref_box_from_mesh(fn)


def test_macro_tangent_lookup():
    from deepbnd_b200.multiscale_misc import Chom_multiscale, Chom_precomputed_dist
    rng = np.random.RandomState(0)
    tang = rng.rand(5, 3, 3)
    mp = rng.randint(0, 5, 12)
    ch = Chom_multiscale(tang, mp)
    assert np.array_equal(ch.eval_cell(3), tang[mp[3]].flatten()) and ch().shape == (12, 3, 3)
    pts = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [2, 0], [2, 1.0]])
    cells = np.array([[0, 1, 2], [1, 3, 2], [1, 4, 3], [4, 5, 3]])
    centers = np.array([[0.3, 0.3], [1.7, 0.6]])
    cd = Chom_precomputed_dist(tang[:2], centers, pts, cells)
    assert cd.map.tolist() == [0, 0, 1, 1]


@pytest.mark.gpu
def test_cuda_path_on_elliptical_inclusions_matches_oracle():
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    from oracle.rve_oracle import OracleRVE
    par = _ellipse_param()
    mesh = buildRVEmesh_arrays(par, size="reduced")
    b = RVEBatch().set_meshes([mesh], PARAM, model="per").assemble()
    b.solve(np.eye(3)[None].copy(), rtol=1e-11)
    C = b.tangent()[0]
    Co = OracleRVE(*mesh, param=PARAM, model="per").compute_tangent()
    assert np.abs(C - Co).max() < 1e-8 * np.abs(Co).max()
    b.close()
