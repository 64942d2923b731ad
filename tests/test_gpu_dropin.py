"""GPU tests of the script-level drop-ins: paramRVEdataset.hd5 (+ bcs.hd5) in, snapshots.hd5 / tangents.hd5 out,
in the reference's dataset layout (simulation_snapshots.py:80-83, tangents_prediction.py:56-57), checked row by
row against the CPU oracle on the same cached meshes."""
import os

import numpy as np
import pytest

from conftest import smooth_uD, synthetic_param

pytestmark = pytest.mark.gpu
PARAM = [0.3, 10.0, 0.3, 1.0]


def test_buildSnapshots_writes_reference_layout_and_matches_oracle(tmp_path, monkeypatch):
    monkeypatch.setenv("DEEPBND_KEEP_MESHES", "1")        # keep the meshes of createMesh=True: the oracle solves the same ones
    import deepbnd_b200.wrapper_h5py as myhd
    from deepbnd_b200.simulation_snapshots import buildSnapshots, LABELS
    from deepbnd_b200.mesh.mesh_io import load_mesh
    from oracle.rve_oracle import OracleRVE, vref_points
    ns = 3
    par = synthetic_param(ns, seed=13)
    pfile, sfile = str(tmp_path / "paramRVEdataset.hd5"), str(tmp_path / "snapshots_{0}.hd5")
    myhd.savehd5(pfile, par, 'param', 'w')
    mname = str(tmp_path / "meshes" / "mesh_temp_{0}.xdmf")
    for run in range(2):                                   # two "ranks" on the same GPU: ids 0,2 and 1
        buildSnapshots(PARAM, ["", pfile, sfile.format(run), mname.format(run)], 'per', True, run, 2, None, device=0)
    pts = vref_points(-1, -1, 2, 2, 21)
    for run, ids in ((0, [0, 2]), (1, [1])):
        f = {l: myhd.loadhd5(sfile.format(run), l) for l in LABELS}
        assert np.array_equal(f['id'], ids)
        n = len(ids)
        assert f['solutions_S'].shape == (n, 162) and f['sigma_A'].shape == (n, 3) and f['a_S'].shape == (n, 2)
        assert f['B_A'].shape == (n, 2, 2) and f['sigmaTotal_S'].shape == (n, 3)
    # oracle on the cached mesh of snapshot id 1
    mesh = load_mesh(mname.format(1).replace(".xdmf", "_1.xdmf"))
    so = OracleRVE(*mesh, param=tuple(PARAM), model='per').snapshot(pts)
    f = {l: myhd.loadhd5(sfile.format(1), l) for l in LABELS}
    for l in LABELS[1:]:
        ref = np.asarray(so[l]).ravel()
        assert np.abs(f[l][0].ravel() - ref).max() < 1e-8 * max(np.abs(ref).max(), 1e-3), l
    # merge like the reference (wrapper_h5py.py:105-122)
    merged = str(tmp_path / "snapshots.hd5")
    myhd.merge([sfile.format(0), sfile.format(1)], merged, LABELS, LABELS)
    assert myhd.loadhd5(merged, 'solutions_A').shape == (3, 162)


@pytest.mark.parametrize("modelBnd", ["dnn", "lin", "per", "MR"])
def test_predictTangents_matches_oracle(tmp_path, modelBnd, monkeypatch):
    monkeypatch.setenv("DEEPBND_KEEP_MESHES", "1")
    import deepbnd_b200.wrapper_h5py as myhd
    from deepbnd_b200.tangents_prediction import predictTangents
    from deepbnd_b200.mesh.mesh_io import load_mesh
    from oracle.rve_oracle import OracleRVE, vref_points
    ns = 3
    par = synthetic_param(ns, seed=5)
    pfile, tfile, bfile = (str(tmp_path / n) for n in ("paramRVEdataset.hd5", "tangents.hd5", "bcs.hd5"))
    myhd.savehd5(pfile, [par, np.arange(ns, dtype=float), np.arange(2.0 * ns).reshape(ns, 2)], ['param', 'id', 'center'], 'w')
    uD = np.stack([np.stack([smooth_uD(81, 10 * s + i) for s in range(ns)]) for i in range(3)])   # (3, ns, 162)
    from deepbnd_b200.vref import COMP_LABEL, COORD_LABEL, vref_order_datasets
    coords, comps = vref_order_datasets()
    myhd.savehd5(bfile, [uD[0], uD[1], uD[2], coords, comps], ['u0', 'u1', 'u2', COORD_LABEL, COMP_LABEL], 'w')
    mname = str(tmp_path / "meshes" / "mesh_micro_{0}_{1}.xdmf")
    predictTangents(0, 1, modelBnd, ["", pfile, tfile, bfile, mname], True, 'reduced', device=0)
    T, ids, cen = myhd.loadhd5(tfile, 'tangent'), myhd.loadhd5(tfile, 'id'), myhd.loadhd5(tfile, 'center')
    assert T.shape == (ns, 3, 3) and np.array_equal(ids, np.arange(ns)) and np.array_equal(cen, np.arange(2.0 * ns).reshape(ns, 2))
    pts = vref_points(-1, -1, 2, 2, 21)
    for s in (0, 2):
        mesh = load_mesh(mname.format(s, 'reduced'))
        o = OracleRVE(*mesh, param=(0.3, 10.0, 0.3, 1.0), model=modelBnd)
        Co = o.compute_tangent(uD_list=[uD[i, s] for i in range(3)] if modelBnd == 'dnn' else None, pts=pts)
        assert np.abs(T[s] - Co).max() < 1e-8 * np.abs(Co).max(), (modelBnd, s)


def test_micro_model_mirrors(reduced_meshes):
    from deepbnd_b200.micro_model import MicroModel, MicroConstitutiveModelDNN
    from oracle.rve_oracle import OracleRVE
    mesh = reduced_meshes[0]
    m = MicroModel(mesh, [0.3, 10.0, 0.3, 1.0], 'per')
    m.compute()
    o = OracleRVE(*mesh, model='per')
    o.compute()
    for i in range(3):
        assert np.abs(m.homogenise([0, 1], i) - o.homogenise([0, 1], i)).max() < 1e-8 * np.abs(o.homogenise([0, 1], i)).max()
        u = o.sol[i].reshape(-1, 2)
        assert np.linalg.norm(m.sol[i][0] - u) < 1e-8 * np.linalg.norm(u)
    d = MicroConstitutiveModelDNN(mesh, [0.3, 10.0, 0.3, 1.0], 'lin')
    d.others['uD0_'] = d.others['uD1_'] = d.others['uD2_'] = np.zeros(162)
    C = d.getTangent()
    Co = OracleRVE(*mesh, model='lin').compute_tangent()
    assert np.abs(C - Co).max() < 1e-8 * np.abs(Co).max() and d.getTangent() is C


@pytest.mark.parametrize("model", ["per", "MR", "dnn"])
def test_micro_model_gen_matches_oracle(full_mesh, model):
    """MicroConstitutiveModelGen.computeHomogenisation on a 'full' mesh (regions 0..3)."""
    from deepbnd_b200.micro_model import MicroConstitutiveModelGen
    from oracle.rve_oracle import OracleRVE, vref_points
    mesh = full_mesh
    if model == "dnn":                                   # dnn prescribes the boundary of the mesh itself: use the 6x6 box as Vref
        pts = vref_points(-3, -3, 6, 6, 21)
        uD = [smooth_uD(81, 40 + i) for i in range(3)]
    mm = MicroConstitutiveModelGen(mesh, [0.3, 10.0, 0.3, 1.0], model)
    if model == "dnn":
        for i in range(3):
            mm.others['uD%d_' % i] = uD[i]
    if model == "dnn":
        # the library takes the Vref box from the batch: the mirror uses the mesh bounding box by default
        Hg = mm.computeHomogenisation()
        Ho = OracleRVE(*mesh, model="dnn").compute_homogenisation(uD_list=uD, pts=pts)
    else:
        Hg = mm.computeHomogenisation()
        Ho = OracleRVE(*mesh, model=model).compute_homogenisation()
    for k in ("sigma", "sigmaL", "eps", "epsL", "tangent", "tangentL"):
        assert np.abs(Hg[k] - Ho[k]).max() < 1e-8 * np.abs(Ho[k]).max(), (model, k)
    assert mm.getHomogenisation() is Hg


def test_predictTangents_with_morphed_meshes(tmp_path):
    """mesher='morph' (one template per rank, fixed topology): tangents equal the oracle's on the same morphed
    meshes to 1e-8, and differ from the re-meshed run only at the level of the discretisation error."""
    import deepbnd_b200.wrapper_h5py as myhd
    from deepbnd_b200.tangents_prediction import predictTangents
    from deepbnd_b200.mesh.rve_morph import MorphTemplate, buildRVEmesh_morphed
    from oracle.rve_oracle import OracleRVE
    ns = 5
    par = synthetic_param(ns, seed=7)
    pfile, bfile = str(tmp_path / "paramRVEdataset.hd5"), str(tmp_path / "bcs.hd5")
    myhd.savehd5(pfile, par, 'param', 'w')
    mname = str(tmp_path / "meshes" / "mesh_micro_{0}_{1}.xdmf")
    out = {}
    for mesher in ("morph", "delaunay"):
        tfile = str(tmp_path / ("tangents_%s.hd5" % mesher))
        predictTangents(0, 1, 'per', ["", pfile, tfile, bfile, mname], True, 'reduced', device=0, mesher=mesher)
        out[mesher] = myhd.loadhd5(tfile, 'tangent')
    tpl = MorphTemplate(par[0], size='reduced')
    for s in (0, 3):
        Co = OracleRVE(*buildRVEmesh_morphed(par[s].copy(), tpl), param=tuple(PARAM), model='per').compute_tangent()
        assert np.abs(out["morph"][s] - Co).max() < 1e-8 * np.abs(Co).max()
    assert np.abs(out["morph"] - out["delaunay"]).max() < 5e-3 * np.abs(out["delaunay"]).max()
