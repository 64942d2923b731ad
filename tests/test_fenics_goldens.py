"""Pinned parity against the reference itself — active only when tests/golden/fenics_c1_<size>.npz exist
(produced by tools/export_goldens_fenics.py on a machine that has FEniCS + multiphenics + micmacsfenics).
Solves the reference's own gmsh mesh with the oracle (CPU) and the CUDA path (GPU) and compares with the
FEniCS outputs to the 1e-8 of BASELINE.json.  Until the files exist the oracle stays "parity unpinned"."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAM = (0.3, 10.0, 0.3, 1.0)


def _load(size):
    fn = os.path.join(GOLD, "fenics_c1_%s.npz" % size)
    if not os.path.exists(fn):
        pytest.skip("no FEniCS goldens (run tools/export_goldens_fenics.py where the reference runs)")
    return np.load(fn)


@pytest.mark.parametrize("model", ["per", "MR", "lin"])
def test_oracle_against_fenics_reduced(model):
    g = _load("reduced")
    from oracle.rve_oracle import OracleRVE
    o = OracleRVE(g["points"][:, :2], g["triangles"], g["region"], param=PARAM, model=model)
    C = o.compute_tangent()
    assert np.abs(C - g["tangent_" + model]).max() < 1e-8 * np.abs(g["tangent_" + model]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("size", ["reduced", "full"])
@pytest.mark.parametrize("model", ["per", "MR", "lin"])
def test_cuda_path_against_fenics(size, model):
    g = _load(size)
    from deepbnd_b200.batch import RVEBatch
    mesh = (g["points"][:, :2], g["triangles"].astype(np.int32), g["region"].astype(np.int32))
    b = RVEBatch().set_meshes([mesh], PARAM, model=model, vref_nb=21, vref_box=(-1, -1, 2, 2)).assemble()
    b.solve(np.eye(3)[None].copy(), rtol=1e-11, maxit=8000)
    C = b.tangent()[0]
    ref = g["tangent_" + model]
    assert np.abs(C - ref).max() < 1e-8 * np.abs(ref).max()
    snap = b.snapshot()
    for j, lab in ((2, "S"), (0, "A")):
        ref = g["sigma_%s_%s" % (model, lab)]
        assert np.abs(snap["sigma"][0, j] - ref).max() < 1e-8 * np.abs(ref).max()
        # boundary samples: compare through the DOF coordinates the exporter recorded (settles H5)
        xy = g["vref_dof_coordinates"]
        mine = snap["sol"][0, j].reshape(-1, 2)
        from oracle.rve_oracle import vref_points
        pts = vref_points(-1, -1, 2, 2, 21)
        refv = g["solutions_%s_%s" % (model, lab)]
        for d in range(len(refv)):
            k = int(np.argmin(np.abs(pts - xy[d]).sum(1)))
            comp = d % 2                           # dolfin interleaves the components of a vector CG1 space per node
            assert abs(mine[k, comp] - refv[d]) < 1e-8 * max(np.abs(refv).max(), 1e-12), (d, k, comp)
    b.close()


def test_dnn_export_settles_the_two_ext_switches():
    """tangent_dnn of the exporter (smooth non-zero uD) against the four oracle variants: exactly one combination of
    dirichlet_semantics x zero_average_for_dirichlet may reproduce it to 1e-8; the test fails, naming the best variant,
    if it is not the library's choice ('p1_vertex', no zero-average block)."""
    g = _load("reduced")
    if "tangent_dnn" not in g.files:
        pytest.skip("export predates the dnn case")
    from deepbnd_b200.vref import permutation_to_library
    from oracle.rve_oracle import OracleRVE, vref_points
    perm = permutation_to_library(g["vref_dof_coordinates"], g["vref_dof_component"])
    uD = [g["dnn_uD%d_dolfin_order" % i][perm] for i in range(3)]
    pts = vref_points(-1, -1, 2, 2, 21)
    ref = g["tangent_dnn"]
    errs = {}
    for sem in ("p1_vertex", "p2_nodal"):
        for za in (False, True):
            o = OracleRVE(g["points"][:, :2], g["triangles"], g["region"], param=PARAM, model="dnn",
                          dirichlet_semantics=sem, zero_average_for_dirichlet=za)
            errs[(sem, za)] = np.abs(o.compute_tangent(uD, pts) - ref).max() / np.abs(ref).max()
    best = min(errs, key=errs.get)
    assert errs[best] < 1e-8, errs
    assert best == ("p1_vertex", False), "the reference runs %s: switch the library (k_bnd_values) accordingly; %s" % (best, errs)


def test_vref_dof_order_of_the_export_is_a_permutation():
    g = _load("reduced")
    if "vref_dof_component" not in g.files:
        pytest.skip("export predates the DOF-order sidecar")
    from deepbnd_b200.vref import permutation_to_library
    perm = permutation_to_library(g["vref_dof_coordinates"], g["vref_dof_component"])
    assert sorted(perm.tolist()) == list(range(162))
