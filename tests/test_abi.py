"""CPU tests of the C-ABI boundary: the shared library loads, exports every symbol include/rve_b200.h
declares, fails loudly without a GPU (no CPU fallback), and its host-side symbolic phase (P2 numbering,
constraint map, SELL/halo/restriction layout) agrees with the oracle's discrete problem."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from deepbnd_b200 import _lib as L
    from deepbnd_b200.build import build_lib
    build_lib()
    return L, L.load()


def test_header_symbols_exported():
    L, lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "rve_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(rve_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(L.SYMBOLS) == declared


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L, lib = _lib()
    h = ctypes.c_void_p()
    rc = lib.rve_create(0, ctypes.byref(h))
    assert rc == -2 and not h                       # RVE_E_CUDA
    assert b"CUDA" in lib.rve_last_error(None)
    from deepbnd_b200.batch import RVEBatch
    with pytest.raises(L.RVEError):
        RVEBatch(0)


@pytest.mark.parametrize("model", ["lin", "per", "MR"])
def test_host_symbolic_matches_oracle_pattern(coarse_reduced_meshes, model):
    """rows = the oracle's active nodes, block pattern = pattern of the oracle's constrained matrix,
    and the device layout passes the library's own SELL/halo/restriction consistency check."""
    from deepbnd_b200.batch import host_symbolic
    from oracle.rve_oracle import OracleRVE, periodic_map
    mesh = coarse_reduced_meshes[1]
    r = host_symbolic(*mesh, model=model, vref_nb=21)
    o = OracleRVE(*mesh, model=model)
    o._factor()
    m = o.mesh
    if model == "lin":
        active = o.free[0::2] // 2
        A = o.K[o.free][:, o.free]
    elif model == "per":
        active = np.unique(periodic_map(m))
        A = o.P.T @ o.K @ o.P
    else:
        active = np.arange(m.nn)
        A = o.K
    assert r["n_rows"] == len(active) and r["n_nodes"] == m.nn
    ed = {(int(a), int(b)): m.nv + i for i, (a, b) in enumerate(m.edges)}
    rown = np.array([a if a == b else ed[(int(a), int(b))] for a, b in zip(r["row_va"], r["row_vb"])])
    assert sorted(rown.tolist()) == sorted(active.tolist())
    pos = {int(n): i for i, n in enumerate(active)}
    perm = np.array([pos[int(n)] for n in rown])
    # node-block pattern of the oracle matrix in library row order
    Ab = (abs(A[0::2][:, 0::2]) + abs(A[0::2][:, 1::2]) + abs(A[1::2][:, 0::2]) + abs(A[1::2][:, 1::2])).tocsr()
    Ab = Ab[perm][:, perm].tocsr()
    Ab.sort_indices()
    lens = np.diff(r["rowptr"])
    for row in range(r["n_rows"]):
        lib_cols = r["bcol"][r["rowptr"][row]:r["rowptr"][row + 1]]
        ora_cols = Ab.indices[Ab.indptr[row]:Ab.indptr[row + 1]]
        ora_vals = Ab.data[Ab.indptr[row]:Ab.indptr[row + 1]]
        # the library pattern is the structural one (a superset of the numerically non-zero blocks)
        assert set(ora_cols[ora_vals > 0].tolist()) <= set(lib_cols.tolist())
        assert len(lib_cols) == len(set(lib_cols.tolist())) == lens[row]
    # every node maps to a row consistently
    assert (r["node_row"] < r["n_rows"]).all()
    assert ((r["node_row"] >= 0).sum() >= r["n_rows"])


def test_vref_sample_location(coarse_reduced_meshes):
    from deepbnd_b200.batch import host_symbolic
    from oracle.rve_oracle import vref_points
    mesh = coarse_reduced_meshes[0]
    r = host_symbolic(*mesh, model="per", vref_nb=21)
    xy, tri = mesh[0], mesh[1]
    pts = vref_points(xy[:, 0].min(), xy[:, 1].min(), np.ptp(xy[:, 0]), np.ptp(xy[:, 1]), 21)
    tri_ccw = tri.copy()
    a = xy[tri[:, 1]] - xy[tri[:, 0]]
    b = xy[tri[:, 2]] - xy[tri[:, 0]]
    neg = (a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) < 0
    tri_ccw[neg, 1], tri_ccw[neg, 2] = tri[neg, 2], tri[neg, 1]
    for k in range(81):
        t = r["samp_elem"][k]
        l1, l2 = r["samp_lam"][k]
        p = (1 - l1 - l2) * xy[tri_ccw[t, 0]] + l1 * xy[tri_ccw[t, 1]] + l2 * xy[tri_ccw[t, 2]]
        assert np.abs(p - pts[k]).max() < 1e-12
        assert min(l1, l2, 1 - l1 - l2) > -1e-9
