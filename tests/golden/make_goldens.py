"""Generates tests/golden/c1_oracle.npz: outputs of the CPU oracle (sparse LU) for BASELINE config 1 —
the single RVE of the reference's parameter table (tests/test_meshes.py:63-98, stored in
paramRVE_test_meshes.txt), material nu=0.3, E = 10/1 (simulation_snapshots.py:120-123).

PARITY UNPINNED (see oracle/rve_oracle.py): these are oracle outputs, not FEniCS outputs; they pin the
oracle against regressions and give the GPU tests a committed fixture.  Run:  python tests/golden/make_goldens.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays          # noqa: E402
from oracle.rve_oracle import OracleRVE, vref_points                    # noqa: E402

PARAM = (0.3, 10.0, 0.3, 1.0)


def main():
    table = np.loadtxt(os.path.join(HERE, "paramRVE_test_meshes.txt"))
    pts = vref_points(-1, -1, 2, 2, 21)
    out = {}
    p = table.copy()
    reduced = buildRVEmesh_arrays(p, size="reduced")
    for model in ("lin", "per", "MR"):
        out["tangent_reduced_" + model] = OracleRVE(*reduced, param=PARAM, model=model).compute_tangent()
    p = table.copy()
    full = buildRVEmesh_arrays(p, size="full")
    for model in ("per", "MR"):
        o = OracleRVE(*full, param=PARAM, model=model)
        for k, v in o.snapshot(pts).items():
            out["full_%s_%s" % (model, k)] = np.asarray(v)
    np.savez(os.path.join(HERE, "c1_oracle.npz"), **out)
    for k in sorted(out):
        print(k, out[k].shape)


if __name__ == "__main__":
    main()
