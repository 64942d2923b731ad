#!/usr/bin/env python
"""Parity sweep beyond the unit tests: N random `reduced` RVEs (lcar = 2/30, the reference resolution) x the four boundary
models, tangents of the CUDA path (shipped settings: matrix-free operator, two-level preconditioner, rtol = 1e-10) against
the sparse-LU oracle; plus M `full` RVEs (per, snapshot outputs).  Prints / stores the largest relative deviations.
The oracle runs in worker processes forked before CUDA is initialised.  Diagnostics: tests/ hold the asserted cases."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PARAM = (0.3, 10.0, 0.3, 1.0)


def _oracle_tangent(args):
    mesh, model, uD = args
    from oracle.rve_oracle import OracleRVE, vref_points
    o = OracleRVE(*mesh, param=PARAM, model=model)
    if model == "dnn":
        return o.compute_tangent(list(uD), vref_points(-1, -1, 2, 2, 21))
    return o.compute_tangent()


def _oracle_snapshot(mesh):
    from oracle.rve_oracle import OracleRVE, vref_points
    return OracleRVE(*mesh, param=PARAM, model="per").snapshot(vref_points(-1, -1, 2, 2, 21))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--nfull", type=int, default=4)
    ap.add_argument("--seed", type=int, default=101)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_sweep.json"))
    args = ap.parse_args()
    from deepbnd_b200.synthetic import synthetic_param, smooth_uD
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    from bench import make_meshes
    ncpu = os.cpu_count() or 1
    par = synthetic_param(args.n, seed=args.seed)
    meshes = make_meshes(par, "reduced", None, ncpu)
    fulls = make_meshes(synthetic_param(args.nfull, seed=args.seed + 1), "full", None, ncpu)
    uD = np.stack([np.stack([smooth_uD(81, 1000 * s + i) for i in range(3)]) for s in range(args.n)])
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(ncpu) as pool:
        ref = {m: pool.map(_oracle_tangent, [(meshes[s], m, uD[s]) for s in range(args.n)]) for m in ("lin", "per", "MR", "dnn")}
        ref_full = pool.map(_oracle_snapshot, fulls)
    t_oracle = time.perf_counter() - t0
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.elasticity import macro_strain_voigt
    res = {"n_reduced": args.n, "n_full": args.nfull, "seed": args.seed, "rtol": 1e-10, "oracle_s": t_oracle, "models": {}}
    eps3 = np.broadcast_to(np.eye(3), (args.n, 3, 3)).copy()
    for m in ("lin", "per", "MR", "dnn"):
        b = RVEBatch().set_meshes(meshes, PARAM, model=m, vref_box=(-1.0, -1.0, 2.0, 2.0)).assemble()
        st, it = b.solve(eps3, uD=uD if m == "dnn" else None, rtol=1e-10)
        C = b.tangent()
        b.close()
        err = [float(np.abs(C[s] - ref[m][s]).max() / np.abs(ref[m][s]).max()) for s in range(args.n)]
        res["models"][m] = {"max_rel_err": max(err), "median_rel_err": float(np.median(err)), "iters_mean": float(it.mean()),
                            "iters_max": int(it.max()), "status_max": int(st.max())}
        print(m, res["models"][m], flush=True)
    eps2 = np.broadcast_to(np.stack([macro_strain_voigt(2), macro_strain_voigt(0)]), (args.nfull, 2, 3)).copy()
    b = RVEBatch().set_meshes(fulls, PARAM, model="per", vref_box=(-1.0, -1.0, 2.0, 2.0)).assemble()
    st, it = b.solve(eps2, rtol=1e-10)
    out = b.snapshot()
    b.close()
    worst = {}
    for s in range(args.nfull):
        for j, lab in enumerate(("S", "A")):
            for key, mine in (("solutions_", out["sol"]), ("sigma_", out["sigma"]), ("sigmaTotal_", out["sigmaT"]), ("a_", out["a"]), ("B_", out["B"])):
                r = np.asarray(ref_full[s][key + lab]).ravel()
                e = float(np.abs(mine[s, j].ravel() - r).max() / max(np.abs(r).max(), 1e-3))
                worst[key[:-1]] = max(worst.get(key[:-1], 0.0), e)
    res["full_per_snapshot"] = {"max_rel_err": worst, "iters_mean": float(it.mean()), "status_max": int(st.max())}
    print("full per snapshot", res["full_per_snapshot"], flush=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
