#!/usr/bin/env python
"""Experiment: device-resident throughput of ONE handle holding ns RVEs against TWO handles holding ns/2 each, solved
concurrently by two host threads (their kernels interleave on the GPU).  Diagnostics only."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, PARAM_MAT, make_meshes, flatten   # noqa: E402
from deepbnd_b200.synthetic import synthetic_param, smooth_uD   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2_snapshots_1k_full_per")
    ap.add_argument("--ns", type=int, default=1000)
    ap.add_argument("--unique", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--parts", type=int, nargs="+", default=[1, 2])
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    nu = min(args.ns, args.unique)
    meshes = make_meshes(synthetic_param(nu, seed=2), wl["size"], wl.get("lcar"), os.cpu_count() or 1)
    meshes = [meshes[i % nu] for i in range(args.ns)]
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.elasticity import macro_strain_voigt
    from concurrent.futures import ThreadPoolExecutor
    import torch
    nl = len(wl["loads"])
    eps1 = np.stack([macro_strain_voigt(i) for i in wl["loads"]])
    for parts in args.parts:
        hs, sl = [], []
        for k in range(parts):
            a, z = args.ns * k // parts, args.ns * (k + 1) // parts
            b = RVEBatch(0)
            b.set_batch(*flatten(meshes[a:z]), param=PARAM_MAT, model=wl["model"], vref_nb=21, vref_box=(-1.0, -1.0, 2.0, 2.0))
            hs.append(b); sl.append((a, z))
        def run(k):
            a, z = sl[k]
            eps = np.broadcast_to(eps1, (z - a, nl, 3)).copy()
            uD = None
            if wl["model"] == "dnn":
                uD = np.broadcast_to(np.stack([smooth_uD(81, 17 + i) for i in range(3)]), (z - a, 3, 162)).copy()
            hs[k].assemble()
            st, it = hs[k].solve(eps, uD=uD, rtol=1e-10)
            return hs[k].tangent() if nl == 3 else hs[k].snapshot()
        pool = ThreadPoolExecutor(max_workers=parts)
        list(pool.map(run, range(parts)))
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(args.steps):
            list(pool.map(run, range(parts)))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / args.steps
        print("parts %d: %.1f ms per step, %.0f RVE/s" % (parts, 1e3 * dt, args.ns / dt), flush=True)
        del hs


if __name__ == "__main__":
    main()
