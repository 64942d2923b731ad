#!/usr/bin/env python
"""One set_batch -> assemble -> solve of a bench workload, nothing else: the short program ncu is pointed at
(`ncu --set full -k regex:k_update --launch-skip 20 --launch-count 1 python tools/profile_solve.py ...`) and the
RVE_KTIME=1 per-kernel-class timer.  Diagnostics only; bench.py is the measurement."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import WORKLOADS, PARAM_MAT, make_meshes, flatten   # noqa: E402
from deepbnd_b200.synthetic import synthetic_param, smooth_uD   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2_snapshots_1k_full_per", choices=sorted(WORKLOADS))
    ap.add_argument("--ns", type=int, default=250)
    ap.add_argument("--unique", type=int, default=32)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--lcar", type=float, default=0.0)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--operator", default="matfree", choices=["matfree", "assembled"])
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    nu = min(args.ns, args.unique)
    meshes = make_meshes(synthetic_param(nu, seed=2), wl["size"], args.lcar or None, os.cpu_count() or 1)
    meshes = [meshes[i % nu] for i in range(args.ns)]
    voff, toff, xy, tri, reg = flatten(meshes)
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.elasticity import macro_strain_voigt
    nl = len(wl["loads"])
    eps = np.broadcast_to(np.stack([macro_strain_voigt(i) for i in wl["loads"]]), (args.ns, nl, 3)).copy()
    uD = None
    if wl["model"] == "dnn":
        uD = np.broadcast_to(np.stack([smooth_uD(81, 17 + i) for i in range(3)]), (args.ns, 3, 162)).copy()
    b = RVEBatch(0)
    b.set_operator(args.operator)
    b.set_batch(voff, toff, xy, tri, reg, param=PARAM_MAT, model=wl["model"], vref_nb=21, vref_box=(-1.0, -1.0, 2.0, 2.0))
    for _ in range(args.reps):
        b.assemble()
        status, iters = b.solve(eps, uD=uD, rtol=args.rtol)
        out = b.tangent() if nl == 3 else b.snapshot()
    t, c = b.timings()
    print("iters mean %.2f max %d" % (iters.mean(), iters.max()), {k: round(v, 3) for k, v in t.items()}, file=sys.stderr)


if __name__ == "__main__":
    main()
