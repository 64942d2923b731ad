#!/usr/bin/env python
"""Oracle spot check of the C5 campaign (tools/run_c5_campaign.py): rebuilds the morphed refined mesh of the stored rows
(template of the rank that solved them) and compares the snapshot outputs with the sparse-LU oracle.  CPU, minutes/row."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from deepbnd_b200.mesh.rve_morph import MorphTemplate, buildRVEmesh_morphed   # noqa: E402
from deepbnd_b200.synthetic import synthetic_param                              # noqa: E402
from oracle.rve_oracle import OracleRVE, vref_points                           # noqa: E402

spot = np.load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "c5_spot.npz"))
world, ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8, int(sys.argv[3]) if len(sys.argv) > 3 else 16000
nrows = int(sys.argv[4]) if len(sys.argv) > 4 else len(spot["ids"])
param_all = synthetic_param(ns, seed=2)
pts = vref_points(-1, -1, 2, 2, 21)
out = []
for k, sid in enumerate(spot["ids"][:nrows]):
    rank = int(sid) % world
    tpl = MorphTemplate(param_all[rank], size="full", lcar=float(spot["lcar"]))      # buildSnapshots: template from the rank's first row
    mesh = buildRVEmesh_morphed(spot["param"][k].copy(), tpl)
    t0 = time.perf_counter()
    ref = OracleRVE(*mesh, param=(0.3, 10.0, 0.3, 1.0), model="per").snapshot(pts)
    errs = {}
    for key in ref:
        r = np.asarray(ref[key]).ravel()
        errs[key] = float(np.abs(spot[key][k].ravel() - r).max() / max(np.abs(r).max(), 1e-3))
    out.append({"id": int(sid), "rank": rank, "oracle_s": time.perf_counter() - t0, "max_rel_err": max(errs.values()), "errs": errs})
    print(json.dumps(out[-1]))
json.dump(out, open(os.path.join(ROOT, "profiles", os.environ.get("SPOT_OUT", "r02_c5_campaign_spotcheck.json")), "w"), indent=1)
