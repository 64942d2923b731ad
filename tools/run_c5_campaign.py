#!/usr/bin/env python
"""BASELINE config 5 as a real campaign: `torchrun --nproc-per-node 8 tools/run_c5_campaign.py --per-gpu 2000` runs the
drop-in snapshot driver (deepbnd_b200.simulation_snapshots.buildSnapshots, mesher='morph', refined meshes lcar = 1/30)
on every rank exactly like the reference's `mpiexec -n R python simulation_snapshots.py`: paramRVEdataset.hd5 in,
one snapshots_<run>.hd5 per rank out, then merged (wrapper_h5py.merge).  Prints one JSON line with the wall-clock
rate INCLUDING meshing and file I/O, and stores a few rows (parameters + outputs) for an oracle spot check."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=2000)
    ap.add_argument("--lcar", type=float, default=1.0 / 30.0)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--dir", default="/tmp/c5_campaign")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c5_campaign.json"))
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("gloo")
    import deepbnd_b200.wrapper_h5py as myhd
    from deepbnd_b200.simulation_snapshots import buildSnapshots, LABELS
    from deepbnd_b200.synthetic import synthetic_param
    from deepbnd_b200.mesh.degenerated_rectangle_mesh import degeneratedBoundaryRectangleMesh
    ns = args.per_gpu * world
    pfile = os.path.join(args.dir, "paramRVEdataset.hd5")
    bfile = os.path.join(args.dir, "boundaryMesh.xdmf")
    if rank == 0:
        os.makedirs(args.dir, exist_ok=True)
        myhd.savehd5(pfile, synthetic_param(ns, seed=2), 'param', 'w')
        degeneratedBoundaryRectangleMesh(-1.0, -1.0, 2.0, 2.0, 21).write(bfile, 'fenics')
    if world > 1:
        dist.barrier()
    sfile = os.path.join(args.dir, "snapshots_%d.hd5" % rank)
    t0 = time.perf_counter()
    buildSnapshots([0.3, 10.0, 0.3, 1.0], [bfile, pfile, sfile, os.path.join(args.dir, "meshes", "mesh_temp_%d.xdmf" % rank)],
                   'per', True, rank, world, None, batch_size=args.batch, mesher="morph", lcar=args.lcar * 1.0)
    t_rank = time.perf_counter() - t0
    from deepbnd_b200 import simulation_snapshots as _ss
    phases = dict(_ss.LAST_TIMINGS)
    tt = torch.tensor([t_rank] + [phases.get(k, 0.0) for k in ("files_and_template_s", "cuda_context_and_handles_s", "pipeline_s", "close_s")],
                      dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        t1 = time.perf_counter()
        merged = os.path.join(args.dir, "snapshots.hd5")
        if os.path.exists(merged):
            os.remove(merged)
        myhd.merge([os.path.join(args.dir, "snapshots_%d.hd5" % r) for r in range(world)], merged, LABELS, LABELS, axis=0, mode='w-')
        t_merge = time.perf_counter() - t1
        ids = myhd.loadhd5(merged, 'id').astype(int)
        assert sorted(ids.tolist()) == list(range(ns)), "merged ids are not a permutation of the dataset"
        sol = myhd.loadhd5(merged, 'solutions_S')
        assert np.isfinite(sol).all() and np.abs(sol).max() > 0
        # rows for the oracle spot check (run where there is time for a 290k-DOF sparse LU)
        param = myhd.loadhd5(pfile, 'param')
        pick = [0, ns // 2 + 1]
        rows = [int(np.flatnonzero(ids == k)[0]) for k in pick]
        spot = {"ids": np.array(pick), "param": param[pick], "lcar": args.lcar}
        for lab in LABELS[1:]:
            spot[lab] = myhd.loadhd5(merged, lab)[rows]
        np.savez(os.path.join(os.path.dirname(args.out), "c5_spot.npz"), **spot)
        line = {"workload": "C5 campaign through buildSnapshots (files in, files out)", "n_gpus": world, "rves": ns, "per_gpu": args.per_gpu,
                "lcar": args.lcar, "mesher": "morph (shared topology)", "batch_size": args.batch,
                "wall_s_max_over_ranks": float(tt[0]), "rves_per_s_incl_meshing_and_file_io": ns / float(tt[0]),
                "merge_s": t_merge, "file_bytes_per_rank": os.path.getsize(sfile),
                "phases_max_over_ranks_s": {"files_and_template": float(tt[1]), "cuda_context_and_handles": float(tt[2]),
                                            "double_buffered_pipeline": float(tt[3]), "close": float(tt[4])}}
        print(json.dumps(line))
        open(args.out, "w").write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
