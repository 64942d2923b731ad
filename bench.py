#!/usr/bin/env python
"""bench.py — RVE solves/sec of the batched B200 path (and of the CPU reference arm).

Contract (task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A "step" is one pass of the hot path (assemble -> multi-RHS PCG -> constraint post-processing ->
fused epilogues [-> RB projection]) over one batch of synthetic RVEs.  Default workload = BASELINE.json
configs[1] ("C2"): 1k extended 6x6 ('full') RVEs, periodic model, loads S and A, internal-boundary
DOF extraction + projection onto a (synthetic, orthonormal) Wbasis.

  value : whole-job RVE solves/s with the batch already resident in HBM (device-resident loop)
  e2e   : same metric through the C-ABI with HOST mesh buffers: symbolic phase, H2D of the batch,
          device work, D2H of every output inside the timed region
  roofline : the SpMV kernel (k_spmv_dot) — algorithmic CSR bytes B_spmv(k) / CUDA-event time of the
          launches inside the timed PCG loops, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the CPU oracle (scipy sparse LU restatement of the reference path) timed on the
          host cores, rank-strided worker processes like the reference's MPI launch

`--impl reference` times only that CPU arm.  Multi-GPU (torchrun): RVEs shard by rank, no data-path
collective, one final gather of the per-RVE outputs ("weak" scaling: per-GPU batch fixed).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
_OUT = sys.stdout
sys.path.insert(0, ROOT)

PARAM_MAT = (0.3, 10.0, 0.3, 1.0)           # simulation_snapshots.py:120-123
WORKLOADS = {
    # name: (ns per GPU, mesh size, model, loads (Voigt ids), project?)
    "C2_snapshots_1k_full_per": dict(ns=1000, size="full", model="per", loads=(2, 0), project=True),
    "C2_snapshots_1k_full_MR": dict(ns=1000, size="full", model="MR", loads=(2, 0), project=True),
    "C3_tangents_10k_reduced_dnn": dict(ns=10000, size="reduced", model="dnn", loads=(0, 1, 2), project=False),
    "C4_tangents_10k_reduced_per": dict(ns=10000, size="reduced", model="per", loads=(0, 1, 2), project=False),
    "C4_tangents_10k_reduced_lin": dict(ns=10000, size="reduced", model="lin", loads=(0, 1, 2), project=False),
    "C4_tangents_10k_reduced_MR": dict(ns=10000, size="reduced", model="MR", loads=(0, 1, 2), project=False),
}


def _mesh_worker(args):
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    row, size, lcar = args
    return buildRVEmesh_arrays(row.copy(), size=size, lcar=lcar)


def make_meshes(param, size, lcar, nproc):
    jobs = [(param[i], size, lcar) for i in range(len(param))]
    if nproc <= 1 or len(jobs) < 4:
        return [_mesh_worker(j) for j in jobs]
    with mp.get_context("fork").Pool(nproc) as pool:
        return pool.map(_mesh_worker, jobs, chunksize=max(1, len(jobs) // (8 * nproc)))


def flatten(meshes):
    n = len(meshes)
    voff = np.zeros(n + 1, dtype=np.int64)
    toff = np.zeros(n + 1, dtype=np.int64)
    for s, (p, t, r) in enumerate(meshes):
        voff[s + 1] = voff[s] + len(p)
        toff[s + 1] = toff[s] + len(t)
    xy = np.ascontiguousarray(np.concatenate([m[0] for m in meshes]), dtype=np.float64)
    tri = np.ascontiguousarray(np.concatenate([m[1] for m in meshes]), dtype=np.int32)
    reg = np.ascontiguousarray(np.concatenate([m[2] for m in meshes]), dtype=np.int32)
    return voff, toff, xy, tri, reg


def _oracle_worker(args):
    """one reference-style serial worker: K once, sparse LU once, every load, epilogue."""
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle.rve_oracle import OracleRVE, vref_points
    meshes, model, loads, uDs = args
    pts = vref_points(-1, -1, 2, 2, 21)
    t0 = time.perf_counter()
    for k, m in enumerate(meshes):
        o = OracleRVE(*m, param=PARAM_MAT, model=model)
        if len(loads) == 3:
            o.compute_tangent(uD_list=None if uDs is None else uDs[k], pts=pts)
        else:
            o.snapshot(pts)
    return time.perf_counter() - t0


def cpu_reference_rate(meshes, model, loads, uDs, nproc):
    """RVE/s of P independent serial workers with rank-strided ids (simulation_snapshots.py:76)."""
    jobs = [(meshes[r::nproc], model, loads, None if uDs is None else uDs[r::nproc]) for r in range(nproc)]
    jobs = [j for j in jobs if len(j[0])]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        pool.map(_oracle_worker, jobs)
    dt = time.perf_counter() - t0
    return len(meshes) / dt, dt


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = "/tmp/bench_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            # median of the samples under load (top half)
            srt = sorted(sm)
            out["sm_mhz"] = srt[len(srt) // 2]
            out["sm_max_mhz"] = max(smax)
        out["reasons"] = sorted(reasons)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def main():
    # the contract is ONE JSON line on stdout: anything a library prints there (NCCL's version banner, ...)
    # is sent to stderr instead; the line itself goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2_snapshots_1k_full_per", choices=sorted(WORKLOADS))
    ap.add_argument("--ns", type=int, default=0, help="override RVEs per GPU")
    ap.add_argument("--lcar", type=float, default=0.0, help="override mesh size (default 2/30)")
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--precond", default="twolevel", choices=["twolevel", "jacobi"])
    ap.add_argument("--unique", type=int, default=256, help="distinct synthetic RVEs meshed per rank (cycled to --ns); 0 = all")
    ap.add_argument("--chunks", type=int, default=4, help="chunks per step in the pipelined e2e loop")
    ap.add_argument("--operator", default="matfree", choices=["matfree", "assembled"],
                    help="q = A p inside the PCG: element-wise (default) or the assembled SELL-32 SpMV")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="RVEs in the CPU baseline sample (default: one per core)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.ns:
        wl["ns"] = args.ns
    lcar = args.lcar if args.lcar > 0 else None
    ncpu = os.cpu_count() or 1
    nproc_local = max(1, ncpu // max(1, world))
    if world > 1:                                   # ranks share the host cores of the box
        os.environ.setdefault("RVE_HOST_THREADS", str(max(1, ncpu // world)))
        if ncpu // world < 8:
            os.environ.setdefault("RVE_BLOCKING_SYNC", "1")   # few cores per rank: the kernel-feeding thread sleeps while it waits
    from deepbnd_b200.synthetic import synthetic_param, smooth_uD

    if args.impl == "reference":
        # CPU reference arm: rank 0 alone, bounded sample per step (one RVE per host core)
        if rank != 0:
            return
        nsamp = args.cpu_sample or ncpu
        param = synthetic_param(nsamp, seed=2)
        meshes = make_meshes(param, wl["size"], lcar, ncpu)
        uDs = None
        if wl["model"] == "dnn":
            uDs = [[smooth_uD(81, 1000 * s + i) for i in range(3)] for s in range(nsamp)]
        times = []
        for st in range(args.warmup + args.steps):
            rate, dt = cpu_reference_rate(meshes, wl["model"], wl["loads"], uDs, ncpu)
            if st >= args.warmup:
                times.append(dt)
            if st >= 1 and sum(times) > 240:      # keep the whole run within minutes
                break
        tot = sum(times) if times else dt
        nst = max(len(times), 1)
        val = nsamp * nst / tot
        line = {"impl": "reference", "metric": "RVE solves/sec", "value": val, "unit": "RVE/s", "n_gpus": args.gpus,
                "steps": nst, "warmup": args.warmup, "ms_per_step": 1e3 * tot / nst, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "mesh": wl["size"], "model": wl["model"], "loads": list(wl["loads"]),
                           "lcar": lcar or 2 / 30},
                "cpu_baseline": {"value": val, "unit": "RVE/s", "cores": ncpu, "kind": "port",
                                 "sample": "%d RVEs per step, %d serial worker processes, scipy splu" % (nsamp, ncpu)},
                "e2e": {"value": val, "unit": "RVE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _OUT.write(json.dumps(line) + "\n")
        _OUT.flush()
        return

    # ---------------- B200 arm ----------------
    ns = wl["ns"]
    # meshing is outside the path ("generated once and cached"): at most --unique distinct RVEs are meshed per
    # rank and cycled to the batch size; every copy is still set up and solved as an independent system
    nuniq = min(ns, args.unique) if args.unique > 0 else ns
    param = synthetic_param(nuniq, seed=2 + 1000 * rank)
    t0 = time.perf_counter()
    meshes = make_meshes(param, wl["size"], lcar, nproc_local)   # fork BEFORE any CUDA initialisation
    meshes = [meshes[i % nuniq] for i in range(ns)]
    t_mesh = time.perf_counter() - t0
    voff, toff, xy, tri, reg = flatten(meshes)

    nl = len(wl["loads"])
    uD = None
    if wl["model"] == "dnn":
        rng = np.random.RandomState(5 + rank)
        base = np.stack([smooth_uD(81, 17 + i) for i in range(3)])
        uD = base[None] * (1.0 + 0.1 * rng.randn(ns, 1, 1))
    # CPU baseline first (worker processes are forked before CUDA is initialised)
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        nsamp = min(ns, args.cpu_sample or ncpu)
        uDs = None if uD is None else [list(uD[s]) for s in range(nsamp)]
        rate, dt = cpu_reference_rate(meshes[:nsamp], wl["model"], wl["loads"], uDs, ncpu)
        cpu = {"value": rate, "unit": "RVE/s", "cores": ncpu, "kind": "port",
               "sample": "first %d RVEs of the same batch, %d serial worker processes (scipy splu), %.1f s" % (nsamp, ncpu, dt)}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from deepbnd_b200.batch import RVEBatch
    from deepbnd_b200.elasticity import macro_strain_voigt

    eps = np.broadcast_to(np.stack([macro_strain_voigt(i) for i in wl["loads"]]), (ns, nl, 3)).copy()
    vref_box = (-1.0, -1.0, 2.0, 2.0)
    W = M = None
    if wl["project"]:
        rng = np.random.RandomState(0)
        Wq = np.linalg.qr(rng.randn(162, 160))[0].T                   # synthetic orthonormal basis (160,162)
        W = np.stack([Wq] * nl)
        n = 80
        M = np.zeros((162, 162))
        for k in range(n):                                             # boundary mass of Vref (h = 0.1)
            k2 = (k + 1) % n
            for c in range(2):
                M[2 * k + c, 2 * k + c] += 0.1 / 3; M[2 * k2 + c, 2 * k2 + c] += 0.1 / 3
                M[2 * k + c, 2 * k2 + c] += 0.1 / 6; M[2 * k2 + c, 2 * k + c] += 0.1 / 6

    # two handles: while one batch is being solved, a second host thread uploads the next one and runs its symbolic phase
    # one on the other handle from a second host thread (ctypes releases the GIL)
    from concurrent.futures import ThreadPoolExecutor
    hb = [RVEBatch(local_rank).set_operator(args.operator), RVEBatch(local_rank).set_operator(args.operator)]
    b = hb[0]
    pool = ThreadPoolExecutor(max_workers=2)
    maxit = 4000

    def device_pass(bb=None):
        bb = bb or b
        bb.assemble()
        status, iters = bb.solve(eps, uD=uD, rtol=args.rtol, maxit=maxit, precond=args.precond)
        if nl == 3:
            out = bb.tangent()
        else:
            out = bb.snapshot()
            if wl["project"]:
                out["Y"] = bb.project(W, M, translate=out["a"])
        return out, iters

    def set_pass(bb, c=None):
        if c is None:
            bb.set_batch(voff, toff, xy, tri, reg, param=PARAM_MAT, model=wl["model"], vref_nb=21, vref_box=vref_box)
        else:
            bb.set_batch(*chunks[c][:5], param=PARAM_MAT, model=wl["model"], vref_nb=21, vref_box=vref_box)

    # e2e streams every step's batch through the two handles in `nchunk` chunks, so that the pipeline
    # fill/drain costs one chunk, not one batch
    nchunk = max(1, min(args.chunks, ns // 64))
    cb = [ns * c // nchunk for c in range(nchunk + 1)]
    chunks = []
    for c in range(nchunk):
        a, z = cb[c], cb[c + 1]
        chunks.append((voff[a:z + 1] - voff[a], toff[a:z + 1] - toff[a], xy[voff[a]:voff[z]], tri[toff[a]:toff[z]],
                       reg[toff[a]:toff[z]], slice(a, z)))

    def chunk_pass(bb, c):
        sl_ = chunks[c][5]
        bb.assemble()
        status, iters = bb.solve(eps[sl_], uD=None if uD is None else uD[sl_], rtol=args.rtol, maxit=maxit, precond=args.precond)
        if nl == 3:
            out = bb.tangent()
        else:
            out = bb.snapshot()
            if wl["project"]:
                out["Y"] = bb.project(W, M, translate=out["a"])
        return out, iters

    def e2e_steps(nsteps):
        """nsteps full passes from HOST mesh arrays.  Two host workers, one per handle, each run set_batch (H2D of the
        raw meshes + GPU symbolic phase) -> assemble -> solve -> epilogues -> D2H for every other chunk; the kernels of
        the two handles interleave on the GPU (deepbnd_b200.batch.DoubleBuffer does the same for the drop-in scripts)."""
        jobs = [c for _ in range(nsteps) for c in range(nchunk)]
        outs = [None] * len(jobs)
        dbg = os.environ.get("BENCH_E2E_DEBUG") == "1"
        acc = [{}, {}]

        def worker(k):
            for i in range(k, len(jobs), 2):
                tc0 = time.perf_counter()
                set_pass(hb[k], jobs[i])
                tc1 = time.perf_counter()
                outs[i] = chunk_pass(hb[k], jobs[i])
                if dbg:                                        # where the wall time of a chunk goes (stderr only)
                    for k_, v_ in hb[k].timings()[0].items():
                        acc[k][k_] = acc[k].get(k_, 0.0) + v_
                    acc[k]["wall_set_batch"] = acc[k].get("wall_set_batch", 0.0) + 1e3 * (tc1 - tc0)
                    acc[k]["wall_chunk_pass"] = acc[k].get("wall_chunk_pass", 0.0) + 1e3 * (time.perf_counter() - tc1)

        futs = [pool.submit(worker, k) for k in range(min(2, len(jobs)))]
        for f_ in futs:
            f_.result()
        if dbg:
            tot = {k_: acc[0].get(k_, 0.0) + acc[1].get(k_, 0.0) for k_ in set(acc[0]) | set(acc[1])}
            print("[e2e debug] per step (ms, both workers): " + json.dumps({k_: round(v_ / nsteps, 2) for k_, v_ in sorted(tot.items())}), file=sys.stderr)
        last = outs[-nchunk:]
        iters = np.concatenate([o[1] for o in last])
        if nl == 3:
            return np.concatenate([o[0] for o in last]), iters
        return {k_: np.concatenate([o[0][k_] for o in last]) for k_ in last[0][0]}, iters

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, *a):
        barrier()
        t = time.perf_counter()
        r = fn(*a)
        barrier()
        return time.perf_counter() - t, r

    def counters():
        t0_, c0_ = hb[0].timings()
        t1_, c1_ = hb[1].timings()
        return {k_: c0_[k_] + c1_[k_] for k_ in c0_}

    # e2e loop (host buffers in, outputs out)
    timed(e2e_steps, max(3, args.warmup))
    cnt0 = counters()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_e2e, (out, iters) = timed(e2e_steps, args.steps)
    cnt1 = counters()
    # device-resident loop (batch already in HBM): `value`
    set_pass(b)
    timed(device_pass)
    spmv_ms, spmv_bytes, launches0 = 0.0, 0.0, b.timings()[1]["kernel_launches"]
    barrier()
    t = time.perf_counter()
    phase = {}
    for _ in range(args.steps):
        device_pass()
        tt, cc = b.timings()
        spmv_ms += tt["spmv"]; spmv_bytes += cc["spmv_alg_bytes"]
        for k_, v_ in tt.items():
            phase[k_] = phase.get(k_, 0.0) + v_ / args.steps
        spmv_launches = cc["spmv_launches"]
    barrier()
    t_dev = time.perf_counter() - t
    clocks = sampler.stop()
    launches = b.timings()[1]["kernel_launches"] - launches0

    tmax = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        # the only exchange of the path: final gather of the per-RVE outputs on rank 0
        flat = torch.from_numpy(np.ascontiguousarray(
            out.reshape(ns, -1) if nl == 3 else np.concatenate([out[k].reshape(ns, -1) for k in ("sol", "sigma", "a", "B", "sigmaT")], 1))).cuda()
        gathered = [torch.empty_like(flat) for _ in range(world)] if rank == 0 else None
        dist.gather(flat, gathered, dst=0)
    t_dev_max, t_e2e_max = tmax.tolist()
    total = ns * world
    value = total * args.steps / t_dev_max
    e2e_val = total * args.steps / t_e2e_max
    h2d = (cnt1["h2d_bytes"] - cnt0["h2d_bytes"]) / args.steps
    d2h = (cnt1["d2h_bytes"] - cnt0["d2h_bytes"]) / args.steps

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else 0.0
    # DRAM traffic of one SpMV launch from the committed ncu --set full capture (per RVE, scaled to this batch)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(args.workload)
        if tj and not args.lcar:
            traffic = tj["dram_bytes_per_rve_per_launch"] * ns

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    dof = int(np.mean([2 * b.sizes(s)["n_rows"] for s in range(min(ns, 8))]))
    line = {"metric": "RVE solves/sec", "value": value, "unit": "RVE/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "rves_per_gpu": ns, "mesh": wl["size"], "lcar": lcar or 2 / 30,
                       "model": wl["model"], "loads": list(wl["loads"]), "dofs_per_rve": dof, "rtol": args.rtol,
                       "precond": args.precond, "mean_pcg_iters": float(np.mean(iters)),
                       "l2": "per-iteration working set >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "unique_meshes": nuniq, "mesh_gen_s": t_mesh, "phase_ms": phase},
            "e2e": {"value": e2e_val, "unit": "RVE/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pipeline": "2 handles / 2 host workers, %d chunks per step: each worker runs H2D of the raw meshes + GPU symbolic phase + solve + D2H for every other chunk; the two streams interleave on the GPU" % nchunk},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_spmv", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "launches_timed": int(spmv_launches * args.steps),
                         "alg_bytes_per_launch": spmv_bytes / max(1.0, spmv_launches * args.steps)},
            "cpu_baseline": cpu}
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
