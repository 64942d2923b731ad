#!/usr/bin/env python
"""bench.py — RVE solves/sec of the batched B200 path (and of the CPU reference arm).

Contract (task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A "step" is one pass of the hot path (preconditioner set-up -> multi-RHS PCG -> constraint post-processing ->
fused epilogues [-> RB projection]) over one batch of synthetic RVEs.  Default workload = BASELINE.json
configs[1] ("C2"): 1k extended 6x6 ('full') RVEs, periodic model, loads S and A, internal-boundary
DOF extraction + projection onto a (synthetic, orthonormal) Wbasis.

  value : whole-job RVE solves/s with the batch already resident in HBM (device-resident loop)
  e2e   : same metric through the C-ABI with HOST mesh buffers: H2D of the raw meshes, GPU symbolic phase,
          device work, D2H of every output inside the timed region
  roofline : the operator application q = A p inside the PCG (k_apply, matrix-free; k_spmv with
          --operator assembled) — algorithmic CSR bytes B_spmv(k) of SURVEY 8d / CUDA-event time of its launches
          inside the timed solves, against MEASURED_PEAKS.json hbm_gbs.  `roofline.spmv_assembled` is the stored
          SELL-32 SpMV kernel timed alone in the same run (rve_bench_spmv).
  cpu_baseline : the CPU oracle (scipy sparse LU restatement of the reference path) timed on the
          host cores, rank-strided worker processes like the reference's MPI launch
  extra_workloads : the other BASELINE configs (C3 dnn 10k, C4 lin, C2 MR, C5 resolution) measured in the same run
          with fewer steps (N = 1 only; --no-extra skips them)

`--impl reference` times only the CPU arm.  Multi-GPU (torchrun): RVEs shard by rank, no data-path
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
    # name: RVEs per GPU, mesh size, model, loads (Voigt ids), project?, [lcar, mesher, e2e chunk size]
    "C2_snapshots_1k_full_per": dict(ns=1000, size="full", model="per", loads=(2, 0), project=True),
    "C2_snapshots_1k_full_MR": dict(ns=1000, size="full", model="MR", loads=(2, 0), project=True),
    "C3_tangents_10k_reduced_dnn": dict(ns=10000, size="reduced", model="dnn", loads=(0, 1, 2), project=False),
    "C4_tangents_10k_reduced_per": dict(ns=10000, size="reduced", model="per", loads=(0, 1, 2), project=False),
    "C4_tangents_10k_reduced_lin": dict(ns=10000, size="reduced", model="lin", loads=(0, 1, 2), project=False),
    "C4_tangents_10k_reduced_MR": dict(ns=10000, size="reduced", model="MR", loads=(0, 1, 2), project=False),
    # BASELINE config 5 per GPU: "refined" = lcar 1/30 (4x the DOFs of C2), meshes morphed from one template,
    # streamed through the two handles in device-sized chunks
    "C5_snapshots_refined_per": dict(ns=512, size="full", model="per", loads=(2, 0), project=True, lcar=1.0 / 30.0,
                                     mesher="morph", chunk=128),
}
EXTRA = {"C3_tangents_10k_reduced_dnn": 10000, "C4_tangents_10k_reduced_lin": 4000, "C2_snapshots_1k_full_MR": 1000,
         "C5_snapshots_refined_per": 256}


def _mesh_worker(args):
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    row, size, lcar = args
    return buildRVEmesh_arrays(row.copy(), size=size, lcar=lcar)


MORPH = {}                                          # template + (l, e, theta) of the last make_meshes(mesher="morph")


def make_meshes(param, size, lcar, nproc, mesher="delaunay"):
    if mesher == "morph":                           # one template, every RVE morphed from it (rve_morph.py)
        from deepbnd_b200.mesh.rve_morph import MorphTemplate, buildRVEmesh_morphed, morph_parameters
        tpl = MorphTemplate(param[0], size=size, lcar=lcar)
        MORPH["template"], MORPH["prm"] = tpl, morph_parameters(param.copy(), tpl)
        return [buildRVEmesh_morphed(param[i].copy(), tpl) for i in range(len(param))]
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
            srt = sorted(sm)
            out["sm_mhz"] = srt[len(srt) // 2]
            out["sm_max_mhz"] = max(smax)
        out["reasons"] = sorted(reasons)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def synthetic_inputs(wl, ns, nuniq, rank, lcar, nproc):
    """meshes (at most `nuniq` distinct ones, cycled to ns — every copy is set up and solved as an independent
    system), macro strains, Dirichlet data, projection basis of one workload"""
    from deepbnd_b200.synthetic import synthetic_param, smooth_uD
    from deepbnd_b200.elasticity import macro_strain_voigt
    param = synthetic_param(nuniq, seed=2 + 1000 * rank)
    t0 = time.perf_counter()
    meshes = make_meshes(param, wl["size"], lcar, nproc, wl.get("mesher", "delaunay"))
    meshes = [meshes[i % nuniq] for i in range(ns)]
    morph = None
    if wl.get("mesher") == "morph":
        morph = (MORPH["template"], np.ascontiguousarray(MORPH["prm"][np.arange(ns) % nuniq]))
    t_mesh = time.perf_counter() - t0
    nl = len(wl["loads"])
    uD = None
    if wl["model"] == "dnn":
        rng = np.random.RandomState(5 + rank)
        base = np.stack([smooth_uD(81, 17 + i) for i in range(3)])
        uD = base[None] * (1.0 + 0.1 * rng.randn(ns, 1, 1))
    eps = np.broadcast_to(np.stack([macro_strain_voigt(i) for i in wl["loads"]]), (ns, nl, 3)).copy()
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
    return dict(meshes=meshes, flat=flatten(meshes), eps=eps, uD=uD, W=W, M=M, t_mesh=t_mesh, nl=nl, morph=morph)


class Runner:
    """device-resident and end-to-end passes of one workload on the two handles of this rank"""

    def __init__(self, hb, pool, wl, inp, args, barrier):
        self.hb, self.pool, self.wl, self.inp, self.args, self.barrier = hb, pool, wl, inp, args, barrier
        self.ns = len(inp["meshes"])
        voff, toff, xy, tri, reg = inp["flat"]
        csz = wl.get("chunk", 0)
        nchunk = max(1, -(-self.ns // csz)) if csz else max(1, min(args.chunks, self.ns // 64))
        cb = [self.ns * c // nchunk for c in range(nchunk + 1)]
        self.chunks = [(voff[a:z + 1] - voff[a], toff[a:z + 1] - toff[a], xy[voff[a]:voff[z]], tri[toff[a]:toff[z]],
                        reg[toff[a]:toff[z]], slice(a, z)) for a, z in zip(cb[:-1], cb[1:])]
        self.nchunk = nchunk
        self.vref_box = (-1.0, -1.0, 2.0, 2.0)
        self.shared = None
        if wl.get("mesher") == "morph" and not args.no_shared_topology:
            self.shared = inp["morph"]

    def set_pass(self, bb, c=None):
        if self.shared is not None:                      # morphed meshes: one symbolic phase, coordinates made on the GPU
            tpl, prm = self.shared
            bb.set_morphed(tpl, prm if c is None else prm[self.chunks[c][5]], param=PARAM_MAT, model=self.wl["model"],
                           vref_nb=21, vref_box=self.vref_box)
            return
        arr = self.inp["flat"] if c is None else self.chunks[c][:5]
        bb.set_batch(*arr, param=PARAM_MAT, model=self.wl["model"], vref_nb=21, vref_box=self.vref_box)

    def solve_pass(self, bb, sl_=slice(None)):
        inp, wl = self.inp, self.wl
        bb.assemble()
        status, iters = bb.solve(inp["eps"][sl_], uD=None if inp["uD"] is None else inp["uD"][sl_], rtol=self.args.rtol,
                                 maxit=4000, precond=self.args.precond)
        if inp["nl"] == 3:
            out = bb.tangent()
        else:
            out = bb.snapshot()
            if wl["project"]:
                out["Y"] = bb.project(inp["W"], inp["M"], translate=out["a"])
        return out, iters

    def e2e_steps(self, nsteps):
        """nsteps full passes from HOST mesh arrays.  Two host workers, one per handle, each run set_batch (H2D of the
        raw meshes + GPU symbolic phase) -> set-up -> solve -> epilogues -> D2H for every other chunk; the kernels of
        the two handles interleave on the GPU (deepbnd_b200.batch.DoubleBuffer does the same for the drop-in scripts)."""
        jobs = [c for _ in range(nsteps) for c in range(self.nchunk)]
        outs = [None] * len(jobs)

        def worker(k):
            for i in range(k, len(jobs), 2):
                self.set_pass(self.hb[k], jobs[i])
                outs[i] = self.solve_pass(self.hb[k], self.chunks[jobs[i]][5])

        futs = [self.pool.submit(worker, k) for k in range(min(2, len(jobs)))]
        for f_ in futs:
            f_.result()
        last = outs[-self.nchunk:]
        iters = np.concatenate([o[1] for o in last])
        if self.inp["nl"] == 3:
            return np.concatenate([o[0] for o in last]), iters
        return {k_: np.concatenate([o[0][k_] for o in last]) for k_ in last[0][0]}, iters

    def timed(self, fn, *a):
        self.barrier()
        t = time.perf_counter()
        r = fn(*a)
        self.barrier()
        return time.perf_counter() - t, r

    def counters(self):
        c0_, c1_ = self.hb[0].timings()[1], self.hb[1].timings()[1]
        return {k_: c0_[k_] + c1_[k_] for k_ in c0_}

    def measure(self, steps, warmup, do_e2e=True):
        res = {}
        b = self.hb[0]
        if do_e2e:
            self.timed(self.e2e_steps, max(3, warmup))
            cnt0 = self.counters()
            res["t_e2e"], (res["out"], res["iters"]) = self.timed(self.e2e_steps, steps)
            cnt1 = self.counters()
            res["h2d"] = (cnt1["h2d_bytes"] - cnt0["h2d_bytes"]) / steps
            res["d2h"] = (cnt1["d2h_bytes"] - cnt0["d2h_bytes"]) / steps
        # device-resident loop (batch already in HBM): `value`
        self.set_pass(b)
        for _ in range(1 if do_e2e else max(1, warmup)):
            self.timed(self.solve_pass, b)
        launches0 = b.timings()[1]["kernel_launches"]
        spmv_ms = spmv_bytes = 0.0
        phase = {}
        self.barrier()
        t = time.perf_counter()
        for _ in range(steps):
            out, iters = self.solve_pass(b)
            tt, cc = b.timings()
            spmv_ms += tt["spmv"]; spmv_bytes += cc["spmv_alg_bytes"]
            for k_, v_ in tt.items():
                phase[k_] = phase.get(k_, 0.0) + v_ / steps
            spmv_launches = cc["spmv_launches"]
        self.barrier()
        res["t_dev"] = time.perf_counter() - t
        res.update(spmv_ms=spmv_ms, spmv_bytes=spmv_bytes, spmv_launches=spmv_launches, phase=phase,
                   launches=b.timings()[1]["kernel_launches"] - launches0)
        if not do_e2e:
            res["out"], res["iters"] = out, iters
        res["dof"] = int(np.mean([2 * b.sizes(s)["n_rows"] for s in range(min(self.ns, 8))]))
        return res


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
    ap.add_argument("--lcar", type=float, default=0.0, help="override mesh size (default 2/30; C5: 1/30)")
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--precond", default="twolevel", choices=["twolevel", "jacobi"])
    ap.add_argument("--unique", type=int, default=256, help="distinct synthetic RVEs meshed per rank (cycled to --ns); 0 = all")
    ap.add_argument("--chunks", type=int, default=2, help="chunks per step in the pipelined e2e loop (measured r02h: e2e / value 0.896, 0.892, 0.874, 0.829 for 1, 2, 4, 8)")
    ap.add_argument("--operator", default="matfree", choices=["matfree", "assembled"],
                    help="q = A p inside the PCG: element-wise (default) or the assembled SELL-32 SpMV")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads block")
    ap.add_argument("--mesher", default="", choices=["", "delaunay", "morph"], help="override the workload's mesher")
    ap.add_argument("--no-shared-topology", action="store_true", help="morph mesher: still run the symbolic phase per RVE")
    ap.add_argument("--cpu-sample", type=int, default=0, help="RVEs in the CPU baseline sample (default: one per core)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.ns:
        wl["ns"] = args.ns
    if args.mesher:
        wl["mesher"] = args.mesher
    lcar = args.lcar if args.lcar > 0 else wl.get("lcar")
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
                "config": {"workload": args.workload, "rves_per_gpu": wl["ns"], "mesh": wl["size"], "model": wl["model"],
                           "loads": list(wl["loads"]), "lcar": lcar or 2 / 30},
                "cpu_baseline": {"value": val, "unit": "RVE/s", "cores": ncpu, "kind": "port",
                                 "sample": "%d RVEs per step (one per core) of the %d-RVE workload, %d serial worker processes, "
                                           "oracle port: scipy splu on the bordered KKT system" % (nsamp, wl["ns"], ncpu)},
                "e2e": {"value": val, "unit": "RVE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _OUT.write(json.dumps(line) + "\n")
        _OUT.flush()
        return

    # ---------------- B200 arm ----------------
    ns = wl["ns"]
    # meshing is outside the path ("generated once and cached"): at most --unique distinct RVEs are meshed per
    # rank and cycled to the batch size; every copy is still set up and solved as an independent system
    nuniq = min(ns, args.unique) if args.unique > 0 else ns
    inp = synthetic_inputs(wl, ns, nuniq, rank, lcar, nproc_local)       # forks BEFORE any CUDA initialisation
    extra_inp = {}
    do_extra = world == 1 and not args.no_extra and args.workload == "C2_snapshots_1k_full_per" and not args.ns
    if do_extra:
        for name, n2 in EXTRA.items():
            w2 = WORKLOADS[name]
            extra_inp[name] = synthetic_inputs(w2, n2, min(n2, args.unique or n2), rank, w2.get("lcar"), nproc_local)
    # CPU baseline first (worker processes are forked before CUDA is initialised)
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        nsamp = min(ns, args.cpu_sample or ncpu)
        uDs = None if inp["uD"] is None else [list(inp["uD"][s]) for s in range(nsamp)]
        rate, dt = cpu_reference_rate(inp["meshes"][:nsamp], wl["model"], wl["loads"], uDs, ncpu)
        cpu = {"value": rate, "unit": "RVE/s", "cores": ncpu, "kind": "port",
               "sample": "first %d RVEs of the same batch, %d serial worker processes (oracle port: scipy splu on the "
                         "bordered KKT system), %.1f s" % (nsamp, ncpu, dt)}

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from deepbnd_b200.batch import RVEBatch
    from concurrent.futures import ThreadPoolExecutor

    # two handles: while one chunk is being solved, a second host thread uploads the next one and runs its symbolic
    # phase on the other handle (ctypes releases the GIL)
    hb = [RVEBatch(local_rank).set_operator(args.operator), RVEBatch(local_rank).set_operator(args.operator)]
    pool = ThreadPoolExecutor(max_workers=2)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run = Runner(hb, pool, wl, inp, args, barrier)
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = run.measure(args.steps, args.warmup, do_e2e=True)
    clocks = sampler.stop()

    # assembled SELL-32 SpMV of the same batch, timed alone (the stored-matrix kernel BASELINE's "SpMV HBM GB/s" names)
    spmv_as = None
    if rank == 0:
        try:
            bs = hb[1]
            nsp = min(ns, 256 if wl["size"] == "full" else 2000)
            bs.set_operator("assembled")
            bs.set_batch(*flatten(inp["meshes"][:nsp]), param=PARAM_MAT, model=wl["model"], vref_nb=21, vref_box=run.vref_box)
            bs.assemble()
            kk = 3 if (wl["model"] == "MR" or inp["nl"] == 3) else inp["nl"]
            ms_, by_ = bs.bench_spmv(n_rhs=kk, reps=20)
            spmv_as = {"kernel": "k_spmv<%d>" % kk, "rves": nsp, "ms_per_launch": ms_, "alg_bytes_per_launch": by_,
                       "achieved": by_ / (ms_ * 1e-3) / 1e9, "unit": "GB/s"}
            bs.set_operator(args.operator)
        except Exception as e:                       # noqa: BLE001 - diagnostics only
            spmv_as = {"error": str(e)}

    tmax = torch.tensor([res["t_dev"], res["t_e2e"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        # the only exchange of the path: final gather of the per-RVE outputs on rank 0
        out = res["out"]
        flat = torch.from_numpy(np.ascontiguousarray(
            out.reshape(ns, -1) if inp["nl"] == 3 else np.concatenate([out[k].reshape(ns, -1) for k in ("sol", "sigma", "a", "B", "sigmaT")], 1))).cuda()
        gathered = [torch.empty_like(flat) for _ in range(world)] if rank == 0 else None
        dist.gather(flat, gathered, dst=0)
    t_dev_max, t_e2e_max = tmax.tolist()
    total = ns * world
    value = total * args.steps / t_dev_max
    e2e_val = total * args.steps / t_e2e_max

    extras = []
    if do_extra and rank == 0:
        for name, n2 in EXTRA.items():
            r2 = Runner(hb, pool, WORKLOADS[name], extra_inp[name], args, barrier).measure(2, 1, do_e2e=True)
            extras.append({"workload": name, "rves_per_gpu": n2, "value": n2 * 2 / r2["t_dev"], "e2e": n2 * 2 / r2["t_e2e"], "unit": "RVE/s", "steps": 2, "warmup": 1,
                           "ms_per_step": 1e3 * r2["t_dev"] / 2, "mean_pcg_iters": float(np.mean(r2["iters"])),
                           "dofs_per_rve": r2["dof"], "operator_alg_gbs": r2["spmv_bytes"] / (r2["spmv_ms"] * 1e-3) / 1e9,
                           "phase_ms": {k_: round(v_, 2) for k_, v_ in r2["phase"].items()}})

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = res["spmv_bytes"] / (res["spmv_ms"] * 1e-3) / 1e9 if res["spmv_ms"] > 0 else 0.0
    # DRAM bytes of one operator launch from the committed ncu --set full capture (per RVE): reported next to the
    # algorithmic figure, NOT measured in this run
    traffic_ncu = None
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(args.workload + ":" + args.operator)
        if tj and not args.lcar:
            traffic_ncu = dict(tj)
            ms_launch = res["spmv_ms"] / max(1.0, res["spmv_launches"] * args.steps)
            traffic_ncu["dram_gbs_at_this_run"] = tj["dram_bytes_per_rve_per_launch"] * ns / (ms_launch * 1e-3) / 1e9
            # the kernel's distance from the HBM roofline in bytes really moved (the CSR-accounted `frac` above is a format comparison)
            traffic_ncu["frac_of_peak_in_dram_bytes"] = traffic_ncu["dram_gbs_at_this_run"] / peak
    if spmv_as and "achieved" in spmv_as:
        spmv_as["frac"] = spmv_as["achieved"] / peak

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    opname = "k_apply (matrix-free element-wise operator)" if args.operator == "matfree" else "k_spmv (assembled SELL-32)"
    line = {"metric": "RVE solves/sec", "value": value, "unit": "RVE/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "rves_per_gpu": ns, "mesh": wl["size"], "lcar": lcar or 2 / 30,
                       "model": wl["model"], "loads": list(wl["loads"]), "dofs_per_rve": res["dof"], "rtol": args.rtol,
                       "precond": args.precond, "operator": args.operator, "mean_pcg_iters": float(np.mean(res["iters"])),
                       "l2": "per-iteration working set >> 126 MB L2 (inputs larger than L2, no flush needed)",
                       "unique_meshes": nuniq, "mesher": wl.get("mesher", "delaunay"),
                       "shared_topology": run.shared is not None, "mesh_gen_s": inp["t_mesh"],
                       "phase_ms": res["phase"]},
            "e2e": {"value": e2e_val, "unit": "RVE/s", "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                    "pipeline": "2 handles / 2 host workers, %d chunks per step: each worker runs H2D of the %s + solve + D2H for every other chunk; the two streams interleave on the GPU" % (
                        run.nchunk, "inclusion parameters (the meshes are morphed from one template on the GPU) + ONE GPU symbolic phase per chunk" if run.shared is not None else "raw meshes + GPU symbolic phase")},
            "gpu_launches": int(res["launches"]),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": opname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "alg_bytes_model": "B_spmv(K) = 12 nnz + 4 (n+1) + 16 n K per RVE (scalar-CSR accounting of SURVEY 8d); the "
                                            "matrix-free kernel moves several times fewer bytes, so frac > 1 is the format advantage, not a bandwidth claim",
                         "traffic_ncu": traffic_ncu,
                         "launches_timed": int(res["spmv_launches"] * args.steps),
                         "alg_bytes_per_launch": res["spmv_bytes"] / max(1.0, res["spmv_launches"] * args.steps),
                         "spmv_assembled": spmv_as},
            "cpu_baseline": cpu,
            "extra_workloads": extras}
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
