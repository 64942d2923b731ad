"""RVEBatch — numpy-facing wrapper of the C-ABI (include/rve_b200.h): one object = one batch of RVEs
on one GPU.  This is the layer the mirror classes (micro_model.py) and the drop-in drivers
(simulation_snapshots.py, tangents_prediction.py) are written on."""
import ctypes

import numpy as np

from . import _lib as L
from .elasticity import getLameTable


class RVEBatch:
    def __init__(self, device=0):
        self.lib = L.load()
        self.h = ctypes.c_void_p()
        rc = self.lib.rve_create(int(device), ctypes.byref(self.h))
        if rc != 0:
            raise L.RVEError(rc, self.lib.rve_last_error(None).decode())
        self.n_sys = 0
        self.nvref = 0
        self.n_rhs = 0

    def close(self):
        if self.h:
            self.lib.rve_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise L.RVEError(rc, self.lib.rve_last_error(self.h).decode())
        return rc

    # ---- batch set-up -------------------------------------------------------------------------
    def set_meshes(self, meshes, param=(0.3, 10.0, 0.3, 1.0), model="lin", vref_nb=21, vref_box=None, mat=None):
        """meshes: list of (points (nv,2), triangles (nt,3), region (nt,)).  param = [nu1,E1,nu2,E2]
        like the reference's paramMaterial (simulation_snapshots.py:120-123)."""
        n = len(meshes)
        voff = np.zeros(n + 1, dtype=np.int64)
        toff = np.zeros(n + 1, dtype=np.int64)
        for s, (p, t, r) in enumerate(meshes):
            voff[s + 1] = voff[s] + len(p)
            toff[s + 1] = toff[s] + len(t)
        xy = L.f64(np.concatenate([np.asarray(m[0], dtype=np.float64).reshape(-1, 2) for m in meshes]))
        tri = L.i32(np.concatenate([np.asarray(m[1]).reshape(-1, 3) for m in meshes]))
        reg = L.i32(np.concatenate([np.asarray(m[2]).ravel() for m in meshes]))
        return self.set_batch(voff, toff, xy, tri, reg, param=param, model=model, vref_nb=vref_nb, vref_box=vref_box, mat=mat)

    def set_batch(self, voff, toff, xy, tri, reg, param=(0.3, 10.0, 0.3, 1.0), model="lin", vref_nb=21, vref_box=None, mat=None):
        mat = L.f64(getLameTable(*param) if mat is None else mat)
        vb = None if vref_box is None else L.f64(vref_box)
        self.n_sys = len(voff) - 1
        self.model = model
        self.nvref = 4 * (vref_nb - 1) + 1 if vref_nb else 0
        self._keep = (voff, toff, xy, tri, reg, mat, vb)
        self._check(self.lib.rve_set_batch(self.h, self.n_sys, L.lptr(voff), L.lptr(toff), L.dptr(xy), L.iptr(tri),
                                           L.iptr(reg), L.dptr(mat), L.MODEL[model], int(vref_nb), L.dptr(vb)))
        return self

    def set_shared(self, points, tri, region, param=(0.3, 10.0, 0.3, 1.0), model="lin", vref_nb=21, vref_box=None, mat=None):
        """n_sys meshes with ONE topology: points (n_sys, nv, 2), triangles (nt, 3), region (nt,) — what
        deepbnd_b200.mesh.rve_morph produces.  The symbolic phase runs once (rve_set_batch_shared)."""
        mat = L.f64(getLameTable(*param) if mat is None else mat)
        vb = None if vref_box is None else L.f64(vref_box)
        xy = L.f64(points)
        if xy.ndim != 3 or xy.shape[2] != 2:
            raise ValueError("points must be (n_sys, nv, 2)")
        tri, reg = L.i32(np.asarray(tri).reshape(-1, 3)), L.i32(np.asarray(region).ravel())
        self.n_sys = xy.shape[0]
        self.model = model
        self.nvref = 4 * (vref_nb - 1) + 1 if vref_nb else 0
        self._keep = (xy, tri, reg, mat, vb)
        self._check(self.lib.rve_set_batch_shared(self.h, self.n_sys, xy.shape[1], len(tri), L.dptr(xy), L.iptr(tri), L.iptr(reg),
                                                  L.dptr(mat), L.MODEL[model], int(vref_nb), L.dptr(vb)))
        return self

    def set_morphed(self, template, prm, param=(0.3, 10.0, 0.3, 1.0), model="lin", vref_nb=21, vref_box=None, mat=None):
        """n_sys RVEs morphed ON THE GPU from a deepbnd_b200.mesh.rve_morph.MorphTemplate: prm (n_sys, n_incl, 3) =
        (l, e, theta) of every inclusion in the mesher's (permuted) order; only these travel to the device."""
        mat = L.f64(getLameTable(*param) if mat is None else mat)
        vb = None if vref_box is None else L.f64(vref_box)
        prm = L.f64(prm)
        if prm.ndim != 3 or prm.shape[1:] != (template.n_incl, 3):
            raise ValueError("prm must be (n_sys, %d, 3)" % template.n_incl)
        xy, tri, reg = L.f64(template.points), L.i32(template.tri), L.i32(template.region)
        mv, inc = L.i32(template.moving), L.i32(template.k)
        d, u, c = L.f64(template.d), L.f64(template.dir), L.f64(template.centres)
        self.n_sys = prm.shape[0]
        self.model = model
        self.nvref = 4 * (vref_nb - 1) + 1 if vref_nb else 0
        self._keep = (xy, tri, reg, mat, vb, prm, mv, inc, d, u, c)
        self._check(self.lib.rve_set_batch_morphed(self.h, self.n_sys, len(xy), len(tri), L.dptr(xy), L.iptr(tri), L.iptr(reg), L.dptr(mat),
                                                   L.MODEL[model], int(vref_nb), L.dptr(vb), len(mv), L.iptr(mv), L.iptr(inc), L.dptr(d),
                                                   L.dptr(u), len(c), L.dptr(c), float(template.r0), float(template.R), L.dptr(prm)))
        return self

    def assemble(self):
        self._check(self.lib.rve_assemble(self.h))
        return self

    # ---- solve + epilogues --------------------------------------------------------------------
    def solve(self, eps, uD=None, rtol=1e-10, maxit=4000, precond="twolevel", raise_on_noconv=True):
        """eps: (n_sys, n_rhs, 3) Voigt macro strains; uD: (n_sys, n_rhs, 2*nvref) for 'dnn'."""
        eps = L.f64(eps).reshape(self.n_sys, -1, 3)
        self.n_rhs = eps.shape[1]
        if uD is not None:
            uD = L.f64(uD).reshape(self.n_sys, self.n_rhs, 2 * self.nvref)
        status = np.zeros(self.n_sys, dtype=np.int32)
        iters = np.zeros(self.n_sys, dtype=np.int32)
        rc = self.lib.rve_solve(self.h, self.n_rhs, L.dptr(eps), L.dptr(uD), float(rtol), int(maxit),
                                L.PRECOND[precond], L.iptr(status), L.iptr(iters))
        self._check(rc, allow=() if raise_on_noconv else (-5,))
        return status, iters

    def tangent(self):
        C = np.zeros((self.n_sys, 3, 3))
        self._check(self.lib.rve_epilogue_tangent(self.h, L.dptr(C)))
        return C

    def snapshot(self):
        n, k = self.n_sys, self.n_rhs
        out = dict(sol=np.zeros((n, k, 2 * self.nvref)) if self.nvref else None, sigma=np.zeros((n, k, 3)),
                   a=np.zeros((n, k, 2)), B=np.zeros((n, k, 2, 2)), sigmaT=np.zeros((n, k, 3)))
        self._check(self.lib.rve_epilogue_snapshot(self.h, L.dptr(out["sol"]), L.dptr(out["sigma"]), L.dptr(out["a"]),
                                                   L.dptr(out["B"]), L.dptr(out["sigmaT"])))
        return out

    def set_option(self, option, value):
        self._check(self.lib.rve_set_option(self.h, int(option), int(value)))
        return self

    def set_operator(self, name):
        """'matfree' (default: element-wise application, the matrix is never stored) or 'assembled' (SELL-32 SpMV);
        call before assemble()."""
        return self.set_option(L.OPT_OPERATOR, L.OPERATOR[name])

    def apply_operator(self, p):
        """test accessor: q = A p for the whole batch; p (total rows, n_rhs, 2)."""
        p = L.f64(p)
        q = np.zeros_like(p)
        self._check(self.lib.rve_apply_operator(self.h, p.shape[1], L.dptr(p), L.dptr(q)))
        return q

    def homogenisation(self):
        """sigmaL, sigmaT (n,k,3), gradL, gradT (n,k,2,2), eps_eff (n,k,3) of the last solve
        (micro_model_gen.py:139-147)."""
        n, k = self.n_sys, self.n_rhs
        out = dict(sigmaL=np.zeros((n, k, 3)), sigmaT=np.zeros((n, k, 3)), gradL=np.zeros((n, k, 2, 2)),
                   gradT=np.zeros((n, k, 2, 2)), eps_eff=np.zeros((n, k, 3)))
        self._check(self.lib.rve_epilogue_homogenisation(self.h, L.dptr(out["sigmaL"]), L.dptr(out["sigmaT"]),
                                                         L.dptr(out["gradL"]), L.dptr(out["gradT"]), L.dptr(out["eps_eff"])))
        return out

    def project(self, W, M, translate=None):
        """W: (n_rhs, nmax, nh) or (nmax, nh) shared; M: (nh, nh).  Returns Y (n_sys, n_rhs, nmax)."""
        W = L.f64(W)
        if W.ndim == 2:
            W = L.f64(np.broadcast_to(W, (self.n_rhs,) + W.shape))
        nmax, nh = W.shape[1], W.shape[2]
        M = L.f64(M)
        tr = None if translate is None else L.f64(translate).reshape(self.n_sys, self.n_rhs, 2)
        Y = np.zeros((self.n_sys, self.n_rhs, nmax))
        self._check(self.lib.rve_project(self.h, nmax, nh, L.dptr(W), L.dptr(M), L.dptr(tr), L.dptr(Y)))
        return Y

    # ---- accessors for tests / diagnostics ------------------------------------------------------
    def sizes(self, sys):
        v = [ctypes.c_int64() for _ in range(4)]
        self._check(self.lib.rve_get_sizes(self.h, sys, *[ctypes.byref(x) for x in v]))
        return dict(n_nodes=v[0].value, n_rows=v[1].value, n_blocks=v[2].value, n_elems=v[3].value)

    def csr(self, sys):
        import scipy.sparse as sp
        s = self.sizes(sys)
        n, nnz = 2 * s["n_rows"], 4 * s["n_blocks"]
        rp, col, val = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        self._check(self.lib.rve_get_csr(self.h, sys, L.iptr(rp), L.iptr(col), L.dptr(val)))
        return sp.csr_matrix((val, col, rp), shape=(n, n))

    def rowmap(self, sys):
        s = self.sizes(sys)
        va, vb = np.zeros(s["n_rows"], np.int32), np.zeros(s["n_rows"], np.int32)
        self._check(self.lib.rve_get_rowmap(self.h, sys, L.iptr(va), L.iptr(vb)))
        return va, vb

    def nodemap(self, sys):
        s = self.sizes(sys)
        va, vb = np.zeros(s["n_nodes"], np.int32), np.zeros(s["n_nodes"], np.int32)
        self._check(self.lib.rve_get_nodemap(self.h, sys, L.iptr(va), L.iptr(vb)))
        return va, vb

    def rhs(self, sys):
        s = self.sizes(sys)
        k = 3 if self.model == "MR" else self.n_rhs
        b = np.zeros((s["n_rows"], k, 2))
        self._check(self.lib.rve_get_rhs(self.h, sys, L.dptr(b)))
        return b

    def solution(self, sys):
        s = self.sizes(sys)
        u = np.zeros((s["n_nodes"], self.n_rhs, 2))
        self._check(self.lib.rve_get_solution(self.h, sys, L.dptr(u)))
        return u

    def coarse_levels(self):
        nl = ctypes.c_int32()
        self._check(self.lib.rve_get_coarse_dim(self.h, 0, None, ctypes.byref(nl)))
        dims = []
        for l in range(nl.value):
            g = ctypes.c_int32()
            self._check(self.lib.rve_get_coarse_dim(self.h, l, ctypes.byref(g), None))
            dims.append(g.value)
        return dims

    def coarse_dense(self, sys, level):
        G = self.coarse_levels()[level]
        A = np.zeros((2 * G, 2 * G))
        self._check(self.lib.rve_get_coarse_dense(self.h, sys, level, L.dptr(A)))
        return A

    def timings(self):
        t, c = np.zeros(8), np.zeros(8)
        self._check(self.lib.rve_timings(self.h, L.dptr(t), L.dptr(c)))
        keys = ["set_batch_numbering", "set_batch_pattern", "assemble", "precond_setup", "rhs", "pcg", "spmv", "epilogue"]
        return dict(zip(keys, t.tolist())), dict(spmv_launches=c[0], spmv_alg_bytes=c[1], pcg_iters_launched=c[2],
                                                 sum_iters=c[3], kernel_launches=c[4], post_ms=c[5], h2d_bytes=c[6],
                                                 d2h_bytes=c[7])

    def bench_spmv(self, n_rhs=3, reps=20):
        ms, by = ctypes.c_double(), ctypes.c_double()
        self._check(self.lib.rve_bench_spmv(self.h, n_rhs, reps, ctypes.byref(ms), ctypes.byref(by)))
        return ms.value, by.value


class DoubleBuffer:
    """Two handles on one GPU, one host worker thread each.  Worker k takes jobs k, k+2, ... and runs
    `prepare(job, handle)` (meshing on the host; H2D + symbolic phase on the GPU) and `solve(job, handle, prepared)`
    (assemble, PCG, epilogues, D2H) back to back; the kernels of the two handles interleave on the GPU, so the host
    gaps of one worker (meshing, result copies, synchronisation points) and the tails of its kernels are filled by
    the other.  `store(job, result)` runs on the caller's thread in job order (file writes stay serial and ordered).
    ctypes releases the GIL during the library calls."""

    def __init__(self, device=0, handles=None):
        # `handles`: two ready-made handle objects (the CPU tests pass stand-ins to check ordering / error propagation)
        self.h = list(handles) if handles is not None else [RVEBatch(device), RVEBatch(device)]

    def run(self, jobs, prepare, solve, store=None):
        import threading
        from concurrent.futures import Future
        jobs = list(jobs)
        futs = [Future() for _ in jobs]

        def worker(k):
            for i in range(k, len(jobs), 2):
                try:
                    futs[i].set_result(solve(jobs[i], self.h[k], prepare(jobs[i], self.h[k])))
                except BaseException as e:   # noqa: BLE001 - handed to the caller's thread
                    for j in range(i, len(jobs), 2):
                        futs[j].set_exception(e)
                    return

        threads = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(min(2, len(jobs)))]
        for t in threads:
            t.start()
        results = []
        try:
            for job, fut in zip(jobs, futs):
                res = fut.result()
                results.append(store(job, res) if store is not None else res)
        finally:
            for t in threads:
                t.join()
        return results

    def close(self):
        for b in self.h:
            b.close()


def host_symbolic(xy, tri, region, model="lin", vref_nb=0):
    """CPU-only test hook (rve_host_symbolic): symbolic phase of one system."""
    lib = L.load()
    xy, tri, region = L.f64(xy), L.i32(tri), L.i32(region)
    nv, nt = len(xy), len(tri)
    cap_rows, cap_blocks = nv + 3 * nt, 40 * (nv + 3 * nt)
    out = [ctypes.c_int32() for _ in range(4)]
    rowptr = np.zeros(cap_rows + 1, np.int32)
    bcol = np.zeros(cap_blocks, np.int32)
    va, vb, node_row = np.zeros(cap_rows, np.int32), np.zeros(cap_rows, np.int32), np.zeros(cap_rows, np.int32)
    nvref = 4 * (vref_nb - 1) + 1 if vref_nb else 0
    se, sl = np.zeros(max(nvref, 1), np.int32), np.zeros(2 * max(nvref, 1))
    rc = lib.rve_host_symbolic(L.dptr(xy), nv, L.iptr(tri), L.iptr(region), nt, L.MODEL[model], vref_nb, cap_rows,
                               cap_blocks, *[ctypes.byref(o) for o in out], L.iptr(rowptr), L.iptr(bcol), L.iptr(va),
                               L.iptr(vb), L.iptr(node_row), L.iptr(se), L.dptr(sl))
    if rc != 0:
        raise L.RVEError(rc, "rve_host_symbolic failed")
    nn, na, nb, nc = [o.value for o in out]
    return dict(n_nodes=nn, n_rows=na, n_blocks=nb, nc=nc, rowptr=rowptr[:na + 1], bcol=bcol[:nb], row_va=va[:na],
                row_vb=vb[:na], node_row=node_row[:nn], samp_elem=se[:nvref], samp_lam=sl[:2 * nvref].reshape(-1, 2))
