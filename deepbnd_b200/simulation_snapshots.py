"""Drop-in for creation_model/dataset/simulation_snapshots.py (reference :42-107): builds
snapshots{suffix}_{run}.hd5 — for each RVE of paramRVEdataset.hd5 the fluctuation on the internal
boundary (Vref samples), (a, B) affine data and homogenised stresses for the shear 'S' (Voigt 2) and
axial 'A' (Voigt 0) loads — but solves the RVEs in BATCHES on one B200 through the C-ABI instead of
one FEniCS/multiphenics LU at a time.

Same call signature, same rank-strided split (`ids_run = arange(run, ns, num_runs)`, :76), same
datasets and shapes (:80-83), same in-place permutation of the param rows by the mesher (:97-98).
Extra keyword arguments only add what a GPU run needs (device, batch size, mesh cache directory).
"""
import multiprocessing as mp
import os
from timeit import default_timer as timer

import numpy as np

from . import wrapper_h5py as myhd
from .batch import DoubleBuffer
from .elasticity import macro_strain_voigt
from .mesh.mesh_io import cache_name, load_mesh, read_xdmf_points, save_mesh_cache
from .mesh.rve_mesher import buildRVEmesh_arrays, paramRVE_default
from .vref import COMP_LABEL, COORD_LABEL, vref_order_datasets

LABELS = ['id', 'solutions_S', 'sigma_S', 'a_S', 'B_S', 'sigmaTotal_S',
          'solutions_A', 'sigma_A', 'a_A', 'B_A', 'sigmaTotal_A']


def vref_box_from_mesh(bndMeshname):
    """Bounding box (x0, y0, Lx, Ly) and Nb of the internal-boundary reference mesh
    (degeneratedBoundaryRectangleMesh, core/mesh/degenerated_rectangle_mesh.py:3-35).  The reference's constants
    (paramRVE_default: [-1,1]^2, Nb = 21) are used ONLY when no file is given / the file does not exist; a file
    that exists but cannot be read raises (a silently different Vref box would change nh and every sample)."""
    p = paramRVE_default()
    if not bndMeshname or not os.path.exists(bndMeshname):
        return (p.x0L, p.y0L, p.LxL, p.LyL), 21
    pts = read_xdmf_points(bndMeshname)                         # no cell markers needed for Vref
    x0, y0 = pts.min(axis=0)
    Lx, Ly = np.ptp(pts[:, 0]), np.ptp(pts[:, 1])
    if (len(pts) - 1) % 4:
        raise ValueError("%s: %d vertices is not 4 (Nb - 1) + 1: not a degeneratedBoundaryRectangleMesh" % (bndMeshname, len(pts)))
    return (float(x0), float(y0), float(Lx), float(Ly)), (len(pts) - 1) // 4 + 1


def _mesh_job(args):
    """createMesh=True: mesh and hand the arrays back (the reference overwrites ONE temporary mesh per rank; nothing
    is left on disk here unless DEEPBND_KEEP_MESHES=1).  createMesh=False: the cached / XDMF mesh of that name, meshed
    and cached on first use ("generated once and cached")."""
    row, meshname, size, createMesh = args
    have = os.path.exists(cache_name(meshname)) or os.path.exists(meshname)
    if createMesh or not have:
        mesh = buildRVEmesh_arrays(row, isOrdered=False, size=size)        # permutes `row` in place
        if not createMesh or os.environ.get("DEEPBND_KEEP_MESHES") == "1":
            save_mesh_cache(meshname, mesh)
        return mesh, row
    return load_mesh(meshname), row


def make_mesh_pool(nproc=None):
    """Process pool for the meshing jobs, to be created BEFORE the GPU handles and their host worker threads exist
    (the workers are forked from a single-threaded, CUDA-free parent and re-used for every batch).  None = serial."""
    nproc = nproc or default_nproc()
    return mp.get_context("fork").Pool(nproc) if nproc > 1 else None


def default_nproc():
    """meshing workers of this rank: the host cores divided among the ranks of the node"""
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return max(1, (os.cpu_count() or 1) // max(1, local_world))


def default_device(run):
    """CUDA device of rank `run`: LOCAL_RANK under torchrun, else the rank modulo the visible devices"""
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:                                               # noqa: BLE001
        n = 0
    return int(run) % n if n > 0 else int(run)


def get_meshes(paramRVEdata, meshnames, size, createMesh, nproc=None, template=None, pool=None):
    """Mesh every row (multi-process, results cached as .npz next to the reference's mesh name).  With a
    `template` (deepbnd_b200.mesh.rve_morph.MorphTemplate) the meshes are morphed from it instead: fixed topology,
    ~0.1 ms each, nothing cached.  `pool`: a pool from make_mesh_pool (otherwise a temporary one is forked here)."""
    if template is not None:
        from .mesh.rve_morph import buildRVEmesh_morphed
        return [buildRVEmesh_morphed(paramRVEdata[i], template) for i in range(len(meshnames))]
    jobs = [(paramRVEdata[i], meshnames[i], size, createMesh) for i in range(len(meshnames))]
    nproc = nproc or min(len(jobs), default_nproc())
    if pool is not None and len(jobs) >= 4:
        res = pool.map(_mesh_job, jobs, chunksize=max(1, len(jobs) // (4 * nproc)))
    elif nproc <= 1 or len(jobs) < 4:
        res = [_mesh_job(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(nproc) as tmp:
            res = tmp.map(_mesh_job, jobs, chunksize=max(1, len(jobs) // (4 * nproc)))
    for i, (_, row) in enumerate(res):
        paramRVEdata[i] = row                                               # keep the reference's in-place permutation
    return [m for m, _ in res]


def set_batch_meshes(batch, meshes, template, param, model, nb, box):
    """morphed meshes share the template's topology: one symbolic phase for the whole batch (rve_set_batch_shared)"""
    if template is not None:
        batch.set_shared(np.stack([m[0] for m in meshes]), template.tri, template.region, param, model=model, vref_nb=nb, vref_box=box)
    else:
        batch.set_meshes(meshes, param, model=model, vref_nb=nb, vref_box=box)


def solve_snapshot_batch(rows, batch, datasets, rtol=1e-10):
    """Batch twin of solve_snapshot (reference :42-64) on a batch already set on the GPU: fills rows `rows`
    of the 11 datasets."""
    start = timer()
    out = _solve_snapshots(batch, rtol)
    _store_snapshots(rows, out, datasets)
    print("batch of", batch.n_sys, "snapshots concluded in ", timer() - start)
    return out


def _solve_snapshots(batch, rtol):
    """loads S (Voigt 2) and A (Voigt 0) like the reference's loop `for j_voigt, sol, ... in zip([2, 0], ...)` (:52)"""
    batch.assemble()
    n = batch.n_sys
    eps = np.broadcast_to(np.stack([macro_strain_voigt(2), macro_strain_voigt(0)]), (n, 2, 3)).copy()
    batch.solve(eps, rtol=rtol)
    return batch.snapshot()


def _store_snapshots(rows, out, datasets):
    ids, sol_S, sigma_S, a_S, B_S, sigmaT_S, sol_A, sigma_A, a_A, B_A, sigmaT_A = datasets
    for k, (sol, sigma, a, B, sigmaT) in enumerate(((sol_S, sigma_S, a_S, B_S, sigmaT_S),
                                                     (sol_A, sigma_A, a_A, B_A, sigmaT_A))):
        sol[rows, :] = out["sol"][:, k]
        sigma[rows, :] = out["sigma"][:, k]
        a[rows, :] = out["a"][:, k]
        B[rows, :, :] = out["B"][:, k]
        sigmaT[rows, :] = out["sigmaT"][:, k]


LAST_TIMINGS = {}          # wall-clock phases of the last buildSnapshots call of this process (diagnostics)


def buildSnapshots(paramMaterial, filesnames, opModel, createMesh, run, num_runs, comm_self=None,
                   device=None, batch_size=256, rtol=1e-10, nproc=None, mesher="delaunay", lcar=None):
    t_begin = timer()
    bndMeshname, paramRVEname, snapshotsname, meshname = filesnames
    box, nb = vref_box_from_mesh(bndMeshname)
    nh = 2 * (4 * (nb - 1) + 1)                                             # Vref.dim()

    ns = len(myhd.loadhd5(paramRVEname, 'param')[:, 0])
    ids_run = np.arange(run, ns, num_runs).astype('int')
    nrun = len(ids_run)

    if os.path.exists(snapshotsname):
        os.remove(snapshotsname)
    snapshots, fsnaps = myhd.zeros_openFile(filename=snapshotsname,
                                            shape=[(nrun,)] + 2 * [(nrun, nh), (nrun, 3), (nrun, 2), (nrun, 2, 2), (nrun, 3)],
                                            label=LABELS, mode='w-')
    ids = snapshots[0]
    # DOF order of the solutions_* vectors, self-described in the file (deepbnd_b200/vref.py)
    coords, comps = vref_order_datasets(box, nb)
    myhd.addDataset(fsnaps, coords, COORD_LABEL)
    myhd.addDataset(fsnaps, comps, COMP_LABEL)
    paramRVEdata = myhd.loadhd5(paramRVEname, 'param')[ids_run]

    template = None
    if mesher == "morph":                     # one template mesh per rank, every snapshot morphed from it (rve_morph.py)
        from .mesh.rve_morph import MorphTemplate, morph_parameters
        template = MorphTemplate(paramRVEdata[0], size='full', lcar=lcar)          # lcar: None = the reference's 2/30
    elif mesher != "delaunay":
        raise ValueError("mesher must be 'delaunay' or 'morph'")

    def prepare(rows, batch):
        # the reference reuses ONE temporary mesh name per rank; a batch needs one cache entry per snapshot
        names = [meshname.replace(".xdmf", "_%d.xdmf" % int(ids_run[i])) for i in rows]
        if template is not None:              # morphing on the GPU from the template: only (l, e, theta) per inclusion travel
            sub = paramRVEdata[rows]
            prm = morph_parameters(sub, template)
            paramRVEdata[rows] = sub              # keep the reference's in-place permutation of the rows
            batch.set_morphed(template, prm, tuple(paramMaterial), model=opModel, vref_nb=nb, vref_box=box)
            return
        meshes = get_meshes(paramRVEdata[rows], names, 'full', createMesh, nproc, template, pool)
        set_batch_meshes(batch, meshes, template, tuple(paramMaterial), opModel, nb, box)

    def solve(rows, batch, _):
        print("Solving snapshots", int(ids_run[rows[0]]), "..", int(ids_run[rows[-1]]))
        return _solve_snapshots(batch, rtol)

    def store(rows, out):
        ids[rows] = ids_run[rows]
        _store_snapshots(rows, out, snapshots)
        fsnaps.flush(rows)

    pool = make_mesh_pool(nproc) if template is None else None             # forked before the GPU threads exist
    t_files = timer()
    db = DoubleBuffer(default_device(run) if device is None else device)
    t_ctx = timer()
    try:
        db.run([np.arange(b0, min(nrun, b0 + batch_size)) for b0 in range(0, nrun, batch_size)], prepare, solve, store)
    finally:
        t_run = timer()
        db.close()
        if pool is not None:
            pool.close(); pool.join()
    fsnaps.close()
    LAST_TIMINGS.clear()
    LAST_TIMINGS.update(files_and_template_s=t_files - t_begin, cuda_context_and_handles_s=t_ctx - t_files,
                        pipeline_s=t_run - t_ctx, close_s=timer() - t_run)


if __name__ == '__main__':
    import sys
    run = int(os.environ.get("RANK", sys.argv[1] if len(sys.argv) > 1 else 0))
    num_runs = int(os.environ.get("WORLD_SIZE", sys.argv[2] if len(sys.argv) > 2 else 1))
    folder = os.environ.get("DEEPBND_DATA", "./deepBND_data") + "/deepBND/dataset/"
    suffix, opModel, createMesh = "_validation", 'per', True
    contrast, E2, nu = 10.0, 1.0, 0.3
    paramMaterial = [nu, E2 * contrast, nu, E2]
    filesnames = [folder + 'boundaryMesh.xdmf', folder + 'paramRVEdataset{0}.hd5'.format(suffix),
                  folder + 'snapshots{0}_{1}.hd5'.format(suffix, run), folder + "meshes/mesh_temp_{0}.xdmf".format(run)]
    buildSnapshots(paramMaterial, filesnames, opModel, createMesh, run, num_runs)
