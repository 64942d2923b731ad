"""Drop-in for creation_model/prediction/tangents_prediction.py (reference :30-100): homogenised
tangents of every RVE of paramRVEdataset.hd5 under the boundary model `modelBnd` ('dnn' with the
predicted fluctuations of bcs.hd5, 'lin' = zero fluctuation, 'per', 'MR'), written to tangents.hd5 as
`id (n,)`, `tangent (n,3,3)`, `center (n,2)` (:56-57) — batched on one B200 per rank.

The reference reads the rank from a module-level `run`; here `num` is that rank (the signature keeps the
reference's positional order)."""
import os
import sys
from timeit import default_timer as timer

import numpy as np

from . import wrapper_h5py as myhd
from .batch import DoubleBuffer
from .elasticity import macro_strain_voigt
from .simulation_snapshots import default_device, get_meshes, make_mesh_pool, set_batch_meshes, vref_box_from_mesh
from .vref import load_vref_vectors


def predictTangents(num, num_runs, modelBnd, namefiles, createMesh, meshSize, device=None, batch_size=2048,
                    rtol=1e-10, nproc=None, mesher="delaunay"):
    nameMeshRefBnd, paramRVEname, tangentName, BCname, meshMicroName = namefiles
    run = num
    start = timer()
    vref = vref_box_from_mesh(nameMeshRefBnd)

    paramRVEdata = myhd.loadhd5(paramRVEname, 'param')[run::num_runs]
    ns = len(paramRVEdata)                                            # per rank
    if myhd.checkExistenceDataset(paramRVEname, 'id'):
        ids = myhd.loadhd5(paramRVEname, 'id')[run::num_runs].astype('int')
        centers = myhd.loadhd5(paramRVEname, 'center')[ids, :]
    else:                                                             # dummy (reference :52-54)
        ids = np.zeros(ns)
        centers = np.zeros((ns, 2))

    if os.path.exists(tangentName):
        os.remove(tangentName)
    (Iid, Itangent, Icenter), f = myhd.zeros_openFile(tangentName, [(ns,), (ns, 3, 3), (ns, 2)],
                                                      ['id', 'tangent', 'center'], mode='w')
    uD_all = None
    if modelBnd == 'dnn':
        # bcs.hd5 in the library's Vref DOF order: permuted if the file (or a sidecar) describes another order, refused
        # if it describes none (deepbnd_b200/vref.py)
        u012 = load_vref_vectors(BCname, ['u0', 'u1', 'u2'], box=vref[0], Nb=vref[1])
        uD_all = np.stack([u[run::num_runs, :] for u in u012], axis=1)

    contrast, E2, nu = 10.0, 1.0, 0.3
    param = [nu, E2 * contrast, nu, E2]
    eps3 = np.stack([macro_strain_voigt(i) for i in range(3)])
    box, nb = vref

    template = None
    if mesher == "morph":
        from .mesh.rve_morph import MorphTemplate, morph_parameters
        template = MorphTemplate(paramRVEdata[0], size=meshSize)
    elif mesher != "delaunay":
        raise ValueError("mesher must be 'delaunay' or 'morph'")

    def prepare(rows, batch):
        names = [meshMicroName.format(int(ids[i]) if np.any(ids) else int(i), meshSize) for i in rows]
        if template is not None:              # morphing on the GPU from the template: only (l, e, theta) per inclusion travel
            sub = paramRVEdata[rows]
            prm = morph_parameters(sub, template)
            paramRVEdata[rows] = sub              # keep the reference's in-place permutation of the rows
            batch.set_morphed(template, prm, tuple(param), model=modelBnd, vref_nb=nb, vref_box=box)
            return
        meshes = get_meshes(paramRVEdata[rows], names, meshSize, createMesh, nproc, template, pool)
        set_batch_meshes(batch, meshes, template, tuple(param), modelBnd, nb, box)

    def solve(rows, batch, _):
        print(run, rows[0], rows[-1], ids[rows[0]])
        batch.assemble()
        batch.solve(np.broadcast_to(eps3, (len(rows), 3, 3)).copy(), uD=uD_all[rows] if modelBnd == 'dnn' else None, rtol=rtol)
        return batch.tangent()

    def store(rows, tangents):
        Iid[rows] = ids[rows]
        Itangent[rows, :, :] = tangents
        Icenter[rows, :] = centers[rows, :]
        f.flush(rows)
        sys.stdout.flush()

    pool = make_mesh_pool(nproc) if template is None else None             # forked before the GPU threads exist
    db = DoubleBuffer(default_device(run) if device is None else device)
    try:
        db.run([np.arange(b0, min(ns, b0 + batch_size)) for b0 in range(0, ns, batch_size)], prepare, solve, store)
    finally:
        db.close()
        if pool is not None:
            pool.close(); pool.join()
    f.close()
    print("tangents of", ns, "RVEs concluded in", timer() - start)


if __name__ == '__main__':
    run = int(sys.argv[1]) if len(sys.argv) > 1 else int(os.environ.get("RANK", 0))
    num_runs = int(sys.argv[2]) if len(sys.argv) > 2 else int(os.environ.get("WORLD_SIZE", 1))
    folder = os.environ.get("DEEPBND_DATA", "./deepBND_data") + "/deepBND/"
    namefiles = [folder + 'dataset/boundaryMesh.xdmf', folder + 'prediction/paramRVEdataset.hd5',
                 folder + 'prediction/tangents_{0}.hd5'.format(run), folder + 'prediction/bcs.hd5',
                 folder + 'prediction/meshes/mesh_micro_{0}_{1}.xdmf']
    predictTangents(run, num_runs, 'dnn', namefiles, True, 'reduced')
