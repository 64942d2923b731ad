#!/usr/bin/env python
"""FEniCS-side golden exporter.  RUNS ONLY WHERE THE REFERENCE RUNS: needs
fenics/dolfin 2019.1, multiphenics, micmacsfenics and the deepBND package on PYTHONPATH (none of which
exist in the build image — this script has not been executed there; it is written against the reference's
call sites and kept small so a maintainer can fix API drift in minutes).

What it does, for the reference's parameter table (tests/test_meshes.py:63-98 = tests/golden/
paramRVE_test_meshes.txt) and the material of simulation_snapshots.py:120-123:

  * meshes the RVE with the REFERENCE mesher (gmsh) for size 'reduced' and 'full' and dumps points, triangles
    and region tags, so that this repository can solve on the identical mesh;
  * runs MicroModel.compute() for the models 'per', 'MR', 'lin' and dumps the nodal fluctuation of loads 0 and 2
    at the mesh vertices and edge midpoints (evaluation by point, independent of dolfin's DOF order), the
    Vref samples exactly as simulation_snapshots.py:56-58 stores them, sigma/sigmaTotal, (a, B);
  * runs MicroConstitutiveModelDNN.getTangent() for 'per', 'MR', 'lin' (zero uD) and for 'dnn' with a smooth NON-ZERO uD
    (the same synthetic field deepbnd_b200.synthetic.smooth_uD builds, evaluated at the dolfin DOF coordinates), keeping
    B, Eps and the tangent of every load: this one case decides (i) how a P1 function on the boundary mesh becomes
    Dirichlet data on the P2 RVE boundary (oracle switch dirichlet_semantics: 'p1_vertex' vs 'p2_nodal', relative
    tangent delta 1e-5) and (ii) whether micmacsfenics keeps the zero-average block for lin/dnn
    (zero_average_for_dirichlet, delta > 1e-6): tests/test_fenics_goldens.py tries the variants and reports which fits;
  * dumps the Frobenius norms of the assembled block matrix A and the load vectors F per model (multiplier scaling
    and block layout show up there before any solve);
  * dumps Vref.tabulate_dof_coordinates() and the component of every DOF — written also as `vref_dof_order.hd5`
    (datasets vref_dof_coordinates / vref_dof_component), the sidecar deepbnd_b200.vref.load_vref_vectors looks for
    next to a reference-produced bcs.hd5 / Wbasis.hd5 (settles SURVEY hypothesis H5).

Output: fenics_c1_<size>.npz.  Drop the files into tests/golden/ — tests/test_fenics_goldens.py picks them
up (it is skipped while they are absent) and turns "parity unpinned" into a pinned comparison.

Usage:  python tools/export_goldens_fenics.py <out_dir>
"""
import os
import sys

import numpy as np


def main(out_dir):
    import dolfin as df                                                        # noqa: F401
    from deepBND.core.multiscale.mesh_RVE import buildRVEmesh, paramRVE_default
    import deepBND.core.multiscale.micro_model as mscm
    from deepBND.core.multiscale.micro_model_dnn import MicroConstitutiveModelDNN
    import deepBND.core.multiscale.misc as mtsm
    from deepBND.core.fenics_tools.enriched_mesh import EnrichedMesh
    from deepBND.core.mesh.degenerated_rectangle_mesh import degeneratedBoundaryRectangleMesh

    here = os.path.dirname(os.path.abspath(__file__))
    table = np.loadtxt(os.path.join(here, "..", "tests", "golden", "paramRVE_test_meshes.txt"))
    nu, E2, contrast = 0.3, 1.0, 10.0
    param = [nu, E2 * contrast, nu, E2]
    os.makedirs(out_dir, exist_ok=True)

    p = paramRVE_default()
    bnd = os.path.join(out_dir, "boundaryMesh.xdmf")
    meshRef = degeneratedBoundaryRectangleMesh(x0=p.x0L, y0=p.y0L, Lx=p.LxL, Ly=p.LyL, Nb=21)
    meshRef.generate()
    meshRef.write(bnd, 'fenics')
    Mref = EnrichedMesh(bnd)
    Vref = df.VectorFunctionSpace(Mref, "CG", 1)
    usol = df.Function(Vref)

    for size in ("reduced", "full"):
        rows = table.copy()
        meshname = os.path.join(out_dir, "mesh_c1_%s.xdmf" % size)
        buildRVEmesh(rows, meshname, isOrdered=False, size=size)
        comp = np.zeros(Vref.dim())
        for c in range(2):
            comp[Vref.sub(c).dofmap().dofs()] = c
        out = {"param_rows_after_permutation": rows, "vref_dof_coordinates": Vref.tabulate_dof_coordinates(),
               "vref_dof_component": comp}
        try:                                          # sidecar for reference-produced bcs.hd5 / Wbasis.hd5
            import h5py
            with h5py.File(os.path.join(out_dir, "vref_dof_order.hd5"), "w") as f5:
                f5["vref_dof_coordinates"] = out["vref_dof_coordinates"]
                f5["vref_dof_component"] = comp
        except ImportError:
            pass
        mesh = EnrichedMesh(meshname)
        out["points"] = mesh.coordinates().copy()
        out["triangles"] = mesh.cells().copy()
        out["region"] = np.asarray(mesh.subdomains.array(), dtype=np.int64).copy()
        # evaluation points: vertices and edge midpoints (= the P2 nodes of this repository)
        mesh.init(1)
        e2v = np.array([e.entities(0) for e in df.edges(mesh)])
        mids = 0.5 * (out["points"][e2v[:, 0]] + out["points"][e2v[:, 1]])
        out["edges"] = e2v
        evalpts = np.vstack([out["points"], mids])
        for model in ("per", "MR", "lin"):
            mm = mscm.MicroModel(meshname, param, model)
            mm.compute()
            for j, lab in ((2, "S"), (0, "A")):
                u = mm.sol[j][0]
                u.set_allow_extrapolation(True)
                out["u_%s_%s" % (model, lab)] = np.array([u(pt) for pt in evalpts])
                T, a, B = mtsm.getAffineTransformationLocal(u, mm.mesh, [0, 1], justTranslation=False)
                out["a_%s_%s" % (model, lab)] = np.asarray(a)
                out["B_%s_%s" % (model, lab)] = np.asarray(B)
                usol.interpolate(u)
                out["solutions_%s_%s" % (model, lab)] = usol.vector().get_local().copy()
                out["sigma_%s_%s" % (model, lab)] = mm.homogenise([0, 1], j).flatten()[[0, 3, 2]]
                if size == "full":
                    out["sigmaTotal_%s_%s" % (model, lab)] = mm.homogenise([0, 1, 2, 3], j).flatten()[[0, 3, 2]]
            # norms of the assembled system (micro_model.py:56-62): block layout / multiplier scaling
            try:
                import multiphenics as mp
                Eps_ = df.Constant(((0., 0.), (0., 0.)))
                a_form, f_form, bcs_, W_ = mm.multiscaleModel(mm.mesh, mm.sigmaLaw, Eps_, mm.others)()   # micro_model.py:56-57
                A_ = mp.block_assemble(a_form)
                out["A_norm_%s" % model] = A_.norm("frobenius")
                out["n_dofs_%s" % model] = A_.size(0)
                for i in range(3):
                    Eps_.assign(df.Constant(mscm.mscm.macro_strain(i)))                                 # :72
                    out["F_norm_%s_%d" % (model, i)] = mp.block_assemble(f_form).norm("l2")
            except Exception as err:                                          # noqa: BLE001 - diagnostics only
                print("norm dump skipped for", model, ":", err)
            mc = MicroConstitutiveModelDNN(meshname, param, model)
            if model == "lin":
                mc.others['uD'] = df.Function(Vref)
                for i in range(3):
                    mc.others['uD%d_' % i] = np.zeros(Vref.dim())
            out["tangent_%s" % model] = np.asarray(mc.getTangent())
        # 'dnn' with smooth non-zero boundary data (amplitude 1e-2, 3 Fourier modes per component; the same field as
        # deepbnd_b200.synthetic.smooth_uD(81, seed), here evaluated at the dolfin DOF coordinates)
        def smooth_field(seed, xy, amp=1e-2):
            rng = np.random.RandomState(seed)
            x0, y0, L = p.x0L, p.y0L, p.LxL
            s_param = np.zeros(len(xy))                                       # CCW arc parameter of the boundary point
            for d, (x, y) in enumerate(xy):
                if abs(y - y0) < 1e-9: s_param[d] = (x - x0) / L
                elif abs(x - x0 - L) < 1e-9: s_param[d] = 1 + (y - y0) / L
                elif abs(y - y0 - L) < 1e-9: s_param[d] = 2 + (x0 + L - x) / L
                elif abs(x - x0) < 1e-9: s_param[d] = 3 + (y0 + L - y) / L
                else: s_param[d] = np.nan                                     # centre node
            ang = 2 * np.pi * s_param / 4.0
            u = np.zeros((len(xy), 2))
            for c in range(2):
                for k in range(1, 4):
                    u[:, c] += amp * rng.randn() * np.cos(k * ang) + amp * rng.randn() * np.sin(k * ang)
            return np.nan_to_num(u)
        mc = MicroConstitutiveModelDNN(meshname, param, "dnn")
        mc.others['uD'] = df.Function(Vref)
        xy = out["vref_dof_coordinates"]
        for i in range(3):
            uv = smooth_field(40 + i, xy)
            vec = np.where(comp == 0, uv[:, 0], uv[:, 1])
            mc.others['uD%d_' % i] = vec
            out["dnn_uD%d_dolfin_order" % i] = vec
        out["tangent_dnn"] = np.asarray(mc.getTangent())
        np.savez(os.path.join(out_dir, "fenics_c1_%s.npz" % size), **out)
        print("wrote", os.path.join(out_dir, "fenics_c1_%s.npz" % size))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "fenics_goldens")
