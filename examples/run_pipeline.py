#!/usr/bin/env python
"""End-to-end example of the drop-in chain on synthetic data (one GPU):

  paramRVEdataset.hd5  --buildSnapshots-->  snapshots_0.hd5  --build_inout (ops 0-3)-->  Wbasis.hd5, XY.hd5
  XY.hd5 --(least-squares stand-in for the NN training)--> net_X_weights.npz --predictBCs--> bcs.hd5
  paramRVEdataset.hd5 (+ bcs.hd5)  --predictTangents (lin / per / MR / dnn)-->  tangents_<model>.hd5

  python examples/run_pipeline.py [out_dir] [n_snapshots] [n_tangents]

File names and dataset layouts are the reference's (creation_model/dataset/simulation_snapshots.py,
creation_model/RB/build_inout.py, creation_model/prediction/tangents_prediction.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fit_linear_net(nameXY, label, nrb, prefix):
    """least-squares fit of a Dense(36 -> nrb, linear) 'network' on XY.hd5 in MinMax-scaled variables; writes the
    weight and scaler files NetArch expects and returns the NetArch"""
    import deepbnd_b200.wrapper_h5py as myhd
    from deepbnd_b200.bc_prediction import NetArch
    X, Y = myhd.loadhd5(nameXY, ['X', 'Y_%s' % label])
    X, Y = np.asarray(X)[:, :36], np.asarray(Y)[:, :nrb]
    lim = np.zeros((max(36, nrb), 4))
    lim[:36, 0], lim[:36, 1] = X.min(0) - 1e-3, X.max(0) + 1e-3
    lim[:nrb, 2], lim[:nrb, 3] = Y.min(0) - 1e-12, Y.max(0) + 1e-12
    Xs = (X - lim[:36, 0]) / (lim[:36, 1] - lim[:36, 0])
    Ys = (Y - lim[:nrb, 2]) / (lim[:nrb, 3] - lim[:nrb, 2])
    A = np.hstack([Xs, np.ones((len(Xs), 1))])
    sol = np.linalg.lstsq(A, Ys, rcond=None)[0]
    net = NetArch([], ['linear'])
    net.nX, net.nY = 36, nrb
    net.files['weights'], net.files['scaler'] = prefix + "_weights.npz", prefix + "_scaler.txt"
    np.savez(net.files['weights'], W0=sol[:36], b0=sol[36])
    np.savetxt(net.files['scaler'], lim)
    return net


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "pipeline_out"
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    nt = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    os.makedirs(os.path.join(out, "meshes"), exist_ok=True)
    from deepbnd_b200.synthetic import synthetic_param
    import deepbnd_b200.wrapper_h5py as myhd
    from deepbnd_b200.simulation_snapshots import buildSnapshots
    from deepbnd_b200.tangents_prediction import predictTangents
    import deepbnd_b200.build_inout as bio
    import deepbnd_b200.bc_prediction as bp

    param = [0.3, 10.0, 0.3, 1.0]
    p_snap, p_tan = os.path.join(out, "paramRVEdataset.hd5"), os.path.join(out, "paramRVEdataset_tangents.hd5")
    myhd.savehd5(p_snap, synthetic_param(ns, seed=13), 'param', 'w')
    myhd.savehd5(p_tan, synthetic_param(nt, seed=2), 'param', 'w')

    t = time.perf_counter()
    snaps = os.path.join(out, "snapshots_0.hd5")
    buildSnapshots(param, ["", p_snap, snaps, os.path.join(out, "meshes", "mesh_temp_0.xdmf")], 'per', True, 0, 1, None,
                   device=0, batch_size=256)
    print("snapshots: %d full RVEs in %.1f s (meshing included)" % (ns, time.perf_counter() - t))

    Nmax = min(160, ns - 2)
    wb, yl, xy = (os.path.join(out, n) for n in ("Wbasis.hd5", "Y.h5", "XY.hd5"))
    bio.translateSolution(snaps)
    bio.computingBasis(snaps, wb, Nmax)
    bio.extractAlpha(snaps, wb, Nmax, yl)
    bio.createXY(p_snap, yl, xy)
    print("RB stage: sig_A[:4] =", myhd.loadhd5(wb, 'sig_A')[:4], " XY:", myhd.loadhd5(xy, 'Y_A').shape)

    # boundary-condition prediction: a linear least-squares stand-in for the trained networks (the reference trains
    # MLPs on XY.hd5 with Keras, creation_model/training/NN_training.py — out of scope), then bcs.hd5 -> 'dnn' tangents
    bcs = os.path.join(out, "bcs.hd5")
    net = {l: fit_linear_net(xy, l, min(20, Nmax), os.path.join(out, "net_%s" % l)) for l in ("A", "S")}
    bp.predictBCs(["", wb, p_tan, bcs], net, device="cuda")
    print("bcs.hd5: u0", myhd.loadhd5(bcs, 'u0').shape, " max |u0| = %.3e" % np.abs(myhd.loadhd5(bcs, 'u0')).max())

    for model in ("lin", "per", "MR", "dnn"):
        t = time.perf_counter()
        tf = os.path.join(out, "tangents_%s.hd5" % model)
        predictTangents(0, 1, model, ["", p_tan, tf, bcs, os.path.join(out, "meshes", "mesh_micro_{0}_{1}.xdmf")],
                        model == "lin", 'reduced', device=0)
        C = myhd.loadhd5(tf, 'tangent')
        print("tangents %-3s: %d reduced RVEs in %.1f s; mean C11 = %.5f" % (model, nt, time.perf_counter() - t, C[:, 0, 0].mean()))


if __name__ == "__main__":
    main()
