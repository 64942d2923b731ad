"""RB stage (build_inout.py drop-in) against the numpy restatement of RB_utils.py; runs on the CPU (torch cpu) here
and on the GPU (torch cuda) on the B200 box."""
import os

import numpy as np
import pytest


def _fake_snapshots(path, ns=300, seed=0):
    """smooth random boundary fields with decaying spectrum, in the snapshots.hd5 layout"""
    import deepbnd_b200.wrapper_h5py as myhd
    rng = np.random.RandomState(seed)
    s = 2 * np.pi * np.arange(80) / 80
    data = {}
    for lab in ("A", "S"):
        U = np.zeros((ns, 162))
        for k in range(1, 40):
            amp = rng.randn(ns, 4) / k ** 2
            U[:, 0:160:2] += amp[:, [0]] * np.cos(k * s) + amp[:, [1]] * np.sin(k * s)
            U[:, 1:160:2] += amp[:, [2]] * np.cos(k * s) + amp[:, [3]] * np.sin(k * s)
        U[:, 160:] = rng.randn(ns, 2)
        data["solutions_" + lab] = U
        data["a_" + lab] = 0.01 * rng.randn(ns, 2)
        data["B_" + lab] = np.zeros((ns, 2, 2))
    from deepbnd_b200.vref import COMP_LABEL, COORD_LABEL, vref_order_datasets
    data[COORD_LABEL], data[COMP_LABEL] = vref_order_datasets()
    myhd.savehd5(path, list(data.values()), list(data.keys()), 'w')
    return data


def test_build_inout_pipeline_matches_numpy_restatement(tmp_path, device="cpu"):
    import deepbnd_b200.wrapper_h5py as myhd
    import deepbnd_b200.build_inout as bio
    from oracle import rb_oracle
    from oracle.rve_oracle import vref_points, vref_mass
    snaps, wb, yl, xy, pr = (str(tmp_path / n) for n in ("snapshots.hd5", "Wbasis.hd5", "Y.h5", "XY.hd5", "param.hd5"))
    data = _fake_snapshots(snaps)
    ns = data["a_A"].shape[0]
    myhd.savehd5(pr, np.random.RandomState(1).rand(ns, 36, 5), 'param', 'w')
    Nmax = 60
    bio.translateSolution(snaps)
    bio.computingBasis(snaps, wb, Nmax, device=device)
    bio.extractAlpha(snaps, wb, Nmax, yl, device=device)
    bio.createXY(pr, yl, xy)
    M = vref_mass(vref_points(-1, -1, 2, 2, 21))
    assert np.abs(myhd.loadhd5(wb, 'massMat') - M).max() < 1e-15
    assert not os.path.exists(yl)
    for lab in ("A", "S"):
        It = rb_oracle.translate(data["solutions_" + lab], data["a_" + lab])
        assert np.abs(myhd.loadhd5(snaps, 'solutions_trans_partial_' + lab) - It).max() < 1e-14
        Wo, sig2o = rb_oracle.computingBasis_svd(It, M, Nmax)
        W, sig2 = myhd.loadhd5(wb, 'Wbasis_' + lab), myhd.loadhd5(wb, 'sig_' + lab)
        assert W.shape == (Nmax, 162) and np.allclose(sig2, sig2o[:Nmax], rtol=1e-9)
        # singular vectors are defined up to a sign: compare after aligning signs, for the well-separated modes
        for i in range(20):
            sgn = np.sign(W[i] @ M @ Wo[i])
            assert np.abs(sgn * W[i] - Wo[i]).max() < 1e-7 * np.abs(Wo[i]).max(), (lab, i)
        # M-orthonormality of the basis and the projections
        G = W @ M @ W.T
        assert np.abs(G - np.eye(Nmax)).max() < 1e-8
        Y = myhd.loadhd5(xy, 'Y_' + lab)
        Yo = rb_oracle.getAlphas_fast(W, M, It, Nmax)
        assert Y.shape == (ns, Nmax) and np.abs(Y - Yo).max() < 1e-10 * np.abs(Yo).max()
    assert np.array_equal(myhd.loadhd5(xy, 'X'), myhd.loadhd5(pr, 'param')[:, :, 2])
    # product and oracle agree on the Vref space
    from deepbnd_b200.vref import vref_mass as vm, vref_points as vp
    assert np.array_equal(vp(), vref_points()) and np.abs(vm(vp()) - M).max() < 1e-15


@pytest.mark.gpu
def test_build_inout_pipeline_on_the_gpu(tmp_path):
    import torch
    assert torch.cuda.is_available()
    test_build_inout_pipeline_matches_numpy_restatement(tmp_path, device="cuda")


def test_build_inout_has_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import deepbnd_b200.build_inout as bio
    with pytest.raises(RuntimeError):
        bio.getAlphas_fast((np.eye(4), np.eye(4)), np.ones((2, 4)), 2)
