"""Seeded synthetic inputs of the benchmark / smoke / example workloads: paramRVEdataset rows like
build_snapshots_param.py:21-41 writes them and smooth boundary fields standing in for bcs.hd5."""
import numpy as np


def synthetic_param(ns, seed=13):
    """Synthetic paramRVEdataset rows: 6x6 unit grid in [-3,3]^2, e=1, theta=0,
    radii exp(mu + sig*theta), theta ~ centred LHS in [-1,1] per column (build_snapshots_param.py:21-41)."""
    rng = np.random.RandomState(seed)
    cx = np.linspace(-2.5, 2.5, 6)
    X, Y = np.meshgrid(cx, cx)
    p = np.zeros((ns, 36, 5))
    p[:, :, 0] = X.ravel()
    p[:, :, 1] = Y.ravel()
    p[:, :, 3] = 1.0
    r0, r1 = 0.2, 0.4
    mu, sig = 0.5 * np.log(r1 * r0), 0.5 * np.log(r1 / r0)
    for j in range(36):
        theta = -1.0 + 2.0 * (rng.permutation(ns) + 0.5) / ns
        p[:, j, 2] = np.exp(mu + sig * theta)
    return p


def smooth_uD(nvref, seed, amp=1e-2):
    """small smooth random boundary field on Vref (our node order), SURVEY 8d C3"""
    rng = np.random.RandomState(seed)
    nb = nvref - 1
    s = 2 * np.pi * np.arange(nb) / nb
    u = np.zeros((nvref, 2))
    for c in range(2):
        for k in range(1, 4):
            u[:nb, c] += amp * rng.randn() * np.cos(k * s) + amp * rng.randn() * np.sin(k * s)
    return u.ravel()
