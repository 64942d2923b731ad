import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


from deepbnd_b200.synthetic import synthetic_param, smooth_uD  # noqa: E402,F401  (tests import them from here)


@pytest.fixture(scope="session")
def reduced_meshes():
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    par = synthetic_param(4, seed=13)
    return [buildRVEmesh_arrays(par[i].copy(), size="reduced") for i in range(3)]


@pytest.fixture(scope="session")
def coarse_reduced_meshes():
    """cheap meshes (lcar = 0.1) for the many-system tests"""
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    par = synthetic_param(8, seed=7)
    return [buildRVEmesh_arrays(par[i].copy(), size="reduced", lcar=0.1) for i in range(8)]


@pytest.fixture(scope="session")
def full_mesh():
    from deepbnd_b200.mesh.rve_mesher import buildRVEmesh_arrays
    par = synthetic_param(2, seed=13)
    return buildRVEmesh_arrays(par[0].copy(), size="full", lcar=0.1)
