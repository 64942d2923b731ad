"""deepbnd_b200.bc_prediction (reference creation_model/prediction/bc_prediction.py, NN_elast.py):
scalers, MLP inference against a numpy restatement, the Piola rotation on the library's Vref order, and the
bcs.hd5 file the tangent driver reads.  CPU only (device='cpu' passed explicitly)."""
import numpy as np
import pytest

from deepbnd_b200 import bc_prediction as bp
from deepbnd_b200 import wrapper_h5py as myhd
from deepbnd_b200.vref import vref_points


def _np_forward(weights, acts, X):
    f = {'swish': lambda x: x / (1 + np.exp(-x)), 'linear': lambda x: x, 'tanh': np.tanh, 'relu': lambda x: np.maximum(x, 0),
         'sigmoid': lambda x: 1 / (1 + np.exp(-x)), 'leaky_relu': lambda x: np.where(x > 0, x, 0.2 * x),
         'elu': lambda x: np.where(x > 0, x, np.exp(np.minimum(x, 0)) - 1)}
    x = X
    for (W, b), a in zip(weights, acts):
        x = f[a](x @ W + b)
    return x


def _random_net(rs, nX, nY, tmp_path, label, fmt):
    net = bp.NetArch([17, 11], ['swish', 'tanh', 'linear'])
    net.nX, net.nY = nX, nY
    weights = [(rs.randn(*s) * 0.4, rs.randn(s[1]) * 0.1) for s in net.layer_shapes()]
    wfile = str(tmp_path / ("weights_%s.%s" % (label, fmt)))
    if fmt == 'npz':
        np.savez(wfile, **{'W%d' % i: W for i, (W, b) in enumerate(weights)}, **{'b%d' % i: b for i, (W, b) in enumerate(weights)})
    else:
        myhd.savehd5(wfile, [a for Wb in weights for a in Wb], [n % i for i in range(len(weights)) for n in ('W%d', 'b%d')], mode='w')
    sfile = str(tmp_path / ("scaler_%s.txt" % label))
    tab = np.zeros((max(nX, nY), 4))
    tab[:nX, 0], tab[:nX, 1] = 0.2, 0.4                                     # radii range
    tab[:nY, 2], tab[:nY, 3] = -1.0 - rs.rand(nY), 1.0 + rs.rand(nY)
    np.savetxt(sfile, tab)
    net.files['weights'], net.files['scaler'] = wfile, sfile
    return net, weights, tab


def test_scalers_follow_the_reference_formulas(tmp_path):
    tab = np.array([[0.0, 2.0, -1.0, 3.0], [1.0, 5.0, 0.0, 0.5], [9.0, 9.5, 0.0, 0.0]])
    f = str(tmp_path / "s.txt")
    np.savetxt(f, tab)
    sx, sy = bp.importScale(f, 3, 2, 'MinMax')
    X = np.array([[1.0, 3.0, 9.25]])
    assert np.allclose(sx.transform(X), [[0.5, 0.5, 0.5]])
    assert np.allclose(sy.inverse_transform(np.array([[0.5, 0.5]])), [[1.0, 0.25]])
    assert np.allclose(sx.inverse_transform(sx.transform(X)), X)
    nx, ny = bp.importScale(f, 2, 2, 'Normalisation')                        # columns are (mean, std)
    assert np.allclose(nx.transform(np.array([[2.0, 6.0]])), [[1.0, 1.0]])
    assert np.allclose(ny.inverse_transform(np.array([[1.0, 2.0]])), [[2.0, 1.0]])


@pytest.mark.parametrize("fmt", ["npz", "hd5"])
def test_mlp_inference_matches_numpy(tmp_path, fmt):
    rs = np.random.RandomState(3)
    net, weights, _ = _random_net(rs, 36, 20, tmp_path, 'A', fmt)
    X = rs.rand(50, 36)
    Y = net.getModel(device='cpu')(X)
    assert Y.shape == (50, 20)
    assert np.abs(Y - _np_forward(weights, net.activations, X)).max() < 1e-12
    bad = bp.NetArch([17, 12], ['swish', 'tanh', 'linear'])                  # architecture / file mismatch is an error
    bad.nX, bad.nY, bad.files['weights'] = 36, 20, net.files['weights']
    with pytest.raises(ValueError):
        bad.load_weights()


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    net = bp.NetArch([4], ['tanh', 'linear'])
    net.nX, net.nY = 3, 2
    with pytest.raises(RuntimeError, match="no CUDA device"):
        net.getModel(weights=[(np.zeros((3, 4)), np.zeros(4)), (np.zeros((4, 2)), np.zeros(2))])


def test_piola_quarter_turn_is_a_signed_permutation():
    theta = 3 * np.pi / 2
    P = bp.PiolaTransform_rotation_matricial(theta)
    nh = P.shape[0]
    assert nh == 162
    assert np.allclose(P @ P.T, np.eye(nh)) and np.allclose(np.linalg.matrix_power(P, 4), np.eye(nh))
    # a linear field u(x) = G x goes to B G B^T x (the reference's Piola(sol)(x) = B sol(B^T x), fenics_utils.py:56-65)
    pts = vref_points()
    s, c = np.sin(theta), np.cos(theta)
    B = np.array([[c, -s], [s, c]])
    G = np.array([[0.3, -1.2], [0.7, 0.4]])
    u = (pts @ G.T).ravel()
    assert np.allclose(P @ u, (pts @ (B @ G @ B.T).T).ravel())
    # an arbitrary nodal field: value at x_k is B times the value at the node B^T x_k
    rs = np.random.RandomState(0)
    v = rs.randn(nh)
    w = (P @ v).reshape(-1, 2)
    for k in (0, 7, 33, 79, 80):
        kp = np.argmin(((pts - pts[k] @ B) ** 2).sum(1))
        assert np.allclose(w[k], B @ v.reshape(-1, 2)[kp])
    with pytest.raises(ValueError):
        bp.PiolaTransform_rotation_matricial(0.3)


def test_permY_is_a_quarter_turn_of_the_6x6_grid():
    p = bp.NNElast.getPermY()
    assert sorted(p) == list(range(36))
    assert np.array_equal(p[p[p[p]]], np.arange(36)) and not np.array_equal(p[p], np.arange(36))
    assert p[0] == 30 and p[5] == 0 and p[35] == 5                           # (i, j) -> (Nx - j - 1, i)


def test_predictBCs_writes_bcs_file(tmp_path):
    rs = np.random.RandomState(5)
    nrb, ns = 12, 9
    net = {}
    keep = {}
    for l in ('A', 'S'):
        net[l], w, tab = _random_net(rs, 36, nrb, tmp_path, l, 'npz')
        keep[l] = (w, tab)
    Wb = {l: rs.randn(20, 162) for l in ('A', 'S')}
    nameW = str(tmp_path / "Wbasis.hd5")
    from deepbnd_b200.vref import COMP_LABEL, COORD_LABEL, vref_order_datasets
    coords, comps = vref_order_datasets()
    myhd.savehd5(nameW, [Wb['A'], Wb['S'], coords, comps], ['Wbasis_A', 'Wbasis_S', COORD_LABEL, COMP_LABEL], mode='w')
    param = np.zeros((ns, 36, 5))
    param[:, :, 2] = 0.2 + 0.2 * rs.rand(ns, 36)
    nameP = str(tmp_path / "paramRVEdataset.hd5")
    myhd.savehd5(nameP, param, 'param', mode='w')
    out = str(tmp_path / "bcs.hd5")
    bp.predictBCs([None, nameW, nameP, out], net, device='cpu')
    u0, u1, u2 = myhd.loadhd5(out, ['u0', 'u1', 'u2'])
    assert u0.shape == u1.shape == u2.shape == (ns, 162)

    def direct(l, X):
        w, tab = keep[l]
        Xs = (X - tab[:36, 0]) / (tab[:36, 1] - tab[:36, 0])
        Y = _np_forward(w, net[l].activations, Xs) * (tab[:nrb, 3] - tab[:nrb, 2]) + tab[:nrb, 2]
        return Y @ Wb[l][:nrb]
    X = param[:, :, 2]
    assert np.allclose(u0, direct('A', X)) and np.allclose(u2, direct('S', X))
    P = bp.PiolaTransform_rotation_matricial(3 * np.pi / 2)
    assert np.allclose(u1, direct('A', X[:, bp.NNElast.getPermY()]) @ P.T)
