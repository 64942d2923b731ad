"""Drop-in for creation_model/prediction/bc_prediction.py + NN_elast.py (reference bc_prediction.py:32-50,
NN_elast.py:21-62), the producer of `bcs.hd5` on the near side of the tangent batch: the trained
MLPs map the 36 inclusion radii of an RVE to reduced-basis coefficients of the boundary fluctuation for the axial
load ('A') and the shear load ('S'); `u0 = Y_A W_A`, `u2 = Y_S W_S`, and `u1` (axial load in y) is the 'A' network
evaluated on the radii of the RVE rotated by a quarter turn, rotated back by the Piola transform (NN_elast.py:48-55).

What changes against the reference:
  * inference only, through torch (library calls; this stage is not part of the hand-written hot path).  The device
    is "cuda" unless the caller passes "cpu" explicitly — no silent fallback (same rule as build_inout.py);
  * Keras is not needed: weights come from a flat file `W0, b0, W1, b1, ...` (`.npz` or `.hd5`, fp32/fp64,
    `Wk` of shape (n_in, n_out) like a Keras Dense kernel); `tools/export_keras_weights.py` writes it from the
    reference's `model_weights_X.hdf5` where TensorFlow exists.  Dropout / GaussianDropout layers of the reference
    architecture (net_arch.py:43-61) are the identity at inference;
  * the Piola rotation is built in closed form on the library's Vref node order (deepbnd_b200/vref.py): for the
    square Vref centred at the origin a quarter turn maps nodes onto nodes, so the 162 x 162 matrix the reference
    assembles column by column with dolfin interpolations (fenics_utils.py:56-82) is a signed permutation."""
import os

import numpy as np

from . import wrapper_h5py as myhd
from .simulation_snapshots import vref_box_from_mesh
from .vref import COMP_LABEL, COORD_LABEL, load_vref_vectors, vref_order_datasets, vref_points

DEVICE = "cuda"


def _dev(device=None):
    import torch
    d = torch.device(device or DEVICE)
    if d.type == "cuda" and not torch.cuda.is_available():
        raise RuntimeError("deepbnd_b200.bc_prediction: no CUDA device (pass device='cpu' explicitly to run the inference on the host)")
    return d


# ---- scalers (core/data_manipulation/utils.py:38-201) -------------------------------------------------------------
class myMinMaxScaler:
    """x -> (x - min) / (max - min) per column (utils.py:62-90)."""

    def load_param(self, param):
        self.a, self.b = np.asarray(param[:, 0], float), np.asarray(param[:, 1], float)
        self.n = len(self.a)

    def set_n(self, n):
        self.n, self.a, self.b = n, self.a[:n], self.b[:n]

    def transform(self, x):
        return (np.asarray(x, float)[:, :self.n] - self.a) / (self.b - self.a)

    def inverse_transform(self, x):
        return (self.b - self.a) * np.asarray(x, float)[:, :self.n] + self.a


class myNormalisationScaler(myMinMaxScaler):
    """x -> (x - mean) / std per column (utils.py:114-138); the two file columns are (mean, std)."""

    def transform(self, x):
        return (np.asarray(x, float)[:, :self.n] - self.a) / self.b

    def inverse_transform(self, x):
        return np.asarray(x, float)[:, :self.n] * self.b + self.a


def importScale(filenameIn, nX, nY, scalerType='MinMax'):
    """utils.py:192-201: text file with columns (X param 0, X param 1, Y param 0, Y param 1), max(nX, nY) rows."""
    cls = {'MinMax': myMinMaxScaler, 'Normalisation': myNormalisationScaler}[scalerType]
    tab = np.loadtxt(filenameIn, ndmin=2)
    scalerX, scalerY = cls(), cls()
    scalerX.load_param(tab[:, 0:2]); scalerY.load_param(tab[:, 2:4])
    scalerX.set_n(nX); scalerY.set_n(nY)
    return scalerX, scalerY


# ---- network (creation_model/training/net_arch.py:21-61, wrapper_tensorflow.py:18-24) -----------------------------
def _activation(name):
    import torch
    import torch.nn.functional as F
    return {'tanh': torch.tanh, 'sigmoid': torch.sigmoid, 'linear': lambda x: x, 'relu': torch.relu,
            'leaky_relu': lambda x: F.leaky_relu(x, 0.2),            # tf.nn.leaky_relu default alpha
            'swish': lambda x: x * torch.sigmoid(x), 'elu': F.elu}[name]


class NetArch:
    """Inference view of the reference's NetArch: `Neurons` hidden widths, `activations` one per Dense layer
    (hidden + output), `nX` inputs, `nY` outputs, `files['weights']`, `files['scaler']`."""

    def __init__(self, Neurons, activations, lr=5.0e-4, decay=1.0, drps=0.0, reg=0.0):
        self.Neurons, self.activations = list(Neurons), list(activations)
        self.lr, self.decay, self.drps, self.reg = lr, decay, drps, reg
        self.nY = self.nX = self.archId = None
        self.files = {'weights': None, 'net_settings': None, 'scaler': None, 'XY': None, 'XY_val': None}

    def layer_shapes(self):
        widths = [self.nX] + self.Neurons + [self.nY]
        return [(widths[i], widths[i + 1]) for i in range(len(widths) - 1)]

    def load_weights(self, filename=None):
        """flat `W0, b0, W1, b1, ...` from .npz or .hd5; shapes are checked against the architecture."""
        filename = filename or self.files['weights']
        shapes = self.layer_shapes()
        if filename.endswith('.npz'):
            z = np.load(filename)
            arrs = [(z['W%d' % i], z['b%d' % i]) for i in range(len(shapes))]
        else:
            arrs = [tuple(myhd.loadhd5(filename, ['W%d' % i, 'b%d' % i])) for i in range(len(shapes))]
        for i, ((W, b), shp) in enumerate(zip(arrs, shapes)):
            if W.shape != shp or b.shape != (shp[1],):
                raise ValueError("layer %d: weights %s / %s do not match the architecture %s" % (i, W.shape, b.shape, shp))
        return [(np.asarray(W, float), np.asarray(b, float)) for W, b in arrs]

    def getModel(self, weights=None, device=None):
        """callable X (n, nX) -> Y (n, nY), fp64"""
        import torch
        dev = _dev(device)
        weights = weights if weights is not None else self.load_weights()
        Ws = [(torch.from_numpy(W).to(dev), torch.from_numpy(b).to(dev)) for W, b in weights]
        acts = [_activation(a) for a in self.activations]
        if len(acts) != len(Ws):
            raise ValueError("one activation per Dense layer expected")

        def predict(X):
            with torch.no_grad():
                x = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
                for (W, b), act in zip(Ws, acts):
                    x = act(x @ W + b)
                return x.cpu().numpy()
        return predict


standardNets = {'small': NetArch([40, 40, 40], 3 * ['swish'] + ['linear'], 5.0e-4, 0.8, [0.0] + 3 * [0.001] + [0.0], 1.0e-8),
                'big': NetArch([300, 300, 300], 3 * ['swish'] + ['linear'], 5.0e-4, 0.1, [0.0] + 3 * [0.005] + [0.0], 1.0e-8)}


# ---- Piola transform of a rotation on Vref (core/elasticity/fenics_utils.py:56-82) --------------------------------
def PiolaTransform_rotation_matricial(theta, nameMeshRefBnd=None, box=None, nb=None):
    """Matrix P of u -> (x -> B u(B^T x)), B the rotation by theta, in the library's Vref DOF order.  theta must be
    a multiple of pi/2 and Vref centred at the origin (the reference's [-1,1]^2), so that B^T maps nodes to nodes."""
    if box is None:
        box, nb = vref_box_from_mesh(nameMeshRefBnd)
    pts = vref_points(*box, nb)
    s, c = np.sin(theta), np.cos(theta)
    B = np.array([[c, -s], [s, c]])
    src = pts @ B                                     # rows: B^T x_k
    d2 = ((src[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    perm = d2.argmin(1)
    if d2[np.arange(len(pts)), perm].max() > 1e-20 * max(box[2], box[3]) ** 2 + 1e-24:
        raise ValueError("the rotation does not map the Vref nodes onto themselves (theta must be a multiple of pi/2, "
                         "Vref a square centred at the origin)")
    nh = 2 * len(pts)
    P = np.zeros((nh, nh))
    for k, kp in enumerate(perm):
        P[2 * k:2 * k + 2, 2 * kp:2 * kp + 2] = B
    return P


class NNElast:
    """NN_elast.py:21-62: one network + scalers + basis per load label ('A', 'S')."""

    def __init__(self, nameWbasis, net, Nrb, scalertype='MinMax', device=None):
        self.Nrb = Nrb
        self.labels = list(net.keys())
        self.Wbasis, self.scalerX, self.scalerY, self.model = {}, {}, {}, {}
        for l in self.labels:
            self.Wbasis[l] = load_vref_vectors(nameWbasis, 'Wbasis_%s' % l)     # library Vref DOF order (deepbnd_b200/vref.py)
            self.scalerX[l], self.scalerY[l] = importScale(net[l].files['scaler'], net[l].nX, net[l].nY, scalertype)
            self.model[l] = net[l].getModel(device=device)

    def predict(self, X, nameMeshRefBnd=None, box=None, nb=None):
        X = np.asarray(X, float)
        S_p = {}
        for l in self.labels:
            Y_p = self.scalerY[l].inverse_transform(self.model[l](self.scalerX[l].transform(X)))
            S_p[l] = Y_p @ self.Wbasis[l][:self.Nrb, :]
        X_axialY_s = self.scalerX['A'].transform(X[:, self.getPermY()])     # permY performs a counterclockwise rotation
        Y_p_axialY = self.scalerY['A'].inverse_transform(self.model['A'](X_axialY_s))
        theta = 3 * np.pi / 2.0                                               # 'minus HalfPi'
        piola_mat = PiolaTransform_rotation_matricial(theta, nameMeshRefBnd, box, nb)
        S_p_axialY = Y_p_axialY @ self.Wbasis['A'][:self.Nrb, :] @ piola_mat.T
        return [S_p['A'], S_p_axialY, S_p['S']]

    @staticmethod
    def getPermY(n=6):
        """radii ordered by rows and columns (bottom to top, left to right): (i, j) -> (Nx - j - 1, i)"""
        return np.array([[(n - 1 - j) * n + i for j in range(n)] for i in range(n)]).flatten()


def predictBCs(namefiles, net, device=None):
    """bc_prediction.py:32-50: paramRVEdataset.hd5:param[:, :, 2] (radii) -> bcs.hd5: u0, u1, u2 (ns, 162)."""
    nameMeshRefBnd, nameWbasis, paramRVEname, bcs_namefile = namefiles
    paramRVEdata = myhd.loadhd5(paramRVEname, 'param')
    model = NNElast(nameWbasis, net, net['A'].nY, device=device)
    S_p = model.predict(paramRVEdata[:, :, 2], nameMeshRefBnd)
    if os.path.exists(bcs_namefile):
        os.remove(bcs_namefile)
    coords, comps = vref_order_datasets(*vref_box_from_mesh(nameMeshRefBnd))
    myhd.savehd5(bcs_namefile, S_p + [coords, comps], ['u0', 'u1', 'u2', COORD_LABEL, COMP_LABEL], mode='w')
    return S_p
