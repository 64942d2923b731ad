#!/usr/bin/env python3
"""Runs where TensorFlow / h5py exist (NOT in the B200 image): converts the reference's Keras weight files
(`model_weights_A.hdf5`, `model_weights_S.hdf5`, written by NetArch.training through ModelCheckpoint,
creation_model/training/net_arch.py:95-140) into the flat layout deepbnd_b200.bc_prediction.NetArch.load_weights
reads: datasets W0, b0, W1, b1, ... (Dense kernels (n_in, n_out) and biases in layer order).

    python tools/export_keras_weights.py model_weights_A.hdf5 weights_A.npz
"""
import sys

import numpy as np


def dense_layers(h5file):
    """(kernel, bias) of every Dense layer in model order, read straight from the HDF5 groups Keras writes."""
    import h5py
    with h5py.File(h5file, 'r') as f:
        g = f['model_weights'] if 'model_weights' in f else f
        names = [n.decode() if isinstance(n, bytes) else n for n in g.attrs['layer_names']]
        out = []
        for name in names:
            wn = [n.decode() if isinstance(n, bytes) else n for n in g[name].attrs['weight_names']]
            if not wn:
                continue                                   # Dropout / GaussianDropout / Input carry no weights
            ker = [n for n in wn if 'kernel' in n]
            bia = [n for n in wn if 'bias' in n]
            out.append((np.array(g[name][ker[0]]), np.array(g[name][bia[0]])))
    return out


if __name__ == '__main__':
    src, dst = sys.argv[1], sys.argv[2]
    layers = dense_layers(src)
    np.savez(dst, **{('W%d' % i): W for i, (W, b) in enumerate(layers)}, **{('b%d' % i): b for i, (W, b) in enumerate(layers)})
    print("wrote", dst, [W.shape for W, _ in layers])
