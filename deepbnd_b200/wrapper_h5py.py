"""Numpy-facing dataset I/O with the call signatures of the reference's
core/data_manipulation/wrapper_h5py.py (loadhd5 :38-49, savehd5 :64-72, addDataset :75-78,
zeros_openFile :82-96, checkExistenceDataset :159-161, merge :105-122), implemented on the in-tree
HDF5 reader/writer (deepbnd_b200/io/minih5.py) because h5py is not part of this image.  When h5py
is importable it is used instead, with the reference's exact dataset creation properties."""
import os

import numpy as np

from .io import minih5

try:                                    # pragma: no cover - not available in the build image
    import h5py as _h5py
except Exception:                       # noqa: BLE001
    _h5py = None

defaultCompression = {'dtype': 'f8', 'compression': "gzip", 'compression_opts': 0, 'shuffle': False,
                      'chunks': (5, 100)}
toChunk = lambda a: tuple([1] + [a[i] for i in range(1, len(a))])


def loadhd5(filename, label):
    with minih5.File(filename) as f:
        if isinstance(label, str):
            return np.array(f[label])
        return [np.array(f[l]) for l in label]


def checkExistenceDataset(filename, label):
    with minih5.File(filename) as f:
        return label in f


def _existing(filename):
    if not os.path.exists(filename):
        return {}
    with minih5.File(filename) as f:
        return {k: np.array(f[k]) for k in f.keys()}


def savehd5(filename, X, label, mode):
    if mode in ('w-', 'x') and os.path.exists(filename):
        raise OSError("Unable to create file (file exists): %s" % filename)
    items = _existing(filename) if mode in ('a', 'r+') else {}
    if isinstance(label, str):
        items[label] = np.asarray(X, dtype=np.float64)
    else:
        for x, l in zip(X, label):
            items[l] = np.asarray(x, dtype=np.float64)
    _write(filename, items)


def addDataset(f, X, label):
    if isinstance(f, ZerosFile):
        f.arrays[label] = np.asarray(X, dtype=np.float64)
        return
    items = _existing(f)
    if label in items:
        raise ValueError("Unable to create dataset (name already exists)")
    items[label] = np.asarray(X, dtype=np.float64)
    _write(f, items)


def _write(filename, items):
    if _h5py is not None:               # pragma: no cover
        with _h5py.File(filename, 'w') as f:
            for k, v in items.items():
                v = np.asarray(v)
                if v.ndim >= 1 and v.size:
                    f.create_dataset(k, data=v, dtype='f8', compression='gzip', compression_opts=0, shuffle=False,
                                     chunks=toChunk(v.shape))
                else:
                    f.create_dataset(k, data=v)
        return
    tmp = filename + ".tmp%d" % os.getpid()
    with minih5.Writer(tmp) as w:
        for k, v in items.items():
            w.create_dataset(k, v)
    os.replace(tmp, filename)


class ZerosFile:
    """What zeros_openFile hands back as the file handle: numpy arrays filled in place by the caller,
    written on flush()/close() (the reference flushes per snapshot / per 10 tangents)."""

    def __init__(self, filename, arrays):
        self.filename = filename
        self.arrays = arrays
        self.closed = False

    def flush(self):
        _write(self.filename, self.arrays)

    def close(self):
        if not self.closed:
            self.flush()
            self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def zeros_openFile(filename, shape, label, mode='w-'):
    if mode in ('w-', 'x') and os.path.exists(filename):
        raise OSError("Unable to create file (file exists): %s" % filename)
    if isinstance(label, str):
        arrays = {label: np.zeros(shape)}
        f = ZerosFile(filename, arrays)
        return arrays[label], f
    arrays = {l: np.zeros(s) for s, l in zip(shape, label)}
    return [arrays[l] for l in label], ZerosFile(filename, arrays)


def merge(filenameInputs, filenameOutput, InputLabels, OutputLabels, axis=0, mode='w-'):
    if mode in ('w-', 'x') and os.path.exists(filenameOutput):
        raise OSError("Unable to create file (file exists): %s" % filenameOutput)
    items = _existing(filenameOutput) if mode in ('a', 'r+') else {}
    for li, lo in zip(InputLabels, OutputLabels):
        items[lo] = np.concatenate([loadhd5(fn, li) for fn in filenameInputs], axis=axis)
    _write(filenameOutput, items)
