"""Numpy-facing dataset I/O with the call signatures of the reference's
core/data_manipulation/wrapper_h5py.py (loadhd5 :38-49, savehd5 :64-72, addDataset :75-78,
zeros_openFile :82-96, checkExistenceDataset :159-161, merge :105-122), implemented on the in-tree
HDF5 reader/writer (deepbnd_b200/io/minih5.py) because h5py is not part of this image: reads always go through
minih5; writes use h5py with the reference's exact dataset creation properties when it is importable, else
minih5.Writer — contiguous datasets by default, the reference's chunked (1, ...) + deflate-0 layout
(wrapper_h5py.py:30-34) with DEEPBND_H5_LAYOUT=chunked."""
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
    w = minih5.Writer(tmp, layout=os.environ.get("DEEPBND_H5_LAYOUT", "contiguous"))
    for k, v in items.items():
        w.create_dataset(k, v)
    w.close()
    os.replace(tmp, filename)
    return w.data_addr


class ZerosFile:
    """What zeros_openFile hands back as the file handle: numpy arrays filled in place by the caller,
    written on flush()/close() (the reference flushes per snapshot / per 10 tangents)."""

    def __init__(self, filename, arrays):
        self.filename = filename
        self.arrays = arrays
        self.closed = False
        self._addr = None                  # dataset -> file offset of its (contiguous) data after the first write
        self._rowwise = list(arrays)       # the datasets the caller fills row by row (addDataset adds static ones)

    def flush(self, rows=None):
        """rows=None: (re)write the whole file; rows = indices along axis 0 that changed since the last flush: only
        those rows are written in place (a 100k-snapshot campaign flushes per batch without rewriting the file)"""
        if rows is None or not self._addr or not os.path.exists(self.filename):
            self._addr = _write(self.filename, self.arrays)
            return
        rows = np.unique(np.asarray(rows, dtype=np.int64))
        if len(rows) == 0:
            return
        runs = np.split(rows, np.flatnonzero(np.diff(rows) != 1) + 1)
        with open(self.filename, "r+b") as fh:
            for k in self._rowwise:
                a = self.arrays[k]
                if k not in self._addr or a.ndim == 0 or a.dtype != np.float64:
                    self._addr = _write(self.filename, self.arrays)      # layout not updatable in place
                    return
                rb = a[0:1].nbytes
                for r in runs:
                    fh.seek(self._addr[k] + int(r[0]) * rb)
                    fh.write(np.ascontiguousarray(a[r[0]:r[-1] + 1]).tobytes())

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
