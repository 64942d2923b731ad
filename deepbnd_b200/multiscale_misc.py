"""Macro-scale consumers of tangents.hd5: numpy twins of the two dolfin UserExpressions that turn the homogenised
tangents into a piecewise-constant elasticity tensor on a macro mesh (core/multiscale/misc.py:9-44).  The reference's
own cook.py / barMultiscale.py read our tangents.hd5 unchanged; these helpers give the same per-cell lookup without
dolfin (cell index -> 3x3 Voigt tangent)."""
import numpy as np


class Chom_multiscale:
    """misc.py:9-20: values[cell] = tangent[mapping[cell]] (3x3, flattened by eval_cell)."""

    def __init__(self, tangent, mapping):
        self.tangent = np.asarray(tangent, dtype=float).reshape(-1, 3, 3)
        self.map = np.asarray(mapping).astype(int)

    def eval_cell(self, cell_index):
        return self.tangent[self.map[cell_index], :, :].flatten()

    def __call__(self, cells=None):
        """(ncells, 3, 3) tangents of all (or the given) macro cells"""
        return self.tangent[self.map if cells is None else self.map[np.asarray(cells)]]

    @staticmethod
    def value_shape():
        return (3, 3)


class Chom_precomputed_dist(Chom_multiscale):
    """misc.py:25-44: every macro cell takes the tangent of the RVE whose `center` (tangents.hd5 / paramRVEdataset.hd5
    dataset 'center') is closest to the cell midpoint."""

    def __init__(self, tangent, center, points, cells):
        mid = np.asarray(points, dtype=float)[np.asarray(cells)].mean(axis=1)[:, :2]
        center = np.asarray(center, dtype=float)
        d = np.linalg.norm(mid[:, None, :] - center[None, :, :], axis=2)
        super().__init__(tangent, d.argmin(axis=1))
