"""The internal-boundary reference mesh Vref: twin of core/mesh/degenerated_rectangle_mesh.py:3-35
(`degeneratedBoundaryRectangleMesh(x0, y0, Lx, Ly, Nb)`: four fan surfaces corner-corner-centre meshed with a huge
characteristic length and Nb transfinite nodes per outer side, so the mesh is 4 (Nb - 1) boundary vertices, the centre
and 4 (Nb - 1) fan triangles).  gmsh is not needed: the mesh is written directly, as the XDMF + HDF5 triplet the
reference's `write(name, 'fenics')` produces (wrapper_io.py:78-134), with the side tags 'side0..3' on the boundary
lines (physicalNaming :33-35).  Vertex order = the library's Vref node order (deepbnd_b200.vref.vref_points)."""
import numpy as np

from ..vref import vref_points
from .mesh_io import write_xdmf_mesh


class degeneratedBoundaryRectangleMesh:
    def __init__(self, x0, y0, Lx, Ly, Nb):
        self.x0, self.y0, self.Lx, self.Ly, self.Nb = x0, y0, Lx, Ly, Nb
        self.points = self.triangles = self.region = self.lines = self.line_tags = None

    def generate(self):
        n = 4 * (self.Nb - 1)
        self.points = vref_points(self.x0, self.y0, self.Lx, self.Ly, self.Nb)
        k = np.arange(n)
        self.triangles = np.stack([k, (k + 1) % n, np.full(n, n)], 1).astype(np.int32)      # CCW fans around the centre
        self.region = np.zeros(n, dtype=np.int32)                                           # one volume tag ('vol')
        self.lines = np.stack([k, (k + 1) % n], 1).astype(np.int32)
        self.line_tags = (k // (self.Nb - 1)).astype(np.int32)                              # side0 .. side3
        return self

    def write(self, savefile, opt='fenics'):
        if self.points is None:
            self.generate()
        if opt != 'fenics':
            raise ValueError("only the 'fenics' (XDMF + HDF5) export exists here")
        write_xdmf_mesh(savefile, self.points, self.triangles, self.region, self.lines, self.line_tags)
