"""Mesh files of the drop-in path.

* `read_xdmf_mesh`: reads the triplet the reference writes with meshio
  (core/fenics_tools/wrapper_io.py:78-134 `exportMeshHDF5_fromGMSH`: `mesh.xdmf` + `mesh.h5` with
  `data0` = points, `data1` = triangles, and `mesh_regions.xdmf` + `mesh_regions.h5` holding the
  per-triangle physical tags; `mesh_faces.*` is not needed by the RVE models) without meshio/h5py:
  the XML is parsed for its `DataItem` references and the heavy data is read with minih5.
* `save_mesh_cache` / `load_mesh_cache`: the array triple our own mesher produces, cached next to the
  name the reference would have used ("generated once and cached", BASELINE.json north_star).
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from ..io import minih5


def _dataitems(xdmf_path):
    root = ET.parse(xdmf_path).getroot()
    out = []
    for it in root.iter("DataItem"):
        txt = (it.text or "").strip()
        if it.attrib.get("Format", "").upper() in ("HDF", "HDF5") and ":" in txt:
            fn, ds = txt.split(":", 1)
            out.append((os.path.join(os.path.dirname(xdmf_path), fn), ds, it.attrib.get("Dimensions", "")))
    return out


def read_xdmf_points(meshFile):
    """vertex coordinates (nv, 2) of an XDMF mesh; no cell markers needed (the boundary reference mesh Vref)"""
    for fn, ds, dims in _dataitems(meshFile):
        with minih5.File(fn) as f:
            a = np.array(f[ds])
        if a.dtype.kind == "f" and a.ndim == 2:
            return np.ascontiguousarray(a[:, :2].astype(np.float64))
    raise ValueError("read_xdmf_points: no coordinate array in %s" % meshFile)


def read_xdmf_mesh(meshFile):
    """-> (points (nv,2) f8, triangles (nt,3) i4, region (nt,) i4)."""
    items = _dataitems(meshFile)
    pts = tri = None
    for fn, ds, dims in items:
        with minih5.File(fn) as f:
            a = np.array(f[ds])
        if a.dtype.kind == "f" and pts is None:
            pts = a[:, :2].astype(np.float64)
        elif a.dtype.kind in "iu" and a.ndim == 2 and a.shape[1] == 3 and tri is None:
            tri = a.astype(np.int32)
    reg = None
    rfile = meshFile[:-5] + "_regions.xdmf"
    if os.path.exists(rfile):
        for fn, ds, dims in _dataitems(rfile):
            try:
                with minih5.File(fn) as f:
                    a = np.array(f[ds])           # data1 of *_regions.h5 is an ExternalLink back to mesh.h5 (resolved)
            except KeyError:
                continue                      # link target missing: the geometry is read from mesh.h5 anyway
            if a.dtype.kind in "iu" and (a.ndim == 1 or (a.ndim == 2 and a.shape[1] == 1)) and tri is not None \
                    and a.shape[0] == tri.shape[0]:
                reg = a.reshape(-1).astype(np.int32)
    if pts is None or tri is None or reg is None:
        raise ValueError("read_xdmf_mesh: could not find points/triangles/regions in %s" % meshFile)
    return np.ascontiguousarray(pts), np.ascontiguousarray(tri), np.ascontiguousarray(reg)


def cache_name(meshname):
    return (meshname[:-5] if meshname.endswith(".xdmf") else meshname) + ".npz"


def save_mesh_cache(meshname, mesh):
    os.makedirs(os.path.dirname(os.path.abspath(meshname)), exist_ok=True)
    np.savez(cache_name(meshname), points=mesh[0], triangles=mesh[1], region=mesh[2])


def load_mesh(meshname):
    """cached array triple if present, else the reference's XDMF triplet."""
    cn = cache_name(meshname)
    if os.path.exists(cn):
        z = np.load(cn)
        return z["points"], z["triangles"], z["region"]
    if os.path.exists(meshname) and meshname.endswith(".xdmf"):
        return read_xdmf_mesh(meshname)
    raise FileNotFoundError("no cached mesh %s and no XDMF mesh %s (call with createMesh=True)" % (cn, meshname))
