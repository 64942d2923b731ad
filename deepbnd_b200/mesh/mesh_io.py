"""Mesh files of the drop-in path.

* `read_xdmf_mesh`: reads the triplet the reference writes with meshio
  (core/fenics_tools/wrapper_io.py:78-134 `exportMeshHDF5_fromGMSH`: `mesh.xdmf` + `mesh.h5` with
  `data0` = points, `data1` = triangles, and `mesh_regions.xdmf` + `mesh_regions.h5` holding the
  per-triangle physical tags; `mesh_faces.*` is not needed by the RVE models) without meshio/h5py:
  the XML is parsed for its `DataItem` references and the heavy data is read with minih5.
* `write_xdmf_mesh`: the same triplet, written (for meshes that must leave this package: the boundary reference
  mesh `boundaryMesh.xdmf`, RVE meshes handed to the reference's FEniCS scripts).
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


def _xdmf(path, h5name, npts, ncell, cell, ncol, attr=None):
    a = ""
    if attr:
        a = ('<Attribute AttributeType="Scalar" Center="Cell" Name="%s"><DataItem DataType="Int" Dimensions="%d" Format="HDF" '
             'Precision="8">%s:/data2</DataItem></Attribute>' % (attr, ncell, h5name))
    with open(path, "w") as f:
        f.write('<Xdmf Version="3.0"><Domain><Grid Name="Grid"><Geometry GeometryType="XY"><DataItem DataType="Float" '
                'Dimensions="%d 2" Format="HDF" Precision="8">%s:/data0</DataItem></Geometry><Topology NumberOfElements="%d" '
                'TopologyType="%s"><DataItem DataType="Int" Dimensions="%d %d" Format="HDF" Precision="8">%s:/data1</DataItem>'
                '</Topology>%s</Grid></Domain></Xdmf>\n' % (npts, h5name, ncell, cell, ncell, ncol, h5name, a))


def write_xdmf_mesh(meshFile, points, triangles, region, lines=None, line_tags=None):
    """The triplet exportMeshHDF5_fromGMSH leaves on disk (wrapper_io.py:78-134), readable by readXDMF_with_markers
    (:31-49): `mesh.xdmf/.h5` (data0 points, data1 triangles), `mesh_regions.xdmf/.h5` (data1 triangles, data2 the
    physical tag per triangle, named 'regions') and `mesh_faces.xdmf/.h5` (data1 boundary lines, data2 their tags,
    named 'faces').  The reference replaces data1 of the regions file by an ExternalLink to save space; here it is a
    plain copy (old-style groups cannot hold link messages, and every reader resolves both the same way)."""
    rad = meshFile[:-5] if meshFile.endswith(".xdmf") else meshFile
    os.makedirs(os.path.dirname(os.path.abspath(rad)), exist_ok=True)
    base = os.path.basename(rad)
    pts = np.asarray(points, dtype=np.float64)[:, :2]
    tri = np.asarray(triangles, dtype=np.int64)
    with minih5.Writer(rad + ".h5") as w:
        w["data0"] = pts; w["data1"] = tri
    _xdmf(rad + ".xdmf", base + ".h5", len(pts), len(tri), "Triangle", 3)
    with minih5.Writer(rad + "_regions.h5") as w:
        w["data0"] = np.zeros((1, 2)); w["data1"] = tri; w["data2"] = np.asarray(region, dtype=np.int64).ravel()
    _xdmf(rad + "_regions.xdmf", base + "_regions.h5", 1, len(tri), "Triangle", 3, "regions")
    if lines is None:                                    # boundary edges of the triangulation, one tag
        e = np.vstack([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
        es = np.sort(e, axis=1)
        key, idx, cnt = np.unique(es[:, 0] * (len(pts) + 1) + es[:, 1], return_index=True, return_counts=True)
        lines = e[idx[cnt == 1]]
        line_tags = np.full(len(lines), int(np.max(region)) + 1)
    lines = np.asarray(lines, dtype=np.int64)
    with minih5.Writer(rad + "_faces.h5") as w:
        w["data0"] = np.zeros((1, 2)); w["data1"] = lines; w["data2"] = np.asarray(line_tags, dtype=np.int64).ravel()
    _xdmf(rad + "_faces.xdmf", base + "_faces.h5", 1, len(lines), "Polyline", 2, "faces")


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
