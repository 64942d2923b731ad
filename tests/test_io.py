"""CPU tests of the dataset I/O layer (the reference's wrapper_h5py.py call surface on the in-tree
HDF5 reader/writer) and of the mesh cache."""
import os
import struct

import numpy as np
import pytest

import deepbnd_b200.wrapper_h5py as myhd

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reader_on_a_file_written_by_libhdf5():
    """tests/golden/libhdf5_written_testdouble.mat is scipy's MATLAB-v7.3 test file (BSD): written by
    libhdf5 with a 512-byte user block, superblock v0, old-style root group, one f8 dataset."""
    from deepbnd_b200.io import minih5
    f = minih5.File(os.path.join(GOLD, "libhdf5_written_testdouble.mat"))
    assert f.keys() == ["testdouble"]
    a = np.array(f["testdouble"])
    assert a.shape == (9, 1) and a.dtype == np.float64
    assert np.allclose(a.ravel(), np.arange(9) * np.pi / 4)


def test_writer_reader_roundtrip_many_datasets(tmp_path):
    from deepbnd_b200.io import minih5
    rng = np.random.RandomState(0)
    data = {"d%02d" % i: rng.randn(4, 3, i + 1) for i in range(21)}       # > 2 symbol-table nodes
    data.update(id=np.arange(7, dtype=np.int64), i32=np.arange(6, dtype=np.int32).reshape(2, 3),
                scalar=np.float64(2.5), empty=np.zeros((0, 3)), f32=rng.randn(5).astype(np.float32))
    fn = str(tmp_path / "t.hd5")
    with minih5.Writer(fn) as w:
        for k, v in data.items():
            w.create_dataset(k, v)
    raw = open(fn, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0            # superblock v0
    g = minih5.File(fn)
    assert sorted(g.keys()) == sorted(data)
    for k, v in data.items():
        a = np.array(g[k])
        assert a.shape == np.shape(v) and a.dtype == np.asarray(v).dtype and np.array_equal(a, v), k
    assert "nope" not in g
    with pytest.raises(KeyError):
        g["nope"]


def test_wrapper_h5py_call_surface(tmp_path):
    import deepbnd_b200.wrapper_h5py as myhd
    fn = str(tmp_path / "snap.hd5")
    X, f = myhd.zeros_openFile(fn, [(3,), (3, 162), (3, 2, 2)], ['id', 'solutions_S', 'B_S'], mode='w-')
    X[0][:] = [4, 5, 6]
    X[1][1, :] = 1.5
    f.flush()
    assert np.array_equal(myhd.loadhd5(fn, 'id'), [4, 5, 6])
    X[2][2] = np.eye(2)
    f.close()
    a, b = myhd.loadhd5(fn, ['solutions_S', 'B_S'])
    assert a.shape == (3, 162) and a[1, 7] == 1.5 and np.array_equal(b[2], np.eye(2))
    with pytest.raises(OSError):
        myhd.zeros_openFile(fn, (2,), 'x', mode='w-')                 # reference semantics: 'w-' refuses to overwrite
    assert myhd.checkExistenceDataset(fn, 'id') and not myhd.checkExistenceDataset(fn, 'center')
    myhd.addDataset(fn, np.ones((3, 2)), 'center')                      # build_inout.py:39,50 appends datasets
    assert myhd.loadhd5(fn, 'center').shape == (3, 2) and np.array_equal(myhd.loadhd5(fn, 'id'), [4, 5, 6])
    fn2 = str(tmp_path / "snap2.hd5")
    myhd.savehd5(fn2, [np.arange(3.0) + 10, np.zeros((3, 162)), np.zeros((3, 2, 2))], ['id', 'solutions_S', 'B_S'], 'w')
    out = str(tmp_path / "merged.hd5")
    myhd.merge([fn, fn2], out, ['id', 'solutions_S'], ['id', 'solutions_S'])
    assert np.array_equal(myhd.loadhd5(out, 'id'), [4, 5, 6, 10, 11, 12])
    assert myhd.loadhd5(out, 'solutions_S').shape == (6, 162)


def test_mesh_cache_roundtrip(tmp_path, coarse_reduced_meshes):
    from deepbnd_b200.mesh.mesh_io import save_mesh_cache, load_mesh, cache_name
    name = str(tmp_path / "meshes" / "mesh_micro_3_reduced.xdmf")
    save_mesh_cache(name, coarse_reduced_meshes[0])
    assert os.path.exists(cache_name(name))
    p, t, r = load_mesh(name)
    assert np.array_equal(p, coarse_reduced_meshes[0][0]) and np.array_equal(t, coarse_reduced_meshes[0][1])
    assert np.array_equal(r, coarse_reduced_meshes[0][2])
    with pytest.raises(FileNotFoundError):
        load_mesh(str(tmp_path / "missing.xdmf"))


def test_xdmf_triplet_reader(tmp_path, coarse_reduced_meshes):
    """a meshio-3.3-style XDMF/HDF5 triplet (wrapper_io.py:78-134) written with minih5 is read back."""
    from deepbnd_b200.io import minih5
    from deepbnd_b200.mesh.mesh_io import read_xdmf_mesh
    p, t, r = coarse_reduced_meshes[0]
    base = str(tmp_path / "mesh")
    with minih5.Writer(base + ".h5") as w:
        w.create_dataset("data0", p)
        w.create_dataset("data1", t.astype(np.int64))
    with minih5.Writer(base + "_regions.h5") as w:
        w.create_dataset("data2", r.astype(np.int64))
    xd = ('<Xdmf Version="3.0"><Domain><Grid Name="Grid"><Geometry GeometryType="XY"><DataItem DataType="Float" '
          'Dimensions="%d 2" Format="HDF" Precision="8">mesh.h5:/data0</DataItem></Geometry><Topology NumberOfElements="%d" '
          'TopologyType="Triangle"><DataItem DataType="Int" Dimensions="%d 3" Format="HDF" Precision="8">mesh.h5:/data1'
          '</DataItem></Topology></Grid></Domain></Xdmf>') % (len(p), len(t), len(t))
    open(base + ".xdmf", "w").write(xd)
    xr = ('<Xdmf Version="3.0"><Domain><Grid Name="Grid"><Topology TopologyType="Triangle"><DataItem Format="HDF">'
          'mesh_regions.h5:/data1</DataItem></Topology><Attribute AttributeType="Scalar" Center="Cell" Name="regions">'
          '<DataItem DataType="Int" Dimensions="%d" Format="HDF" Precision="8">mesh_regions.h5:/data2</DataItem>'
          '</Attribute></Grid></Domain></Xdmf>') % len(t)
    open(base + "_regions.xdmf", "w").write(xr)
    p2, t2, r2 = read_xdmf_mesh(base + ".xdmf")
    assert np.array_equal(p2, p) and np.array_equal(t2, t) and np.array_equal(r2, r)


# ---- interop with files this package did not write (N1, a17) -------------------------------------------

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reader_on_hand_assembled_chunked_deflate_file():
    """tests/golden/make_h5_fixtures.py builds this file byte by byte from the HDF5 specification (not with
    minih5.Writer): chunks (1, ...) + deflate level 0, chunk B-tree v1 — the layout of wrapper_h5py.py:30-34."""
    e = np.load(os.path.join(GOLD, "h5_fixtures_expected.npz"))
    fn = os.path.join(GOLD, "chunked_deflate0.h5")
    assert sorted(myhd.minih5.File(fn).keys()) == ["B_S", "id", "solutions_S"]
    for k in ("solutions_S", "B_S", "id"):
        a = myhd.loadhd5(fn, k)
        assert a.shape == e[k].shape and np.array_equal(a, e[k]), k
    assert myhd.checkExistenceDataset(fn, "id") and not myhd.checkExistenceDataset(fn, "nope")


@pytest.mark.parametrize("ver", [1, 2])
def test_reader_follows_external_links_in_new_style_groups(ver):
    """the `_regions.h5` of wrapper_io.py:124-127: link messages (v1 and v2 object headers), data1 -> mesh.h5:/data1"""
    e = np.load(os.path.join(GOLD, "h5_fixtures_expected.npz"))
    f = myhd.minih5.File(os.path.join(GOLD, "link_regions_v%d.h5" % ver))
    assert sorted(f.keys()) == ["data0", "data1", "data2"]
    assert np.array_equal(np.array(f["data1"]), e["tri"])          # through the external link, chunked + deflate 4
    assert np.array_equal(np.array(f["data2"]), e["tags"])
    assert np.array(f["data0"]).shape == (1, 2)


def test_xdmf_triplet_with_external_link_is_read(tmp_path):
    """meshio-3.3-style triplet around the hand-assembled files: mesh.xdmf -> link_mesh.h5, mesh_regions.xdmf ->
    link_regions.h5 whose topology item is the ExternalLink"""
    import shutil
    from deepbnd_b200.mesh.mesh_io import read_xdmf_mesh
    e = np.load(os.path.join(GOLD, "h5_fixtures_expected.npz"))
    shutil.copy(os.path.join(GOLD, "link_mesh.h5"), tmp_path / "link_mesh.h5")
    shutil.copy(os.path.join(GOLD, "link_regions_v1.h5"), tmp_path / "mesh_regions.h5")
    (tmp_path / "mesh.xdmf").write_text(
        '<Xdmf Version="3.0"><Domain><Grid Name="Grid"><Geometry GeometryType="XY"><DataItem DataType="Float" Dimensions="9 2" '
        'Format="HDF" Precision="8">link_mesh.h5:/data0</DataItem></Geometry><Topology NumberOfElements="7" TopologyType="Triangle">'
        '<DataItem DataType="Int" Dimensions="7 3" Format="HDF" Precision="8">link_mesh.h5:/data1</DataItem></Topology></Grid></Domain></Xdmf>')
    (tmp_path / "mesh_regions.xdmf").write_text(
        '<Xdmf Version="3.0"><Domain><Grid Name="Grid"><Geometry GeometryType="XY"><DataItem DataType="Float" Dimensions="1 2" '
        'Format="HDF" Precision="8">mesh_regions.h5:/data0</DataItem></Geometry><Topology NumberOfElements="7" TopologyType="Triangle">'
        '<DataItem DataType="Int" Dimensions="7 3" Format="HDF" Precision="8">mesh_regions.h5:/data1</DataItem></Topology>'
        '<Attribute AttributeType="Scalar" Center="Cell" Name="regions"><DataItem DataType="Int" Dimensions="7" Format="HDF" '
        'Precision="8">mesh_regions.h5:/data2</DataItem></Attribute></Grid></Domain></Xdmf>')
    pts, tri, reg = read_xdmf_mesh(str(tmp_path / "mesh.xdmf"))
    assert np.array_equal(pts, e["pts"]) and np.array_equal(tri, e["tri"]) and np.array_equal(reg, e["tags"])


def test_chunked_writer_layout_roundtrip_and_structure(tmp_path, monkeypatch):
    """DEEPBND_H5_LAYOUT=chunked writes the reference's dataset creation properties (chunks (1, ...), deflate 0): the
    file reads back, and its dataset header carries the same layout / filter messages as the hand-assembled fixture"""
    monkeypatch.setenv("DEEPBND_H5_LAYOUT", "chunked")
    rng = np.random.RandomState(0)
    A, B = rng.randn(200, 5), rng.randn(3, 2, 2)
    fn = str(tmp_path / "c.hd5")
    myhd.savehd5(fn, [A, B], ["A", "B"], "w")
    assert np.array_equal(myhd.loadhd5(fn, "A"), A) and np.array_equal(myhd.loadhd5(fn, "B"), B)

    def layout_and_filter(path, name):
        ds = myhd.minih5.File(path)[name]
        lay = [bytes(d) for t, d in ds.msgs if t == 0x0008][0]
        fil = [bytes(d) for t, d in ds.msgs if t == 0x000B][0]
        return lay[:3], lay[11:], fil[:20]          # (version, class, rank), chunk dims, filter description
    mine = layout_and_filter(fn, "B")
    ref = layout_and_filter(os.path.join(GOLD, "chunked_deflate0.h5"), "B_S")
    assert mine[0] == ref[0] == bytes([3, 2, 4]) and mine[2] == ref[2]
    assert struct.unpack("<4I", mine[1][:16]) == (1, 2, 2, 8)


def test_zerosfile_flushes_rows_in_place(tmp_path):
    """buildSnapshots / predictTangents flush per batch: only the rows of the batch are rewritten"""
    fn = str(tmp_path / "z.hd5")
    (a, b), f = myhd.zeros_openFile(fn, [(6, 3), (6,)], ["a", "b"], mode="w")
    myhd.addDataset(f, np.arange(4.0).reshape(2, 2), "static")
    a[0:2] = 1.0; b[0:2] = 2.0
    f.flush([0, 1])                                   # first flush: whole file
    size0 = os.path.getsize(fn)
    a[4:6] = 5.0; b[4] = 7.0
    f.flush([4, 5])
    assert os.path.getsize(fn) == size0
    assert np.array_equal(myhd.loadhd5(fn, "a"), a) and np.array_equal(myhd.loadhd5(fn, "b"), b)
    assert np.array_equal(myhd.loadhd5(fn, "static"), np.arange(4.0).reshape(2, 2))
    f.close()


def test_vref_vectors_are_permuted_or_refused(tmp_path, monkeypatch):
    """a file with Vref vectors but no DOF-order description is refused; with a description (in the file or
    the exporter's sidecar) the vectors come back in the library's order"""
    from deepbnd_b200 import vref
    rng = np.random.RandomState(3)
    U = rng.randn(4, 162)
    coords, comps = vref.vref_order_datasets()
    shuffle = rng.permutation(162)                     # a "dolfin" order: file DOF d = library DOF shuffle[d]
    fn = str(tmp_path / "bcs.hd5")
    myhd.savehd5(fn, [U[:, shuffle]], ["u0"], "w")
    with pytest.raises(vref.VrefOrderError):
        vref.load_vref_vectors(fn, "u0")
    monkeypatch.setenv("DEEPBND_ASSUME_LIBRARY_VREF_ORDER", "1")
    assert np.array_equal(vref.load_vref_vectors(fn, "u0"), U[:, shuffle])
    monkeypatch.delenv("DEEPBND_ASSUME_LIBRARY_VREF_ORDER")
    myhd.savehd5(str(tmp_path / vref.SIDECAR), [coords[shuffle], comps[shuffle]], [vref.COORD_LABEL, vref.COMP_LABEL], "w")
    assert np.array_equal(vref.load_vref_vectors(fn, "u0"), U)                      # sidecar of the FEniCS exporter
    fn2 = str(tmp_path / "sub" / "Wbasis.hd5")
    os.makedirs(os.path.dirname(fn2))
    myhd.savehd5(fn2, [U, coords, comps], ["Wbasis_A", vref.COORD_LABEL, vref.COMP_LABEL], "w")
    assert np.array_equal(vref.load_vref_vectors(fn2, ["Wbasis_A"])[0], U)          # self-described, identity
