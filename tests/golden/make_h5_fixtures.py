#!/usr/bin/env python
"""Hand-assembled HDF5 fixtures, written byte by byte from the HDF5 File Format Specification (v3.0) and NOT through
deepbnd_b200.io.minih5.Writer, in the two layouts the reference's files really have and our own writer never produces:

  chunked_deflate0.h5   what core/data_manipulation/wrapper_h5py.py:30-34 writes: a float64 dataset with chunks
                        (1, ...) and the deflate filter at level 0 (zlib "stored" blocks), indexed by a version-1
                        B-tree of chunk records; plus an int64 `id` dataset with the same layout.
  link_regions.h5       what wrapper_io.py:124-127 leaves behind: a group whose entries are LINK MESSAGES (the library
  link_mesh.h5          converts the root group to the new style when an ExternalLink is inserted): `data0` hard link,
                        `data1` external link to link_mesh.h5:/data1, `data2` hard link (the region tags).  One
                        variant with a version-1 object header, one with a version-2 ("OHDR") header.

The expected contents are stored next to the files (h5_fixtures_expected.npz).  Run once; the outputs are committed.
"""
import os
import struct
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


def pad8(b):
    return b + b"\x00" * (-len(b) % 8)


class Img:
    """growing file image with 8-byte aligned allocation"""

    def __init__(self):
        self.b = bytearray()

    def alloc(self, n):
        self.b += b"\x00" * (-len(self.b) % 8)
        a = len(self.b)
        self.b += b"\x00" * n
        return a

    def put(self, a, data):
        self.b[a:a + len(data)] = data

    def add(self, data):
        a = self.alloc(len(data))
        self.put(a, data)
        return a


def msg_v1(mtype, data, flags=0):
    data = pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def ohdr_v1(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\x00" * 4 + body


def dataspace_v1(shape):
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def datatype_f8():
    return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def datatype_i8():
    return struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 8) + struct.pack("<HH", 0, 64)


def filter_pipeline_deflate(level):
    # version 1: version, nfilters, 6 reserved; filter: id, name length, flags, n client values, [name], values, pad
    return struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", level) + b"\x00" * 4


def chunked_dataset(img, arr, dtype_msg, level=0):
    """object header of a dataset with chunks (1, shape[1:]) and deflate(level); returns its address"""
    shape = arr.shape
    rank = len(shape)
    cdims = (1,) + shape[1:]
    esize = arr.dtype.itemsize
    # chunk data + keys
    keys = []
    for i in range(shape[0]):
        raw = np.ascontiguousarray(arr[i:i + 1]).tobytes()
        comp = zlib.compress(raw, level)
        a = img.add(comp)
        keys.append((len(comp), 0, (i,) + (0,) * (rank - 1) + (0,), a))
    # one leaf node of the chunk B-tree (type 1): signature, type, level, entries used, left, right siblings, then
    # key/child pairs and the final key (chunk size, filter mask, rank+1 offsets)
    node = b"TREE" + struct.pack("<BBH", 1, 0, len(keys)) + struct.pack("<QQ", UNDEF, UNDEF)
    for nbytes, mask, offs, a in keys:
        node += struct.pack("<II", nbytes, mask) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", a)
    node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in ((shape[0],) + (0,) * rank))
    bt = img.add(node)
    layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", bt) + b"".join(struct.pack("<I", c) for c in cdims + (esize,))
    fill = struct.pack("<BBBB", 2, 1, 0, 0)                         # fill value v2: allocate late... undefined value
    msgs = [msg_v1(0x0001, dataspace_v1(shape)), msg_v1(0x0003, dtype_msg, 1), msg_v1(0x0005, fill),
            msg_v1(0x000B, filter_pipeline_deflate(level)), msg_v1(0x0008, layout)]
    return img.add(ohdr_v1(msgs))


def contiguous_dataset(img, arr, dtype_msg):
    raw = np.ascontiguousarray(arr).tobytes()
    a = img.add(raw)
    layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", a, len(raw))
    msgs = [msg_v1(0x0001, dataspace_v1(arr.shape)), msg_v1(0x0003, dtype_msg, 1), msg_v1(0x0008, layout)]
    return img.add(ohdr_v1(msgs))


def superblock_v0(root_ohdr, btree, heap, eof):
    sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    # root symbol table entry: name offset, object header, cache type 1 (B-tree + heap cached) / 0, scratch
    ctype = 1 if btree is not None else 0
    scratch = struct.pack("<QQ", btree, heap) if btree is not None else b"\x00" * 16
    sb += struct.pack("<QQII", 0, root_ohdr, ctype, 0) + scratch
    return sb


def old_style_file(path, datasets):
    """root group = symbol table (B-tree v1 + SNOD + local heap); datasets: name -> builder(img) -> ohdr address"""
    img = Img()
    img.alloc(96)                                                  # superblock v0 with 8-byte addresses
    addrs = {n: build(img) for n, build in datasets.items()}
    names = sorted(addrs)
    heap_data = b"\x00" * 8
    noff = {}
    for n in names:
        noff[n] = len(heap_data)
        heap_data += pad8(n.encode() + b"\x00")
    heap_data += b"\x00" * 32
    hd = img.add(heap_data)
    heap = img.add(b"HEAP" + struct.pack("<B3x", 0) + struct.pack("<QQQ", len(heap_data), len(heap_data) - 32, hd))
    # free block at the end of the heap data: next-free offset (1 = none), size
    img.put(hd + len(heap_data) - 32, struct.pack("<QQ", 1, 32))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", noff[n], addrs[n], 0, 0)
    snod += b"\x00" * (40 * (2 * 4 - len(names)))                   # node capacity 2K entries (K = 4)
    sn = img.add(snod)
    bt = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<QQQ", 0, sn, noff[names[-1]])
    bt += b"\x00" * (16 * 2 * 16)                                   # room for 2K = 32 children
    btree = img.add(bt)
    root = img.add(ohdr_v1([msg_v1(0x0011, struct.pack("<QQ", btree, heap))]))
    img.put(0, superblock_v0(root, btree, heap, len(img.b)))
    open(path, "wb").write(bytes(img.b))


def link_message(name, kind, payload):
    """Link message v1: flags bit 3 = link type present, bits 0-1 = size of the name-length field (0: 1 byte)"""
    nb = name.encode()
    if kind == "hard":
        return struct.pack("<BB", 1, 0) + struct.pack("<B", len(nb)) + nb + struct.pack("<Q", payload)
    if kind == "external":
        fn, obj = payload
        info = b"\x00" + fn.encode() + b"\x00" + obj.encode() + b"\x00"
        return struct.pack("<BBB", 1, 0x08, 64) + struct.pack("<B", len(nb)) + nb + struct.pack("<H", len(info)) + info
    raise ValueError(kind)


def link_info_message():
    # version 0, flags 0, fractal heap address and name-index B-tree address undefined (compact storage)
    return struct.pack("<BB", 0, 0) + struct.pack("<QQ", UNDEF, UNDEF)


def lookup3(data):
    """Bob Jenkins' lookup3 hashlittle (the checksum of version-2 object headers), initval 0"""
    def rot(x, k):
        return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF
    length = len(data)
    a = b = c = (0xdeadbeef + length) & 0xFFFFFFFF
    i = 0
    while length > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & 0xFFFFFFFF
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & 0xFFFFFFFF
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= rot(c, 4); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= rot(a, 6); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= rot(b, 8); b = (b + a) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= rot(c, 16); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= rot(a, 19); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= rot(b, 4); b = (b + a) & 0xFFFFFFFF
        i += 12; length -= 12
    tail = data[i:] + b"\x00" * (12 - len(data[i:]))
    if length > 0:
        a = (a + int.from_bytes(tail[0:4], "little")) & 0xFFFFFFFF
        b = (b + int.from_bytes(tail[4:8], "little")) & 0xFFFFFFFF
        c = (c + int.from_bytes(tail[8:12], "little")) & 0xFFFFFFFF
        c ^= b; c = (c - rot(b, 14)) & 0xFFFFFFFF
        a ^= c; a = (a - rot(c, 11)) & 0xFFFFFFFF
        b ^= a; b = (b - rot(a, 25)) & 0xFFFFFFFF
        c ^= b; c = (c - rot(b, 16)) & 0xFFFFFFFF
        a ^= c; a = (a - rot(c, 4)) & 0xFFFFFFFF
        b ^= a; b = (b - rot(a, 14)) & 0xFFFFFFFF
        c ^= b; c = (c - rot(b, 24)) & 0xFFFFFFFF
    return c


def ohdr_v2(msgs):
    """version-2 object header: 'OHDR', version 2, flags (bits 0-1: size of the chunk-0 size field = 1 -> 2 bytes),
    chunk size, messages (type 1 byte, size 2, flags 1), checksum"""
    body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in msgs)
    hdr = b"OHDR" + struct.pack("<BB", 2, 0x01) + struct.pack("<H", len(body)) + body
    return hdr + struct.pack("<I", lookup3(hdr))


def new_style_file(path, links, v2):
    img = Img()
    img.alloc(96)
    lmsgs = []
    for name, kind, payload in links:
        if kind == "hard":
            payload = payload(img)
        lmsgs.append((name, kind, payload))
    if v2:
        root = img.add(ohdr_v2([(0x02, link_info_message())] + [(0x06, link_message(*l)) for l in lmsgs]))
    else:
        root = img.add(ohdr_v1([msg_v1(0x0002, link_info_message())] + [msg_v1(0x0006, link_message(*l)) for l in lmsgs]))
    img.put(0, superblock_v0(root, None, None, len(img.b)))
    open(path, "wb").write(bytes(img.b))


def main():
    rng = np.random.RandomState(7)
    sol = rng.randn(5, 6)
    B = rng.randn(5, 2, 2)
    ids = np.arange(5, dtype=np.int64) * 3
    old_style_file(os.path.join(HERE, "chunked_deflate0.h5"), {
        "solutions_S": lambda img: chunked_dataset(img, sol, datatype_f8(), 0),
        "B_S": lambda img: chunked_dataset(img, B, datatype_f8(), 0),
        "id": lambda img: chunked_dataset(img, ids, datatype_i8(), 6),
    })
    tri = rng.randint(0, 9, size=(7, 3)).astype(np.int64)
    pts = rng.randn(9, 2)
    tags = rng.randint(0, 4, size=7).astype(np.int64)
    old_style_file(os.path.join(HERE, "link_mesh.h5"), {
        "data0": lambda img: contiguous_dataset(img, pts, datatype_f8()),
        "data1": lambda img: chunked_dataset(img, tri, datatype_i8(), 4),
    })
    for v2 in (False, True):
        new_style_file(os.path.join(HERE, "link_regions_v%d.h5" % (2 if v2 else 1)), [
            ("data0", "hard", lambda img: contiguous_dataset(img, np.zeros((1, 2)), datatype_f8())),
            ("data1", "external", ("link_mesh.h5", "/data1")),
            ("data2", "hard", lambda img: contiguous_dataset(img, tags, datatype_i8())),
        ], v2)
    np.savez(os.path.join(HERE, "h5_fixtures_expected.npz"), solutions_S=sol, B_S=B, id=ids, tri=tri, pts=pts, tags=tags)
    print("wrote fixtures in", HERE)


if __name__ == "__main__":
    main()
