"""minih5 — a small pure-numpy HDF5 reader/writer for the DeepBND dataset files.

The reference does all its dataset I/O through h5py (core/data_manipulation/wrapper_h5py.py:30-96:
`loadhd5`, `savehd5`, `zeros_openFile`): float64 datasets in the root group, written chunked
`(1, ...)` with deflate level 0, read back whole with `np.array(f[label])`.  Neither h5py nor libhdf5
exists in this image, so this module speaks the on-disk format directly (HDF5 File Format
Specification v1/v2 subset):

  writer  superblock v0, old-style root group (symbol-table B-tree v1 + local heap + SNODs),
          v1 object headers, CONTIGUOUS little-endian f8/f4/i8/i4/u1 datasets in the root group —
          the most conservative layout; h5py/libhdf5 open it in 'r' and 'a' mode and can add
          datasets to it (build_inout.py:39,50 does that to snapshots.hd5).
  reader  superblock v0/v1, v1 and v2 object headers with continuation blocks, old-style groups (nested) and
          new-style groups with compact link storage (hard, soft and EXTERNAL links: the `_regions.h5` files of
          wrapper_io.py:124-127), datatypes int/float of any size/endianness, layouts compact / contiguous /
          chunked (B-tree v1) with the deflate, shuffle and fletcher32 filters — what h5py's default
          (libver='earliest') writes for the reference's files.  tests/golden/make_h5_fixtures.py assembles
          such files byte by byte from the format specification, independently of the writer below.

Only whole-dataset reads and writes are supported (that is all the reference does).
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


# =================================================================================================
# writer
# =================================================================================================

def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 8:
            prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            signpos = 63
        else:
            prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            signpos = 31
        return struct.pack("<BBBBI", 0x11, 0x20, signpos, 0, dt.itemsize) + prop
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError("minih5 writer: unsupported dtype %s" % dt)


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _object_header(messages):
    body = b"".join(messages)
    # v1 prefix: version, reserved, nmsgs, refcount, header size, then 4 bytes of alignment padding
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\x00" * 4 + body


def _chunk_btree(out_append, keys, rank, end_offsets):
    """version-1 B-tree of chunk records (node type 1) over `keys` = [(nbytes, offsets, address)], at most 64 entries
    per node (the default 'indexed storage internal node K' = 32); `out_append(bytes) -> address`; returns the root."""
    CAP = 64
    level = 0
    nodes = [(k[1], k[0], k[2]) for k in keys]            # (first chunk offsets, nbytes of the first chunk, child address)
    while True:
        groups = [nodes[i:i + CAP] for i in range(0, len(nodes), CAP)]
        addrs, sizes = [], []
        for g in groups:
            sizes.append(24 + len(g) * (8 + 8 * (rank + 1) + 8) + 8 + 8 * (rank + 1))
        base = out_append(b"\x00" * 0)                     # current end (aligned)
        pos = base
        for sz in sizes:
            addrs.append(pos)
            pos += sz + (-sz % 8)
        new_nodes = []
        for gi, g in enumerate(groups):
            left = addrs[gi - 1] if gi > 0 else _UNDEF
            right = addrs[gi + 1] if gi + 1 < len(groups) else _UNDEF
            node = b"TREE" + struct.pack("<BBH", 1, level, len(g)) + struct.pack("<QQ", left, right)
            for offs, nbytes, child in g:
                node += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", child)
            last = groups[gi + 1][0][0] if gi + 1 < len(groups) else end_offsets
            node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in last)
            a = out_append(_pad8(node))
            assert a == addrs[gi]
            new_nodes.append((g[0][0], g[0][1], a))
        if len(new_nodes) == 1:
            return new_nodes[0][2]
        nodes = new_nodes
        level += 1


class Writer:
    """Collects root-group datasets and writes the file on close().  layout='contiguous' (default) or 'chunked' =
    the reference's dataset creation properties (wrapper_h5py.py:30-34): chunks (1, ...), deflate level 0."""

    def __init__(self, path, layout="contiguous"):
        if layout not in ("contiguous", "chunked"):
            raise ValueError("layout must be 'contiguous' or 'chunked'")
        self.path = path
        self.items = {}
        self.layout = layout
        self.data_addr = {}               # name -> file offset of the raw data (contiguous layout), after close()

    def create_dataset(self, name, data):
        a = np.asarray(data)
        if a.dtype.kind == "f" and a.dtype.itemsize not in (4, 8):
            a = a.astype(np.float64)
        if a.dtype.kind == "b":
            a = a.astype(np.uint8)
        shape = a.shape
        a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"), copy=False)).reshape(shape)
        if "/" in name.strip("/"):
            raise ValueError("minih5 writer: only root-group datasets")
        self.items[name.strip("/")] = a

    __setitem__ = create_dataset

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()

    def close(self):
        names = sorted(self.items, key=lambda s: s.encode())
        # ---- local heap data: "" at offset 0, then the names (8-byte aligned)
        heap = bytearray(b"\x00" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode() + b"\x00")
        heap_data_size = max(len(heap) + 32, 128)
        heap += b"\x00" * (heap_data_size - len(heap))
        # one free block covering the tail keeps libhdf5's heap bookkeeping happy when it appends names
        used = 8 + sum(len(_pad8(n.encode() + b"\x00")) for n in names)
        free_off = used
        struct.pack_into("<QQ", heap, free_off, 1, heap_data_size - free_off)   # next = H5HL_FREE_NULL, size

        LEAF_K, INT_K = 4, 16
        groups = [names[i:i + 2 * LEAF_K] for i in range(0, len(names), 2 * LEAF_K)] or [[]]
        if len(groups) > 2 * INT_K:
            raise ValueError("minih5 writer: too many datasets for a single-level group B-tree")

        # ---- layout of the file
        pos = 96                                          # superblock v0 with 8-byte offsets
        root_hdr = pos
        root_hdr_bytes = _object_header([_message(0x0011, struct.pack("<QQ", 0, 0)),      # patched below
                                         _message(0x0000, b"\x00" * 96)])                  # NIL: room to grow
        pos += len(root_hdr_bytes)
        btree = pos
        btree_size = 24 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8
        pos += btree_size
        heap_hdr = pos
        pos += 32
        heap_data = pos
        pos += heap_data_size
        snod_size = 8 + 2 * LEAF_K * 40
        snods = []
        for _ in groups:
            snods.append(pos)
            pos += snod_size
        obj_hdr, obj_bytes, data_addr, chunked = {}, {}, {}, {}
        for n in names:
            a = self.items[n]
            obj_hdr[n] = pos
            shape = a.shape if a.ndim else ()
            space = struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)
            chunked[n] = self.layout == "chunked" and a.ndim >= 1 and a.size > 0
            if chunked[n]:
                cd = (1,) + tuple(shape[1:]) + (a.dtype.itemsize,)
                lay = struct.pack("<BBBQ", 3, 2, len(cd), 0) + b"".join(struct.pack("<I", c) for c in cd)   # B-tree address patched below
                filt = struct.pack("<BB6x", 1, 1) + struct.pack("<HHHHI4x", 1, 0, 0, 1, 0)                   # deflate, level 0
                msgs = [_message(0x0001, space), _message(0x0003, _dtype_message(a.dtype), flags=1),
                        _message(0x0005, struct.pack("<BBBB", 2, 1, 0, 0)), _message(0x000B, filt),
                        _message(0x0008, lay), _message(0x0000, b"\x00" * 64)]
            else:
                msgs = [_message(0x0001, space), _message(0x0003, _dtype_message(a.dtype), flags=1),
                        _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),
                        _message(0x0008, struct.pack("<BBQQ", 3, 1, 0, 0)),                # patched below
                        _message(0x0000, b"\x00" * 64)]                                    # NIL: room for attributes
            obj_bytes[n] = bytearray(_object_header(msgs))
            pos += len(obj_bytes[n])
        for n in names:
            if chunked[n]:
                continue
            pos += -pos % 8
            data_addr[n] = pos if self.items[n].nbytes else _UNDEF
            pos += self.items[n].nbytes
        eof = pos

        out = bytearray(eof)
        tail = bytearray()                                # chunk data and chunk B-trees follow the fixed part

        def append(b):
            tail.extend(b"\x00" * (-(eof + len(tail)) % 8))
            a = eof + len(tail)
            tail.extend(b)
            return a

        for n in names:
            if not chunked[n]:
                continue
            a = self.items[n]
            rank = a.ndim
            keys = []
            for i in range(a.shape[0]):
                comp = zlib.compress(a[i:i + 1].tobytes(), 0)
                keys.append((len(comp), (i,) + (0,) * rank, append(comp)))
            data_addr[n] = _chunk_btree(append, keys, rank, (a.shape[0],) + (0,) * rank)
        # ---- superblock v0
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INT_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap_hdr)
        assert len(sb) == 96
        out[0:96] = sb
        # ---- root object header (symbol table message)
        rh = bytearray(root_hdr_bytes)
        struct.pack_into("<QQ", rh, 16 + 8, btree, heap_hdr)
        out[root_hdr:root_hdr + len(rh)] = rh
        # ---- group B-tree node (level 0): keys are heap offsets of names
        bt = bytearray(btree_size)
        bt[0:4] = b"TREE"
        struct.pack_into("<BBHQQ", bt, 4, 0, 0, len(groups) if names else 0, _UNDEF, _UNDEF)
        o = 24
        struct.pack_into("<Q", bt, o, 0)
        o += 8
        for g, addr in zip(groups, snods):
            if not names:
                break
            struct.pack_into("<Q", bt, o, addr)
            struct.pack_into("<Q", bt, o + 8, name_off[g[-1]])
            o += 16
        out[btree:btree + btree_size] = bt
        # ---- local heap
        out[heap_hdr:heap_hdr + 32] = b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, heap_data_size, free_off, heap_data)
        out[heap_data:heap_data + heap_data_size] = heap
        # ---- symbol table nodes
        for g, addr in zip(groups, snods):
            sn = bytearray(snod_size)
            sn[0:4] = b"SNOD"
            struct.pack_into("<BBH", sn, 4, 1, 0, len(g))
            for i, n in enumerate(g):
                struct.pack_into("<QQII", sn, 8 + 40 * i, name_off[n], obj_hdr[n], 0, 0)
            out[addr:addr + snod_size] = sn
        # ---- dataset headers + raw data
        for n in names:
            a = self.items[n]
            ob = obj_bytes[n]
            if chunked[n]:
                idx = ob.find(struct.pack("<BBB", 3, 2, a.ndim + 1), 16)
                struct.pack_into("<Q", ob, idx + 3, data_addr[n])
                out[obj_hdr[n]:obj_hdr[n] + len(ob)] = ob
                continue
            idx = ob.find(struct.pack("<HH", 0x0008, 24))
            struct.pack_into("<QQ", ob, idx + 8 + 2, data_addr[n], a.nbytes)
            out[obj_hdr[n]:obj_hdr[n] + len(ob)] = ob
            if a.nbytes:
                out[data_addr[n]:data_addr[n] + a.nbytes] = a.tobytes()
        if tail:
            struct.pack_into("<Q", out, 24 + 16, eof + len(tail))          # end-of-file address of the superblock
        self.data_addr = {n: data_addr[n] for n in names if not chunked[n]}
        with open(self.path, "wb") as f:
            f.write(out)
            f.write(tail)


# =================================================================================================
# reader
# =================================================================================================

class _Dataset:
    def __init__(self, f, addr):
        self.f = f
        self.addr = addr
        self.msgs = f._read_header(addr)
        self.shape = None
        self.dtype = None
        for t, d in self.msgs:
            if t == 0x0001:
                self.shape = f._parse_dataspace(d)
            elif t == 0x0003:
                self.dtype = f._parse_datatype(d)

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, key):
        return self.read()[key]

    def read(self):
        f = self.f
        layout = filters = None
        for t, d in self.msgs:
            if t == 0x0008:
                layout = d
            elif t == 0x000B:
                filters = f._parse_filters(d)
        if layout is None or self.dtype is None or self.shape is None:
            raise ValueError("minih5: not a dataset")
        count = int(np.prod(self.shape, dtype=np.int64)) if len(self.shape) else 1
        nbytes = count * self.dtype.itemsize
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 0:                                   # compact
                size = struct.unpack_from("<H", layout, 2)[0]
                raw = bytes(layout[4:4 + size])
            elif cls == 1:                                 # contiguous
                addr, size = struct.unpack_from("<QQ", layout, 2)
                raw = b"\x00" * nbytes if addr == _UNDEF else f._read(addr, nbytes)
            elif cls == 2:                                 # chunked, B-tree v1
                ndim = layout[2]
                btree = struct.unpack_from("<Q", layout, 3)[0]
                cdims = struct.unpack_from("<%dI" % ndim, layout, 11)
                return self._read_chunked(btree, cdims[:-1], filters)
            else:
                raise NotImplementedError("minih5: layout class %d" % cls)
        elif ver in (1, 2):
            ndim, cls = layout[1], layout[2]
            o = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", layout, o)[0]
                o += 8
            dims = struct.unpack_from("<%dI" % ndim, layout, o)
            o += 4 * ndim
            if cls == 1:
                raw = b"\x00" * nbytes if addr == _UNDEF else f._read(addr, nbytes)
            elif cls == 2:
                return self._read_chunked(addr, dims[:-1], filters)
            else:
                size = struct.unpack_from("<I", layout, o)[0]
                raw = bytes(layout[o + 4:o + 4 + size])
        else:
            raise NotImplementedError("minih5: data layout message version %d" % ver)
        return np.frombuffer(raw, dtype=self.dtype, count=count).reshape(self.shape).astype(self.dtype.newbyteorder("="))

    def _read_chunked(self, btree, cdims, filters):
        f = self.f
        out = np.zeros(self.shape, dtype=self.dtype.newbyteorder("="))
        if btree == _UNDEF:
            return out
        rank = len(self.shape)
        csize = int(np.prod(cdims, dtype=np.int64)) * self.dtype.itemsize

        def walk(addr):
            hd = f._read(addr, 24)
            if hd[:4] != b"TREE":
                raise ValueError("minih5: bad chunk B-tree node")
            ntype, level, used = struct.unpack_from("<BBH", hd, 4)
            if ntype != 1:
                raise ValueError("minih5: expected a raw-data B-tree")
            ksize = 8 + 8 * (rank + 1)
            body = f._read(addr + 24, used * (ksize + 8) + ksize)
            for i in range(used):
                ko = i * (ksize + 8)
                nbytes, fmask = struct.unpack_from("<II", body, ko)
                offs = struct.unpack_from("<%dQ" % (rank + 1), body, ko + 8)
                child = struct.unpack_from("<Q", body, ko + ksize)[0]
                if level > 0:
                    walk(child)
                    continue
                raw = f._read(child, nbytes)
                if filters:
                    for k, (fid, cd) in reversed(list(enumerate(filters))):
                        if fmask & (1 << k):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:
                            es = cd[0] if cd else self.dtype.itemsize
                            n = len(raw) // es
                            raw = np.frombuffer(raw[:n * es], np.uint8).reshape(es, n).T.tobytes() + raw[n * es:]
                        elif fid == 3:
                            raw = raw[:-4]
                        else:
                            raise NotImplementedError("minih5: filter id %d" % fid)
                chunk = np.frombuffer(raw[:csize], dtype=self.dtype).reshape(cdims)
                sl_out, sl_in = [], []
                for d in range(rank):
                    n = min(cdims[d], self.shape[d] - offs[d])
                    sl_out.append(slice(offs[d], offs[d] + n))
                    sl_in.append(slice(0, n))
                out[tuple(sl_out)] = chunk[tuple(sl_in)]

        walk(btree)
        return out


class File:
    """Read-only view: `f[name]` -> dataset (np.array(f[name]) reads it) or sub-group; `keys()`."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as fh:
            self.buf = fh.read()
        # the superblock may sit at 0, 512, 1024, ... (user block, e.g. MATLAB v7.3 files)
        base = 0
        while self.buf[base:base + 8] != _SIG:
            base = 512 if base == 0 else base * 2
            if base >= len(self.buf):
                raise ValueError("minih5: %s is not an HDF5 file" % path)
        ver = self.buf[base + 8]
        if ver not in (0, 1):
            raise NotImplementedError("minih5: superblock version %d (write the file with libver='earliest')" % ver)
        so, sl = self.buf[base + 13], self.buf[base + 14]
        if so != 8 or sl != 8:
            raise NotImplementedError("minih5: only 8-byte offsets/lengths")
        o = base + 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", self.buf, o)[0]
        if self.base == 0 and base:
            self.base = base
        o += 32
        name_off, hdr, ctype = struct.unpack_from("<QQI", self.buf, o)
        self.root = _Group(self, hdr)

    def _read(self, addr, n):
        a = self.base + addr
        if a + n > len(self.buf):
            raise ValueError("minih5: truncated file")
        return self.buf[a:a + n]

    # ---- object headers -------------------------------------------------------------------------
    def _read_header(self, addr):
        hd = self._read(addr, 16)
        if hd[:4] == b"OHDR":
            return self._read_header_v2(addr)
        ver, _, nmsgs, refc, hsize = struct.unpack_from("<BBHII", hd, 0)
        if ver != 1:
            raise ValueError("minih5: bad object header")
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsgs:
            a, size = blocks.pop(0)
            blk = self._read(a, size)
            o = 0
            while o + 8 <= size and len(msgs) < nmsgs:
                mtype, msize, mflags = struct.unpack_from("<HHB", blk, o)
                data = blk[o + 8:o + 8 + msize]
                o += 8 + msize
                if mtype == 0x0010:
                    ca, cs = struct.unpack_from("<QQ", data, 0)
                    blocks.append((ca, cs))
                msgs.append((mtype, data))
        return msgs

    def _read_header_v2(self, addr):
        """version-2 object header ('OHDR'; groups that hold link messages, files written with libver='latest')"""
        hd = self._read(addr, 6)
        if hd[4] != 2:
            raise ValueError("minih5: bad version-2 object header")
        flags = hd[5]
        o = addr + 6
        if flags & 0x20:
            o += 16                                    # access / modification / change / birth times
        if flags & 0x10:
            o += 4                                     # max compact / min dense attributes
        nsz = 1 << (flags & 3)
        size = int.from_bytes(self._read(o, nsz), "little")
        o += nsz
        msgs = []
        blocks = [(o, size)]
        while blocks:
            a, size = blocks.pop(0)
            blk = self._read(a, size)
            p = 0
            while p + 4 <= size:
                mtype = blk[p]
                msize, mflags = struct.unpack_from("<HB", blk, p + 1)
                p += 4 + (2 if flags & 0x04 else 0)    # creation order
                data = blk[p:p + msize]
                p += msize
                if mtype == 0x10:                      # continuation: 'OCHK' + messages + checksum
                    ca, cs = struct.unpack_from("<QQ", data, 0)
                    blocks.append((ca + 4, cs - 8))
                elif mtype != 0:
                    msgs.append((mtype, data))
        return msgs

    @staticmethod
    def _parse_link(d):
        """Link message -> (name, kind, payload): ('hard', address) | ('soft', path) | ('external', (file, path))"""
        ver, flags = d[0], d[1]
        o = 2
        ltype = 0
        if flags & 0x08:
            ltype = d[o]; o += 1
        if flags & 0x04:
            o += 8                                     # creation order
        if flags & 0x10:
            o += 1                                     # character set
        nsz = 1 << (flags & 3)
        nlen = int.from_bytes(d[o:o + nsz], "little")
        o += nsz
        name = bytes(d[o:o + nlen]).decode()
        o += nlen
        if ltype == 0:
            return name, "hard", struct.unpack_from("<Q", d, o)[0]
        ilen = struct.unpack_from("<H", d, o)[0]
        info = bytes(d[o + 2:o + 2 + ilen])
        if ltype == 1:
            return name, "soft", info.decode()
        if ltype == 64:
            fn, obj = info[1:].split(b"\x00")[:2]
            return name, "external", (fn.decode(), obj.decode())
        raise NotImplementedError("minih5: link type %d" % ltype)

    @staticmethod
    def _parse_dataspace(d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            o = 8
        elif ver == 2:
            o = 4
            if d[3] == 2:
                return ()
        else:
            raise NotImplementedError("minih5: dataspace version %d" % ver)
        return tuple(struct.unpack_from("<%dQ" % rank, d, o)) if rank else ()

    @staticmethod
    def _parse_datatype(d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        size = struct.unpack_from("<I", d, 4)[0]
        order = ">" if d[1] & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if d[1] & 0x08 else "u", size))
        if cls == 1:
            return np.dtype("%sf%d" % (order, size))
        raise NotImplementedError("minih5: datatype class %d" % cls)

    @staticmethod
    def _parse_filters(d):
        ver, nf = d[0], d[1]
        o = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = struct.unpack_from("<H", d, o)[0]
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", d, o + 2)[0]
                o += 4
            else:
                nlen = 0
                o += 2
            flags, ncd = struct.unpack_from("<HH", d, o)
            o += 4
            o += nlen + (-nlen % 8 if ver == 1 else 0)
            cd = struct.unpack_from("<%dI" % ncd, d, o)
            o += 4 * ncd
            if ver == 1 and ncd % 2:
                o += 4
            out.append((fid, cd))
        return out

    # ---- mapping interface ------------------------------------------------------------------------
    def keys(self):
        return self.root.keys()

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        node = self.root
        for part in [p for p in name.split("/") if p]:
            if not isinstance(node, _Group):
                raise KeyError(name)
            node = node[part]
        return node

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


class _Group:
    """old-style groups (symbol table) and new-style groups with compact link storage (link messages in the object
    header — what libhdf5 turns a group into when an ExternalLink is added, wrapper_io.py:124-127)"""

    def __init__(self, f, addr):
        self.f = f
        self.entries = {}                 # name -> object header address
        self.links = {}                   # name -> ('soft', path) | ('external', (file, path))
        for t, d in f._read_header(addr):
            if t == 0x0011:
                btree, heap = struct.unpack_from("<QQ", d, 0)
                self._load(btree, heap)
            elif t == 0x0002:
                o = 2 + (8 if d[1] & 1 else 0)
                if struct.unpack_from("<Q", d, o)[0] != _UNDEF:
                    raise NotImplementedError("minih5: groups with dense link storage (fractal heap)")
            elif t == 0x0006:
                name, kind, payload = f._parse_link(d)
                if kind == "hard":
                    self.entries[name] = payload
                else:
                    self.links[name] = (kind, payload)

    def _load(self, btree, heap):
        f = self.f
        hh = f._read(heap, 32)
        if hh[:4] != b"HEAP":
            raise ValueError("minih5: bad local heap")
        dsize, _, daddr = struct.unpack_from("<QQQ", hh, 8)
        hdata = f._read(daddr, dsize)

        def walk(addr):
            hd = f._read(addr, 24)
            if hd[:4] != b"TREE":
                raise ValueError("minih5: bad group B-tree node")
            ntype, level, used = struct.unpack_from("<BBH", hd, 4)
            body = f._read(addr + 24, used * 16 + 8)
            for i in range(used):
                child = struct.unpack_from("<Q", body, 8 + 16 * i)[0]
                if level > 0:
                    walk(child)
                    continue
                sn = f._read(child, 8)
                if sn[:4] != b"SNOD":
                    raise ValueError("minih5: bad symbol table node")
                nsym = struct.unpack_from("<H", sn, 6)[0]
                ents = f._read(child + 8, 40 * nsym)
                for k in range(nsym):
                    noff, ohdr = struct.unpack_from("<QQ", ents, 40 * k)
                    end = hdata.index(b"\x00", noff)
                    self.entries[hdata[noff:end].decode()] = ohdr

        if btree != _UNDEF:
            walk(btree)

    def keys(self):
        return list(self.entries) + list(self.links)

    def __getitem__(self, name):
        if name in self.links:
            kind, payload = self.links[name]
            if kind == "soft":
                return self.f[payload]
            import os
            fn, obj = payload                               # external link: relative to the directory of this file
            target = fn if os.path.isabs(fn) else os.path.join(os.path.dirname(os.path.abspath(self.f.path)), fn)
            if not os.path.exists(target):
                raise KeyError("%s -> external link to missing file %s" % (name, target))
            return File(target)[obj]
        if name not in self.entries:
            raise KeyError(name)
        addr = self.entries[name]
        msgs = self.f._read_header(addr)
        if any(t in (0x0011, 0x0002, 0x0006) for t, _ in msgs):
            return _Group(self.f, addr)
        return _Dataset(self.f, addr)
