"""Minimal HDF5 writer for the VTKHDF snapshots (the reference links HDF5.jl; no HDF5 library
exists in this image, so the container is written directly from the published file format,
"HDF5 File Format Specification Version 1.1/2.0"; section numbers below refer to it).

What is written is the oldest, universally readable subset -- the layout libhdf5 1.6 itself
produced: superblock version 0 (II.A), version-1 object headers (IV.A.1.a), groups as a
version-1 B-tree of symbol-table nodes plus a local heap (III.A.1, III.B, III.D), contiguous
little-endian datasets (IV.A.2.i layout version 3, class 1), version-1 dataspace, datatype and
attribute messages.  No chunking, filters, fill data, modification times or free-space management.

    f = File(path); g = f.create_group("/VTKHDF"); g.attrs["Version"] = np.array([1, 0])
    f.write("/VTKHDF/Points", array); f.close()

Arrays are stored C-ordered with their numpy shape (HDF5.jl writes a Julia array of size (3, N)
as an HDF5 dataset of shape (N, 3): callers pass the (N, 3) array).  Supported element types:
float64, float32, int64, int32, uint8 and fixed-length ASCII strings (scalar attributes).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 16            # symbol-table nodes hold up to 2*_LEAF_K entries (superblock field, II.A)
_INTERNAL_K = 16        # B-tree nodes hold up to 2*_INTERNAL_K children


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _datatype(dt):
    """Datatype message body (IV.A.2.d), version 1."""
    if isinstance(dt, tuple) and dt[0] == "ascii":
        # class 3 string: null-terminated padding, ASCII character set (HDF5.jl datatype(::String)
        # is H5T_C_S1 resized to the string's length; IO.jl:30-33 then selects H5T_CSET_ASCII)
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt[1])
    dt = np.dtype(dt)
    if dt.kind == "f":
        prec = dt.itemsize * 8
        exp_bits, man_bits, bias = (11, 52, 1023) if dt.itemsize == 8 else (8, 23, 127)
        # class 1: little-endian, mantissa normalisation "msb implied" (2 << 4), sign bit location
        return (struct.pack("<BBBBI", 0x11, 0x20, prec - 1, 0, dt.itemsize)
                + struct.pack("<HHBBBBI", 0, prec, man_bits, exp_bits, 0, man_bits, bias))
    if dt.kind in "iu":
        flags = 0x08 if dt.kind == "i" else 0x00        # bit 3: two's complement signed
        return struct.pack("<BBBBI", 0x10, flags, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    raise TypeError(f"unsupported element type {dt}")


def _dataspace(shape):
    """Dataspace message body (IV.A.2.b), version 1, no maximum dimensions; rank 0 = scalar."""
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _object_header(messages):
    """Version-1 object header: 12-byte prefix padded to 16, then the 8-byte aligned messages."""
    data = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(data)) + b"\0" * 4 + data


def _attribute(name, value):
    """Attribute message body (IV.A.2.m), version 1."""
    if isinstance(value, str):
        raw = value.encode("ascii")
        dt, ds, data = _datatype(("ascii", max(len(raw), 1))), _dataspace(()), raw or b"\0"
    else:
        a = np.ascontiguousarray(value)
        if a.dtype.kind in "iu" and a.dtype.itemsize != 1:
            a = a.astype("<i8")
        elif a.dtype.kind == "f":
            a = a.astype("<f8")
        dt, ds, data = _datatype(a.dtype), _dataspace(a.shape), a.tobytes()
    nm = name.encode("ascii") + b"\0"
    return (struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data)


class _Group:
    def __init__(self, file, path):
        self.file, self.path = file, path
        self.children = {}          # name -> _Group | _Dataset
        self.attrs = {}


class _Dataset:
    def __init__(self, array):
        self.array = array
        self.attrs = {}


class File:
    """Everything is kept in memory and laid out by close(): the format needs the addresses of
    children before a parent's B-tree can be written."""

    def __init__(self, path):
        self.path = path
        self.root = _Group(self, "/")
        self.closed = False

    # ------------------------------------------------------------------ building the tree
    def _walk(self, path, create):
        parts = [p for p in path.split("/") if p]
        g = self.root
        for p in parts:
            if p not in g.children:
                if not create:
                    raise KeyError(path)
                g.children[p] = _Group(self, g.path.rstrip("/") + "/" + p)
            g = g.children[p]
            if not isinstance(g, _Group):
                raise ValueError(f"{p} in {path} is a dataset, not a group")
        return g

    def create_group(self, path):
        return self._walk(path, True)

    def write(self, path, data):
        """`write(file, "A/B/name", data)` of HDF5.jl: intermediate groups are created."""
        if self.closed:
            raise ValueError("file is closed")
        parent, _, name = path.rstrip("/").rpartition("/")
        g = self._walk(parent, True)
        if name in g.children:
            raise ValueError(f"{path} exists")      # HDF5.jl: "name already exists"
        a = np.ascontiguousarray(data)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        if a.dtype.kind not in "fiu" or a.dtype.itemsize not in (1, 4, 8):
            raise TypeError(f"unsupported element type {a.dtype}")
        a = a.astype(a.dtype.newbyteorder("<"), copy=False)
        g.children[name] = _Dataset(a)
        return g.children[name]

    # ------------------------------------------------------------------ layout
    def close(self):
        if self.closed:
            return
        self.closed = True
        self._chunks = []           # (address, bytes)
        self._eof = 96              # superblock (version 0, 8-byte offsets and lengths)
        root_hdr, root_btree, root_heap = self._emit_group(self.root)
        sb = (b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
              + struct.pack("<HHI", _LEAF_K, _INTERNAL_K, 0)
              + struct.pack("<QQQQ", 0, UNDEF, self._eof, UNDEF)
              # root group symbol-table entry (III.C): name offset 0, header, cache type 1 + scratch
              + struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_btree, root_heap))
        assert len(sb) == 96
        with open(self.path, "wb") as fh:
            fh.write(sb)
            for addr, blob in sorted(self._chunks, key=lambda c: c[0]):
                fh.seek(addr)
                if isinstance(blob, np.ndarray):
                    blob.tofile(fh)
                else:
                    fh.write(blob)
            fh.truncate(self._eof)

    def _alloc(self, nbytes):
        addr = self._eof
        self._eof += nbytes + (-nbytes % 8)
        return addr

    def _put(self, blob, nbytes=None):
        addr = self._alloc(len(blob) if nbytes is None else nbytes)
        self._chunks.append((addr, blob))
        return addr

    def _emit_dataset(self, d):
        a = d.array
        raw = self._put(a, a.nbytes) if a.nbytes else UNDEF
        msgs = [
            _message(0x0001, _dataspace(a.shape)),
            _message(0x0003, _datatype(a.dtype), flags=1),                      # constant message
            # fill value (IV.A.2.f) version 2: allocate late, write if set, default (size 0) value
            _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),
            # layout version 3, class 1 contiguous: address and size of the raw data
            _message(0x0008, struct.pack("<BBQQ", 3, 1, raw, a.nbytes)),
        ]
        msgs += [_message(0x000C, _attribute(k, v)) for k, v in d.attrs.items()]
        return self._put(_object_header(msgs))

    def _emit_group(self, g):
        names = sorted(g.children, key=lambda s: s.encode("ascii"))     # strcmp order
        entries = []
        for n in names:
            c = g.children[n]
            if isinstance(c, _Group):
                entries.append((n,) + self._emit_group(c))
            else:
                entries.append((n, self._emit_dataset(c), None, None))
        # local heap (III.D): offset 0 holds the empty string the first B-tree key points at
        heap, offs = bytearray(8), {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("ascii") + b"\0")
        heap_data = self._put(bytes(heap))
        # free-list head 1 = H5HL_FREE_NULL: no free block in the data segment
        heap_addr = self._put(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), 1, heap_data))
        # symbol-table nodes (III.C): 2*_LEAF_K slots of 40 bytes each, names ascending
        per = 2 * _LEAF_K
        chunks = [entries[i:i + per] for i in range(0, len(entries), per)] or [[]]
        if len(chunks) > 2 * _INTERNAL_K:
            raise ValueError(f"{len(entries)} links in one group: more than this writer lays out")
        snods, keys = [], [0]
        for ch in chunks:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(ch))
            for n, hdr, bt, hp in ch:
                if bt is None:
                    body += struct.pack("<QQII", offs[n], hdr, 0, 0) + b"\0" * 16
                else:
                    body += struct.pack("<QQII", offs[n], hdr, 1, 0) + struct.pack("<QQ", bt, hp)
            body += b"\0" * (8 + 40 * per - len(body))
            snods.append(self._put(body))
            keys.append(offs[ch[-1][0]] if ch else 0)
        # B-tree node (III.A.1): type 0 (group), level 0, keys are heap offsets of the largest
        # name of the child to their left
        nchild = len(snods) if entries else 0
        bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, nchild, UNDEF, UNDEF)
        for i in range(nchild):
            bt += struct.pack("<QQ", keys[i], snods[i])
        bt += struct.pack("<Q", keys[nchild])
        bt += b"\0" * (24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8 - len(bt))
        bt_addr = self._put(bt)
        msgs = [_message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
        msgs += [_message(0x000C, _attribute(k, v)) for k, v in g.attrs.items()]
        return self._put(_object_header(msgs)), bt_addr, heap_addr
