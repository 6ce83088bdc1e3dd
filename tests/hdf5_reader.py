"""Test-side reader for the HDF5 subset flou_b200.hdf5min writes, written from the format
specification independently of the writer (no shared helpers): it follows the superblock to the
root group, walks B-tree -> symbol-table nodes -> local heap, decodes version-1 object headers and
returns datasets, groups and attributes.  It is strict about what it understands: unknown
signatures, versions, classes or addresses beyond the end-of-file address raise."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Node:
    def __init__(self):
        self.attrs, self.children, self.data = {}, None, None      # children: dict for groups


def _u(buf, off, n):
    return int.from_bytes(buf[off:off + n], "little")


def _decode_type(b):
    cls, ver = b[0] & 0x0F, b[0] >> 4
    assert ver == 1, f"datatype version {ver}"
    size = _u(b, 4, 4)
    if cls == 0:
        assert b[1] & 1 == 0, "big-endian integer"
        assert _u(b, 8, 2) == 0 and _u(b, 10, 2) == 8 * size
        return np.dtype(("<i" if b[1] & 8 else "<u") + str(size)), 12
    if cls == 1:
        assert b[1] & 1 == 0 and (b[1] >> 4) & 3 == 2, "float layout"
        off, prec, eloc, esize, mloc, msize, bias = struct.unpack("<HHBBBBI", b[8:20])
        assert (size, off, prec, b[2], eloc, esize, mloc, msize, bias) in (
            (8, 0, 64, 63, 52, 11, 0, 52, 1023), (4, 0, 32, 31, 23, 8, 0, 23, 127)), "not IEEE 754"
        return np.dtype("<f" + str(size)), 20
    if cls == 3:
        assert b[1] >> 4 == 0, "character set is not ASCII"
        return np.dtype("S" + str(size)), 8
    raise AssertionError(f"datatype class {cls}")


def _decode_space(b):
    assert b[0] == 1, "dataspace version"
    rank, flags = b[1], b[2]
    assert flags == 0
    return tuple(_u(b, 8 + 8 * i, 8) for i in range(rank))


class Reader:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        b = self.buf
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "signature"
        assert b[8] == 0 and b[9] == 0 and b[10] == 0 and b[12] == 0, "superblock versions"
        assert b[13] == 8 and b[14] == 8, "offset / length sizes"
        self.leaf_k, self.internal_k = _u(b, 16, 2), _u(b, 18, 2)
        base, free, self.eof, drv = struct.unpack("<QQQQ", b[24:56])
        assert base == 0 and free == UNDEF and drv == UNDEF
        assert self.eof == len(b), "end-of-file address"
        name_off, hdr, cache, _ = struct.unpack("<QQII", b[56:80])
        assert name_off == 0 and cache == 1
        self.root = self._object(hdr)
        bt, hp = struct.unpack("<QQ", b[80:96])
        assert (bt, hp) == self.root._stab, "root scratch-pad differs from the symbol-table message"

    def _at(self, addr, n):
        assert addr != UNDEF and addr + n <= self.eof, "address beyond the end of the file"
        return self.buf[addr:addr + n]

    def _heap_string(self, heap_addr, off):
        h = self._at(heap_addr, 32)
        assert h[:4] == b"HEAP" and h[4] == 0
        size, free, data = struct.unpack("<QQQ", h[8:32])
        assert free == 1 or free < size
        seg = self._at(data, size)
        end = seg.index(b"\0", off)
        return seg[off:end].decode("ascii")

    def _group_entries(self, bt_addr, heap_addr):
        t = self._at(bt_addr, 24)
        assert t[:4] == b"TREE" and t[4] == 0, "group B-tree node"
        level, used = t[5], _u(t, 6, 2)
        assert struct.unpack("<QQ", t[8:24]) == (UNDEF, UNDEF) and used <= 2 * self.internal_k
        body = self._at(bt_addr + 24, (2 * used + 1) * 8)
        out, prev_key = [], ""
        for i in range(used):
            key_l, child, key_r = struct.unpack("<QQQ", body[16 * i:16 * i + 24])
            if level > 0:
                sub = self._group_entries(child, heap_addr)
            else:
                s = self._at(child, 8 + 40 * 2 * self.leaf_k)
                assert s[:4] == b"SNOD" and s[4] == 1
                n = _u(s, 6, 2)
                assert n <= 2 * self.leaf_k
                sub = []
                for j in range(n):
                    e = s[8 + 40 * j:48 + 40 * j]
                    noff, hdr, cache, _ = struct.unpack("<QQII", e[:24])
                    sub.append((self._heap_string(heap_addr, noff), hdr, cache, struct.unpack("<QQ", e[24:40])))
            # keys bracket the names of the child: key_l < name <= key_r (strcmp order)
            names = [x[0] for x in sub]
            assert names == sorted(names, key=lambda q: q.encode())
            assert self._heap_string(heap_addr, key_l) == prev_key
            prev_key = self._heap_string(heap_addr, key_r)
            assert names and names[-1] == prev_key and (not out or out[-1][0] < names[0])
            out += sub
        return out

    def _object(self, addr):
        p = self._at(addr, 16)
        assert p[0] == 1, "object header version"
        nmsg, refs, size = _u(p, 2, 2), _u(p, 4, 4), _u(p, 8, 4)
        assert refs == 1
        data = self._at(addr + 16, size)
        node, off = Node(), 0
        shape = dtype = layout = None
        node._stab = None
        for _ in range(nmsg):
            mtype, msize = _u(data, off, 2), _u(data, off + 2, 2)
            assert msize % 8 == 0
            m = data[off + 8:off + 8 + msize]
            off += 8 + msize
            if mtype == 0x0001:
                shape = _decode_space(m)
            elif mtype == 0x0003:
                dtype, _ = _decode_type(m)
            elif mtype == 0x0005:
                assert m[0] == 2 and (m[3] == 0 or _u(m, 4, 4) == 0), "fill value with data"
            elif mtype == 0x0008:
                assert m[0] == 3 and m[1] == 1, "layout is not contiguous version 3"
                layout = struct.unpack("<QQ", m[2:18])
            elif mtype == 0x0011:
                node._stab = struct.unpack("<QQ", m[:16])
            elif mtype == 0x000C:
                assert m[0] == 1, "attribute version"
                nsz, tsz, ssz = _u(m, 2, 2), _u(m, 4, 2), _u(m, 6, 2)
                o = 8
                name = m[o:o + nsz].rstrip(b"\0").decode("ascii"); o += nsz + (-nsz % 8)
                adt, used = _decode_type(m[o:o + tsz]); assert used == tsz; o += tsz + (-tsz % 8)
                ashape = _decode_space(m[o:o + ssz]); o += ssz + (-ssz % 8)
                cnt = int(np.prod(ashape)) if ashape else 1
                val = np.frombuffer(m[o:o + cnt * adt.itemsize], dtype=adt).reshape(ashape)
                node.attrs[name] = val[()].decode("ascii") if adt.kind == "S" and not ashape else val.copy()
            elif mtype != 0:
                raise AssertionError(f"unexpected header message 0x{mtype:04x}")
        assert off == size
        if node._stab is not None:
            node.children = {}
            for name, hdr, cache, scratch in self._group_entries(*node._stab):
                child = self._object(hdr)
                if cache == 1:
                    assert scratch == child._stab
                node.children[name] = child
        else:
            assert shape is not None and dtype is not None and layout is not None, "dataset messages"
            nbytes = int(np.prod(shape)) * dtype.itemsize if shape else dtype.itemsize
            assert layout[1] == nbytes
            raw = self._at(layout[0], nbytes) if nbytes else b""
            node.data = np.frombuffer(raw, dtype=dtype).reshape(shape).copy()
        return node

    def get(self, path):
        n = self.root
        for p in [q for q in path.split("/") if q]:
            n = n.children[p]
        return n

    def tree(self, node=None, prefix=""):
        """{path: array} of every dataset."""
        node = node or self.root
        out = {}
        for name, c in node.children.items():
            if c.children is None:
                out[prefix + "/" + name] = c.data
            else:
                out.update(self.tree(c, prefix + "/" + name))
        return out
