"""Minimal HDF5 reader/writer for the reference's training files (h5py is not in this image).

The reference reads its training sets with `h5py.File(name)['data'|'label'|'pid'][:]` (S3DIS/DataIO_S3DIS.py:38-43,
ShapeNet/DataIO_ShapeNet.py:315-324).  Those files (PointNet's `indoor3d_sem_seg_hdf5_data/ply_data_all_*.h5` and
`hdf5_data/ply_data_{train,val,test}*.h5`) were written by h5py with the library's oldest ("earliest") format, as
Networks/dgcnn/utils/data_prep_util.py:59-103 does: `create_dataset(name, data=..., compression='gzip', compression_opts=4|1)`.
That subset of the HDF5 file format is what this module restates from the published HDF5 File Format Specification (v1.x /
2.0, sections III.A superblock v0, III.B v1 B-trees, III.C symbol-table nodes, III.D local heaps, IV.A v1 object headers):

  superblock v0/v1 -> root symbol-table entry -> group B-tree (node type 0) -> SNOD leaves -> object headers (v1) with
  dataspace (0x01), datatype (0x03, fixed-point / IEEE float), layout (0x08, v3: compact / contiguous / chunked),
  filter pipeline (0x0B: deflate, shuffle, fletcher32) and chunk B-trees (node type 1).

`read(path)` returns {dataset name: ndarray} for the datasets of the root group; `File(path)[name][:]` mirrors the h5py idiom
the reference uses.  When h5py is importable it is used instead (`read` / `File` dispatch to it).  `write(path, {...})` emits
the same on-disk subset (chunked + gzip, one-level B-trees) so loaders can be exercised without h5py.

PARITY: unpinned — no HDF5 library or HDF5 file exists in this image to check the restatement against; the only check is the
writer/reader round trip (tests/test_dataio_cpu.py).  Anything outside the subset raises NotImplementedError loudly.
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


def _have_h5py():
    try:
        import h5py
        return hasattr(h5py, "Dataset")                          # a real h5py, not a stand-in built on this module
    except Exception:
        return False


class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file (signature at offset 0 missing; user blocks are not supported)")
        ver = buf[8]
        if ver not in (0, 1):
            raise NotImplementedError("HDF5 superblock version %d (only the 'earliest' format v0/v1 is restated here)" % ver)
        self.O, self.L = buf[13], buf[14]
        if self.O != 8 or self.L != 8:
            raise NotImplementedError("HDF5 offsets/lengths of %d/%d bytes" % (self.O, self.L))
        p = 24 + (4 if ver == 1 else 0)                          # v1 adds indexed-storage K + reserved
        self.base, _free, _eof, _drv = struct.unpack_from("<4Q", buf, p)
        p += 32
        _name_off, self.root_hdr, cache, _r = struct.unpack_from("<QQII", buf, p)
        self.root_scratch = struct.unpack_from("<QQ", buf, p + 24) if cache == 1 else None

    # ---- object headers (v1) -------------------------------------------------------------------------------------------
    def messages(self, addr):
        b = self.b
        addr += self.base
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError("HDF5 object header version %d (v2 'OHDR' headers need libver='latest' files)" % ver)
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:                                # continuation
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off + self.base, ln))
                out.append((mtype, body))
        return out

    # ---- groups ---------------------------------------------------------------------------------------------------------
    def _heap_data(self, addr):
        b = self.b
        addr += self.base
        if b[addr:addr + 4] != b"HEAP":
            raise ValueError("HDF5 local heap signature missing")
        _size, _free, data = struct.unpack_from("<QQQ", b, addr + 8)
        return data + self.base

    def _group_leaves(self, addr):
        b = self.b
        addr += self.base
        if b[addr:addr + 4] == b"SNOD":
            yield addr
            return
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 0:
            raise ValueError("HDF5 group B-tree node expected")
        n = struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 8 + 16
        for i in range(n):
            child = struct.unpack_from("<Q", b, p + 8 + i * 16)[0]
            yield from self._group_leaves(child)

    def links(self, btree, heap):
        b = self.b
        hd = self._heap_data(heap)
        out = {}
        for leaf in self._group_leaves(btree):
            n = struct.unpack_from("<H", b, leaf + 6)[0]
            for i in range(n):
                name_off, hdr = struct.unpack_from("<QQ", b, leaf + 8 + 40 * i)
                e = hd + name_off
                while b[e]:
                    e += 1
                out[bytes(b[hd + name_off:e]).decode()] = hdr
        return out

    def root_links(self):
        if self.root_scratch is not None:
            return self.links(*self.root_scratch)
        for mtype, body in self.messages(self.root_hdr):
            if mtype == 0x11:
                return self.links(*struct.unpack_from("<QQ", body, 0))
        raise NotImplementedError("HDF5 root group without a symbol table (link-message groups need libver='latest')")

    # ---- datasets ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(body):
        cls, ver = body[0] & 15, body[0] >> 4
        bits0 = body[1]
        size = struct.unpack_from("<I", body, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if bits0 & 8 else "u", size))
        if cls == 1:
            if size not in (2, 4, 8):
                raise NotImplementedError("HDF5 float of %d bytes" % size)
            return np.dtype("%sf%d" % (order, size))
        raise NotImplementedError("HDF5 datatype class %d (v%d); the reference's files hold fixed-point and float only" % (cls, ver))

    @staticmethod
    def _shape(body):
        ver, rank = body[0], body[1]
        p = 8 if ver == 1 else 4
        return struct.unpack_from("<%dQ" % rank, body, p) if rank else ()

    @staticmethod
    def _filters(body):
        ver, n = body[0], body[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", body, p)[0]
            if ver == 1 or fid >= 256:
                nlen, flags, ncd = struct.unpack_from("<HHH", body, p + 2)
                p += 8
                p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            else:
                flags, ncd = struct.unpack_from("<HH", body, p + 2)
                p += 6
            cd = struct.unpack_from("<%dI" % ncd, body, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunks(self, addr, rank):
        """Yield (offsets, size, filter_mask, address) of every chunk under a type-1 B-tree node."""
        b = self.b
        addr += self.base
        if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 1:
            raise ValueError("HDF5 chunk B-tree node expected")
        level, n = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        ksz = 8 + 8 * (rank + 1)
        p = addr + 24
        for i in range(n):
            q = p + i * (ksz + 8)
            size, mask = struct.unpack_from("<II", b, q)
            offs = struct.unpack_from("<%dQ" % rank, b, q + 8)
            child = struct.unpack_from("<Q", b, q + ksz)[0]
            if level:
                yield from self._chunks(child, rank)
            else:
                yield offs, size, mask, child + self.base

    def dataset(self, hdr):
        shape = dtype = layout = None
        filters = []
        for mtype, body in self.messages(hdr):
            if mtype == 0x01:
                shape = self._shape(body)
            elif mtype == 0x03:
                dtype = self._dtype(body)
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                filters = self._filters(body)
        if shape is None or dtype is None or layout is None:
            return None                                          # not a dataset (a sub-group)
        if layout[0] != 3:
            raise NotImplementedError("HDF5 data layout message version %d" % layout[0])
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        cls = layout[1]
        if cls == 0:                                             # compact
            n = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout[4:4 + n], dtype, count).reshape(shape).copy()
        if cls == 1:                                             # contiguous
            addr, n = struct.unpack_from("<QQ", layout, 2)
            if addr == _UNDEF:
                return np.zeros(shape, dtype)
            return np.frombuffer(self.b, dtype, count, addr + self.base).reshape(shape).copy()
        if cls != 2:
            raise NotImplementedError("HDF5 layout class %d" % cls)
        rank = layout[2] - 1
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<%dI" % rank, layout, 11)
        out = np.zeros(shape, dtype)
        if btree == _UNDEF:
            return out
        for offs, size, mask, addr in self._chunks(btree, rank):
            raw = bytes(self.b[addr:addr + size])
            for j in range(len(filters) - 1, -1, -1):            # undo the pipeline back to front
                if mask >> j & 1:
                    continue
                fid, cd = filters[j]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    a = np.frombuffer(raw, np.uint8)
                    m = a.size // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]                               # fletcher32 checksum trails the chunk
                else:
                    raise NotImplementedError("HDF5 filter id %d" % fid)
            chunk = np.frombuffer(raw, dtype, int(np.prod(cdims))).reshape(cdims)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = chunk[sl_in]
        return out


def read(path):
    """{name: ndarray} for every dataset of the root group (h5py when importable, else the restated subset)."""
    if _have_h5py():
        import h5py
        with h5py.File(path, "r") as f:
            return {k: f[k][:] for k in f.keys() if isinstance(f[k], h5py.Dataset)}
    with open(path, "rb") as fh:
        buf = memoryview(fh.read())
    r = _Reader(buf)
    out = {}
    for name, hdr in r.root_links().items():
        a = r.dataset(hdr)
        if a is not None:
            out[name] = a.astype(a.dtype.newbyteorder("="), copy=False)
    return out


class _Dataset:
    def __init__(self, a):
        self._a = a
        self.shape, self.dtype = a.shape, a.dtype

    def __getitem__(self, k):
        return self._a[k]


class File:
    """`File(name, 'r')['data'][:]` — the h5py idiom of the reference's loaders (read-only)."""

    def __init__(self, path, mode="r"):
        if mode != "r":
            raise ValueError("wspc _h5.File is read-only; use _h5.write")
        self._d = read(path)

    def __getitem__(self, k):
        return _Dataset(self._d[k])

    def keys(self):
        return self._d.keys()

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ---- writer (same subset: superblock v0, one SNOD leaf, chunked + deflate datasets with one-level chunk B-trees) ----------

def _msg(mtype, body, flags=0):
    body = body + b"\0" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = (8 if dt.kind == "i" else 0)
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 4:
            prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            b0, b1 = 0x20, 31
        else:
            prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            b0, b1 = 0x20, 63
        return struct.pack("<BBBBI", 0x10 | 1, b0, b1, 0, dt.itemsize) + prop
    raise NotImplementedError("dtype %s" % dt)


def write(path, arrays, chunk_rows=64, level=4, chunk_shape=None, shuffle=False, node_entries=64):
    """Write {name: ndarray} as chunked, gzip-compressed datasets.  Default chunks span `chunk_rows` of axis 0 and all of the
    rest; `chunk_shape` (a tuple per rank, e.g. {3: (8, 100, 2)}) gives multi-dimensional chunks with ragged edges like
    h5py's auto-chunking of data_prep_util.py:64-77; `shuffle` adds the byte-shuffle filter in front of deflate; more than
    `node_entries` chunks make a two-level chunk B-tree (HDF5's own nodes hold 64 entries)."""
    names = sorted(arrays)
    blob = bytearray(b"\0" * 96)                                 # superblock v0 (56 + 40-byte root entry)

    def alloc(data):
        while len(blob) % 8:
            blob.append(0)
        at = len(blob)
        blob.extend(data)
        return at

    hdr_addr = {}
    for name in names:
        a = np.ascontiguousarray(arrays[name])
        a = a.astype(a.dtype.newbyteorder("<"), copy=False)
        rank = a.ndim
        cdims = (min(chunk_rows, max(a.shape[0], 1)),) + a.shape[1:]
        if chunk_shape and rank in chunk_shape:
            cdims = tuple(chunk_shape[rank])
        entries = []
        grid = [range(0, max(s_, 1), c) for s_, c in zip(a.shape, cdims)]
        for offs in np.ndindex(*[len(g) for g in grid]):
            o = tuple(g[i] for g, i in zip(grid, offs))
            ch = np.zeros(cdims, a.dtype)
            part = a[tuple(slice(x, x + c) for x, c in zip(o, cdims))]
            ch[tuple(slice(0, n) for n in part.shape)] = part
            raw = ch.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, a.dtype.itemsize).T.tobytes()
            z = zlib.compress(raw, level)
            entries.append((o, len(z), alloc(z)))

        def key(offs, size):
            return struct.pack("<II", size, 0) + struct.pack("<%dQ" % (rank + 1), *offs, 0)

        closing = key(tuple(a.shape), 0)

        def leaf(part, last_key):
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(part), _UNDEF, _UNDEF))
            for offs, size, addr in part:
                node += key(offs, size) + struct.pack("<Q", addr)
            return alloc(node + last_key)

        if len(entries) <= node_entries:
            node = None
            bt_addr = leaf(entries, closing)
        else:                                                    # level-1 node over leaves of `node_entries` chunks
            groups = [entries[i:i + node_entries] for i in range(0, len(entries), node_entries)]
            kids = []
            for gi, g in enumerate(groups):
                nxt = key(groups[gi + 1][0][0], groups[gi + 1][0][1]) if gi + 1 < len(groups) else closing
                kids.append((g[0], leaf(g, nxt)))
            node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 1, len(kids), _UNDEF, _UNDEF))
            for (offs, size, _a), addr in kids:
                node += key(offs, size) + struct.pack("<Q", addr)
            bt_addr = alloc(node + closing)
        bt = bt_addr
        msgs = _msg(0x01, struct.pack("<BBB5x", 1, rank, 1) + struct.pack("<%dQ" % rank, *a.shape)
                    + struct.pack("<%dQ" % rank, *a.shape))
        msgs += _msg(0x03, _dtype_msg(a.dtype), flags=1)
        deflate = struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + struct.pack("<II", level, 0)
        if shuffle:
            msgs += _msg(0x0B, struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\0"
                         + struct.pack("<II", a.dtype.itemsize, 0) + deflate)
        else:
            msgs += _msg(0x0B, struct.pack("<BB6x", 1, 1) + deflate)
        msgs += _msg(0x08, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", bt)
                     + struct.pack("<%dI" % (rank + 1), *cdims, a.dtype.itemsize))
        hdr_addr[name] = alloc(struct.pack("<BBHII4x", 1, 0, 4, 1, len(msgs)) + msgs)

    heap = bytearray(b"\0" * 8)                                  # offset 0 = the empty name of the root entry
    name_off = {}
    for name in names:
        name_off[name] = len(heap)
        heap += name.encode() + b"\0"
        heap += b"\0" * (-len(heap) % 8)
    heap_data = alloc(heap)
    heap_addr = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), _UNDEF, heap_data))
    snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
    for name in names:
        snod += struct.pack("<QQII16x", name_off[name], hdr_addr[name], 0, 0)
    snod_addr = alloc(snod)
    gt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, _UNDEF, _UNDEF) + struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]])
    gt_addr = alloc(gt)
    root_msgs = _msg(0x11, struct.pack("<QQ", gt_addr, heap_addr))
    root_hdr = alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(root_msgs)) + root_msgs)
    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, len(blob), _UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", gt_addr, heap_addr)
    blob[:len(sb)] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(blob))
