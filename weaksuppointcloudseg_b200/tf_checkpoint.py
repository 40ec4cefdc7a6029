"""Reader (and test-only writer) for TensorFlow "V2" checkpoints, the format `tf.train.Saver.save` writes in the reference
(S3DIS/S3DIS_DGCNN_trainer.py:586-590, ShapeNet/ShapeNet_DGCNN_trainer.py:601-605): `<prefix>.index` +
`<prefix>.data-00000-of-00001` (+ `.meta`, the graph, which is not needed).

TensorFlow is not installable in this image, so the format is restated from its published layout:
  * `.index` is a LevelDB-style sorted string table (tensorflow/core/lib/io/table_builder.cc, format.cc): data blocks,
    a meta-index block, an index block, and a 48-byte footer [meta-index handle | index handle | padding | magic
    0xdb4775248b80fb57].  A block is a run of prefix-compressed entries [shared varint32 | non_shared varint32 |
    value_len varint32 | key suffix | value], a restart array of uint32 and its uint32 count; every block is followed by a
    5-byte trailer (compression type, masked crc32c) that its handle does not count.
  * the key "" holds a BundleHeaderProto (num_shards = 1, endianness = 2, version = 3); every other key is a variable name and
    its value a BundleEntryProto (tensor_bundle.proto): dtype = 1, shape = 2 (TensorShapeProto: repeated dim {size = 1}),
    shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32), slices = 7.
  * the data shard holds each tensor's bytes, little-endian, row-major, at [offset, offset + size).

`read(prefix)` -> {variable name: ndarray}.  Only uncompressed blocks, unpartitioned variables and numeric dtypes are
handled; anything else raises.  The tensor crc32c is not verified (sizes and shapes are).

PARITY: unpinned -- no TensorFlow and no TF checkpoint exists in this image; the check is the writer/reader round trip of
tests/test_tf_checkpoint_cpu.py.  The variable names of the reference's graphs are the ones `VariableStore.export()` uses
(`<scope>/weights`, `<scope>/biases`, `<scope>/bn/{beta,gamma,pop_mean,pop_var}`, `Variable` = global step;
tf_util.py:100-101, :515-519, S3DIS_DGCNN_trainer.py:30), Adam slots are `<var>/Adam` and `<var>/Adam_1`.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


def _varint(buf, p):
    out = shift = 0
    while True:
        b = buf[p]
        p += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, p
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _block_entries(buf, offset, size):
    """(key, value) pairs of the block at [offset, offset + size); the trailer byte after it must say 'no compression'."""
    if buf[offset + size] != 0:
        raise NotImplementedError("compressed checkpoint index block (type %d)" % buf[offset + size])
    n_restarts = struct.unpack_from("<I", buf, offset + size - 4)[0]
    end = offset + size - 4 - 4 * n_restarts
    p, key = offset, b""
    while p < end:
        shared, p = _varint(buf, p)
        non_shared, p = _varint(buf, p)
        vlen, p = _varint(buf, p)
        key = key[:shared] + bytes(buf[p:p + non_shared])
        p += non_shared
        yield key, bytes(buf[p:p + vlen])
        p += vlen


def _proto_fields(b):
    """Minimal protobuf wire decoder: yields (field number, wire type, value)."""
    p = 0
    while p < len(b):
        tag, p = _varint(b, p)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, p = _varint(b, p)
        elif wt == 1:
            v = b[p:p + 8]
            p += 8
        elif wt == 2:
            n, p = _varint(b, p)
            v = b[p:p + n]
            p += n
        elif wt == 5:
            v = b[p:p + 4]
            p += 4
        else:
            raise ValueError("protobuf wire type %d" % wt)
        yield f, wt, v


def _entry(value):
    e = dict(dtype=0, shape=(), shard=0, offset=0, size=0, sliced=False)
    for f, wt, v in _proto_fields(value):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            dims = []
            for f2, _, v2 in _proto_fields(v):
                if f2 == 2:                                      # repeated Dim
                    size = 0
                    for f3, _, v3 in _proto_fields(v2):
                        if f3 == 1:
                            size = v3
                    dims.append(size)
                elif f2 == 3 and v2:
                    raise NotImplementedError("tensor of unknown rank in a checkpoint")
            e['shape'] = tuple(dims)
        elif f == 3:
            e['shard'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 7:
            e['sliced'] = True
    return e


def index(prefix):
    """{variable name: entry dict} from `<prefix>.index` (the "" header entry is checked and dropped)."""
    with open(prefix + '.index', 'rb') as fh:
        buf = fh.read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != _MAGIC:
        raise ValueError("%s.index is not a TensorFlow V2 checkpoint index (table magic missing)" % prefix)
    p = len(buf) - 48
    _mo, p = _varint(buf, p)
    _ms, p = _varint(buf, p)
    io, p = _varint(buf, p)
    isz, p = _varint(buf, p)
    out = {}
    for _sep, handle in _block_entries(buf, io, isz):
        bo, q = _varint(handle, 0)
        bs, q = _varint(handle, q)
        for key, value in _block_entries(buf, bo, bs):
            if key == b"":
                hdr = {f: v for f, _, v in _proto_fields(value)}
                if hdr.get(1, 1) != 1:
                    raise NotImplementedError("checkpoint with %d data shards" % hdr[1])
                if hdr.get(2, 0) != 0:
                    raise NotImplementedError("big-endian checkpoint")
                continue
            out[key.decode()] = _entry(value)
    return out


def read(prefix, names=None):
    """{name: ndarray} for every (or the named) variable of the checkpoint `<prefix>`."""
    idx = index(prefix)
    data_path = prefix + '.data-00000-of-00001'
    out = {}
    with open(data_path, 'rb') as fh:
        for name, e in idx.items():
            if names is not None and name not in names:
                continue
            if e['sliced'] or e['shard'] != 0:
                raise NotImplementedError("partitioned variable %s" % name)
            if e['dtype'] not in _DTYPES:
                raise NotImplementedError("dtype %d of %s" % (e['dtype'], name))
            dt = np.dtype(_DTYPES[e['dtype']])
            count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
            if count * dt.itemsize != e['size']:
                raise ValueError("%s: %d bytes stored, shape %s of %s needs %d" % (name, e['size'], e['shape'], dt,
                                                                                 count * dt.itemsize))
            fh.seek(e['offset'])
            raw = fh.read(e['size'])
            if len(raw) != e['size']:
                raise ValueError("%s: data shard truncated" % name)
            out[name] = np.frombuffer(raw, dt.newbyteorder('<'), count).reshape(e['shape']).astype(dt)
    return out


def exists(prefix):
    return os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')


# ---- writer of the same subset (TEST FIXTURES ONLY: block and tensor checksums are written as zero, which TensorFlow's own
#      BundleReader rejects -- these files are for this module's reader, not for TF tooling) ----------------------------------

def _field(f, wt, payload):
    tag = _put_varint(f << 3 | wt)
    if wt == 0:
        return tag + _put_varint(payload)
    if wt == 2:
        return tag + _put_varint(len(payload)) + payload
    return tag + payload


def _block(pairs, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write(prefix, arrays, entries_per_block=32):
    """Write {name: ndarray} as `<prefix>.index` + `<prefix>.data-00000-of-00001` (block trailers carry a zero crc)."""
    names = sorted(arrays, key=lambda s: s.encode())
    data, pairs = bytearray(), [(b"", _field(1, 0, 1) + _field(3, 2, _field(1, 0, 1)))]
    for name in names:
        a = np.asarray(arrays[name]).copy(order='C')             # (ascontiguousarray would turn a scalar into (1,))
        a = a.astype(a.dtype.newbyteorder('<'), copy=False)
        shape = b"".join(_field(2, 2, _field(1, 0, int(d))) for d in a.shape)
        e = _field(1, 0, _DTYPE_IDS[np.dtype(a.dtype.name)]) + _field(2, 2, shape)
        if len(data):
            e += _field(4, 0, len(data))
        e += _field(5, 0, a.nbytes) + _field(6, 5, struct.pack("<I", 0))
        pairs.append((name.encode(), e))
        data += a.tobytes()
    blob, handles = bytearray(), []
    for i in range(0, len(pairs), entries_per_block):
        chunk = pairs[i:i + entries_per_block]
        blk = _block(chunk)
        handles.append((chunk[-1][0], len(blob), len(blk)))
        blob += blk + b"\0" * 5
    meta = _block([])
    meta_handle = (len(blob), len(meta))
    blob += meta + b"\0" * 5
    idx_blk = _block([(k, _put_varint(o) + _put_varint(s)) for k, o, s in handles], restart_interval=1)
    idx_handle = (len(blob), len(idx_blk))
    blob += idx_blk + b"\0" * 5
    footer = _put_varint(meta_handle[0]) + _put_varint(meta_handle[1]) + _put_varint(idx_handle[0]) + _put_varint(idx_handle[1])
    blob += footer + b"\0" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    with open(prefix + '.index', 'wb') as fh:
        fh.write(bytes(blob))
    with open(prefix + '.data-00000-of-00001', 'wb') as fh:
        fh.write(bytes(data))


# ---- mapping onto the flat variable store ------------------------------------------------------------------------------------

def to_store_blob(tf_vars, trainable_names, state_names, shapes, allow_missing_optimizer_state=False):
    """Arrange the variables of a reference checkpoint the way `S3DIS_Trainer.RestoreCheckPoint` consumes its own `.npz`:
    every trainable / state variable by name (shape-checked), `Variable` (global step) and the Adam moments as flat buffers
    in the store's order (`<var>/Adam`, `<var>/Adam_1`).  A checkpoint without the global step or without Adam slots would
    silently restart the learning-rate / batch-norm schedules and Adam's bias correction: that is an error unless
    `allow_missing_optimizer_state` is set (then they start from zero, with a warning)."""
    missing = [k + slot for k in trainable_names for slot in ('/Adam', '/Adam_1') if k + slot not in tf_vars]
    if 'Variable' not in tf_vars:
        missing.append('Variable (global step)')
    if missing:
        msg = ("TensorFlow checkpoint lacks optimiser state (%d entries, e.g. %s): resuming would restart the schedules and "
               "Adam's moments" % (len(missing), missing[0]))
        if not allow_missing_optimizer_state:
            raise KeyError(msg + "; pass allow_missing_optimizer_state=True to load the weights only")
        import warnings
        warnings.warn(msg)
    blob = {}
    for k in list(trainable_names) + list(state_names):
        if k not in tf_vars:
            raise KeyError("variable %s is missing from the TensorFlow checkpoint" % k)
        a = np.asarray(tf_vars[k], np.float32)
        if tuple(a.shape) != tuple(shapes[k]):
            if a.size != int(np.prod(shapes[k])):
                raise ValueError("variable %s: checkpoint shape %s, graph shape %s" % (k, a.shape, tuple(shapes[k])))
            a = a.reshape(shapes[k])                             # conv kernels (1,1,Cin,Cout) vs (Cin,Cout)
        blob[k] = a
    blob['Variable'] = np.asarray(int(np.asarray(tf_vars.get('Variable', 0)).reshape(-1)[0]), np.int64)
    for slot, key in (('/Adam', '__adam_m'), ('/Adam_1', '__adam_v')):
        parts = []
        for k in trainable_names:
            n = int(np.prod(shapes[k]))
            part = np.zeros((n + 3) // 4 * 4, np.float32)        # VariableStore keeps every view 16-byte aligned
            a = tf_vars.get(k + slot)
            if a is not None:
                part[:n] = np.asarray(a, np.float32).reshape(-1)
            parts.append(part)
        blob[key] = np.concatenate(parts) if parts else np.zeros(0, np.float32)
    return blob
