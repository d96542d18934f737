"""TensorFlow tensor-bundle (checkpoint V2) writer / reader without TensorFlow.

`holE.py` saves and restores its model with `tf.train.Saver` (holE.py:308, 313-314, 359):
`<prefix>.index` (a LevelDB-format SSTable of `BundleHeaderProto` / `BundleEntryProto`
records), `<prefix>.data-00000-of-00001` (the tensors back to back in key order) and a
`checkpoint` text proto.  Format decoded from the archived runs; the
writer reproduces `holE-20170712/model.ckpt.index` and `holE-20170724/model.ckpt.index`
byte for byte when given their entries (tests/test_tf_bundle.py).
"""
import os
import struct

import numpy as np

DT_FLOAT, DT_INT32, DT_STRING, DT_INT64 = 1, 3, 7, 9
_NP = {DT_FLOAT: np.float32, DT_INT32: np.int32, DT_INT64: np.int64}
_MAGIC = bytes.fromhex("57fb808b247547db")
_RESTART_INTERVAL = 16
_MASK_DELTA = 0xA282EAD8


# ------------------------------------------------------------------ crc32c
def _crc32c_py(data, crc=0):
    tab = _crc32c_py.table
    if tab is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _crc32c_py.table = tab
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = (c >> 8) ^ tab[(c ^ b) & 0xFF]
    return c ^ 0xFFFFFFFF


_crc32c_py.table = None


def crc32c(data, crc=0):
    """CRC32C of bytes / a C-contiguous numpy array (native helper for large buffers)."""
    if isinstance(data, np.ndarray):
        if data.nbytes > 4096:
            try:
                import ctypes
                from . import _lib
                return int(_lib.load().hole_crc32c(crc, data.ctypes.data_as(ctypes.c_void_p), data.nbytes))
            except Exception:
                pass
        data = data.tobytes()
    return _crc32c_py(data, crc)


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------ protobuf (the 2 messages we need)
def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def encode_header(num_shards=1, producer=1):
    """BundleHeaderProto{num_shards=1; version{producer=1}} (endianness LITTLE = default)."""
    ver = b"\x08" + _varint(producer)
    return b"\x08" + _varint(num_shards) + b"\x1a" + _varint(len(ver)) + ver


def encode_entry(dtype, shape, offset, size, crc_masked, shard_id=0):
    """BundleEntryProto: dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32).
    Zero-valued scalar fields are omitted, as proto3 does."""
    dims = b"".join(b"\x12" + _varint(len(d)) + d for d in (b"\x08" + _varint(int(s)) for s in shape))
    out = b"\x08" + _varint(dtype) + b"\x12" + _varint(len(dims)) + dims
    if shard_id:
        out += b"\x18" + _varint(shard_id)
    if offset:
        out += b"\x20" + _varint(offset)
    if size:
        out += b"\x28" + _varint(size)
    out += b"\x35" + struct.pack("<I", crc_masked)
    return out


def decode_entry(buf):
    pos, e = 0, {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
            e[{1: "dtype", 3: "shard_id", 4: "offset", 5: "size"}.get(field, f"f{field}")] = v
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
            if field == 6:
                e["crc32c"] = v
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            sub = buf[pos:pos + ln]
            pos += ln
            if field == 2:                       # TensorShapeProto: repeated dim{size}
                p2 = 0
                while p2 < len(sub):
                    t2, p2 = _read_varint(sub, p2)
                    l2, p2 = _read_varint(sub, p2)
                    d = sub[p2:p2 + l2]
                    p2 += l2
                    if t2 >> 3 == 2:
                        if d:
                            _, q = _read_varint(d, 0)
                            sz, _ = _read_varint(d, q)
                        else:
                            sz = 0
                        e["shape"].append(sz)
        else:
            raise ValueError(f"unsupported wire type {wt}")
    return e


# ------------------------------------------------------------------ SSTable (LevelDB table format)
def _block(entries):
    """Prefix-compressed block with restart points every 16 entries."""
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % _RESTART_INTERVAL == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block):
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


def _short_successor(key):
    """leveldb BytewiseComparator::FindShortSuccessor."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_index(path, entries):
    """entries: list of (key bytes, value bytes) sorted by key, key b"" (header) first."""
    data = _block(entries)
    meta = _block([])
    out = bytearray(_with_trailer(data))
    meta_off = len(out)
    out += _with_trailer(meta)
    index = _block([(_short_successor(entries[-1][0]), _varint(0) + _varint(len(data)))])
    index_off = len(out)
    out += _with_trailer(index)
    footer = _varint(meta_off) + _varint(len(meta)) + _varint(index_off) + _varint(len(index))
    out += footer + b"\x00" * (40 - len(footer)) + _MAGIC
    with open(path, "wb") as f:
        f.write(out)


def read_index(path):
    """-> list of (key bytes, value bytes) of the (single) data block, in file order."""
    buf = open(path, "rb").read()
    if buf[-8:] != _MAGIC:
        raise ValueError(f"{path}: not an SSTable (bad magic)")
    foot = buf[-48:]
    pos = 0
    _, pos = _read_varint(foot, pos); _, pos = _read_varint(foot, pos)
    ioff, pos = _read_varint(foot, pos); isz, pos = _read_varint(foot, pos)

    def parse(block):
        n_rest = struct.unpack_from("<I", block, len(block) - 4)[0]
        end = len(block) - 4 - 4 * n_rest
        pos, last, out = 0, b"", []
        while pos < end:
            sh, pos = _read_varint(block, pos); ns, pos = _read_varint(block, pos)
            vl, pos = _read_varint(block, pos)
            key = last[:sh] + block[pos:pos + ns]; pos += ns
            out.append((key, block[pos:pos + vl])); pos += vl
            last = key
        return out

    entries = []
    for _, handle in parse(buf[ioff:ioff + isz]):
        off, p = _read_varint(handle, 0)
        sz, _ = _read_varint(handle, p)
        block = buf[off:off + sz]
        want = struct.unpack_from("<I", buf, off + sz + 1)[0]
        if mask_crc(crc32c(block + buf[off + sz:off + sz + 1])) != want:
            raise ValueError(f"{path}: data block checksum mismatch")
        entries += parse(block)
    return entries


# ------------------------------------------------------------------ bundle level
def save_bundle(prefix, tensors):
    """tensors: dict name -> numpy array (float32 / int32 / int64).  Writes
    `<prefix>.index` and `<prefix>.data-00000-of-00001` (tensors in key order)."""
    names = sorted(tensors)
    entries, offset = [(b"", encode_header())], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in names:
            a = np.ascontiguousarray(tensors[name])
            dt = {np.dtype(np.float32): DT_FLOAT, np.dtype(np.int32): DT_INT32,
                  np.dtype(np.int64): DT_INT64}[a.dtype]
            a.tofile(f)
            entries.append((name.encode(), encode_entry(dt, a.shape, offset, a.nbytes, mask_crc(crc32c(a)))))
            offset += a.nbytes
    write_index(prefix + ".index", entries)


def load_bundle(prefix, names=None, verify=True):
    """-> dict name -> numpy array for the numeric tensors of a bundle."""
    out = {}
    data_path = prefix + ".data-00000-of-00001"
    for key, val in read_index(prefix + ".index"):
        if key == b"":
            continue
        name = key.decode()
        if names is not None and name not in names:
            continue
        e = decode_entry(val)
        if e["dtype"] not in _NP:
            continue
        a = np.fromfile(data_path, dtype=_NP[e["dtype"]], count=e["size"] // np.dtype(_NP[e["dtype"]]).itemsize,
                        offset=e["offset"]).reshape(e["shape"])
        if verify and mask_crc(crc32c(a)) != e["crc32c"]:
            raise ValueError(f"{prefix}: checksum mismatch for tensor {name!r}")
        out[name] = a
    return out


def write_checkpoint_state(output_dir, name="model.ckpt"):
    """The `checkpoint` text proto tf.train.Saver maintains (holE-20170712/checkpoint:1-6)."""
    with open(os.path.join(output_dir, "checkpoint"), "w") as f:
        f.write(f'model_checkpoint_path: "{name}"\nall_model_checkpoint_paths: "{name}"\n')


def write_projector_config(output_dir, tensor_name="embeddings:0", metadata_path=None, tensor_path=None):
    """projector_config.pbtxt (holE.py:334-337; holE-20170712/projector_config.pbtxt:1-3).
    `tensor_path` (a TSV of the vectors) lets TensorBoard load them without TensorFlow."""
    lines = ["embeddings {", f'  tensor_name: "{tensor_name}"']
    if metadata_path:
        lines.append(f'  metadata_path: "{metadata_path}"')
    if tensor_path:
        lines.append(f'  tensor_path: "{tensor_path}"')
    lines.append("}")
    with open(os.path.join(output_dir, "projector_config.pbtxt"), "w") as f:
        f.write("\n".join(lines) + "\n")
