"""TensorBoard event-file writer without TensorFlow: what `tf.summary.FileWriter(output_dir)` +
`summary_writer.add_summary(summary, step)` leave in --output_dir (holE.py:317, 353).

File format: a TFRecord stream -- per record `uint64 length | masked crc32c(length) | data | masked
crc32c(data)` -- of serialized `tensorflow.Event` protos.  Only the fields the reference's summaries use
are encoded (by hand; no protobuf dependency):

    Event   { double wall_time = 1; int64 step = 2; string file_version = 3; Summary summary = 5; }
    Summary { repeated Value value = 1; }
    Value   { string tag = 1; float simple_value = 2; HistogramProto histo = 5; }
    HistogramProto { double min = 1, max = 2, num = 3, sum = 4, sum_squares = 5;
                     repeated double bucket_limit = 6 [packed]; repeated double bucket = 7 [packed]; }

Tags follow the reference's name scopes (`summarize`, holE.py:237-246; the archived graph confirms the
`stddev_1` suffix: holE-20170724/graph.pbtxt:6833).
"""
import os
import socket
import struct
import time

import numpy as np

from .tf_bundle import crc32c, mask_crc


def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num, wire, payload):
    return _varint((num << 3) | wire) + payload


def _len_delim(num, data):
    return _field(num, 2, _varint(len(data)) + data)


def _double(num, x):
    return _field(num, 1, struct.pack("<d", float(x)))


def _default_buckets():
    """TensorFlow's histogram bucket limits (core/lib/histogram/histogram.cc): 1e-12 * 1.1^k up to 1e20,
    mirrored for negatives, plus DBL_MAX."""
    pos = []
    v = 1e-12
    while v < 1e20:
        pos.append(v)
        v *= 1.1
    pos.append(float(np.finfo(np.float64).max))
    neg = [-x for x in reversed(pos)]
    return np.array(neg + [0.0] + pos)


_BUCKETS = None


def encode_histogram(values):
    global _BUCKETS
    if _BUCKETS is None:
        _BUCKETS = _default_buckets()
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    idx = np.searchsorted(_BUCKETS, v, side="left")         # value <= limit[idx]
    counts = np.bincount(idx, minlength=len(_BUCKETS)).astype(np.float64)
    nz = np.flatnonzero(counts)
    lo, hi = (int(nz[0]), int(nz[-1])) if len(nz) else (0, 0)
    limits, buckets = _BUCKETS[lo:hi + 1], counts[lo:hi + 1]
    body = (_double(1, v.min() if len(v) else 0.0) + _double(2, v.max() if len(v) else 0.0) + _double(3, len(v)) +
            _double(4, v.sum()) + _double(5, (v * v).sum()) +
            _len_delim(6, struct.pack("<%dd" % len(limits), *limits)) +
            _len_delim(7, struct.pack("<%dd" % len(buckets), *buckets)))
    return body


def encode_event(wall_time, step=None, file_version=None, scalars=None, histograms=None):
    ev = _double(1, wall_time)
    if step is not None:
        ev += _field(2, 0, _varint(int(step)))
    if file_version is not None:
        ev += _len_delim(3, file_version.encode())
    if scalars or histograms:
        summ = b""
        for tag, val in (scalars or {}).items():
            summ += _len_delim(1, _len_delim(1, tag.encode()) + _field(2, 5, struct.pack("<f", float(val))))
        for tag, vals in (histograms or {}).items():
            summ += _len_delim(1, _len_delim(1, tag.encode()) + _len_delim(5, encode_histogram(vals)))
        ev += _len_delim(5, summ)
    return ev


def _record(data):
    head = struct.pack("<Q", len(data))
    return (head + struct.pack("<I", mask_crc(crc32c(np.frombuffer(head, dtype=np.uint8)))) + data +
            struct.pack("<I", mask_crc(crc32c(np.frombuffer(data, dtype=np.uint8)))))


class EventFileWriter:
    """events.out.tfevents.<time>.<host> in `logdir`, flushed after every summary."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self._f = open(self.path, "ab")
        self._f.write(_record(encode_event(time.time(), file_version="brain.Event:2")))
        self._f.flush()

    def add_summary(self, step, scalars=None, histograms=None):
        self._f.write(_record(encode_event(time.time(), step=step, scalars=scalars, histograms=histograms)))
        self._f.flush()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None


def summarize_tags(scope, values):
    """The scalars + histogram `summarize(var)` adds under a name scope (holE.py:237-246)."""
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    mean = float(v.mean())
    base = scope + "/summaries/"
    return ({base + "mean": mean, base + "stddev_1": float(np.sqrt(((v - mean) ** 2).mean())),
             base + "max": float(v.max()), base + "min": float(v.min())},
            {base + "histogram": v})


# ------------------------------------------------------------------------------------------ reader
def _read_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """(field number, wire type, value) of one serialized message; value = int / 8 bytes / bytes / 4 bytes."""
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + n], pos + n
        elif wire == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError("unsupported wire type %d" % wire)
        yield num, wire, v


def read_events(path, verify=True):
    """Parse an events file back: [{"wall_time", "step", "file_version", "scalars": {tag: value},
    "histograms": {tag: {"min", "max", "num", "sum", "sum_squares", "bucket_limit", "bucket"}}}].
    Checks both CRCs of every record."""
    out = []
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    while pos < len(buf):
        (n,) = struct.unpack_from("<Q", buf, pos)
        (c1,) = struct.unpack_from("<I", buf, pos + 8)
        data = buf[pos + 12:pos + 12 + n]
        (c2,) = struct.unpack_from("<I", buf, pos + 12 + n)
        if verify:
            if c1 != mask_crc(crc32c(np.frombuffer(buf[pos:pos + 8], dtype=np.uint8))):
                raise ValueError("%s: bad length crc at byte %d" % (path, pos))
            if c2 != mask_crc(crc32c(np.frombuffer(data, dtype=np.uint8))):
                raise ValueError("%s: bad data crc at byte %d" % (path, pos))
        pos += 16 + n
        ev = {"wall_time": None, "step": 0, "file_version": None, "scalars": {}, "histograms": {}}
        for num, wire, v in _fields(data):
            if num == 1:
                ev["wall_time"] = struct.unpack("<d", v)[0]
            elif num == 2:
                ev["step"] = v
            elif num == 3:
                ev["file_version"] = v.decode()
            elif num == 5:
                for vn, _, val in _fields(v):
                    if vn != 1:
                        continue
                    tag, simple, histo = None, None, None
                    for fn, _, fv in _fields(val):
                        if fn == 1:
                            tag = fv.decode()
                        elif fn == 2:
                            simple = struct.unpack("<f", fv)[0]
                        elif fn == 5:
                            histo = {}
                            names = {1: "min", 2: "max", 3: "num", 4: "sum", 5: "sum_squares"}
                            for hn, _, hv in _fields(fv):
                                if hn in names:
                                    histo[names[hn]] = struct.unpack("<d", hv)[0]
                                elif hn in (6, 7):
                                    histo["bucket_limit" if hn == 6 else "bucket"] = np.frombuffer(hv, dtype="<f8")
                    if simple is not None:
                        ev["scalars"][tag] = simple
                    if histo is not None:
                        ev["histograms"][tag] = histo
        out.append(ev)
    return out
