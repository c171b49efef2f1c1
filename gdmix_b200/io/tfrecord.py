"""TFRecord framing and the tf.train.Example / SequenceExample protobufs, without TensorFlow.

What the reference reads through ``tf.data.TFRecordDataset`` + ``tf.io.parse_(sequence_)example``
(gdmix-trainer/src/gdmix/io/input_data_pipeline.py:129-332):

  record framing   u64 length | u32 masked_crc32c(length) | payload | u32 masked_crc32c(payload)
  compression      by file suffix: ``.gz`` -> GZIP, ``.deflate`` -> ZLIB, anything else uncompressed
                   (input_data_pipeline.py:62-85)
  payload          Example { Features features = 1 } or
                   SequenceExample { Features context = 1; FeatureLists feature_lists = 2 }
                   Feature { oneof: BytesList = 1 | FloatList = 2 (packed fp32) | Int64List = 3 (packed varint) }
"""
import gzip
import struct
import zlib

import numpy as np

_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78  # CRC-32C (Castagnoli), reflected
        t = np.zeros(256, np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            t[i] = c
        _CRC_TABLE = t
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = int(t[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def compression_of(filename: str) -> str:
    """input_data_pipeline.py:62-85: the suffix decides."""
    suffix = filename.split(".")[-1]
    return {"deflate": "ZLIB", "gz": "GZIP"}.get(suffix, "")


def _read_all(filename, compression=None):
    compression = compression_of(filename) if compression is None else compression
    with open(filename, "rb") as f:
        raw = f.read()
    if compression == "GZIP":
        return gzip.decompress(raw)
    if compression == "ZLIB":
        return zlib.decompress(raw)
    return raw


def read_records(filename, compression=None, verify_crc=False):
    """Yields the payload bytes of every record.  The length CRC is always checked (it is 12 bytes of work and
    catches a wrong compression guess); payload CRCs only on request (pure-Python CRC is slow)."""
    buf = _read_all(filename, compression)
    pos, n = 0, len(buf)
    while pos < n:
        if pos + 12 > n:
            raise ValueError(f"{filename}: truncated record header at byte {pos}")
        (length,) = struct.unpack_from("<Q", buf, pos)
        (lcrc,) = struct.unpack_from("<I", buf, pos + 8)
        if masked_crc32c(buf[pos:pos + 8]) != lcrc:
            raise ValueError(f"{filename}: corrupted record length at byte {pos}")
        start = pos + 12
        if start + length + 4 > n:
            raise ValueError(f"{filename}: truncated record at byte {pos}")
        payload = buf[start:start + length]
        if verify_crc:
            (dcrc,) = struct.unpack_from("<I", buf, start + length)
            if masked_crc32c(payload) != dcrc:
                raise ValueError(f"{filename}: corrupted record payload at byte {pos}")
        yield payload
        pos = start + length + 4


class TFRecordWriter:
    """Writes the framing ``tf.io.TFRecordWriter`` writes (used by the tests to build fixtures the way the
    reference's tests do, test_random_effect_lr_lbfgs_model.py:196-229)."""

    def __init__(self, filename, compression=None):
        self.filename = filename
        self.compression = compression_of(filename) if compression is None else compression
        self.chunks = []

    def write(self, payload: bytes):
        head = struct.pack("<Q", len(payload))
        self.chunks.append(head + struct.pack("<I", masked_crc32c(head)) + payload +
                           struct.pack("<I", masked_crc32c(payload)))

    def close(self):
        raw = b"".join(self.chunks)
        if self.compression == "GZIP":
            raw = gzip.compress(raw)
        elif self.compression == "ZLIB":
            raw = zlib.compress(raw)
        with open(self.filename, "wb") as f:
            f.write(raw)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---- protobuf wire format ----------------------------------------------------------------------------------

def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf, pos, end):
    """Yields (field_number, wire_type, value) where value is an int (varint / fixed) or a (start, end) slice."""
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield fno, wt, v
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            yield fno, wt, (pos, pos + ln)
            pos += ln
        elif wt == 5:
            yield fno, wt, struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            yield fno, wt, struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")


def _packed_varints(buf, start, end):
    """Vectorised decode of a packed repeated varint field -> int64 array."""
    b = np.frombuffer(buf, dtype=np.uint8, count=end - start, offset=start)
    if b.size == 0:
        return np.zeros(0, np.int64)
    last = (b & 0x80) == 0
    if last.all():
        return b.astype(np.int64)
    group = np.concatenate([[0], np.cumsum(last)[:-1]])
    first = np.concatenate([[0], np.flatnonzero(last)[:-1] + 1])
    shift = (np.arange(b.size) - first[group]) * 7
    vals = (b & 0x7F).astype(np.uint64) << shift.astype(np.uint64)
    return np.add.reduceat(vals, first).astype(np.int64)  # two's complement wraps negatives (10-byte varints)


def _parse_feature(buf, start, end):
    """-> ('bytes', [bytes...]) | ('float', float32 array) | ('int64', int64 array) | (None, None)."""
    for fno, wt, v in _fields(buf, start, end):
        if wt != 2:
            continue
        s, e = v
        if fno == 1:
            return "bytes", [bytes(buf[a:b]) for f2, w2, (a, b) in _fields(buf, s, e) if f2 == 1]
        if fno == 2:
            out = []
            for f2, w2, v2 in _fields(buf, s, e):
                if f2 != 1:
                    continue
                if w2 == 2:
                    out.append(np.frombuffer(buf, dtype="<f4", count=(v2[1] - v2[0]) // 4, offset=v2[0]))
                else:
                    out.append(np.array([v2], dtype=np.uint32).view(np.float32))
            return "float", (np.concatenate(out) if out else np.zeros(0, np.float32))
        if fno == 3:
            out = []
            for f2, w2, v2 in _fields(buf, s, e):
                if f2 != 1:
                    continue
                if w2 == 2:
                    out.append(_packed_varints(buf, v2[0], v2[1]))
                else:
                    out.append(np.array([v2], dtype=np.uint64).astype(np.int64))
            return "int64", (np.concatenate(out) if out else np.zeros(0, np.int64))
    return None, None


def _parse_features_map(buf, start, end):
    out = {}
    for fno, wt, v in _fields(buf, start, end):
        if fno != 1 or wt != 2:
            continue
        key, val = None, (None, None)
        for f2, w2, v2 in _fields(buf, v[0], v[1]):
            if f2 == 1:
                key = bytes(buf[v2[0]:v2[1]]).decode("utf-8")
            elif f2 == 2:
                val = _parse_feature(buf, v2[0], v2[1])
        out[key] = val
    return out


def parse_example(payload: bytes):
    """tf.train.Example -> {name: (kind, values)}"""
    for fno, wt, v in _fields(payload, 0, len(payload)):
        if fno == 1 and wt == 2:
            return _parse_features_map(payload, v[0], v[1])
    return {}


def parse_sequence_example(payload: bytes):
    """tf.train.SequenceExample -> (context {name: (kind, values)}, feature_lists {name: [(kind, values), ...]})"""
    context, lists = {}, {}
    for fno, wt, v in _fields(payload, 0, len(payload)):
        if wt != 2:
            continue
        if fno == 1:
            context = _parse_features_map(payload, v[0], v[1])
        elif fno == 2:
            for f2, w2, v2 in _fields(payload, v[0], v[1]):
                if f2 != 1 or w2 != 2:
                    continue
                key, feats = None, []
                for f3, w3, v3 in _fields(payload, v2[0], v2[1]):
                    if f3 == 1:
                        key = bytes(payload[v3[0]:v3[1]]).decode("utf-8")
                    elif f3 == 2:
                        feats = [_parse_feature(payload, a, b)
                                 for f4, w4, (a, b) in _fields(payload, v3[0], v3[1]) if f4 == 1]
                lists[key] = feats
    return context, lists


# ---- encoding (fixtures, round-trip tests, score/partition writers) -----------------------------------------

def _enc_varint(v):
    v &= 0xFFFFFFFFFFFFFFFF
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(fno, payload):
    return _enc_varint((fno << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_feature(values, kind=None):
    """values: list/array of bytes | floats | ints (kind inferred unless given: 'bytes' | 'float' | 'int64')."""
    if kind is None:
        v0 = values[0] if len(values) else 0
        kind = "bytes" if isinstance(v0, (bytes, str)) else \
            "float" if isinstance(v0, (float, np.floating)) else "int64"
    if kind == "bytes":
        inner = b"".join(_ld(1, v.encode() if isinstance(v, str) else v) for v in values)
        return _ld(1, inner)
    if kind == "float":
        return _ld(2, _ld(1, np.asarray(values, dtype="<f4").tobytes()))
    return _ld(3, _ld(1, b"".join(_enc_varint(int(v)) for v in values)))


def _encode_features_map(features):
    out = b""
    for name, feat in features.items():
        enc = feat if isinstance(feat, bytes) else encode_feature(feat)
        out += _ld(1, _ld(1, name.encode()) + _ld(2, enc))
    return out


def encode_example(features: dict) -> bytes:
    return _ld(1, _encode_features_map(features))


def encode_sequence_example(context: dict, feature_lists: dict) -> bytes:
    """feature_lists: {name: [feature, feature, ...]} with each feature a list of values or pre-encoded bytes."""
    fl = b""
    for name, feats in feature_lists.items():
        body = b"".join(_ld(1, f if isinstance(f, bytes) else encode_feature(f)) for f in feats)
        fl += _ld(1, _ld(1, name.encode()) + _ld(2, body))
    return _ld(1, _encode_features_map(context)) + _ld(2, fl)
