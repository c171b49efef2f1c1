"""Minimal Avro object-container reader / writer (no fastavro).

Covers what the path's outputs need: the Photon-ML ``BayesianLinearModelAvro`` model files
(gdmix-trainer/src/gdmix/models/schemas.py:3-51, written at util/io_utils.py:163-212) and the score files
(util/io_utils.py:367-375): records, arrays, unions, named-type references, string/bytes/int/long/float/double/
boolean/null; codecs ``null`` (what fastavro.writer writes by default, i.e. what the reference emits) and
``deflate`` (read and write).  Spec: Avro 1.x "Object Container Files" and "Binary Encoding".
"""
import io
import json
import os
import struct
import zlib

MAGIC = b"Obj\x01"
PRIMITIVES = {"null", "boolean", "int", "long", "float", "double", "bytes", "string"}


def _zigzag_encode(n):
    return (n << 1) ^ (n >> 63)


def write_long(out, n):
    v = _zigzag_encode(int(n)) & 0xFFFFFFFFFFFFFFFF
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return


def read_long(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            break
        shift += 7
    return (result >> 1) ^ -(result & 1), pos


class Schema:
    """Parsed schema with a named-type table, enough to encode/decode generically."""

    def __init__(self, schema):
        self.json = json.loads(schema) if isinstance(schema, str) else schema
        self.named = {}
        self._collect(self.json, None)

    def _collect(self, s, namespace):
        if isinstance(s, dict):
            t = s.get("type")
            if t in ("record", "enum", "fixed"):
                ns = s.get("namespace", namespace)
                name = s["name"]
                self.named[name] = s
                if ns and "." not in name:
                    self.named[f"{ns}.{name}"] = s
                for f in s.get("fields", []):
                    self._collect(f["type"], ns)
            elif t == "array":
                self._collect(s["items"], namespace)
            elif t == "map":
                self._collect(s["values"], namespace)
            elif isinstance(t, (dict, list)):
                self._collect(t, namespace)
        elif isinstance(s, list):
            for b in s:
                self._collect(b, namespace)

    def resolve(self, s):
        if isinstance(s, str) and s not in PRIMITIVES:
            return self.named[s]
        return s


def _type_name(s):
    if isinstance(s, str):
        return s
    if isinstance(s, list):
        return "union"
    t = s["type"]
    return t if isinstance(t, str) else _type_name(t)


def _matches(schema, s, datum):
    s = schema.resolve(s)
    t = _type_name(s)
    if t == "null":
        return datum is None
    if t == "boolean":
        return isinstance(datum, bool)
    if t in ("int", "long"):
        return isinstance(datum, int) and not isinstance(datum, bool) or hasattr(datum, "__index__")
    if t in ("float", "double"):
        return isinstance(datum, (int, float)) and not isinstance(datum, bool) or hasattr(datum, "__float__")
    if t == "string":
        return isinstance(datum, str)
    if t == "bytes":
        return isinstance(datum, (bytes, bytearray))
    if t == "array":
        return isinstance(datum, (list, tuple))
    if t in ("record", "map"):
        return isinstance(datum, dict)
    return False


def encode(schema, s, datum, out):
    s = schema.resolve(s)
    if isinstance(s, list):
        for i, branch in enumerate(s):
            if _matches(schema, branch, datum):
                write_long(out, i)
                encode(schema, branch, datum, out)
                return
        raise ValueError(f"datum {datum!r} matches no branch of union {s}")
    t = s if isinstance(s, str) else s["type"]
    if isinstance(t, (dict, list)):
        return encode(schema, t, datum, out)
    if t == "null":
        return
    if t == "boolean":
        out.append(1 if datum else 0)
    elif t in ("int", "long"):
        write_long(out, int(datum))
    elif t == "float":
        out += struct.pack("<f", float(datum))
    elif t == "double":
        out += struct.pack("<d", float(datum))
    elif t == "string":
        b = datum.encode("utf-8")
        write_long(out, len(b)); out += b
    elif t == "bytes":
        write_long(out, len(datum)); out += datum
    elif t == "array":
        if len(datum):
            write_long(out, len(datum))
            for item in datum:
                encode(schema, s["items"], item, out)
        write_long(out, 0)
    elif t == "map":
        if len(datum):
            write_long(out, len(datum))
            for k, v in datum.items():
                encode(schema, "string", k, out)
                encode(schema, s["values"], v, out)
        write_long(out, 0)
    elif t == "record":
        for f in s["fields"]:
            if f["name"] in datum:
                v = datum[f["name"]]
            elif "default" in f:
                v = f["default"]
            else:
                raise ValueError(f"record {s['name']}: field {f['name']!r} missing and has no default")
            encode(schema, f["type"], v, out)
    else:
        raise ValueError(f"unsupported avro type {t!r}")


def decode(schema, s, buf, pos):
    s = schema.resolve(s)
    if isinstance(s, list):
        i, pos = read_long(buf, pos)
        return decode(schema, s[i], buf, pos)
    t = s if isinstance(s, str) else s["type"]
    if isinstance(t, (dict, list)):
        return decode(schema, t, buf, pos)
    if t == "null":
        return None, pos
    if t == "boolean":
        return bool(buf[pos]), pos + 1
    if t in ("int", "long"):
        return read_long(buf, pos)
    if t == "float":
        return struct.unpack_from("<f", buf, pos)[0], pos + 4
    if t == "double":
        return struct.unpack_from("<d", buf, pos)[0], pos + 8
    if t == "string":
        n, pos = read_long(buf, pos)
        return bytes(buf[pos:pos + n]).decode("utf-8"), pos + n
    if t == "bytes":
        n, pos = read_long(buf, pos)
        return bytes(buf[pos:pos + n]), pos + n
    if t == "array":
        out = []
        while True:
            n, pos = read_long(buf, pos)
            if n == 0:
                return out, pos
            if n < 0:
                n = -n
                _, pos = read_long(buf, pos)
            for _ in range(n):
                v, pos = decode(schema, s["items"], buf, pos)
                out.append(v)
    if t == "map":
        out = {}
        while True:
            n, pos = read_long(buf, pos)
            if n == 0:
                return out, pos
            if n < 0:
                n = -n
                _, pos = read_long(buf, pos)
            for _ in range(n):
                k, pos = decode(schema, "string", buf, pos)
                v, pos = decode(schema, s["values"], buf, pos)
                out[k] = v
    if t == "record":
        rec = {}
        for f in s["fields"]:
            rec[f["name"]], pos = decode(schema, f["type"], buf, pos)
        return rec, pos
    raise ValueError(f"unsupported avro type {t!r}")


class Writer:
    """Container-file writer.  ``write_block(records)`` may be called repeatedly (the reference appends one
    block per batch of 1024 records, util/io_utils.py:299-334)."""

    def __init__(self, path_or_file, schema, codec="null", sync=None):
        self.schema = schema if isinstance(schema, Schema) else Schema(schema)
        self.codec = codec
        self.sync = sync or os.urandom(16)
        self.own = isinstance(path_or_file, (str, os.PathLike))
        if self.own:
            d = os.path.dirname(str(path_or_file))
            if d:
                os.makedirs(d, exist_ok=True)
        self.f = open(path_or_file, "wb") if self.own else path_or_file
        header = bytearray(MAGIC)
        meta = {"avro.schema": json.dumps(self.schema.json).encode(), "avro.codec": codec.encode()}
        write_long(header, len(meta))
        for k, v in meta.items():
            kb = k.encode()
            write_long(header, len(kb)); header += kb
            write_long(header, len(v)); header += v
        write_long(header, 0)
        header += self.sync
        self.f.write(bytes(header))
        self.count = 0

    def write_block(self, records):
        body = bytearray()
        n = 0
        for r in records:
            encode(self.schema, self.schema.json, r, body)
            n += 1
        if n == 0:
            return 0
        data = bytes(body)
        if self.codec == "deflate":
            c = zlib.compressobj(wbits=-15)
            data = c.compress(data) + c.flush()
        elif self.codec != "null":
            raise ValueError(f"unsupported codec {self.codec}")
        head = bytearray()
        write_long(head, n)
        write_long(head, len(data))
        self.f.write(bytes(head) + data + self.sync)
        self.count += n
        return n

    def close(self):
        if self.own:
            self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def write_records(path, schema, records, batch_size=1024, codec="null"):
    """batched_write_avro (util/io_utils.py:299-334): one block per `batch_size` records.  An empty iterator
    still produces a valid, empty container (fastavro would leave no file; an empty file is easier on readers)."""
    with Writer(path, schema, codec) as w:
        batch = []
        for r in records:
            batch.append(r)
            if len(batch) >= batch_size:
                w.write_block(batch)
                batch = []
        w.write_block(batch)
        return w.count


def read_container(path_or_bytes):
    """-> (schema_json, iterator over records)."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            buf = f.read()
    if buf[:4] != MAGIC:
        raise ValueError("not an Avro object container file")
    pos = 4
    meta = {}
    while True:
        n, pos = read_long(buf, pos)
        if n == 0:
            break
        if n < 0:
            n = -n
            _, pos = read_long(buf, pos)
        for _ in range(n):
            kl, pos = read_long(buf, pos)
            k = buf[pos:pos + kl].decode(); pos += kl
            vl, pos = read_long(buf, pos)
            meta[k] = buf[pos:pos + vl]; pos += vl
    sync = buf[pos:pos + 16]
    pos += 16
    schema = Schema(meta["avro.schema"].decode())
    codec = meta.get("avro.codec", b"null").decode()

    def records():
        p = pos
        while p < len(buf):
            n, p = read_long(buf, p)
            size, p = read_long(buf, p)
            data = buf[p:p + size]
            p += size
            if buf[p:p + 16] != sync:
                raise ValueError("avro sync marker mismatch")
            p += 16
            if codec == "deflate":
                data = zlib.decompress(data, wbits=-15)
            elif codec != "null":
                raise ValueError(f"unsupported avro codec {codec!r}")
            q = 0
            for _ in range(n):
                rec, q = decode(schema, schema.json, data, q)
                yield rec

    return schema.json, records()


def read_blocks(path_or_bytes):
    """-> (schema_json, iterator over (record_count, uncompressed block bytes)) -- for the library's block decoders."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            buf = f.read()
    if buf[:4] != MAGIC:
        raise ValueError("not an Avro object container file")
    pos = 4
    meta = {}
    while True:
        n, pos = read_long(buf, pos)
        if n == 0:
            break
        if n < 0:
            n = -n
            _, pos = read_long(buf, pos)
        for _ in range(n):
            kl, pos = read_long(buf, pos)
            k = buf[pos:pos + kl].decode(); pos += kl
            vl, pos = read_long(buf, pos)
            meta[k] = buf[pos:pos + vl]; pos += vl
    sync = buf[pos:pos + 16]
    pos += 16
    codec = meta.get("avro.codec", b"null").decode()

    def blocks():
        p = pos
        while p < len(buf):
            n, p = read_long(buf, p)
            size, p = read_long(buf, p)
            data = buf[p:p + size]
            p += size
            if buf[p:p + 16] != sync:
                raise ValueError("avro sync marker mismatch")
            p += 16
            if codec == "deflate":
                data = zlib.decompress(data, wbits=-15)
            elif codec != "null":
                raise ValueError(f"unsupported avro codec {codec!r}")
            yield n, data

    return json.loads(meta["avro.schema"].decode()), blocks()


def read_records(path_or_bytes):
    return list(read_container(path_or_bytes)[1])
