"""TFRecord files of `tf.train.Example` session records without TensorFlow (the
reader producing the reference's `(features, labels)` contract).

The reference's data prep writes one Example per user - `reviewerID`: a single bytes value,
`asin`: the bytes list of the session's items (examples/BERT4Rec/data_prep/main.py:62-98,
clickstream_transformer/data_utils.py:7-50) - and its input pipeline parses exactly those two
features (examples/BERT4Rec/source/input_pipeline.py:149-156).  This module restates the two
published formats involved, on the host, in plain Python:

* TFRecord framing: uint64 length, uint32 masked CRC32C of the length, payload, uint32 masked
  CRC32C of the payload (little endian; mask = rotate right by 15, + 0xa282ead8);
* protocol-buffer wire format of Example { Features features = 1 }, Features { map<string, Feature>
  feature = 1 }, Feature { oneof { BytesList = 1, FloatList = 2, Int64List = 3 } }, each list
  `repeated value = 1` (packed for floats and ints; the unpacked form is accepted on read).

The encoder exists for tests and for writing fixtures; tests cross-check both directions against
the `protobuf` runtime with descriptors built at run time.  Host-side input plumbing only: nothing
here is on the timed path.
"""
import struct

import numpy as np

# ----------------------------------------------------------------------------- CRC32C (Castagnoli)
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)


def crc32c(data):
    crc = 0xFFFFFFFF
    tab = _TABLE
    for b in data:
        crc = tab[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
    crc = crc32c(data)
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- record framing
def write_records(path, payloads):
    with open(path, "wb") as f:
        for p in payloads:
            head = struct.pack("<Q", len(p))
            f.write(head)
            f.write(struct.pack("<I", masked_crc32c(head)))
            f.write(p)
            f.write(struct.pack("<I", masked_crc32c(p)))


def read_records(path, verify_crc=True):
    """Yields the payload bytes of every record; raises ValueError on a truncated file or (with
    verify_crc) a checksum mismatch."""
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if not head:
                return
            if len(head) < 8:
                raise ValueError(f"{path}: truncated record header")
            crc = f.read(4)
            (n,) = struct.unpack("<Q", head)
            if len(crc) < 4:
                raise ValueError(f"{path}: truncated record header")
            if verify_crc and struct.unpack("<I", crc)[0] != masked_crc32c(head):
                raise ValueError(f"{path}: corrupted record length")
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise ValueError(f"{path}: truncated record")
            if verify_crc and struct.unpack("<I", tail)[0] != masked_crc32c(data):
                raise ValueError(f"{path}: corrupted record payload")
            yield data


# ----------------------------------------------------------------------------- protobuf wire format
def _varint(v):
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


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _len_field(number, payload):
    return _varint((number << 3) | 2) + _varint(len(payload)) + payload


def _fields(buf):
    """Yields (field number, wire type, value) - value: int for varint / fixed, bytes for LEN."""
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _read_varint(buf, pos)
        number, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
            if len(v) < ln:
                raise ValueError("truncated length-delimited field")
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield number, wt, v


def encode_example(features):
    """{name: value} -> serialized tf.train.Example.  A value is bytes / str / int / float or a
    list of one of them (the type dispatch of data_utils.to_feature :22-42: str is UTF-8 encoded;
    an empty list is written as an empty bytes list).  Map entries are written in key order."""
    body = b""
    for name in sorted(features):
        value = features[name]
        if not isinstance(value, (list, tuple, np.ndarray)):
            value = [value]
        value = list(value)
        sample = value[0] if value else b""
        if isinstance(sample, (bytes, str)):
            items = b"".join(_len_field(1, v.encode("utf-8") if isinstance(v, str) else v) for v in value)
            feature = _len_field(1, items)
        elif isinstance(sample, (bool, int, np.integer)):
            feature = _len_field(3, _len_field(1, b"".join(_varint(int(v)) for v in value)))
        elif isinstance(sample, (float, np.floating)):
            feature = _len_field(2, _len_field(1, struct.pack(f"<{len(value)}f", *value)))
        else:
            raise TypeError(f"Encountered unsupported type {type(sample)}")
        entry = _len_field(1, name.encode("utf-8")) + _len_field(2, feature)
        body += _len_field(1, entry)
    return _len_field(1, body)


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _decode_feature(buf):
    for number, wt, v in _fields(buf):
        if wt != 2:
            continue
        if number == 1:
            return [bytes(x) for n, w, x in _fields(v) if n == 1 and w == 2]
        if number == 2:
            out = []
            for n, w, x in _fields(v):
                if n == 1 and w == 2:
                    out.extend(struct.unpack(f"<{len(x) // 4}f", x))
                elif n == 1 and w == 5:
                    out.extend(struct.unpack("<f", x))
            return out
        if number == 3:
            out = []
            for n, w, x in _fields(v):
                if n == 1 and w == 2:
                    pos = 0
                    while pos < len(x):
                        iv, pos = _read_varint(x, pos)
                        out.append(_signed64(iv))
                elif n == 1 and w == 0:
                    out.append(_signed64(x))
            return out
    return []   # a Feature with no kind set


def decode_example(buf):
    """Serialized tf.train.Example -> {name: list of bytes | float | int}."""
    out = {}
    for number, wt, features in _fields(buf):
        if number != 1 or wt != 2:
            continue
        for n, w, entry in _fields(features):
            if n != 1 or w != 2:
                continue
            key, feat = "", b""
            for en, ew, ev in _fields(entry):
                if en == 1 and ew == 2:
                    key = bytes(ev).decode("utf-8")
                elif en == 2 and ew == 2:
                    feat = ev
            out[key] = _decode_feature(feat)
    return out


# ----------------------------------------------------------------------------- sessions
def write_sessions(path, users, sessions, group_key="reviewerID", item_key="asin"):
    """One Example per user, as data_prep/main.py:86-91 writes them."""
    write_records(path, (encode_example({group_key: u, item_key: list(s)})
                         for u, s in zip(users, sessions)))


def read_sessions(paths, group_key="reviewerID", item_key="asin", verify_crc=True):
    """TFRecord file(s) -> (users, sessions): the two features input_pipeline.py:151-156 parses,
    as Python strings, in file order."""
    if isinstance(paths, (str, bytes)):
        paths = [paths]
    users, sessions = [], []
    for p in paths:
        for rec in read_records(p, verify_crc=verify_crc):
            ex = decode_example(rec)
            if group_key not in ex or len(ex[group_key]) != 1:
                raise ValueError(f"{p}: '{group_key}' must hold exactly one value "
                                 "(FixedLenFeature([], tf.string))")
            users.append(ex[group_key][0].decode("utf-8"))
            sessions.append([v.decode("utf-8") for v in ex.get(item_key, [])])
    return users, sessions
