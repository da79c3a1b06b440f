"""``create_dataset`` of spotify/input_pipeline.py:39-49 on the native TFRecord decoder.

Same record format (TFRecord framing, ``tf.train.Example`` with six ``int64_list`` features: three of length 5,
three variable -- :23-30) and the same decoded dict per example (``_decode_fn``, :32-37:
the variable-length features densified).  tf.data is replaced by ``esr_decode_tfrecord_int64`` (C, GIL
released) over whole files.
"""
from __future__ import annotations

import ctypes as C
import glob
import struct

import numpy as np

from .. import _lib as L

_KEYS = ("track_context", "album_context", "artist_context", "next_track", "next_album", "next_artist")


def decode_file(path, keys=_KEYS):
    """Every example of one TFRecord file -> list of dicts of int64 arrays."""
    data = open(path, "rb").read()
    n = len(data)
    max_rec = max(1, n // 16)
    cap = max(16, n)                                  # a varint is >= 1 byte
    nk = len(keys)
    vals = [np.empty(cap, np.int64) for _ in keys]
    offs = [np.zeros(max_rec + 1, np.int64) for _ in keys]
    ckeys = (C.c_char_p * nk)(*[k.encode() for k in keys])
    cvals = (C.c_void_p * nk)(*[v.ctypes.data for v in vals])
    coffs = (C.c_void_p * nk)(*[o.ctypes.data for o in offs])
    ccap = (C.c_int64 * nk)(*[cap] * nk)
    used = C.c_size_t(0)
    buf = (C.c_char * n).from_buffer_copy(data)
    rec = L.lib().esr_decode_tfrecord_int64(C.addressof(buf), n, nk, ckeys, cvals, ccap, coffs, max_rec, C.byref(used))
    if rec < 0 or used.value != n:
        raise L.EsrError("esr_decode_tfrecord_int64: malformed file %s (%d records, %d of %d bytes)" % (path, rec, used.value, n))
    out = []
    for r in range(rec):
        out.append({k: vals[i][offs[i][r]:offs[i][r + 1]].copy() for i, k in enumerate(keys)})
    return out


def create_dataset(pattern: str):
    """Iterator over the decoded examples of every file matching ``pattern`` (input_pipeline.py:39-49; the
    reference returns a tf.data.Dataset whose ``as_numpy_iterator()`` yields the same dicts)."""
    for path in sorted(glob.glob(pattern)):
        yield from decode_file(path)


# ---- writer (tests, synthetic corpora): TFRecord framing with masked crc32c, as tf.io.TFRecordWriter ----
def _crc32c_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_T = _crc32c_table()


def crc32c(b: bytes) -> int:
    c = 0xFFFFFFFF
    for x in b:
        c = _T[(c ^ x) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked(b):
    c = crc32c(b)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(v):
    out = bytearray()
    v = int(v) & 0xFFFFFFFFFFFFFFFF
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features: dict) -> bytes:
    """``tf.train.Example`` with int64_list features (spotify/make_training.py:102-112)."""
    entries = b""
    for k in sorted(features):
        packed = b"".join(_varint(v) for v in features[k])
        feat = _ld(3, _ld(1, packed))                       # Feature.int64_list = 3; Int64List.value = 1 (packed)
        entries += _ld(1, _ld(1, k.encode()) + _ld(2, feat))    # Features.feature map entry
    return _ld(1, entries)                                   # Example.features = 1


def write_tfrecord(path, examples):
    with open(path, "wb") as f:
        for ex in examples:
            data = encode_example(ex)
            hdr = struct.pack("<Q", len(data))
            f.write(hdr + struct.pack("<I", _masked(hdr)) + data + struct.pack("<I", _masked(data)))
