"""``CooccurrenceGenerator`` of wikipedia/cooccurrence_matrix.py:58-115 on the native record decoder.

Same class and method names (``get_item``, ``get_shuffled_items``, ``get_batch``), same on-disk format
(``*.cooccur.pb.b64.bz2`` parts: bz2 stream of base64 lines, one ``CooccurrenceRow`` protobuf each --
SURVEY.md App. B.1) and the same batch layout (``x = [int32 (B,), int32 (B,)]``, ``y = float32 (B,)``,
:103-114).  The reference parses one protobuf per line in Python and yields element by element; here a whole
decompressed block is decoded by ``esr_decode_cooccur_b64`` (C, GIL released) into NumPy arrays, and batches
are sliced from them.  ``get_dataset`` (:108-115, tf.data) is replaced by ``get_batch`` itself: the trainer
consumes NumPy / pinned tensors directly.
"""
from __future__ import annotations

import bz2
import ctypes as C
import glob

import numpy as np

from .. import _lib as L

_BLOCK = 8 << 20      # decompressed bytes handed to the decoder per call


def decode_text(text: bytes, cap=None):
    """All complete lines of ``text`` -> (i int32, j int32, count f32, consumed bytes)."""
    cap = int(cap if cap is not None else max(1024, len(text) // 3))     # >= 5 wire bytes + base64 overhead per triple
    i = np.empty(cap, np.int32)
    j = np.empty(cap, np.int32)
    c = np.empty(cap, np.float32)
    rows = C.c_int64(0)
    used = C.c_size_t(0)
    n = L.lib().esr_decode_cooccur_b64(text, len(text), i.ctypes.data, j.ctypes.data, c.ctypes.data, cap, C.byref(rows),
                                       C.byref(used))
    if n < 0:
        raise L.EsrError("esr_decode_cooccur_b64: malformed record (%d)" % n)
    return i[:n], j[:n], c[:n], used.value


def read_part(path):
    """Yields (i, j, count) array triples for one bz2 part, block by block, in file order."""
    tail = b""
    with bz2.open(path, "rb") as f:
        while True:
            blk = f.read(_BLOCK)
            buf = tail + blk
            if not buf:
                return
            if not blk and not buf.endswith(b"\n"):
                buf += b"\n"                       # last line without newline
            i, j, c, used = decode_text(buf)
            if i.size:
                yield i, j, c
            tail = buf[used:]
            if not blk:
                if tail.strip():
                    raise L.EsrError("undecodable tail in %s" % path)
                return


class CooccurrenceGenerator:
    def __init__(self, input_pattern):
        self._input_files = sorted(glob.glob(input_pattern))
        self._total_files = len(self._input_files)

    def get_arrays(self, loop=True):
        """Block-wise (i, j, count) arrays over all files, forever when ``loop`` (as get_item's ``while True``)."""
        while True:
            for path in self._input_files:
                yield from read_part(path)
            if not loop:
                return

    def get_item(self):
        """Gets a single item of i, j, count (cooccurrence_matrix.py:64-83)."""
        for i, j, c in self.get_arrays():
            for k in range(i.size):
                yield (int(i[k]), int(j[k]), float(c[k]))

    def get_shuffled_items(self, num_items):
        """Pre-fetches and shuffles num_items of stuff (:85-92)."""
        it = self.get_item()
        while True:
            items = [next(it) for _ in range(num_items)]
            np.random.shuffle(items)
            for item in items:
                yield item

    def get_batch(self, batch_size, shuffle_size=0, rng=None):
        """(:94-107) ``x = [token1 int32 (B,), token2 int32 (B,)]``, ``y = float32 (B,)``.  With ``shuffle_size`` the
        stream is shuffled in windows of that many triples (vectorised; the reference shuffles a Python list)."""
        rng = rng if rng is not None else np.random.default_rng()
        window = max(int(shuffle_size), batch_size)
        # one owned buffer per column plus a cursor: a decoded block is appended once (the unread tail, shorter than a
        # window, is the only thing re-copied), and windows are sliced by offset -- O(triples), not O(triples * batches)
        bi = np.empty(0, np.int32); bj = np.empty(0, np.int32); bc = np.empty(0, np.float32)
        pos = 0
        for i, j, c in self.get_arrays():
            bi = np.concatenate([bi[pos:], i]); bj = np.concatenate([bj[pos:], j]); bc = np.concatenate([bc[pos:], c])
            pos = 0
            while bi.size - pos >= window:
                wi, wj, wc = bi[pos:pos + window], bj[pos:pos + window], bc[pos:pos + window]
                if shuffle_size:
                    perm = rng.permutation(window)
                    wi, wj, wc = wi[perm], wj[perm], wc[perm]
                nb = window // batch_size
                for b in range(nb):
                    s = slice(b * batch_size, (b + 1) * batch_size)
                    yield [wi[s].copy(), wj[s].copy()], wc[s].copy()
                used = nb * batch_size
                if shuffle_size and used < window:
                    # the (shuffled) remainder of the window goes back in front of the stream
                    bi[pos + used:pos + window] = wi[used:]
                    bj[pos + used:pos + window] = wj[used:]
                    bc[pos + used:pos + window] = wc[used:]
                pos += used


# ---- writer (tests, synthetic corpora): the exact bytes Spark's saveAsTextFile + BZip2Codec part holds ----
def _varint(v):
    out = bytearray()
    v = int(v)
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def encode_row(index, other_index, count):
    """Serialised ``CooccurrenceRow`` (proto3 packed repeated fields; wikipedia/proto/nlp.proto)."""
    import struct
    other = b"".join(_varint(o) for o in other_index)
    cnt = struct.pack("<%df" % len(count), *count)
    msg = b""
    if index:
        msg += b"\x08" + _varint(index)
    if len(other_index):
        msg += b"\x12" + _varint(len(other)) + other
    if len(count):
        msg += b"\x1a" + _varint(len(cnt)) + cnt
    return msg


def write_part(path, rows):
    """rows: iterable of (index, other_index list, count list)."""
    import base64
    with bz2.open(path, "wb") as f:
        for index, other, count in rows:
            f.write(base64.b64encode(encode_row(index, other, count)) + b"\n")
