"""``TokenDictionary`` of wikipedia/token_dictionary.py:18-119 without the protobuf runtime: same method names,
same on-disk format (``token.tstat.pb.b64.bz2``: bz2 stream of base64 lines, one ``TokenStat`` message each --
proto/nlp.proto:20-31 -- with ``index`` == line number) and the same embedding-index scheme
(0 = mask, 1 + dictionary index, then a 65536-wide min-hash space for out-of-dictionary tokens, :58-70).
``dump_knn`` (wikipedia/train_cooccurence.py:114-126) uses it to print neighbours by name.
"""
from __future__ import annotations

import base64
import binascii
import bz2
import re


def _varint(buf, i):
    v = shift = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, i
        shift += 7


def parse_token_stat(buf: bytes) -> dict:
    """TokenStat {1: token, 2: url, 3: frequency, 4: doc_frequency, 5: index} (proto3: absent fields are defaults)."""
    out = {"token": "", "url": "", "frequency": 0, "doc_frequency": 0, "index": 0}
    names = {1: "token", 2: "url", 3: "frequency", 4: "doc_frequency", 5: "index"}
    i = 0
    while i < len(buf):
        tag, i = _varint(buf, i)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, i = _varint(buf, i)
        elif wire == 2:
            n, i = _varint(buf, i)
            v, i = buf[i:i + n], i + n
            if field in (1, 2):
                v = v.decode("utf-8")
        elif wire == 1:
            v, i = None, i + 8
        elif wire == 5:
            v, i = None, i + 4
        else:
            raise ValueError("unsupported wire type %d" % wire)
        if field in names and v is not None:
            out[names[field]] = v
    return out


def encode_token_stat(token="", url="", frequency=0, doc_frequency=0, index=0) -> bytes:
    def vi(v):
        o = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            o.append(b | 0x80 if v else b)
            if not v:
                return bytes(o)
    msg = b""
    for field, s in ((1, token), (2, url)):
        if s:
            b = s.encode("utf-8")
            msg += vi(field << 3 | 2) + vi(len(b)) + b
    for field, v in ((3, frequency), (4, doc_frequency), (5, index)):
        if v:
            msg += vi(field << 3) + vi(int(v))
    return msg


# The delimiter set of the reference's tokenizer (wikipedia/token_dictionary.py:22) -- part of the data contract: the
# dictionary on disk was built with it.
_DELIMS = ' !@#$%^&*()_+\t\n",.:;/?><|{}\'[]'   # no backslash: the reference's "\\/" is an escaped slash
_SPLITTER = re.compile("[" + re.escape(_DELIMS) + "]")
_HASH_SPACE = 1 << 16          # width of the out-of-dictionary bucket range (:60-66)
_HASH_CHARS = 10               # only the head of a long token is hashed (:50)


def _crc16(chunk: bytes) -> int:
    return binascii.crc32(chunk) & (_HASH_SPACE - 1)


def _read_stats(path):
    """TokenStat dicts of a ``*.tstat.pb.b64.bz2`` file, in file order."""
    with bz2.open(path, "rb") as f:
        for line in f:
            yield parse_token_stat(base64.b64decode(line.rstrip(b"\n")))


class TokenDictionary:
    """Token <-> dictionary index <-> embedding row.  Column storage (one list per TokenStat field actually used)
    instead of the reference's list of protobuf messages; behaviour pinned against the reference's own class
    (tests/test_record_decoders.py, tests/test_ref_golden.py)."""

    def __init__(self, dictionary_file=None):
        self._words = []           # dictionary index -> token
        self._doc_freq = []        # dictionary index -> document frequency
        self._lookup = {}          # token -> dictionary index
        if dictionary_file is not None:
            self.load(dictionary_file)

    # -- file format ------------------------------------------------------------------------
    @staticmethod
    def save(all_tokens, output_filename):
        """``all_tokens``: iterable of dicts with the TokenStat fields (the reference passes protobuf messages)."""
        lines = (base64.b64encode(encode_token_stat(**fields)) + b"\n" for fields in all_tokens)
        with bz2.open(output_filename, "wb") as out:
            out.writelines(lines)

    def load(self, dictionary_file):
        for position, stat in enumerate(_read_stats(dictionary_file), start=len(self._words)):
            if stat["index"] != position:      # the index scheme below relies on index == line number (:94)
                raise AssertionError("token index %d at line %d" % (stat["index"], position))
            self._lookup[stat["token"]] = position
            self._words.append(stat["token"])
            self._doc_freq.append(stat["doc_frequency"])

    # -- text -------------------------------------------------------------------------------
    def simple_tokenize(self, x):
        return [piece.lower() for piece in _SPLITTER.split(x) if piece]

    @staticmethod
    def minhash(token):
        """Smallest 16-bit CRC over the 4-byte windows of the token's head (:40-56).  As in the reference the LENGTH is
        counted in characters while the windows slide over UTF-8 bytes; tokens of up to 4 characters hash whole."""
        raw = token.encode("utf-8") if isinstance(token, str) else bytes(token)
        n_chars = len(token)
        if n_chars <= 4:
            return _crc16(raw)
        starts = range(min(n_chars, _HASH_CHARS) - 4)
        return min((_crc16(raw[k:k + 4]) for k in starts), default=0xFFFFFFFF)

    # -- sizes and lookups --------------------------------------------------------------------
    def get_dictionary_size(self):
        return len(self._lookup)

    def get_embedding_dictionary_size(self):
        """Mask row + dictionary + hash buckets (:68-70)."""
        return self.get_dictionary_size() + _HASH_SPACE + 1

    def get_max_doc_frequency(self):
        return max(self._doc_freq, default=0)

    def get_doc_frequency(self, token_index):
        return self._doc_freq[token_index]

    def get_token_index(self, token):
        return self._lookup.get(token)

    def get_token(self, token_index):
        return self._words[token_index]

    def get_embedding_index(self, token):
        """Row of the embedding table: 0 is the mask, known tokens follow, unknown ones land in a hash bucket (:58-66)."""
        known = self._lookup.get(token)
        if known is None:
            return self.get_dictionary_size() + 1 + self.minhash(token)
        return known + 1

    def get_embedding_indices(self, tokens):
        return list(map(self.get_embedding_index, tokens))

    def get_token_from_embedding_index(self, embedding_index):
        """Inverse of get_embedding_index for printing (:111-118), the reference's bucket label included (it subtracts
        the bucket-range width, not the dictionary size).

        The reference tests ``embedding_index is 0`` (:112), which only holds for a Python ``int`` zero; the NumPy / jax
        scalars its dump_knn passes (train_cooccurence.py:118-123) fall through to ``get_token(0 - 1)``, i.e. row 0 prints
        as the LAST dictionary token.  Reproduced as is, so logged neighbour lines match the reference's verbatim even
        when the (untrained) mask row reaches a top-10."""
        if type(embedding_index) is int and embedding_index == 0:
            return "NULL"
        row = int(embedding_index)
        if row <= self.get_dictionary_size():
            return self._words[row - 1]
        return "MINHASH %d" % (row - _HASH_SPACE - 1)
