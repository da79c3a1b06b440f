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


class TokenDictionary:
    def __init__(self, dictionary_file=None):
        self.__token2index = {}
        self.__max_doc_frequency = 0
        self.__token_stat = []
        self.__filter = re.compile('[ !@#$%^&*()_+\t\n",.:;\\\\/?><|{}\'\\[\\]]')
        if dictionary_file is not None:
            self.load(dictionary_file)

    @staticmethod
    def save(all_tokens, output_filename):
        """all_tokens: iterable of dicts with the TokenStat fields (the reference passes protobuf messages)."""
        with bz2.open(output_filename, "wb") as ofile:
            for item in all_tokens:
                ofile.write(base64.b64encode(encode_token_stat(**item)))
                ofile.write(b"\n")

    def simple_tokenize(self, x):
        tokens = self.__filter.split(x)
        return [t.lower() for t in tokens if len(t) > 0]

    @staticmethod
    def minhash(token):
        """Breaks a string up into chunks of overlapping 4 bytes and returns the smallest (token_dictionary.py:40-56)."""
        count = len(token)
        b = bytes(token, "utf-8") if type(token) is str else token
        minhash = 0xFFFFFFFF
        if count <= 4:
            minhash = binascii.crc32(b) & 0xFFFF
        else:
            count = min(10, count)
            for i in range(count - 4):
                minhash = min(binascii.crc32(b[i:i + 4]) & 0xFFFF, minhash)
        return minhash

    def get_embedding_index(self, token):
        token_index = self.get_token_index(token)
        if token_index is not None:
            return 1 + token_index                      # 0 is reserved for the mask
        return 1 + self.get_dictionary_size() + self.minhash(token)

    def get_embedding_dictionary_size(self):
        return 1 + 65536 + self.get_dictionary_size()

    def get_embedding_indices(self, tokens):
        return [self.get_embedding_index(t) for t in tokens]

    def load(self, dictionary_file):
        count = 0
        with bz2.open(dictionary_file, "rb") as file:
            for line in file:
                ts = parse_token_stat(base64.b64decode(line[:-1]))
                assert ts["index"] == count
                self.__token2index[ts["token"]] = ts["index"]
                self.__max_doc_frequency = max(self.__max_doc_frequency, ts["doc_frequency"])
                self.__token_stat.append(ts)
                count += 1

    def get_dictionary_size(self):
        return len(self.__token2index)

    def get_max_doc_frequency(self):
        return self.__max_doc_frequency

    def get_doc_frequency(self, token_index):
        return self.__token_stat[token_index]["doc_frequency"]

    def get_token_index(self, token):
        return self.__token2index.get(token)

    def get_token(self, token_index):
        return self.__token_stat[token_index]["token"]

    def get_token_from_embedding_index(self, embedding_index):
        if embedding_index == 0:
            return "NULL"
        elif embedding_index <= self.get_dictionary_size():
            return self.get_token(embedding_index - 1)
        return "MINHASH %d" % (embedding_index - 65536 - 1)
