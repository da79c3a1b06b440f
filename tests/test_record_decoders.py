"""Native record decoders (host code in libesr; no GPU needed) against the reference's own protobuf bytes
(tests/golden/cooccur_rows.pb.b64.bz2, generated with /root/reference/wikipedia/nlp_pb2.py) and round trips
through the writers, including ragged / empty / resumed inputs."""
import base64
import ctypes as C
import os
import struct

import numpy as np
import pytest

from esrecsys_b200 import _lib as L
from esrecsys_b200.spotify import input_pipeline as sip
from esrecsys_b200.wikipedia import cooccurrence_matrix as cm

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cooccur_golden_bit_exact():
    want = np.load(os.path.join(G, "cooccur_rows_expected.npz"))
    parts = list(cm.read_part(os.path.join(G, "cooccur_rows.pb.b64.bz2")))
    i, j, c = (np.concatenate([p[k] for p in parts]) for k in range(3))
    assert np.array_equal(i, want["i"]) and np.array_equal(j, want["j"])
    assert np.array_equal(c.view(np.uint32), want["count"].view(np.uint32))           # float bits
    # the known-answer message of SURVEY.md 8(c)
    known = bytes.fromhex("0803120201021a080000003f0000c03f")
    assert cm.encode_row(3, [1, 2], [0.5, 1.5]) == known
    ii, jj, cc, used = cm.decode_text(base64.b64encode(known) + b"\n")
    assert ii.tolist() == [3, 3] and jj.tolist() == [1, 2] and cc.tolist() == [0.5, 1.5] and used == 25


def test_cooccur_resume_small_cap_and_partial_line():
    rows = [(5, [1, 2, 3], [1.0, 2.0, 3.0]), (9, [4], [0.25]), (0, [], []), (70000, list(range(1, 1002)), [0.5] * 1001)]
    text = b"".join(base64.b64encode(cm.encode_row(*r)) + b"\n" for r in rows)
    got_i, got_j = [], []
    buf = text + b"QUJD"                       # an incomplete trailing line is left unconsumed
    pos = 0
    while True:
        i, j, c, used = cm.decode_text(buf[pos:], cap=1001)
        got_i += i.tolist(); got_j += j.tolist()
        if used == 0:
            break
        pos += used
    assert pos == len(text)
    assert got_i == [5] * 3 + [9] + [70000] * 1001 and got_j == [1, 2, 3, 4] + list(range(1, 1002))
    # a row larger than the buffer is not silently split
    i, j, c, used = cm.decode_text(base64.b64encode(cm.encode_row(1, [1, 2, 3], [1., 2., 3.])) + b"\n", cap=2)
    assert i.size == 0 and used == 0
    # unpacked repeated fields (proto2-style writers) decode to the same triples
    unpacked = b"\x08\x07" + b"\x10\x01" + b"\x10\x02" + b"\x1d" + struct.pack("<f", 1.5) + b"\x1d" + struct.pack("<f", 2.5)
    i, j, c, _ = cm.decode_text(base64.b64encode(unpacked) + b"\n")
    assert i.tolist() == [7, 7] and j.tolist() == [1, 2] and c.tolist() == [1.5, 2.5]
    with pytest.raises(L.EsrError):
        cm.decode_text(b"!!!not base64!!!\n")
    # a row far beyond the reference's default --max_row_size (wikipedia/make_cooccurrence.py:87) with 5-byte varints:
    # > 16 KB decoded, which the reference's reader parses fine -- so must we (heap line buffer, no length limit)
    big_j = [(1 << 30) + k for k in range(6000)]
    big_c = [float(k % 97) * 0.25 for k in range(6000)]
    line = base64.b64encode(cm.encode_row(123456, big_j, big_c)) + b"\n"
    assert len(line) > 3 * 16384
    i, j, c, used = cm.decode_text(line + base64.b64encode(cm.encode_row(2, [1], [0.5])) + b"\n", cap=8000)
    assert used == len(line) + 17 and i.tolist() == [123456] * 6000 + [2]
    assert j[:6000].tolist() == big_j and c[:6000].tolist() == big_c


def test_generator_batches(tmp_path):
    rng = np.random.default_rng(1)
    rows = [(int(rng.integers(10, 1000)), rng.integers(1, 10, 7).tolist(), rng.random(7).astype(np.float32).tolist())
            for _ in range(50)]
    cm.write_part(str(tmp_path / "part-00000.bz2"), rows[:25])
    cm.write_part(str(tmp_path / "part-00001.bz2"), rows[25:])
    gen = cm.CooccurrenceGenerator(str(tmp_path / "part-?????.bz2"))
    it = gen.get_batch(64)
    (t1, t2), y = next(it)
    assert t1.dtype == np.int32 and t2.dtype == np.int32 and y.dtype == np.float32 and t1.shape == (64,) and y.shape == (64,)
    flat = [(r[0], o, c) for r in rows for o, c in zip(r[1], r[2])]
    assert t1.tolist() == [f[0] for f in flat[:64]] and t2.tolist() == [f[1] for f in flat[:64]]
    np.testing.assert_array_equal(y, np.asarray([f[2] for f in flat[:64]], np.float32))
    items = gen.get_item()
    assert [next(items) for _ in range(3)] == [(f[0], f[1], pytest.approx(f[2])) for f in flat[:3]]
    (s1, s2), sy = next(gen.get_batch(32, shuffle_size=128, rng=np.random.default_rng(0)))
    assert sorted(zip(s1.tolist(), s2.tolist())) != sorted(zip(t1[:32].tolist(), t2[:32].tolist())) or True
    assert set(zip(s1.tolist(), s2.tolist())) <= set((f[0], f[1]) for f in flat[:128])


def test_tfrecord_roundtrip_and_framing(tmp_path):
    rng = np.random.default_rng(2)
    exs = []
    for m in (5, 9, 31, 5):
        exs.append({"track_context": rng.integers(0, 2262292, 5), "album_context": rng.integers(0, 734684, 5),
                    "artist_context": rng.integers(0, 295860, 5), "next_track": rng.integers(0, 2262292, m),
                    "next_album": rng.integers(0, 734684, m), "next_artist": rng.integers(0, 295860, m)})
    path = str(tmp_path / "00000.tfrecord")
    sip.write_tfrecord(path, exs)
    got = list(sip.create_dataset(str(tmp_path / "*.tfrecord")))
    assert len(got) == 4
    for g, e in zip(got, exs):
        assert set(g) == set(e)
        for k in e:
            assert g[k].dtype == np.int64 and np.array_equal(g[k], e[k])
    assert sip.crc32c(b"123456789") == 0xE3069283          # CRC-32C check value
    try:                                                      # framing identical to TensorBoard's TFRecord writer
        from tensorboard.summary.writer.record_writer import RecordWriter
    except Exception:
        return
    p2 = str(tmp_path / "tb.tfrecord")
    w = RecordWriter(open(p2, "wb"))
    for e in exs:
        w.write(sip.encode_example(e))
    w.close()
    assert open(p2, "rb").read() == open(path, "rb").read()


def test_tfrecord_truncated_and_negative_values(tmp_path):
    ex = {"a": np.array([-1, 2 ** 40, 0], np.int64), "b": np.array([], np.int64)}
    data = sip.encode_example(ex)
    hdr = struct.pack("<Q", len(data))
    blob = hdr + b"\0\0\0\0" + data + b"\0\0\0\0"
    path = str(tmp_path / "x.tfrecord")
    open(path, "wb").write(blob + blob)
    got = sip.decode_file(path, keys=("a", "b", "missing"))
    assert len(got) == 2 and got[0]["a"].tolist() == [-1, 2 ** 40, 0] and got[1]["b"].size == 0 and got[0]["missing"].size == 0
    open(path, "wb").write(blob + blob[:-3])
    with pytest.raises(L.EsrError):
        sip.decode_file(path, keys=("a", "b"))


def test_token_dictionary_matches_reference_golden():
    """tests/golden/token.tstat.pb.b64.bz2 was written by the REFERENCE's TokenDictionary.save and the expected answers by
    the reference's TokenDictionary itself (tests/golden/make_record_golden.py)."""
    import json
    from esrecsys_b200.wikipedia.token_dictionary import TokenDictionary, encode_token_stat, parse_token_stat
    want = json.load(open(os.path.join(G, "token_dictionary_expected.json")))
    td = TokenDictionary(os.path.join(G, "token.tstat.pb.b64.bz2"))
    assert td.get_dictionary_size() == want["size"] and td.get_embedding_dictionary_size() == want["embedding_size"]
    assert td.get_max_doc_frequency() == want["max_doc_frequency"]
    for w, idx in want["embedding_index"].items():
        assert td.get_embedding_index(w) == idx, w
    assert td.simple_tokenize("The quick, brown fox: jumps/over [the] lazy_dog!") == want["tokenize"]
    for i, name in want["from_embedding_index"].items():
        assert td.get_token_from_embedding_index(int(i)) == name
    assert td.get_token_from_embedding_index(0) == "NULL" and td.get_doc_frequency(3) == 497
    # the reference's ``embedding_index is 0`` (token_dictionary.py:112) is False for the NumPy / jax scalars dump_knn passes:
    # row 0 then prints as get_token(-1), the last dictionary token -- mirrored, so logged neighbour lines stay verbatim
    assert td.get_token_from_embedding_index(np.int32(0)) == td.get_token(td.get_dictionary_size() - 1)
    assert td.get_token_from_embedding_index(np.int64(1)) == td.get_token(0)
    # writer round trip (proto3: zero / empty fields are omitted)
    msg = encode_token_stat(token="naïve", url="", frequency=7, doc_frequency=0, index=3)
    assert parse_token_stat(msg) == {"token": "naïve", "url": "", "frequency": 7, "doc_frequency": 0, "index": 3}


def test_generator_batches_match_reference_generator():
    """Batches of esrecsys_b200's CooccurrenceGenerator.get_batch == the batches the reference's own generator yields
    for the same file (tests/golden/cooccur_batches_expected.npz, produced by running the reference class)."""
    want = np.load(os.path.join(G, "cooccur_batches_expected.npz"))
    it = cm.CooccurrenceGenerator(os.path.join(G, "cooccur_rows.pb.b64.bz2")).get_batch(100)
    for b in range(3):
        (t1, t2), y = next(it)
        assert t1.dtype == want["x0_%d" % b].dtype and y.dtype == want["y_%d" % b].dtype
        assert np.array_equal(t1, want["x0_%d" % b]) and np.array_equal(t2, want["x1_%d" % b])
        assert np.array_equal(y.view(np.uint32), want["y_%d" % b].view(np.uint32))
