"""Known-answer bytes for the record decoders, produced with the REFERENCE's own generated protobuf module
(/root/reference/wikipedia/nlp_pb2.py, pure-python protobuf runtime) -- the one piece of the reference that
imports in this image.  Run here (the reference tree does not exist on the GPU box):

    PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python python tests/golden/make_record_golden.py

Writes tests/golden/cooccur_rows.pb.b64.bz2 (what make_cooccurrence.py:98-100 emits per part) and
tests/golden/cooccur_rows_expected.npz (the triples CooccurrenceGenerator.get_item yields for it, obtained by
parsing the same bytes back with nlp_pb2 exactly as cooccurrence_matrix.py:69-83 does)."""
import base64
import bz2
import os
import sys

import numpy as np

os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
sys.path.insert(0, "/root/reference/wikipedia")
import nlp_pb2 as nlp_pb  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(0)
    lines = []
    for r in range(40):
        row = nlp_pb.CooccurrenceRow()
        row.index = int(rng.integers(1, 600000))
        n = int(rng.integers(1, 60)) if r != 7 else 1001          # one maximum-size row (make_cooccurrence.py:87)
        row.other_index.extend(int(v) for v in rng.integers(1, max(2, row.index), n))
        row.count.extend(float(np.float32(v)) for v in rng.lognormal(0, 1.5, n))
        lines.append(base64.b64encode(row.SerializeToString()))
    known = nlp_pb.CooccurrenceRow(index=3, other_index=[1, 2], count=[0.5, 1.5]).SerializeToString()
    assert known.hex() == "080312020102" + "1a08" + "0000003f0000c03f", known.hex()
    lines.append(base64.b64encode(known))
    path = os.path.join(HERE, "cooccur_rows.pb.b64.bz2")
    with bz2.open(path, "wb") as f:
        for ln in lines:
            f.write(ln + b"\n")
    ii, jj, cc = [], [], []
    with bz2.open(path, "rb") as f:                                 # cooccurrence_matrix.py:69-83, verbatim semantics
        for line in f:
            proto = nlp_pb.CooccurrenceRow()
            proto.ParseFromString(base64.b64decode(line[:-1]))
            for k in range(len(proto.other_index)):
                ii.append(proto.index); jj.append(proto.other_index[k]); cc.append(proto.count[k])
    np.savez_compressed(os.path.join(HERE, "cooccur_rows_expected.npz"), i=np.asarray(ii, np.int32), j=np.asarray(jj, np.int32),
                        count=np.asarray(cc, np.float32))
    print("rows", len(lines), "triples", len(ii))


if __name__ == "__main__":
    main()


def token_dictionary_golden():
    """tests/golden/token.tstat.pb.b64.bz2 written with the reference's TokenDictionary.save + nlp_pb2.TokenStat, and the
    answers the reference's TokenDictionary gives for it (token_dictionary.py:58-119)."""
    import json
    import token_dictionary as ref_td          # /root/reference/wikipedia/token_dictionary.py
    words = ["the", "of", "and", "naïve", "zürich", "data-set", "x"] + ["w%03d" % i for i in range(40)]
    stats = []
    for i, w in enumerate(words):
        ts = nlp_pb.TokenStat()
        ts.token, ts.frequency, ts.doc_frequency, ts.index = w, 1000 - i, 500 - i, i
        if i % 5 == 0:
            ts.url = "https://example.org/%d" % i
        stats.append(ts)
    path = os.path.join(HERE, "token.tstat.pb.b64.bz2")
    ref_td.TokenDictionary.save(stats, path)
    td = ref_td.TokenDictionary(path)
    probe = ["the", "zürich", "unknownword", "abc", "Supercalifragilistic", "x"]
    ans = {"size": td.get_dictionary_size(), "embedding_size": td.get_embedding_dictionary_size(),
           "max_doc_frequency": td.get_max_doc_frequency(),
           "embedding_index": {w: td.get_embedding_index(w) for w in probe},
           "tokenize": td.simple_tokenize("The quick, brown fox: jumps/over [the] lazy_dog!"),
           "from_embedding_index": {str(i): td.get_token_from_embedding_index(i) for i in (1, 5, 47, 48, 70000)}}
    json.dump(ans, open(os.path.join(HERE, "token_dictionary_expected.json"), "w"), ensure_ascii=False, indent=1)
    print("token dictionary:", ans["size"], "tokens")


if __name__ == "__main__":
    token_dictionary_golden()


def generator_golden():
    """The first batches the REFERENCE's CooccurrenceGenerator.get_batch yields for cooccur_rows.pb.b64.bz2
    (wikipedia/cooccurrence_matrix.py:58-107).  The module imports tensorflow at the top only for get_dataset
    (:108-115); a stub module object stands in for it here -- get_batch itself is pure Python + nlp_pb2 + NumPy."""
    import types
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    import cooccurrence_matrix as ref_cm       # /root/reference/wikipedia/cooccurrence_matrix.py
    gen = ref_cm.CooccurrenceGenerator(os.path.join(HERE, "cooccur_rows.pb.b64.bz2"))
    it = gen.get_batch(100)
    out = {}
    for b in range(3):
        x, y = next(it)
        out["x0_%d" % b], out["x1_%d" % b], out["y_%d" % b] = x[0], x[1], y
    np.savez_compressed(os.path.join(HERE, "cooccur_batches_expected.npz"), **out)
    print("generator batches:", out["x0_0"].dtype, out["y_0"].dtype, out["x0_0"].shape)


if __name__ == "__main__":
    generator_golden()
