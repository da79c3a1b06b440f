"""Host input pipeline (esrecsys_b200/wikipedia/input_pipeline.py): the threaded readers, the decoded-corpus cache and the
pinned-batch loader must reproduce the single-threaded CooccurrenceGenerator.get_batch stream bit for bit
(wikipedia/cooccurrence_matrix.py:94-107).  CPU only: the blocks are NumPy arrays here, pinned tensors in the product."""
import base64
import bz2
import os

import numpy as np
import pytest

from esrecsys_b200.wikipedia import cooccurrence_matrix as cm
from esrecsys_b200.wikipedia import input_pipeline as ip


def _corpus(tmp_path, parts=5, rows_per_part=400, seed=0):
    rng = np.random.default_rng(seed)
    for p in range(parts):
        lines = []
        for _ in range(rows_per_part):
            k = int(rng.integers(1, 40))
            idx = int(rng.integers(1000, 50000))
            lines.append(base64.b64encode(cm.encode_row(idx, rng.integers(1, idx, k).tolist(),
                                                        rng.random(k).astype(np.float32).tolist())) + b"\n")
        with bz2.open(os.path.join(tmp_path, "part-%05d.bz2" % p), "wb") as f:
            f.write(b"".join(lines))
    return os.path.join(tmp_path, "part-*.bz2")


def _take(it, n):
    out = []
    for x in it:
        out.append(tuple(np.array(a).copy() for a in x))
        if len(out) == n:
            break
    return out


@pytest.mark.parametrize("workers", [1, 3, 8])
def test_parallel_reader_preserves_file_order(tmp_path, workers):
    pat = _corpus(str(tmp_path))
    want = [np.concatenate([b[k] for p in sorted(os.listdir(tmp_path)) for b in cm.read_part(os.path.join(tmp_path, p))])
            for k in range(3)]
    rd = ip.ParallelPartReader(pat, workers=workers, blocks_ahead=2)
    got = [np.concatenate(x) for x in zip(*list(rd))]
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@pytest.mark.parametrize("shuffle", [0, 5000])
def test_batches_equal_generator_batches(tmp_path, shuffle):
    pat = _corpus(str(tmp_path))
    B = 256
    gen = cm.CooccurrenceGenerator(pat)
    ref = []
    for x, y in gen.get_batch(B, shuffle, np.random.default_rng(7)):
        ref.append((x[0], x[1], y))
        if len(ref) == 40:
            break
    rd = ip.ParallelPartReader(pat, workers=3, loop=True)
    got = _take(ip.batch_stream(iter(rd), B, shuffle, np.random.default_rng(7)), 40)
    rd.close()
    for (gi, gj, gc), (ri, rj, rc) in zip(got, ref):
        assert np.array_equal(gi, ri) and np.array_equal(gj, rj) and np.array_equal(gc.view(np.uint32), rc.view(np.uint32))


def test_triple_cache_and_pinned_loader(tmp_path):
    pat = _corpus(str(tmp_path / "c") if (tmp_path / "c").mkdir() is None else "")
    cache = ip.TripleCache.build(pat, str(tmp_path / "cache"), workers=2)
    seq = [np.concatenate(x) for x in zip(*[b for p in sorted(os.listdir(tmp_path / "c"))
                                             for b in cm.read_part(os.path.join(tmp_path / "c", p))])]
    assert cache.n == seq[0].size
    assert np.array_equal(cache.i[:], seq[0]) and np.array_equal(cache.j[:], seq[1]) and np.array_equal(cache.c[:], seq[2])
    B, ring = 512, 3
    staged = []

    def make_block():
        return np.zeros((2, B), np.int32), np.zeros(B, np.float32)
    released = []
    done = [0]                                                   # batches the "trainer" has finished uploading
    import time

    def reusable(k):                                             # GloveTrainer.wait_staged in the product
        released.append(k)
        while done[0] <= k:
            time.sleep(1e-4)
    ld = ip.PinnedBatchLoader(cache.blocks(block=3000), B, make_block, ring=ring, reusable=reusable)
    for ids, cnt in ld:
        staged.append((ids.copy(), cnt.copy()))                  # "submit": the trainer would upload the block here
        done[0] += 1
    n_b = cache.n // B
    assert len(staged) == n_b
    for b, (ids, cnt) in enumerate(staged):
        s = slice(b * B, (b + 1) * B)
        assert np.array_equal(ids[0], seq[0][s]) and np.array_equal(ids[1], seq[1][s]) and np.array_equal(cnt, seq[2][s])
    # a ring block is only rewritten after the loader asked whether the step that used it has been staged
    assert released == list(range(0, n_b - ring))
