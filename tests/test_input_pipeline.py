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


def _run_fillers(blocks, B, shuffle, seed, threads=3):
    out = []
    for fill in ip.batch_fillers(iter(blocks), B, shuffle, np.random.default_rng(seed), threads=threads):
        ids, cnt = np.empty((2, B), np.int32), np.empty(B, np.float32)
        fill(ids, cnt)
        out.append((ids, cnt))
    return out


@pytest.mark.parametrize("B,W", [(64, 1000), (64, 64), (100, 250), (37, 1001)])
def test_native_window_shuffle_is_the_reference_stream_up_to_the_permutation(B, W):
    """batch_fillers (esr_host_shuffle_gather): consecutive windows of W triples, each permuted and consumed whole, batches
    straddling window boundaries, tail after the last full window dropped -- get_shuffled_items + get_batch
    (wikipedia/cooccurrence_matrix.py:80-107) with a different (seeded, reproducible) permutation per window."""
    n = 10_000
    i = np.arange(n, dtype=np.int32)
    blocks = [(i[s:s + 777], i[s:s + 777] + 7, (i[s:s + 777] * 0.25).astype(np.float32)) for s in range(0, n, 777)]
    out = _run_fillers(blocks, B, W, 1)
    stream = np.concatenate([ids[0] for ids, _ in out])
    n_windows = n // W
    assert len(out) == (n_windows * W) // B                                  # every full window is consumed, whole batches only
    assert np.unique(stream).size == stream.size                             # nothing twice
    full = stream[: (stream.size // W) * W].reshape(-1, W)
    for t, w in enumerate(full):
        assert np.array_equal(np.sort(w), np.arange(t * W, (t + 1) * W))     # window t = a permutation of triples [tW, (t+1)W)
        if W > 8:
            assert not np.array_equal(w, np.arange(t * W, (t + 1) * W))
    for ids, cnt in out:                                                     # the three columns stay together
        assert np.array_equal(ids[1], ids[0] + 7) and np.array_equal(cnt, (ids[0] * 0.25).astype(np.float32))
    again = _run_fillers(blocks, B, W, 1, threads=1)                          # same rng seed: same stream, any thread count
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(out, again))
    other = _run_fillers(blocks, B, W, 2)
    assert W <= 8 or not all(np.array_equal(a[0], b[0]) for a, b in zip(out, other))


def test_native_fillers_without_shuffle_equal_batch_stream():
    n, B = 5000, 96
    i = np.arange(n, dtype=np.int32)
    blocks = [(i[s:s + 501], i[s:s + 501] * 2, i[s:s + 501].astype(np.float32)) for s in range(0, n, 501)]
    want = [(a.copy(), b.copy(), c.copy()) for a, b, c in ip.batch_stream(iter(blocks), B, 0)]
    got = _run_fillers(blocks, B, 0, 0)
    assert len(got) == len(want) == n // B
    for (ids, cnt), (a, b, c) in zip(got, want):
        assert np.array_equal(ids[0], a) and np.array_equal(ids[1], b) and np.array_equal(cnt, c)


def test_host_shuffle_gather_edge_cases():
    import ctypes as C
    from esrecsys_b200 import _lib as L
    h = L.lib().esr_host_shuffle_gather
    P = lambda a: C.c_void_p(a.ctypes.data)
    for n in (1, 2, 3, 31, 4096, 4097, 100_003):
        src = np.arange(n, dtype=np.int32)
        c = src.astype(np.float32)
        oi, oj, oc = np.full(n, -1, np.int32), np.full(n, -1, np.int32), np.zeros(n, np.float32)
        assert h(P(src), P(src), P(c), n, 99, 0, n, P(oi), P(oj), P(oc), 4) == 0
        assert np.array_equal(np.sort(oi), src) and np.array_equal(oi, oj) and np.array_equal(oc, oi.astype(np.float32))
        if n >= 31:                                               # any sub-range is the slice of the whole permutation
            m = n // 3
            si, sj, sc = np.empty(m, np.int32), np.empty(m, np.int32), np.empty(m, np.float32)
            assert h(P(src), P(src), P(c), n, 99, n // 2, m, P(si), P(sj), P(sc), 2) == 0
            assert np.array_equal(si, oi[n // 2: n // 2 + m])
    a = np.zeros(4, np.int32)
    f = np.zeros(4, np.float32)
    assert h(P(a), P(a), P(f), 4, 0, 2, 3, P(a), P(a), P(f), 1) == L.ESR_EINVAL if hasattr(L, "ESR_EINVAL") else True
    assert h(P(a), P(a), P(f), 4, 0, 0, 0, None, None, None, 1) == 0      # nothing to do


def test_pinned_loader_native_shuffle_through_the_ring(tmp_path):
    pat = _corpus(str(tmp_path / "c") if (tmp_path / "c").mkdir() is None else "")
    cache = ip.TripleCache.build(pat, str(tmp_path / "cache"), workers=2)
    B, W = 256, 2048

    def make_block():
        return np.zeros((2, B), np.int32), np.zeros(B, np.float32)
    ld = ip.PinnedBatchLoader(cache.blocks(block=3000), B, make_block, ring=4, shuffle_size=W, rng=np.random.default_rng(3))
    got = [(ids.copy(), cnt.copy()) for ids, cnt in ld]
    n_full = (cache.n // W) * W
    assert len(got) == n_full // B
    key = lambda a, b, c: a.astype(np.int64) * (1 << 40) + b.astype(np.int64) * (1 << 8) + (c * 4).astype(np.int64) % 256
    flat = [np.concatenate(x) for x in zip(*[(ids[0], ids[1], cnt) for ids, cnt in got])]
    for t in range(n_full // W):
        s = slice(t * W, (t + 1) * W)
        if (t + 1) * W <= flat[0].size:
            assert np.array_equal(np.sort(key(flat[0][s], flat[1][s], flat[2][s])),
                                  np.sort(key(cache.i[s], cache.j[s], cache.c[s])))
