"""Host input pipeline of the GloVe trainer: what ``CooccurrenceGenerator.get_dataset`` + ``.prefetch`` do in the reference
(wikipedia/cooccurrence_matrix.py:108-115; shuffle buffer wikipedia/train_cooccurence.py:49), at the rate the CUDA step needs.

The reference parses one protobuf per line in Python (GIL-bound, 1e5-1e6 pairs/s).  Here

* ``ParallelPartReader``  decodes the ``*.cooccur.pb.b64.bz2`` parts with W threads (bz2 + ``esr_decode_cooccur_b64`` both
  release the GIL) but hands the blocks out in FILE ORDER, so the stream -- and therefore every batch -- is the
  single-threaded generator's, bit for bit;
* ``TripleCache``         is the decoded corpus as three flat little-endian arrays (``i.int32``, ``j.int32``,
  ``count.float32``) that are memory-mapped: bz2 tops out at a few M pairs/s per core whatever the parser does, the
  cache streams at memory speed (it is what a production run trains from after the first epoch);
* ``PinnedBatchLoader``   assembles batches (optionally window-shuffled like ``get_shuffled_items``) straight into a ring
  of pinned ``(ids (2,B) int32, counts (B,) f32)`` blocks on a background thread, one step ahead of
  ``GloveTrainer.submit`` -- a block is only rewritten after the trainer reports its staging copy complete
  (``GloveTrainer.wait_staged``).

The batch layout is the reference's: ``x = [token1 (B,), token2 (B,)]`` -> ``(2,B) int32``, ``y = (B,) float32``
(wikipedia/cooccurrence_matrix.py:94-107,113-114).
"""
from __future__ import annotations

import glob
import os
import queue
import threading

import numpy as np

from .cooccurrence_matrix import read_part


class ParallelPartReader:
    """(i, j, count) blocks of all parts in file order, decoded by ``workers`` threads (part p by worker p % workers)."""

    def __init__(self, input_pattern, workers=4, blocks_ahead=4, loop=False):
        self.files = sorted(glob.glob(input_pattern))
        if not self.files:
            raise FileNotFoundError(input_pattern)
        self.workers = max(1, min(int(workers), len(self.files)))
        self.loop = bool(loop)
        self._q = [queue.Queue(maxsize=blocks_ahead) for _ in range(self.workers)]
        self._stop = threading.Event()
        self._threads = [threading.Thread(target=self._work, args=(w,), daemon=True) for w in range(self.workers)]
        for t in self._threads:
            t.start()

    def _put(self, w, item):
        while not self._stop.is_set():
            try:
                self._q[w].put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _work(self, w):
        try:
            while True:
                for p in range(w, len(self.files), self.workers):
                    for blk in read_part(self.files[p]):
                        if not self._put(w, blk):
                            return
                    if not self._put(w, None):                  # end of part p
                        return
                if not self.loop:
                    self._put(w, StopIteration)
                    return
        except Exception as e:  # pragma: no cover  (surfaced to the consumer)
            self._put(w, e)

    def __iter__(self):
        while True:
            for p in range(len(self.files)):
                w = p % self.workers
                while True:
                    item = self._q[w].get()
                    if item is None:
                        break
                    if item is StopIteration:
                        return
                    if isinstance(item, Exception):
                        raise item
                    yield item
            if not self.loop:
                return

    def close(self):
        self._stop.set()


class TripleCache:
    """Decoded corpus on disk: ``<dir>/i.int32``, ``j.int32``, ``count.float32`` (+ ``n.txt``), memory-mapped."""

    def __init__(self, directory):
        self.dir = directory
        self.n = int(open(os.path.join(directory, "n.txt")).read())
        self.i = np.memmap(os.path.join(directory, "i.int32"), np.int32, "r", shape=(self.n,))
        self.j = np.memmap(os.path.join(directory, "j.int32"), np.int32, "r", shape=(self.n,))
        self.c = np.memmap(os.path.join(directory, "count.float32"), np.float32, "r", shape=(self.n,))

    @staticmethod
    def build(input_pattern, directory, workers=4):
        """Decode every part once (file order) and append the triples to the three flat files."""
        os.makedirs(directory, exist_ok=True)
        n = 0
        with open(os.path.join(directory, "i.int32"), "wb") as fi, open(os.path.join(directory, "j.int32"), "wb") as fj, \
                open(os.path.join(directory, "count.float32"), "wb") as fc:
            rd = ParallelPartReader(input_pattern, workers=workers)
            for i, j, c in rd:
                fi.write(np.ascontiguousarray(i, np.int32).tobytes())
                fj.write(np.ascontiguousarray(j, np.int32).tobytes())
                fc.write(np.ascontiguousarray(c, np.float32).tobytes())
                n += i.size
            rd.close()
        with open(os.path.join(directory, "n.txt"), "w") as f:
            f.write(str(n))
        return TripleCache(directory)

    def blocks(self, block=1 << 22, loop=False):
        """(i, j, count) views of ``block`` triples in corpus order (the same stream ``ParallelPartReader`` yields)."""
        while True:
            for s in range(0, self.n, block):
                e = min(self.n, s + block)
                yield self.i[s:e], self.j[s:e], self.c[s:e]
            if not loop:
                return


def batch_stream(blocks, batch_size, shuffle_size=0, rng=None):
    """Batches from a stream of (i, j, count) blocks, exactly as ``CooccurrenceGenerator.get_batch`` forms them (same
    windows, same permutations for the same ``rng``).  Yields ``(i (B,), j (B,), count (B,))`` VIEWS: copy before the next
    ``next()``."""
    rng = rng if rng is not None else np.random.default_rng()
    window = max(int(shuffle_size), batch_size)
    bi = np.empty(0, np.int32)
    bj = np.empty(0, np.int32)
    bc = np.empty(0, np.float32)
    pos = 0
    for i, j, c in blocks:
        bi = np.concatenate([bi[pos:], i])
        bj = np.concatenate([bj[pos:], j])
        bc = np.concatenate([bc[pos:], c])
        pos = 0
        while bi.size - pos >= window:
            wi, wj, wc = bi[pos:pos + window], bj[pos:pos + window], bc[pos:pos + window]
            if shuffle_size:
                perm = rng.permutation(window)
                wi, wj, wc = wi[perm], wj[perm], wc[perm]
            nb = window // batch_size
            for b in range(nb):
                s = slice(b * batch_size, (b + 1) * batch_size)
                yield wi[s], wj[s], wc[s]
            used = nb * batch_size
            if shuffle_size and used < window:
                bi[pos + used:pos + window] = wi[used:]
                bj[pos + used:pos + window] = wj[used:]
                bc[pos + used:pos + window] = wc[used:]
            pos += used


class _Accumulator:
    """Incoming (i, j, count) blocks -> contiguous windows of exactly W triples (one copy per window)."""

    def __init__(self):
        self.chunks, self.size = [], 0

    def push(self, i, j, c):
        if i.size:
            self.chunks.append((i, j, c))
            self.size += i.size

    def take(self, W):
        oi, oj, oc = np.empty(W, np.int32), np.empty(W, np.int32), np.empty(W, np.float32)
        pos = 0
        while pos < W:
            i, j, c = self.chunks[0]
            m = min(i.size, W - pos)
            oi[pos:pos + m], oj[pos:pos + m], oc[pos:pos + m] = i[:m], j[:m], c[:m]
            pos += m
            if m == i.size:
                self.chunks.pop(0)
            else:
                self.chunks[0] = (i[m:], j[m:], c[m:])
        self.size -= W
        return oi, oj, oc


def _prefetch(gen, depth=1):
    """Runs ``gen`` on a background thread, ``depth`` items ahead (window assembly is NumPy copies: the GIL is released)."""
    q = queue.Queue(maxsize=depth)
    end = object()

    def work():
        try:
            for item in gen:
                q.put(item)
            q.put(end)
        except Exception as e:  # pragma: no cover  (surfaced to the consumer)
            q.put(e)

    threading.Thread(target=work, daemon=True).start()
    while True:
        item = q.get()
        if item is end:
            return
        if isinstance(item, Exception):
            raise item
        yield item


def batch_fillers(blocks, batch_size, shuffle_size=0, rng=None, threads=None):
    """The input stream of ``get_batch`` as ``fill(ids (2,B) int32, counts (B,) f32)`` callables that write a batch
    straight into its destination (a pinned block of the ring) -- and, for ``shuffle_size`` > 0, shuffle natively: the
    window's permutation is a keyed bijection evaluated on the fly by ``esr_host_shuffle_gather`` on ``threads`` host
    threads (250 M pairs/s on 8 cores against 10 M pairs/s for a NumPy permutation + three fancy-index gathers).

    Semantics are the reference's (``get_shuffled_items`` + ``get_batch``, wikipedia/cooccurrence_matrix.py:80-107):
    consecutive windows of ``shuffle_size`` triples, each shuffled and consumed whole; a batch that starts in the tail of
    one shuffled window ends in the head of the next; the triples after the last full window are dropped.  The
    permutation inside a window comes from ``rng`` (one 63-bit seed per window), not the one ``np.random.shuffle`` would
    draw.  The next window is assembled on a background thread while the current one is consumed."""
    from .. import _lib as L
    import ctypes as C
    rng = rng if rng is not None else np.random.default_rng()
    B = int(batch_size)
    W = max(int(shuffle_size), B) if shuffle_size else B
    threads = int(threads or min(8, len(os.sched_getaffinity(0))))
    h = L.lib().esr_host_shuffle_gather if shuffle_size else None

    def windows():
        acc = _Accumulator()
        for i, j, c in blocks:
            acc.push(np.asarray(i, np.int32), np.asarray(j, np.int32), np.asarray(c, np.float32))
            while acc.size >= W:
                yield acc.take(W) + (int(rng.integers(0, 2 ** 63)) if shuffle_size else 0,)

    def piece(win, k0, m, ids, cnt, d0):
        """elements [k0, k0 + m) of the (shuffled) window -> positions [d0, d0 + m) of the batch"""
        wi, wj, wc, seed = win
        if not shuffle_size:
            ids[0, d0:d0 + m], ids[1, d0:d0 + m], cnt[d0:d0 + m] = wi[k0:k0 + m], wj[k0:k0 + m], wc[k0:k0 + m]
            return
        base = ids.ctypes.data + 4 * d0
        L.check(h(C.c_void_p(wi.ctypes.data), C.c_void_p(wj.ctypes.data), C.c_void_p(wc.ctypes.data), W, seed, k0, m,
                  C.c_void_p(base), C.c_void_p(base + 4 * B), C.c_void_p(cnt.ctypes.data + 4 * d0), threads),
                "esr_host_shuffle_gather")

    def filler(parts):
        def fill(ids, cnt):
            assert ids.flags.c_contiguous and ids.shape == (2, B) and ids.dtype == np.int32 and cnt.dtype == np.float32
            d0 = 0
            for win, k0, m in parts:
                piece(win, k0, m, ids, cnt, d0)
                d0 += m
        return fill

    it = _prefetch(windows(), depth=1)
    cur = next(it, None)
    off = 0
    while cur is not None:
        if off + B <= W:
            yield filler([(cur, off, B)])
            off += B
            if off == W:
                cur, off = next(it, None), 0
        else:
            nxt = next(it, None)
            if nxt is None:
                return                                  # the tail that does not fill a batch is dropped
            yield filler([(cur, off, W - off), (nxt, 0, B - (W - off))])
            cur, off = nxt, B - (W - off)


class PinnedBatchLoader:
    """Background thread: batches of ``blocks`` -> a ring of pinned ``(ids (2,B), counts (B,))`` blocks.

    ``make_block()`` returns one pinned block (``GloveTrainer.pinned_batch`` -- ids and counts adjacent, uploaded with a
    single copy; any ``(ids (2,B) int32, counts (B,) f32)`` pair of writable array-likes works, which is how the CPU tests
    and the host-only throughput probe run it); ``reusable(k)`` blocks until the k-th block handed out may be rewritten
    (the trainer's staging copy of that step has completed) -- ``None``: immediately.
    """

    def __init__(self, blocks, batch_size, make_block, ring=4, shuffle_size=0, rng=None, reusable=None, native=True,
                 threads=None):
        """``native``: batches are written into the ring by ``batch_fillers`` (window shuffle in libesr, host threads);
        ``False``: the NumPy path of ``batch_stream`` (same windows AND same permutations as ``get_batch`` for one rng)."""
        self.B = int(batch_size)
        self.ring = [make_block() for _ in range(int(ring))]
        self._np = [(np.asarray(a), np.asarray(b)) for a, b in self.ring]       # zero-copy NumPy views of the blocks
        self._native = bool(native)
        self._src = (batch_fillers(blocks, self.B, shuffle_size, rng, threads) if self._native
                     else batch_stream(blocks, self.B, shuffle_size, rng))
        self._ready = queue.Queue()
        self._free = queue.Queue()                      # blocks the consumer is done with (FIFO: batch k uses block k % ring)
        for slot in range(len(self.ring)):
            self._free.put(slot)
        self._held = None                               # the block handed out last: released by the NEXT __next__ call
        self._reusable = reusable
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._fill, daemon=True)
        self._thread.start()

    def _fill(self):
        k = 0
        try:
            for item in self._src:
                slot = None
                while slot is None:
                    if self._stop.is_set():
                        return
                    try:
                        slot = self._free.get(timeout=0.1)
                    except queue.Empty:
                        continue
                if k >= len(self.ring) and self._reusable is not None:
                    self._reusable(k - len(self.ring))          # the step that last used this block has been staged
                ids, cnt = self._np[slot]
                if self._native:
                    item(ids, cnt)
                else:
                    ids[0, :], ids[1, :], cnt[:] = item
                self._ready.put(slot)
                k += 1
            self._ready.put(None)
        except Exception as e:  # pragma: no cover
            self._ready.put(e)

    def __iter__(self):
        return self

    def __next__(self):
        """The next batch block.  It stays untouched until the following ``__next__`` call (and, with ``reusable``, until
        the trainer has staged it)."""
        if self._held is not None:
            self._free.put(self._held)
            self._held = None
        slot = self._ready.get()
        if slot is None:
            raise StopIteration
        if isinstance(slot, Exception):
            raise slot
        self._held = slot
        return self.ring[slot]

    def close(self):
        self._stop.set()


def train_from(loader, trainer, steps=None):
    """``for batch in loader: trainer.submit(*batch)`` -- the body of train_epoch (wikipedia/train_cooccurence.py:103-112)
    with the loader one step ahead.  Returns the number of steps submitted."""
    n = 0
    for ids, counts in loader:
        trainer.submit(ids, counts)
        n += 1
        if steps is not None and n >= steps:
            break
    return n
