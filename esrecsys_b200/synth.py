"""Seeded synthetic index streams and batches (host side, NumPy only).

The reference trains on the Wikipedia co-occurrence dump, the Spotify MPD and
the STL image set, none of which exist offline.  These generators produce
batches of the SAME layout the reference's iterators yield
(``((2,B) int32, (B,) float32)`` -- wikipedia/cooccurrence_matrix.py:103-114) with
the index distribution the book itself uses for token frequencies: bounded Zipf
with exponent 1, ``P(k) = (1/k) / H_M`` (book-text/CH2-Math-considerations.tex:15-19).
Ids are frequency ranks (wikipedia/make_dictionary.py:113-116), row 0 is the mask
token (wikipedia/token_dictionary.py:58-64), and ``i > j`` always
(wikipedia/make_cooccurrence.py:48).
"""
from __future__ import annotations

import numpy as np


class ZipfStream:
    """Bounded Zipf(1) over ranks ``1..M`` by inverse CDF; ``uniform=True`` is the no-reuse control."""

    def __init__(self, M: int, seed: int, uniform: bool = False):
        self.M = int(M)
        self.rng = np.random.default_rng(seed)
        self.uniform = uniform
        if not uniform:
            cdf = np.cumsum(1.0 / np.arange(1, self.M + 1, dtype=np.float64))
            self.cdf = cdf / cdf[-1]

    def draw(self, n: int) -> np.ndarray:
        """Ranks in ``[1, M]`` as int64."""
        if self.uniform:
            return self.rng.integers(1, self.M + 1, size=n, dtype=np.int64)
        u = self.rng.random(n)
        return np.minimum(np.searchsorted(self.cdf, u, side="left"), self.M - 1).astype(np.int64) + 1


def glove_counts(rng: np.random.Generator, n: int) -> np.ndarray:
    """Emulates ``count = sum 1/dist`` (wikipedia/make_cooccurrence.py:33-55): LogNormal(0,1.5) in [1/9, 1e4]."""
    c = rng.lognormal(mean=0.0, sigma=1.5, size=n)
    return np.clip(c, 1.0 / 9.0, 1.0e4).astype(np.float32)


def glove_batches(V: int, B: int, n_batches: int, seed: int, uniform: bool = False):
    """``n_batches`` GloVe batches: ``ids int32 (n, 2, B)`` with ``V > i > j >= 1``, ``counts f32 (n, B)``."""
    stream = ZipfStream(V - 1, seed, uniform)
    crng = np.random.default_rng(seed + 7919)
    ids = np.empty((n_batches, 2, B), np.int32)
    counts = np.empty((n_batches, B), np.float32)
    for k in range(n_batches):
        a = stream.draw(B)
        b = stream.draw(B)
        bad = a == b
        while bad.any():
            b[bad] = stream.draw(int(bad.sum()))
            bad = a == b
        ids[k, 0] = np.maximum(a, b)
        ids[k, 1] = np.minimum(a, b)
        counts[k] = glove_counts(crng, B)
    return ids, counts


def pair_batches(Vq: int, Vk: int, B: int, n_batches: int, seed: int, uniform: bool = False):
    """(query, item) id pairs for the in-batch-negative configs; ranks go through a fixed permutation
    (first-seen id order, spotify/make_dictionary.py:41-45)."""
    sq = ZipfStream(Vq, seed, uniform)
    sk = ZipfStream(Vk, seed + 1, uniform)
    pq = np.random.default_rng(seed + 2).permutation(Vq).astype(np.int32)
    pk = np.random.default_rng(seed + 3).permutation(Vk).astype(np.int32)
    q = np.stack([pq[sq.draw(B) - 1] for _ in range(n_batches)])
    k = np.stack([pk[sk.draw(B) - 1] for _ in range(n_batches)])
    return q, k


def init_glove_tables(V: int, D: int, seed: int):
    """flax ``nn.Embed`` defaults: table ~ N(0, 1/D) (variance_scaling fan_in, out_axis=0), bias zeros
    (wikipedia/models.py:15-19)."""
    rng = np.random.default_rng(seed + 1000)
    E = (rng.standard_normal((V, D), dtype=np.float32) / np.float32(np.sqrt(D))).astype(np.float32)
    b = np.zeros(V, np.float32)
    return E, b


def spotify_example(rng: np.random.Generator, m: int, o: int = 64, n_tracks: int = 2262292,
                    n_albums: int = 734684, n_artists: int = 295860):
    """One playlist record in the spotify/input_pipeline.py:23-30 layout plus sampled negatives
    (spotify/train_spotify.py:139-150).  Ids are raw (album ids exceed the 100000-row hashed table)."""
    def ids(hi, n):
        return rng.integers(0, hi, size=n, dtype=np.int64)
    x = {
        "track_context": ids(n_tracks, 5), "album_context": ids(n_albums, 5), "artist_context": ids(n_artists, 5),
        "next_track": ids(n_tracks, m), "next_album": ids(n_albums, m), "next_artist": ids(n_artists, m),
        "neg_track": ids(n_tracks, o), "neg_album": ids(n_albums, o), "neg_artist": ids(n_artists, o),
    }
    # playlists repeat artists/albums: force some duplicates and context hits so the isin boosts,
    # max ties and duplicate-row gradient merges are exercised.
    if m >= 3:
        x["next_artist"][1] = x["artist_context"][0]
        x["next_album"][2] = x["album_context"][1]
        x["next_artist"][m - 1] = x["next_artist"][0]
    x["artist_context"][4] = x["artist_context"][3]
    x["album_context"][4] = x["album_context"][3]
    return x
