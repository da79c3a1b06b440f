"""Oracle for the integer bookkeeping of the sparse update path (TEST INFRASTRUCTURE).

The reference has no index bookkeeping of its own: XLA's gather / scatter-add
(the VJP of ``jnp.take`` behind ``nn.Embed``, wikipedia/models.py:31-34,
spotify/models.py:42-45) hides it.  SURVEY.md section 0.1 row D7 therefore makes
*this file* the bit-exact contract for the CUDA path:

* slot numbering of a GloVe batch (``[i ; j]`` because both roles hit the same
  table, wikipedia/models.py:16-19,31-34),
* stable sort of the slots by table row,
* unique rows + segment offsets,
* cyclic owner routing for the row-sharded table (rows are frequency ranks,
  wikipedia/make_dictionary.py:113-116, so block sharding would hot-spot rank 0).

Parity unpinned (no reference tests exist); everything here is plain integer
NumPy and is pinned by hand-written micro cases in tests/test_oracle_index.py.
"""
from __future__ import annotations

import numpy as np


def slot_keys(i: np.ndarray, j: np.ndarray) -> np.ndarray:
    """Slots ``s in [0, 2B)``: ``key[s] = i[s]`` for ``s < B`` else ``j[s-B]``."""
    return np.concatenate([np.asarray(i, np.int32), np.asarray(j, np.int32)])


def sort_slots(keys: np.ndarray):
    """Stable ascending sort of slots by row id.

    Returns ``(sorted_keys int32[n], perm int32[n])`` with
    ``sorted_keys == keys[perm]``; ties keep ascending slot order.
    """
    keys = np.asarray(keys, np.int32)
    perm = np.argsort(keys, kind="stable").astype(np.int32)
    return keys[perm], perm


def segments(sorted_keys: np.ndarray):
    """Unique rows and segment offsets of a sorted key array.

    Returns ``(uniq int32[U], seg_off int32[U+1])``; slots of ``uniq[u]`` are
    ``sorted positions [seg_off[u], seg_off[u+1])``.
    """
    sorted_keys = np.asarray(sorted_keys, np.int32)
    n = sorted_keys.shape[0]
    if n == 0:
        return np.zeros(0, np.int32), np.zeros(1, np.int32)
    head = np.ones(n, bool)
    head[1:] = sorted_keys[1:] != sorted_keys[:-1]
    starts = np.flatnonzero(head).astype(np.int32)
    return sorted_keys[starts], np.concatenate([starts, np.array([n], np.int32)])


def slot_segment_index(sorted_keys: np.ndarray) -> np.ndarray:
    """``useg[p]`` = index into ``uniq`` of sorted position ``p``."""
    sorted_keys = np.asarray(sorted_keys, np.int32)
    n = sorted_keys.shape[0]
    if n == 0:
        return np.zeros(0, np.int32)
    head = np.ones(n, np.int32)
    head[1:] = sorted_keys[1:] != sorted_keys[:-1]
    return (np.cumsum(head) - 1).astype(np.int32)


# --------------------------------------------------------------------------
# Row-sharded table: cyclic ownership and all-to-all routing plan
# --------------------------------------------------------------------------

def owner_of(rows: np.ndarray, n_ranks: int) -> np.ndarray:
    return (np.asarray(rows, np.int64) % n_ranks).astype(np.int32)


def local_row(rows: np.ndarray, n_ranks: int) -> np.ndarray:
    return (np.asarray(rows, np.int64) // n_ranks).astype(np.int32)


def global_row(local: np.ndarray, rank: int, n_ranks: int) -> np.ndarray:
    return (np.asarray(local, np.int64) * n_ranks + rank).astype(np.int32)


def shard_rows(V: int, rank: int, n_ranks: int) -> int:
    """Number of rows rank ``rank`` owns under cyclic sharding."""
    return (V - rank + n_ranks - 1) // n_ranks


def route_plan(uniq: np.ndarray, n_ranks: int):
    """Bucket a rank's sorted unique rows by owner.

    Returns ``(send_counts int32[n], send_displs int32[n+1], send_local int32[U],
    order int32[U])``.  ``order`` is the stable permutation of ``uniq`` that
    groups rows by owner (ascending row within an owner); ``send_local`` are the
    owner-local row ids in that order -- the payload of the index all-to-all.
    """
    uniq = np.asarray(uniq, np.int32)
    own = owner_of(uniq, n_ranks)
    order = np.argsort(own, kind="stable").astype(np.int32)
    counts = np.bincount(own, minlength=n_ranks).astype(np.int32)
    displs = np.zeros(n_ranks + 1, np.int32)
    displs[1:] = np.cumsum(counts)
    return counts, displs, local_row(uniq[order], n_ranks), order


def exchange_counts(all_send_counts: np.ndarray):
    """Simulate the count all-to-all: ``recv_counts[r][s] = send_counts[s][r]``."""
    return np.ascontiguousarray(np.asarray(all_send_counts, np.int32).T)


def _splitmix64(x):
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


def sample_uniform(seed, step, n, hi):
    """Bit-exact restatement of esr_sample_uniform_i32 (csrc/table_ops.cu): our own counter-based stream that
    replaces jax.random.randint in sample_negative (spotify/train_spotify.py:139-150) -- not threefry."""
    m = (1 << 64) - 1
    base = _splitmix64((seed ^ ((step * 0xD1342543DE82EF95) & m)) & m)
    out = np.empty(n, np.int32)
    for k in range(n):
        r = _splitmix64((base + k) & m)
        out[k] = ((r >> 32) * hi) >> 32
    return out
