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
  wikipedia/make_dictionary.py:113-116, so block sharding would hot-spot rank 0),
* the peer-memory path's owner-side id pull / resolve and source-side emit map
  (csrc/peer_ops.cu; pinned by the conservation checks in tests/test_oracle_index.py).

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


# --------------------------------------------------------------------------
# Peer-memory path (csrc/peer_ops.cu): what every rank derives, on the device, from
# the route plans all ranks published -- no id exchange, no host-known sizes.
# ``all_counts[s][q]`` = rows source s routes to owner q (route_plan counts),
# ``all_send_local[s]`` = source s's owner-local ids in owner-bucket order.
# --------------------------------------------------------------------------

EMIT_SHIFT = 27     # emit_map = owner << 27 | index in the owner's inbox (include/esr.h, EsrGloveCfg.emit_map)


def peer_pull_ids(all_counts, all_send_local, me: int, map_stride: int):
    """Owner ``me``: esr_peer_pull_ids_i32.  Returns ``(recv_ids int32[total], src_meta int32[3n+4],
    slot_map int32[n, map_stride])``: recv_ids is source-major; src_meta[3s:3s+3] = (offset of source s in
    recv_ids, its count, displacement of my bucket inside source s's list), src_meta[3n] = total;
    slot_map[s, x] = position of owner-local row x in source s's bucket, -1 if s does not name x."""
    counts = np.asarray(all_counts, np.int64)
    n = counts.shape[0]
    src_meta = np.zeros(3 * n + 4, np.int32)
    slot_map = np.full((n, map_stride), -1, np.int32)
    parts, off = [], 0
    for s in range(n):
        dsp = int(counts[s, :me].sum())
        cnt = int(counts[s, me])
        src_meta[3 * s: 3 * s + 3] = (off, cnt, dsp)
        ids = np.asarray(all_send_local[s], np.int32)[dsp: dsp + cnt]
        slot_map[s, ids] = np.arange(cnt, dtype=np.int32)
        parts.append(ids)
        off += cnt
    src_meta[3 * n] = off
    recv_ids = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    return recv_ids, src_meta, slot_map


def peer_resolve(n: int, recv_ids, src_meta, slot_map):
    """esr_peer_resolve_i32.  For received entry k (source s, row x): the entry of the FIRST source naming x owns
    the row; ``desc[k, q]`` = index in my inbox of source q's gradient row for x (offset_q + position) for
    q >= s naming x, -1 otherwise; non-owning entries are all -1.  Returns ``(desc int32[total, n], own bool[total])``
    (the kernel's compacted own_list holds the k with own[k], in no particular order)."""
    total = int(src_meta[3 * n])
    desc = np.full((total, n), -1, np.int32)
    own = np.zeros(total, bool)
    offs = [int(src_meta[3 * s]) for s in range(n)]
    for k in range(total):
        s = 0
        while s + 1 < n and k >= offs[s + 1]:
            s += 1
        x = int(recv_ids[k])
        pos = slot_map[:, x]
        if (pos[:s] >= 0).any():
            continue
        own[k] = True
        for q in range(s, n):
            if pos[q] >= 0:
                desc[k, q] = offs[q] + int(pos[q])
    return desc, own


def peer_emit_map(all_counts, me: int, uniq, inv_order, n: int):
    """Source ``me``: esr_peer_emit_plan_i32.  emit_map[u] = owner << 27 | (rows the sources before me send to that
    owner + position of row u inside my bucket for it); inv_order[u] = position of uniq[u] in my owner-bucket order."""
    counts = np.asarray(all_counts, np.int64)
    uniq = np.asarray(uniq, np.int64)
    own = uniq % n
    off = counts[:me, :].sum(axis=0)                     # [owner]
    dsp = np.concatenate([[0], np.cumsum(counts[me])])[:n]
    idx = off[own] + np.asarray(inv_order, np.int64) - dsp[own]
    return ((own << EMIT_SHIFT) | idx).astype(np.int32)


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


# ---- owner-computes pair routing (esr_peer_route_pairs_i32 / esr_peer_collect_pairs_i32, csrc/peer_ops.cu) ---------------
def route_pairs(ids, counts, n: int):
    """Source side.  ``ids`` int32 (2, B) = [i ; j] (wikipedia/cooccurrence_matrix.py:103-114), ``counts`` f32 (B,).
    Stable partition of the pairs by owner(i) = i % n.  Returns ``(per_owner, send_counts)`` with
    ``per_owner[o] = (i, j, x)`` arrays in original pair order and ``send_counts[o] = len``."""
    ids = np.asarray(ids, np.int32)
    counts = np.asarray(counts, np.float32)
    own = ids[0].astype(np.int64) % n
    per_owner, send_counts = [], np.zeros(n, np.int32)
    for o in range(n):
        m = own == o                                  # boolean mask keeps the original order: stable
        per_owner.append((ids[0][m].copy(), ids[1][m].copy(), counts[m].copy()))
        send_counts[o] = int(m.sum())
    return per_owner, send_counts


def collect_pairs(regions, B_cap: int, pad_key: int):
    """Owner side.  ``regions[s] = (i, j, x)`` received from source s; concatenated in SOURCE order into the flat
    [i ; j] key array of capacity 2 * B_cap (padding = pad_key), the counts (padding 0) and n_valid = 2 m.
    Returns ``(keys, counts, n_valid, overflow)``."""
    gi = np.concatenate([np.asarray(r[0], np.int32) for r in regions])
    gj = np.concatenate([np.asarray(r[1], np.int32) for r in regions])
    gx = np.concatenate([np.asarray(r[2], np.float32) for r in regions])
    m_all = gi.size
    m = min(m_all, B_cap)
    keys = np.full(2 * B_cap, pad_key, np.int32)
    cnt = np.zeros(B_cap, np.float32)
    keys[:m] = gi[:m]
    keys[B_cap:B_cap + m] = gj[:m]
    cnt[:m] = gx[:m]
    return keys, cnt, 2 * m, m_all > B_cap


def routed_batches(ids_per_rank, counts_per_rank, n: int):
    """The batch every owner processes after routing: ``out[o] = (i, j, x)`` = pairs of ALL ranks whose row i lives on
    o, source-major, original order inside a source.  The union over o is the global batch (conservation)."""
    routed = [route_pairs(ids_per_rank[r], counts_per_rank[r], n)[0] for r in range(n)]
    return [tuple(np.concatenate([routed[s][o][k] for s in range(n)]) for k in range(3)) for o in range(n)]
