"""Pins oracle.index on hand-computed micro cases (the bit-exact bookkeeping contract)."""
import numpy as np

from oracle import index as oidx


def test_sort_and_segments_micro():
    i = np.array([5, 3, 5, 9], np.int32)
    j = np.array([3, 1, 2, 5], np.int32)
    keys = oidx.slot_keys(i, j)
    assert keys.tolist() == [5, 3, 5, 9, 3, 1, 2, 5]
    sk, perm = oidx.sort_slots(keys)
    assert sk.tolist() == [1, 2, 3, 3, 5, 5, 5, 9]
    assert perm.tolist() == [5, 6, 1, 4, 0, 2, 7, 3]          # stable: ties keep slot order
    uniq, off = oidx.segments(sk)
    assert uniq.tolist() == [1, 2, 3, 5, 9]
    assert off.tolist() == [0, 1, 2, 4, 7, 8]
    assert oidx.slot_segment_index(sk).tolist() == [0, 1, 2, 2, 3, 3, 3, 4]


def test_empty_and_single():
    sk, perm = oidx.sort_slots(np.zeros(0, np.int32))
    uniq, off = oidx.segments(sk)
    assert uniq.size == 0 and off.tolist() == [0]
    uniq, off = oidx.segments(np.array([7, 7, 7], np.int32))
    assert uniq.tolist() == [7] and off.tolist() == [0, 3]


def test_cyclic_routing_micro():
    uniq = np.array([0, 1, 2, 5, 8, 9, 13], np.int32)
    counts, displs, send_local, order = oidx.route_plan(uniq, 4)
    # owners: 0,1,2,1,0,1,1
    assert counts.tolist() == [2, 4, 1, 0]
    assert displs.tolist() == [0, 2, 6, 7, 7]
    assert uniq[order].tolist() == [0, 8, 1, 5, 9, 13, 2]
    assert send_local.tolist() == [0, 2, 0, 1, 2, 3, 0]
    for r in range(4):
        loc = send_local[displs[r]:displs[r + 1]]
        assert oidx.global_row(loc, r, 4).tolist() == uniq[order][displs[r]:displs[r + 1]].tolist()


def test_shard_rows_partition():
    for V in (1, 7, 8, 9, 1000003):
        for n in (1, 2, 3, 8):
            assert sum(oidx.shard_rows(V, r, n) for r in range(n)) == V
            rows = np.arange(min(V, 50))
            assert (oidx.global_row(oidx.local_row(rows, n), 0, n) + oidx.owner_of(rows, n) == rows).all()


def test_exchange_counts_transpose():
    sc = np.arange(9, dtype=np.int32).reshape(3, 3)
    rc = oidx.exchange_counts(sc)
    assert rc[2, 0] == sc[0, 2] and rc.sum() == sc.sum()


def _published(uniqs, n):
    """Route plans of n sources: (all_counts [n][n], all_send_local, orders, inv_orders)."""
    counts, sls, orders, invs = [], [], [], []
    for u in uniqs:
        c, _, sl, order = oidx.route_plan(u, n)
        inv = np.empty(len(u), np.int32)
        inv[order] = np.arange(len(u), dtype=np.int32)
        counts.append(c); sls.append(sl); orders.append(order); invs.append(inv)
    return np.stack(counts), sls, orders, invs


def test_peer_pull_resolve_emit_micro():
    # 2 ranks; source 0 names rows {0,1,2,4}, source 1 names {1,2,3}; owner 0 holds even rows, owner 1 odd rows
    uniqs = [np.array([0, 1, 2, 4], np.int32), np.array([1, 2, 3], np.int32)]
    counts, sls, orders, invs = _published(uniqs, 2)
    assert counts.tolist() == [[3, 1], [1, 2]]
    assert sls[0].tolist() == [0, 1, 2, 0] and sls[1].tolist() == [1, 0, 1]
    # owner 0: source 0 sends local rows 0,1,2 (global 0,2,4), source 1 sends local row 1 (global 2)
    recv, meta, smap = oidx.peer_pull_ids(counts, sls, 0, 4)
    assert recv.tolist() == [0, 1, 2, 1]
    assert meta[:6].tolist() == [0, 3, 0, 3, 1, 0] and meta[6] == 4
    assert smap.tolist() == [[0, 1, 2, -1], [-1, 0, -1, -1]]
    desc, own = oidx.peer_resolve(2, recv, meta, smap)
    assert own.tolist() == [True, True, True, False]            # source 1's entry for local row 1 is owned by source 0's
    assert desc.tolist() == [[0, -1], [1, 3], [2, -1], [-1, -1]]
    # owner 1: source 0 sends local 0 (global 1), displacement 3 in its list; source 1 sends local 0, 1 (global 1, 3)
    recv, meta, smap = oidx.peer_pull_ids(counts, sls, 1, 4)
    assert recv.tolist() == [0, 0, 1] and meta[:6].tolist() == [0, 1, 3, 1, 2, 1]
    desc, own = oidx.peer_resolve(2, recv, meta, smap)
    assert own.tolist() == [True, False, True] and desc.tolist() == [[0, 1], [-1, -1], [-1, 2]]
    # sources: where each unique row's gradient lands (owner << 27 | inbox index)
    em0 = oidx.peer_emit_map(counts, 0, uniqs[0], invs[0], 2)
    em1 = oidx.peer_emit_map(counts, 1, uniqs[1], invs[1], 2)
    S = oidx.EMIT_SHIFT
    assert em0.tolist() == [0, (1 << S) | 0, 1, 2]
    assert em1.tolist() == [(1 << S) | 1, 3, (1 << S) | 2]


def test_peer_scatter_merge_conserves_every_gradient():
    """n virtual ranks: every source scatters one gradient per unique row through emit_map into the owners' inboxes;
    every owner merges through desc.  Each (source, row) gradient must be summed exactly once, in source order."""
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 4, 8):
        V = 997
        sizes = rng.integers(0, 300, size=n)
        sizes[rng.integers(0, n)] = 0 if n > 2 else sizes[0]                     # an empty source
        uniqs = [np.unique(rng.integers(0, V, size=int(sz))).astype(np.int32) for sz in sizes]
        counts, sls, orders, invs = _published(uniqs, n)
        V_max = oidx.shard_rows(V, 0, n)
        inbox_cap = int(sum(len(u) for u in uniqs)) + 1
        inbox = [np.full(inbox_cap, np.nan) for _ in range(n)]
        filled = [np.zeros(inbox_cap, bool) for _ in range(n)]
        for s in range(n):                                                       # the row pass's peer scatter
            em = oidx.peer_emit_map(counts, s, uniqs[s], invs[s], n)
            o, idx = em >> oidx.EMIT_SHIFT, em & ((1 << oidx.EMIT_SHIFT) - 1)
            for u in range(len(uniqs[s])):
                assert o[u] == uniqs[s][u] % n
                assert not filled[o[u]][idx[u]], "two gradients land in one inbox slot"
                filled[o[u]][idx[u]] = True
                inbox[o[u]][idx[u]] = 1000.0 * s + uniqs[s][u]                   # identifies (source, row)
        expect = {}
        for s in range(n):
            for r in uniqs[s]:
                expect.setdefault(int(r), []).append(1000.0 * s + int(r))
        seen = set()
        for me in range(n):                                                      # the owners' merge
            recv, meta, smap = oidx.peer_pull_ids(counts, sls, me, V_max)
            total = int(meta[3 * n])
            assert total == int(counts[:, me].sum()) and filled[me].sum() == total
            desc, own = oidx.peer_resolve(n, recv, meta, smap)
            for k in np.flatnonzero(own):
                g = int(oidx.global_row(recv[k:k + 1], me, n)[0])
                got = [inbox[me][d] for d in desc[k] if d >= 0]
                assert got == expect[g], (n, me, g)                              # exactly the sources naming g, in source order
                assert g not in seen
                seen.add(g)
            assert (desc[~own] == -1).all()
        assert seen == set(expect)


def test_route_pairs_hand_case_and_conservation():
    """Owner-computes pair routing (oracle side of esr_peer_route_pairs_i32 / esr_peer_collect_pairs_i32)."""
    ids = np.array([[5, 4, 9, 2, 7, 4], [1, 3, 2, 1, 6, 0]], np.int32)
    x = np.array([.5, 1., 2., 4., 8., 16.], np.float32)
    per, sc = oidx.route_pairs(ids, x, 2)                        # owner = i % 2
    assert sc.tolist() == [3, 3]
    assert per[0][0].tolist() == [4, 2, 4] and per[0][1].tolist() == [3, 1, 0] and per[0][2].tolist() == [1., 4., 16.]
    assert per[1][0].tolist() == [5, 9, 7] and per[1][1].tolist() == [1, 2, 6] and per[1][2].tolist() == [.5, 2., 8.]
    keys, cnt, nv, over = oidx.collect_pairs([per[0], per[1]], 8, 100)
    assert nv == 12 and not over
    assert keys.tolist() == [4, 2, 4, 5, 9, 7, 100, 100] + [3, 1, 0, 1, 2, 6, 100, 100]
    assert cnt.tolist() == [1., 4., 16., .5, 2., 8., 0., 0.]
    keys, cnt, nv, over = oidx.collect_pairs([per[0], per[1]], 4, 100)
    assert nv == 8 and over and keys.tolist() == [4, 2, 4, 5, 3, 1, 0, 1]
    # conservation over n ranks: the routed batches are a permutation of the global batch, every i on its owner
    rng = np.random.default_rng(0)
    for n in (2, 3, 8):
        B = 257
        idr = [rng.integers(1, 5000, (2, B)).astype(np.int32) for _ in range(n)]
        xr = [rng.random(B).astype(np.float32) for _ in range(n)]
        out = oidx.routed_batches(idr, xr, n)
        assert sum(o[0].size for o in out) == n * B
        for o, (i, j, c) in enumerate(out):
            assert (i % n == o).all()
        got = sorted(zip(np.concatenate([o[0] for o in out]).tolist(), np.concatenate([o[1] for o in out]).tolist(),
                         np.concatenate([o[2] for o in out]).tolist()))
        want = sorted(zip(np.concatenate([a[0] for a in idr]).tolist(), np.concatenate([a[1] for a in idr]).tolist(),
                          np.concatenate(xr).tolist()))
        assert got == want
