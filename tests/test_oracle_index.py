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
