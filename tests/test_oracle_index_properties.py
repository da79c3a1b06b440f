"""Property tests (hypothesis) of the integer oracle oracle/index.py -- the bit-exact contract of the CUDA index plan, the
cyclic row sharding and the owner-computes pair routing.  Size-independent properties: sortedness, stability, partition,
conservation; they are what the GPU tests assert at full size where the oracle itself is the checker."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import index as oidx  # noqa: E402

ids_arrays = st.integers(1, 4000).flatmap(
    lambda V: st.tuples(st.just(V), st.lists(st.integers(0, V - 1), min_size=0, max_size=300)))


@settings(max_examples=80, deadline=None, derandomize=True)
@given(ids_arrays)
def test_sort_is_the_stable_permutation_and_segments_partition_the_slots(arg):
    V, keys = arg
    keys = np.asarray(keys, np.int32)
    sk, perm = oidx.sort_slots(keys)
    assert np.array_equal(np.sort(perm), np.arange(keys.size))                 # a permutation of the slots
    assert np.array_equal(keys[perm], sk) and np.all(np.diff(sk) >= 0)         # ... that sorts the keys
    same = np.diff(sk) == 0
    assert np.all(np.diff(perm)[same] > 0)                                     # equal keys keep their slot order
    uniq, off = oidx.segments(sk)
    assert np.array_equal(uniq, np.unique(keys))
    assert off[0] == 0 and off[-1] == keys.size and np.all(np.diff(off) > 0) if keys.size else off.tolist() == [0]
    useg = oidx.slot_segment_index(sk)
    assert np.array_equal(uniq[useg], sk)
    for u in range(len(uniq)):
        assert np.all(sk[off[u]:off[u + 1]] == uniq[u])


@settings(max_examples=60, deadline=None, derandomize=True)
@given(ids_arrays, st.integers(1, 8))
def test_cyclic_sharding_is_a_bijection_and_the_route_plan_groups_by_owner(arg, n):
    V, keys = arg
    uniq = np.unique(np.asarray(keys, np.int32))
    own, loc = oidx.owner_of(uniq, n), oidx.local_row(uniq, n)
    for r in range(n):
        assert np.array_equal(oidx.global_row(loc[own == r], r, n), uniq[own == r])
        assert np.all(loc[own == r] < oidx.shard_rows(V, r, n))
    assert sum(oidx.shard_rows(V, r, n) for r in range(n)) == V
    counts, displs, send_local, order = oidx.route_plan(uniq, n)
    assert counts.sum() == uniq.size and displs[-1] == uniq.size
    routed = uniq[order]
    for r in range(n):
        part = routed[displs[r]:displs[r + 1]]
        assert np.all(part % n == r) and np.all(np.diff(part) > 0)             # by owner, ascending inside an owner
        assert np.array_equal(send_local[displs[r]:displs[r + 1]], part // n)


pair_batches = st.integers(2, 500).flatmap(lambda V: st.tuples(
    st.just(V), st.integers(1, 4),
    st.lists(st.tuples(st.integers(0, V - 1), st.integers(0, V - 1), st.floats(1.0, 500.0, width=32)), min_size=1, max_size=60)))


@settings(max_examples=60, deadline=None, derandomize=True)
@given(pair_batches, st.integers(0, 3))
def test_pair_routing_conserves_every_pair_and_keeps_source_order(arg, seed):
    V, n, pairs = arg
    rng = np.random.default_rng(seed)
    B = len(pairs)
    ids, cnt = [], []
    for r in range(n):                                                          # every rank: a shuffle of the same pairs
        p = rng.permutation(B)
        ids.append(np.array([[pairs[k][0] for k in p], [pairs[k][1] for k in p]], np.int32))
        cnt.append(np.array([pairs[k][2] for k in p], np.float32))
    routed = oidx.routed_batches(ids, cnt, n)
    got = sorted((int(i), int(j), float(x)) for o in range(n) for i, j, x in zip(*routed[o]))
    want = sorted((int(i), int(j), float(x)) for r in range(n) for i, j, x in zip(ids[r][0], ids[r][1], cnt[r]))
    assert got == want                                                          # the union of the owners' batches = the global batch
    for o in range(n):
        assert np.all(routed[o][0] % n == o)                                    # every pair sits at the owner of its row i
        # source-major, original order inside a source
        pos = 0
        for r in range(n):
            m = ids[r][0] % n == o
            k = int(m.sum())
            assert np.array_equal(routed[o][0][pos:pos + k], ids[r][0][m])
            assert np.array_equal(routed[o][1][pos:pos + k], ids[r][1][m])
            pos += k
        assert pos == routed[o][0].size
    # collect: capacity padding with the key V, overflow flagged
    per_owner = [oidx.route_pairs(ids[r], cnt[r], n)[0] for r in range(n)]
    for o in range(n):
        regions = [per_owner[s][o] for s in range(n)]
        m_all = sum(r[0].size for r in regions)
        for cap in (m_all + 3, max(1, m_all - 1)):
            keys, c, n_valid, over = oidx.collect_pairs(regions, cap, V)
            m = min(m_all, cap)
            assert n_valid == 2 * m and over == (m_all > cap)
            assert np.all(keys[m:cap] == V) and np.all(keys[cap + m:] == V) and np.all(c[m:] == 0)
            assert np.array_equal(keys[:m], routed[o][0][:m]) and np.array_equal(keys[cap:cap + m], routed[o][1][:m])


@settings(max_examples=40, deadline=None, derandomize=True)
@given(st.integers(0, 2 ** 40), st.integers(0, 1000), st.integers(1, 200), st.integers(1, 10 ** 8))
def test_uniform_sampler_is_in_range_and_a_pure_function_of_seed_and_step(seed, step, n, hi):
    a = oidx.sample_uniform(seed, step, n, hi)
    assert a.dtype == np.int32 and a.shape == (n,) and a.min() >= 0 and a.max() < hi
    assert np.array_equal(a, oidx.sample_uniform(seed, step, n, hi))
    assert np.array_equal(a[:n // 2], oidx.sample_uniform(seed, step, n // 2, hi)) if n >= 2 else True
