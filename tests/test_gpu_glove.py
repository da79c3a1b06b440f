"""GPU parity: the CUDA GloVe path (through the C ABI) vs the NumPy oracle on the same seeded batches.

Integer bookkeeping (sort permutation, unique rows, segment offsets) is bit-exact; floating point is
compared at 1e-5 (abs + rel, fp32) as BASELINE.json's north_star states.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from esrecsys_b200 import synth
from oracle import glove as og
from oracle import index as oidx
from oracle import optim as oopt

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ATOL = 1e-5
# Row-pass variants under test (every one has run on a B200; "auto" is the default).
IMPLS = ["auto", "ldg", "tma", "fifo", "endfirst"]


def _engine():
    from esrecsys_b200 import engine
    return engine


def _batch(V, B, seed, uniform=False, n=1):
    ids, counts = synth.glove_batches(V, B, n, seed, uniform)
    return ids, counts


def _tables(V, D, seed, bias_scale=0.05):
    E, b = synth.init_glove_tables(V, D, seed)
    rng = np.random.default_rng(seed + 5)
    b = (rng.standard_normal(V) * bias_scale).astype(np.float32)   # non-zero biases exercise the broadcast terms
    return E, b


@pytest.mark.parametrize("V,B", [(50, 64), (1000, 2048), (10000, 2048), (200000, 4096)])
def test_plan_bit_exact(V, B):
    eng = _engine()
    ids, _ = _batch(V, B, seed=V + B)
    keys = torch.from_numpy(ids[0].reshape(-1)).cuda()
    plan = eng.IndexPlan(2 * B, V).build(keys)
    sk, perm, uniq, seg_off = plan.host_view()
    osk, operm = oidx.sort_slots(oidx.slot_keys(ids[0, 0], ids[0, 1]))
    ouniq, ooff = oidx.segments(osk)
    assert np.array_equal(sk, osk)
    assert np.array_equal(perm, operm)
    assert np.array_equal(uniq, ouniq)
    assert np.array_equal(seg_off, ooff)
    assert np.array_equal(plan.useg[:2 * B].cpu().numpy(), oidx.slot_segment_index(osk))
    i, j = ids[0]
    pair = operm % B
    partner = np.where(operm >= B, i[pair], j[pair])
    assert np.array_equal(plan.partner[:2 * B].cpu().numpy(), partner)
    # remap: ids rewritten as indices into uniq
    remap = plan.remap_ids().cpu().numpy()
    assert np.array_equal(ouniq[remap], oidx.slot_keys(i, j))


def _check_plan(plan, keys_np, n_real=None):
    """Every output of esr_plan_build_i32 against oracle/index.py (stable argsort): bit for bit."""
    n = keys_np.size
    n_real = n if n_real is None else n_real
    osk, operm = oidx.sort_slots(keys_np)
    sk, perm = plan.sorted_keys[:n].cpu().numpy(), plan.perm[:n].cpu().numpy()
    assert np.array_equal(sk, osk)
    assert np.array_equal(perm, operm)
    ouniq, ooff = oidx.segments(osk[:n_real])
    U = int(plan.n_uniq.item())
    assert U == len(ouniq)
    assert np.array_equal(plan.uniq[:U].cpu().numpy(), ouniq)
    assert np.array_equal(plan.seg_off[:U + 1].cpu().numpy(), ooff)
    assert np.array_equal(plan.useg[:n_real].cpu().numpy(), oidx.slot_segment_index(osk[:n_real]))
    if plan.partner is not None and n % 2 == 0:
        half = n // 2
        other = np.where(operm < half, operm + half, operm - half)
        assert np.array_equal(plan.partner[:n_real].cpu().numpy(), keys_np[other][:n_real])


@pytest.mark.parametrize("sort", ["own", "cub"])
@pytest.mark.parametrize("V,n,dist", [
    (10, 2, "uniform"), (10, 62, "zipf"), (300, 2046, "uniform"), (300, 2048, "zipf"), (300, 2050, "uniform"),
    (2 ** 16, 4098, "zipf"), (2 ** 16 + 1, 6144, "uniform"), (10 ** 6, 100000, "zipf"), (10 ** 6, 2 * 262144, "zipf"),
    (10 ** 6, 2 * 262144, "uniform"), (10 ** 8, 2 * 262144, "zipf"), (10 ** 8, 300002, "uniform"),
    (2 ** 31 - 1, 70000, "uniform"), (5, 2 * 262144, "uniform"), (10 ** 6, 2 * 262144, "same")])
def test_plan_sort_every_shape(V, n, dist, sort, monkeypatch):
    """The two plan builders (libesr's wide sort + fused head pass, cub + the head count / scan / write kernels -- chosen
    by EsrPlan.sort_impl, here forced through ESR_PLAN_SORT): tile tails, one to four digit passes, skewed / uniform /
    constant keys, keys up to 2^31 - 2."""
    eng = _engine()
    monkeypatch.setenv("ESR_PLAN_SORT", sort)
    rng = np.random.default_rng(V % 1000 + n)
    if dist == "uniform":
        keys = rng.integers(0, V, size=n, dtype=np.int64)
    elif dist == "zipf":
        keys = np.minimum(rng.zipf(1.1, size=n) - 1, V - 1)
    else:
        keys = np.full(n, V - 1)
    keys = keys.astype(np.int32)
    plan = eng.IndexPlan(n, V).build(torch.from_numpy(keys).cuda())
    _check_plan(plan, keys)
    # the same plan object again with other keys (workspace counters are re-zeroed by every build)
    keys2 = np.ascontiguousarray(keys[::-1])
    plan.build(torch.from_numpy(keys2).cuda())
    _check_plan(plan, keys2)


@pytest.mark.parametrize("n_real", [0, 1, 2047, 2048, 2049, 40000, 65536])
@pytest.mark.parametrize("sort", ["wide", "library"])
def test_plan_sort_padded(n_real, sort):
    """EsrPlan.n_valid: capacity 65536 slots, the tail padded with the key V (sorts last, no segment of its own)."""
    eng = _engine()
    V, cap = 5000, 65536
    rng = np.random.default_rng(n_real)
    keys = np.full(cap, V, np.int32)
    keys[:n_real] = np.minimum(rng.zipf(1.2, size=n_real) - 1, V - 1)
    rng.shuffle(keys)                                   # padding anywhere in the slot array
    nv = torch.tensor([n_real], dtype=torch.int32, device="cuda")
    plan = eng.IndexPlan(cap, V + 1, with_partner=False, n_valid=nv, sort=sort).build(torch.from_numpy(keys).cuda())
    _check_plan(plan, keys, n_real)


def test_plan_empty_and_all_same():
    eng = _engine()
    plan = eng.IndexPlan(0, 10).build(torch.zeros(0, dtype=torch.int32, device="cuda"))
    assert int(plan.n_uniq.item()) == 0 and int(plan.seg_off[0].item()) == 0
    keys = torch.full((64,), 7, dtype=torch.int32, device="cuda")
    plan = eng.IndexPlan(64, 10).build(keys)
    sk, perm, uniq, seg_off = plan.host_view()
    assert uniq.tolist() == [7] and seg_off.tolist() == [0, 64] and perm.tolist() == list(range(64))


def test_check_ids():
    eng = _engine()
    ids = torch.tensor([0, 5, 9, 10, -1], dtype=torch.int32, device="cuda")
    assert eng.check_ids(ids, 10) == 2
    assert eng.check_ids(ids[:3], 10) == 0


@pytest.mark.parametrize("D", [4, 32, 64, 100, 128, 256, 512])
def test_gather_and_export_bit_exact(D):
    eng = _engine()
    V = 3000
    E, b = _tables(V, D, seed=D)
    t = eng.EmbeddingTable.from_dense(E, b)
    rng = np.random.default_rng(D)
    ids = rng.integers(0, V, size=4097).astype(np.int32)
    out = t.gather(torch.from_numpy(ids)).cpu().numpy()
    assert np.array_equal(out, E[ids])
    assert np.array_equal(t.dense().cpu().numpy(), E)
    assert t.gather(torch.zeros(0, dtype=torch.int32)).shape == (0, D)


def _emit(E, b, ids, counts, bias_mode, chunk=0, impl="auto"):
    eng = _engine()
    V, D = E.shape
    B = ids.shape[1]
    t = eng.EmbeddingTable.from_dense(E, b, sparse=False)
    keys = torch.from_numpy(ids.reshape(-1)).cuda()
    plan = eng.IndexPlan(2 * B, V).build(keys)
    step = eng.GloveStep(t, B, bias_mode=bias_mode, emit_grads=True, chunk=chunk, impl=impl)
    sc = step.run(plan, torch.from_numpy(counts).cuda()).cpu().numpy()
    U = int(plan.n_uniq.item())
    return sc, step.dE[:U].cpu().numpy(), step.db[:U].cpu().numpy(), plan


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("bias_mode", ["reference_broadcast", "per_pair"])
@pytest.mark.parametrize("V,D,B,chunk", [
    (10000, 64, 2048, 0),      # BASELINE config 1 shape
    (10000, 64, 2048, 8),
    (50, 64, 512, 0),          # duplicate-heavy: every row straddles many chunks
    (50, 128, 512, 4),
    (100000, 128, 4096, 16),
    (2000, 256, 1000, 0),
    (2000, 100, 777, 12),      # D not a multiple of 32, ragged B
    (2000, 512, 300, 0),
    (3, 8, 1, 0),              # single pair
])
def test_grads_match_oracle(V, D, B, chunk, bias_mode, impl):
    E, b = _tables(V, D, seed=V + D)
    ids, counts = _batch(V, B, seed=B + D)
    sc, dE, db, plan = _emit(E, b, ids[0], counts[0], bias_mode, chunk, impl)
    gr = og.loss_and_grads(E, b, ids[0, 0], ids[0, 1], counts[0], bias_mode)
    assert np.array_equal(plan.uniq[:len(gr.uniq)].cpu().numpy(), gr.uniq)
    np.testing.assert_allclose(sc[5], gr.loss, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(sc[2], gr.S0, rtol=RTOL)
    np.testing.assert_allclose(sc[3], gr.S1, rtol=1e-4, atol=1e-3)   # a large cancelling sum
    np.testing.assert_allclose(dE, gr.dE, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(db, gr.db, rtol=RTOL, atol=ATOL)


def test_grads_i_equals_j_and_literal_loss():
    """i == j cannot occur in real data (make_cooccurrence.py:48) but both terms must then apply;
    the loss is also checked against the literal (B,B) evaluation of the reference forward."""
    V, D, B = 40, 64, 96
    E, b = _tables(V, D, seed=1)
    ids, counts = _batch(V, B, seed=2)
    ids = ids[0].copy()
    ids[1, :10] = ids[0, :10]
    sc, dE, db, _ = _emit(E, b, ids, counts[0], "reference_broadcast")
    gr = og.loss_and_grads(E, b, ids[0], ids[1], counts[0])
    np.testing.assert_allclose(dE, gr.dE, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(db, gr.db, rtol=RTOL, atol=ATOL)
    lit = og.loss_literal(E.astype(np.float64), b.astype(np.float64), ids[0], ids[1], counts[0].astype(np.float64))
    np.testing.assert_allclose(sc[5], lit, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("bias_mode", ["reference_broadcast", "per_pair"])
@pytest.mark.parametrize("V,D,B,uniform", [(10000, 64, 2048, False), (300, 128, 1024, False), (50000, 128, 8192, True),
                                           (20, 256, 4096, False)])
def test_adagrad_steps_match_oracle(V, D, B, uniform, bias_mode, impl):
    """Several fused sparse-Adagrad steps (north-star rule) vs oracle.glove.step_adagrad."""
    eng = _engine()
    n_steps = 5
    lr = 0.05
    E, b = _tables(V, D, seed=V)
    ids, counts = _batch(V, B, seed=V + 1, uniform=uniform, n=n_steps)
    t = eng.EmbeddingTable.from_dense(E, b, sparse=True)
    step = eng.GloveStep(t, B, lr=lr, bias_mode=bias_mode, impl=impl)
    plan = eng.IndexPlan(2 * B, V)
    Eo, bo = E.copy(), b.copy()
    accE = np.full_like(Eo, oopt.ADAGRAD_INIT_ACC)
    accb = np.full_like(bo, oopt.ADAGRAD_INIT_ACC)
    for k in range(n_steps):
        keys = torch.from_numpy(ids[k].reshape(-1)).cuda()
        plan.build(keys)
        sc = step.run(plan, torch.from_numpy(counts[k]).cuda())
        loss = float(sc[5].item())
        oloss = og.step_adagrad(Eo, bo, accE, accb, ids[k, 0], ids[k, 1], counts[k], lr, bias_mode)
        np.testing.assert_allclose(loss, oloss, rtol=2e-5, atol=ATOL)
    np.testing.assert_allclose(t.dense().cpu().numpy(), Eo, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(t.bias.cpu().numpy(), bo, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(t.acc.cpu().numpy(), accE, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(t.bias_acc.cpu().numpy(), accb, rtol=RTOL, atol=ATOL)
    # untouched rows are bit-identical to the initial table
    touched = np.zeros(V, bool)
    touched[ids.reshape(-1)] = True
    assert np.array_equal(t.dense().cpu().numpy()[~touched], E[~touched])


def test_step_is_deterministic():
    """Sorted segment sums + fixed-order partial combine: two runs give identical bits."""
    eng = _engine()
    V, D, B = 500, 128, 4096
    E, b = _tables(V, D, seed=3)
    ids, counts = _batch(V, B, seed=4)
    outs = []
    for _ in range(2):
        t = eng.EmbeddingTable.from_dense(E, b, sparse=True)
        plan = eng.IndexPlan(2 * B, V).build(torch.from_numpy(ids[0].reshape(-1)).cuda())
        eng.GloveStep(t, B).run(plan, torch.from_numpy(counts[0]).cuda())
        outs.append(t.dense().cpu().numpy())
    assert np.array_equal(outs[0], outs[1])


def test_empty_batch():
    eng = _engine()
    E, b = _tables(16, 64, seed=0)
    t = eng.EmbeddingTable.from_dense(E, b, sparse=True)
    plan = eng.IndexPlan(0, 16).build(torch.zeros(0, dtype=torch.int32, device="cuda"))
    sc = eng.GloveStep(t, 0).run(plan, torch.zeros(0, dtype=torch.float32, device="cuda"))
    assert float(sc[5].item()) == 0.0
    assert np.array_equal(t.dense().cpu().numpy(), E)


def test_optimizer_kernels_match_oracle():
    eng = _engine()
    rng = np.random.default_rng(0)
    n = 100003
    p = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    mu = np.zeros(n, np.float32)
    nu = np.zeros(n, np.float32)
    tp, tmu, tnu = (torch.from_numpy(x.copy()).cuda() for x in (p, mu, nu))
    count = 0
    for k in range(3):
        gk = (g * (k + 1)).astype(np.float32)
        p, mu, nu, count = oopt.adam_update(p, gk, mu, nu, count, 1e-3)
        eng.dense_adam(tp, torch.from_numpy(gk).cuda(), tmu, tnu, 1e-3, count)
    np.testing.assert_allclose(tp.cpu().numpy(), p, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(tnu.cpu().numpy(), nu, rtol=RTOL, atol=1e-9)
    p2 = rng.standard_normal(n).astype(np.float32)
    tr = np.zeros(n, np.float32)
    tp2, ttr = torch.from_numpy(p2.copy()).cuda(), torch.from_numpy(tr.copy()).cuda()
    for k in range(3):
        p2, tr = oopt.sgdm_update(p2, g, tr, 1e-3, 0.98)
        eng.dense_sgdm(tp2, torch.from_numpy(g).cuda(), ttr, 1e-3, 0.98)
    np.testing.assert_allclose(tp2.cpu().numpy(), p2, rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(ttr.cpu().numpy(), tr, rtol=RTOL, atol=1e-6)


def test_sparse_adagrad_and_scatter_rows():
    eng = _engine()
    V, D = 1000, 128
    E, b = _tables(V, D, seed=9)
    t = eng.EmbeddingTable.from_dense(E, b, sparse=True)
    rng = np.random.default_rng(1)
    uniq = np.sort(rng.choice(V, 300, replace=False)).astype(np.int32)
    g = rng.standard_normal((300, D)).astype(np.float32)
    gb = rng.standard_normal(300).astype(np.float32)
    cap = 512
    tu = torch.zeros(cap, dtype=torch.int32, device="cuda")
    tu[:300] = torch.from_numpy(uniq).cuda()
    tg = torch.zeros(cap, D, device="cuda")
    tg[:300] = torch.from_numpy(g).cuda()
    tgb = torch.zeros(cap, device="cuda")
    tgb[:300] = torch.from_numpy(gb).cuda()
    nu = torch.tensor([300], dtype=torch.int32, device="cuda")
    eng.sparse_adagrad(t, tu, nu, tg, tgb, 0.05)
    Eo, bo = E.copy(), b.copy()
    acc = np.full_like(Eo, 0.1)
    accb = np.full_like(bo, 0.1)
    Eo[uniq], acc[uniq] = oopt.adagrad_update(Eo[uniq], g, acc[uniq], 0.05)
    bo[uniq], accb[uniq] = oopt.adagrad_update(bo[uniq], gb, accb[uniq], 0.05)
    np.testing.assert_allclose(t.dense().cpu().numpy(), Eo, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(t.bias.cpu().numpy(), bo, rtol=RTOL, atol=ATOL)
    dense = torch.zeros(V, D, device="cuda")
    eng.scatter_rows(dense, tu, nu, tg)
    ref = np.zeros((V, D), np.float32)
    ref[uniq] = g
    assert np.array_equal(dense.cpu().numpy(), ref)
    eng.scatter_rows(dense, tu, nu, tg, accumulate=True)
    assert np.array_equal(dense.cpu().numpy(), ref * 2)


def test_bench_shape_trainer_matches_oracle():
    """Parity AT THE BENCHMARKED SHAPE (BASELINE configs[1]: V = 1M, D = 128, B = 262 144, Zipf(1)) through the exact
    code path bench.py times: GloveTrainer with CUDA graphs, two streams, the 8/9 persistent row-pass grid and pinned
    host batches -- three graph-replayed steps against oracle.glove.step_adagrad
    (wikipedia/train_cooccurence.py:71-101 with the north-star Adagrad rule)."""
    from esrecsys_b200 import synth
    from esrecsys_b200.trainer import GloveTrainer
    eng = _engine()
    V, D, B, lr, steps = 1_000_000, 128, 262144, 0.05, 3
    E, b = synth.init_glove_tables(V, D, 0)
    b = (np.random.default_rng(3).standard_normal(V) * 0.05).astype(np.float32)
    ids, counts = synth.glove_batches(V, B, steps, 0)
    table = eng.EmbeddingTable.from_dense(E, b, sparse=True)
    tr = GloveTrainer(table, B, lr=lr)                              # bench.py's construction (graphs=True, depth=2)
    assert tr.use_graphs and tr.g_step[0] is not None
    Eo, bo = E.copy(), b.copy()
    aE, ab = np.full_like(Eo, oopt.ADAGRAD_INIT_ACC), np.full_like(bo, oopt.ADAGRAD_INIT_ACC)
    ref = []
    for k in range(steps):
        tr.submit(torch.from_numpy(ids[k]).pin_memory(), torch.from_numpy(counts[k]).pin_memory())
        ref.append(og.step_adagrad(Eo, bo, aE, ab, ids[k, 0], ids[k, 1], counts[k], lr))
    got = tr.losses(0, steps)
    np.testing.assert_allclose(got, np.array(ref, np.float32), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(table.dense().cpu().numpy(), Eo, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(table.bias.cpu().numpy(), bo, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(table.acc.cpu().numpy(), aE, rtol=RTOL, atol=1e-7)
    touched = np.zeros(V, bool)
    touched[np.unique(ids)] = True
    assert np.array_equal(table.dense().cpu().numpy()[~touched], E[~touched])      # untouched rows bit-identical
