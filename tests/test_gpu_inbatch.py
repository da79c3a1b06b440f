"""GPU parity of the tcgen05 in-batch score kernel (esr_inbatch_fwd_bwd_bf16) against oracle/inbatch.py.

Tolerances: the contract rounds Q, K to bf16 first, so every product is exact in fp32 and only the
summation order differs -> scores / losses / gradients within 1e-5 (abs + rel).  A hinge mask bit can
legitimately differ where |margin| < 1e-5 (the two summation orders straddle zero); the backward is
therefore checked against the oracle evaluated on the KERNEL's own mask (exact), and the mask itself
against the oracle's outside that band.  Softmax probabilities are rounded to bf16 by contract: the
backward is checked on the kernel's own bf16 probabilities, and those against the oracle's to 1 bf16 ulp."""
import numpy as np
import pytest
import torch

from oracle import inbatch as oib

pytestmark = pytest.mark.gpu


def _data(Bq, Bk, D, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    Q = (rng.standard_normal((Bq, D)) * spread / np.sqrt(np.sqrt(D))).astype(np.float32)
    K = (rng.standard_normal((Bk, D)) * spread / np.sqrt(np.sqrt(D))).astype(np.float32)
    return Q, K


def _run(Q, K, **kw):
    from esrecsys_b200.engine import InBatchScorer
    kw.setdefault("chunk_rows", 1 << 20)     # one chunk: the whole dL/dS matrix stays inspectable
    sc = InBatchScorer(Q.shape[0], Q.shape[1], Bk=K.shape[0], **kw)
    loss, dQ, dK = sc.run(torch.from_numpy(Q).cuda(), torch.from_numpy(K).cuda())
    torch.cuda.synchronize()
    G, diag, cnt, lse = sc.debug_views()
    return (float(loss.item()), dQ.cpu().numpy(), dK.cpu().numpy(), G.float().cpu().numpy(), diag.cpu().numpy(),
            cnt.cpu().numpy(), lse.cpu().numpy())


SHAPES = [(128, 128, 128, 0), (256, 256, 64, 0), (384, 384, 128, 0), (200, 200, 128, 0), (1024, 1024, 256, 0),
          (128, 512, 128, 256), (333, 777, 192, 100), (2048, 2048, 128, 0),
          (1, 1, 64, 0), (2, 7, 128, 3), (130, 129, 256, 0)]       # single pair; tiny ragged; one tile + 1-2 rows


@pytest.mark.parametrize("Bq,Bk,D,off", SHAPES)
def test_hinge_parity(Bq, Bk, D, off):
    Q, K = _data(Bq, Bk, D, 1)
    loss, dQ, dK, G, diag, cnt, _ = _run(Q, K, loss="hinge", diag_off=off)
    S, Qh, Kh = oib.scores(Q, K)
    pj = np.arange(Bq) + off
    has = pj < Bk                                     # queries whose positive lies outside the item batch score 0
    np.testing.assert_allclose(diag[has], S[np.arange(Bq)[has], pj[has]], rtol=1e-5, atol=1e-5)
    mask, h = oib.hinge_mask(S, off)
    gm = G > 0.5
    assert set(np.unique(G)) <= {0.0, 1.0}
    differ = gm != mask
    assert np.all(np.abs(h[differ]) < 1e-4), "mask differs outside the rounding band: %d cells" % differ.sum()
    assert differ.mean() < 1e-4
    np.testing.assert_array_equal(cnt.sum(0), gm.sum(1))
    ref_loss = np.sum(np.where(gm, h, 0).astype(np.float64)) / Bq
    assert abs(loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    rdQ, rdK = oib.hinge_backward(gm, Qh, Kh, off)
    np.testing.assert_allclose(dQ, rdQ, rtol=1e-5, atol=1e-5 * np.abs(rdQ).max())
    np.testing.assert_allclose(dK, rdK, rtol=1e-5, atol=1e-5 * np.abs(rdK).max())


@pytest.mark.parametrize("Bq,Bk,D,off", SHAPES)
def test_softmax_parity(Bq, Bk, D, off):
    Q, K = _data(Bq, Bk, D, 2)
    loss, dQ, dK, G, diag, _, lse = _run(Q, K, loss="softmax", diag_off=off)
    S, Qh, Kh = oib.scores(Q, K)
    P, rlse = oib.softmax_probs(S)
    np.testing.assert_allclose(lse, rlse, rtol=1e-5, atol=1e-5)
    rloss, _, _, Pb = oib.softmax(Q, K, off)
    assert abs(loss - rloss) <= 1e-5 * max(1.0, abs(rloss))
    # probabilities: bf16 by contract -> within one bf16 ulp (2^-8 relative) of the oracle's
    np.testing.assert_allclose(G, Pb, rtol=2.0 ** -7, atol=1e-30)
    rdQ, rdK = oib.softmax_backward(G, Qh, Kh, off)
    np.testing.assert_allclose(dQ, rdQ, rtol=1e-5, atol=1e-5 * np.abs(rdQ).max())
    np.testing.assert_allclose(dK, rdK, rtol=1e-5, atol=1e-5 * np.abs(rdK).max())
    # and end to end against the oracle's own rounding: a handful of 1-ulp flips of P only
    _, odQ, odK, _ = oib.softmax(Q, K, off)
    assert np.abs(dQ - odQ).max() <= 2e-3 * np.abs(odQ).max()


def test_hinge_margin_scale_norm_and_determinism():
    Q, K = _data(512, 512, 128, 3)
    a = _run(Q, K, loss="hinge", margin=0.25, scale=0.5, b_norm=4096.0)
    b = _run(Q, K, loss="hinge", margin=0.25, scale=0.5, b_norm=4096.0)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    S, Qh, Kh = oib.scores(Q, K)
    mask, h = oib.hinge_mask(S, 0, 0.25, 0.5)
    gm = a[3] > 0.5
    assert np.all(np.abs(h[gm != mask]) < 1e-4)
    rdQ, rdK = oib.hinge_backward(gm, Qh, Kh, 0, 0.5, 4096.0)
    np.testing.assert_allclose(a[1], rdQ, rtol=1e-5, atol=1e-5 * np.abs(rdQ).max())
    np.testing.assert_allclose(a[2], rdK, rtol=1e-5, atol=1e-5 * np.abs(rdK).max())
    ref_loss = np.sum(np.where(gm, h, 0).astype(np.float64)) / 4096.0
    assert abs(a[0] - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))


def test_splitk_matches_single():
    Q, K = _data(640, 640, 128, 5)
    a = _run(Q, K, loss="softmax", splits=1)
    b = _run(Q, K, loss="softmax", splits=4)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("kind", ["hinge", "softmax"])
@pytest.mark.parametrize("Bq,Bk,off,chunk", [(1000, 1000, 0, 256), (700, 1500, 300, 128), (2048, 2048, 0, 512)])
def test_row_chunks_match_single_pass(kind, Bq, Bk, off, chunk):
    """The L2-resident row-chunked schedule (the default at large B) against one pass over the whole matrix:
    same mask / probabilities, only the split of the dK sum differs."""
    Q, K = _data(Bq, Bk, 128, 6)
    a = _run(Q, K, loss=kind, diag_off=off)
    b = _run(Q, K, loss=kind, diag_off=off, chunk_rows=chunk)
    assert abs(a[0] - b[0]) <= 1e-6 * max(1.0, abs(a[0]))
    np.testing.assert_allclose(a[1], b[1], rtol=1e-5, atol=1e-5 * np.abs(a[1]).max())
    np.testing.assert_allclose(a[2], b[2], rtol=1e-5, atol=1e-5 * np.abs(a[2]).max())


def test_full_size_properties():
    """BASELINE configs[2] size (B = 8192, D = 128): size-independent checks -- sum_j dS_ij = 0 per row for
    softmax (dQ of a constant K is 0), and the loss of identical rows is log(B)."""
    B, D = 8192, 128
    Q = np.full((B, D), 0.25, np.float32)
    K = np.full((B, D), 0.5, np.float32)
    loss, dQ, dK, G, *_ = _run(Q, K, loss="softmax", chunk_rows=0)
    assert abs(loss - np.log(B)) < 1e-4
    # p_ij = 1/B = 2^-13 exactly in bf16 -> dQ_i = (sum_j p_ij K_j - K_i)/B = 0
    assert np.abs(dQ).max() < 1e-9 and np.abs(dK).max() < 1e-9
    Q, K = _data(B, B, D, 7)
    from esrecsys_b200.engine import InBatchScorer
    assert InBatchScorer(B, D, chunk_rows=2048).plan()["n_chunks"] == 4
    auto = _run(Q, K, loss="hinge", chunk_rows=2048)
    loss, dQ, dK, G, diag, cnt, _ = _run(Q, K, loss="hinge")
    np.testing.assert_allclose(auto[1], dQ, rtol=1e-5, atol=1e-5 * np.abs(dQ).max())
    np.testing.assert_allclose(auto[2], dK, rtol=1e-5, atol=1e-5 * np.abs(dK).max())
    S, Qh, Kh = oib.scores(Q, K)
    np.testing.assert_allclose(diag, np.diag(S), rtol=1e-5, atol=1e-5)
    gm = G > 0.5
    rdQ, rdK = oib.hinge_backward(gm, Qh, Kh, 0)
    np.testing.assert_allclose(dQ, rdQ, rtol=1e-5, atol=1e-5 * np.abs(rdQ).max())
    np.testing.assert_allclose(dK, rdK, rtol=1e-5, atol=1e-5 * np.abs(rdK).max())


@pytest.mark.parametrize("kind", ["hinge", "softmax"])
def test_shared_table_steps_match_oracle(kind):
    """configs[2] shape (scaled down): 3 consecutive steps, duplicate-heavy Zipf ids."""
    from esrecsys_b200 import engine, synth
    from esrecsys_b200.inbatch import SharedTableInBatch
    V, D, B, lr = 5000, 128, 512, 0.05
    rng = np.random.default_rng(11)
    E = (rng.standard_normal((V, D)) / D ** 0.25).astype(np.float32)
    q, k = synth.pair_batches(V, V, B, 3, 5)
    table = engine.EmbeddingTable.from_dense(E, sparse=False, adagrad=True)
    tr = SharedTableInBatch(table, B, lr=lr, loss=kind)
    Eo, acc = E.copy(), np.full_like(E, 0.1)
    for s in range(3):
        ids = torch.from_numpy(np.stack([q[s], k[s]])).cuda()
        got = float(tr.step(ids).item())
        want = oib.shared_table_step(Eo, acc, q[s], k[s], lr, kind)
        assert abs(got - want) <= 2e-5 * max(1.0, abs(want))
    # Rows match to 1e-5 except where a hinge-mask bit sits inside the rounding band of the two summation orders
    # (or a bf16 probability flips by one ulp): those rows move by ~lr * delta_g / sqrt(acc), bounded below.
    got, gacc = table.rows0.cpu().numpy(), table.acc.cpu().numpy()
    bad = np.abs(got - Eo) > 1e-5 + 1e-5 * np.abs(Eo)
    assert bad.mean() < 2e-3, bad.mean()
    assert np.abs(got - Eo).max() < 1e-3
    np.testing.assert_allclose(gacc, acc, rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("kind", ["hinge", "softmax"])
def test_shared_table_graphed_step_equals_eager(kind):
    """SharedTableInBatch.graphed(): the CUDA-graph replay of the trainer step is the eager step, bit for bit, and the
    warm-up / capture calls leave no trace in the table."""
    from esrecsys_b200 import engine, synth
    from esrecsys_b200.inbatch import SharedTableInBatch
    V, D, B, lr = 5000, 128, 512, 0.05
    rng = np.random.default_rng(11)
    E = (rng.standard_normal((V, D)) / D ** 0.25).astype(np.float32)
    q, k = synth.pair_batches(V, V, B, 3, 5)
    outs = []
    for graphed in (False, True):
        table = engine.EmbeddingTable.from_dense(E, sparse=False, adagrad=True)
        tr = SharedTableInBatch(table, B, lr=lr, loss=kind)
        ids = [torch.from_numpy(np.stack([q[s], k[s]])).cuda() for s in range(3)]
        step = tr.graphed(ids[0]) if graphed else tr.step
        if graphed:
            assert np.array_equal(table.rows0.cpu().numpy(), E)            # capture left the table untouched
        losses = [float(step(ids[s]).item()) for s in range(3)]
        outs.append((losses, table.rows0.cpu().numpy(), table.acc.cpu().numpy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])


def test_two_tower_steps_match_oracle():
    """configs[3] shape (scaled down): id tables + MLP towers + softmax in-batch loss, 3 steps."""
    from esrecsys_b200 import engine, synth
    from esrecsys_b200.inbatch import TwoTowerInBatch
    V, D, B = 3000, 64, 256
    rng = np.random.default_rng(12)
    Es = (rng.standard_normal((V, D)) / D ** 0.5).astype(np.float32)
    Ep = (rng.standard_normal((V, D)) / D ** 0.5).astype(np.float32)
    ts = engine.EmbeddingTable.from_dense(Es, sparse=False, adagrad=True)
    tp = engine.EmbeddingTable.from_dense(Ep, sparse=False, adagrad=True)
    tr = TwoTowerInBatch(ts, tp, B, lr=0.05, tower_lr=1e-3, loss="softmax", scale=4.0, seed=3)
    ps = {k: v.cpu().numpy().copy() for k, v in tr.scene_tower.p.items()}
    pp = {k: v.cpu().numpy().copy() for k, v in tr.product_tower.p.items()}
    mk = lambda p: dict(count=0, mu={k: np.zeros_like(v) for k, v in p.items()}, nu={k: np.zeros_like(v) for k, v in p.items()})
    os_, op_ = mk(ps), mk(pp)
    Eso, Epo = Es.copy(), Ep.copy()
    accs, accp = np.full_like(Es, 0.1), np.full_like(Ep, 0.1)
    s_ids, p_ids = synth.pair_batches(V, V, B, 3, 9)
    for s in range(3):
        got = float(tr.step(torch.from_numpy(s_ids[s]).cuda(), torch.from_numpy(p_ids[s]).cuda()).item())
        want = oib.two_tower_step(Eso, accs, Epo, accp, ps, pp, os_, op_, s_ids[s], p_ids[s], 0.05, 1e-3, "softmax", 1.0, 4.0)
        assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (s, got, want)
    np.testing.assert_allclose(ts.rows0.cpu().numpy(), Eso, rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(tp.rows0.cpu().numpy(), Epo, rtol=1e-3, atol=2e-4)
    for k in ps:
        np.testing.assert_allclose(tr.scene_tower.p[k].cpu().numpy(), ps[k], rtol=1e-3, atol=2e-4)


def test_two_tower_tf32_towers_stay_close_and_restore_the_flag():
    """tower_matmul="tf32": tensor-core GEMMs in the towers (10-bit mantissa inputs) -- loss within 2e-3 of the fp32 oracle
    on the first step, and torch's global TF32 switch is left as it was found."""
    from esrecsys_b200 import engine, synth
    from esrecsys_b200.inbatch import TwoTowerInBatch
    V, D, B = 3000, 64, 256
    rng = np.random.default_rng(12)
    Es = (rng.standard_normal((V, D)) / D ** 0.5).astype(np.float32)
    Ep = (rng.standard_normal((V, D)) / D ** 0.5).astype(np.float32)
    ts = engine.EmbeddingTable.from_dense(Es, sparse=False, adagrad=True)
    tp = engine.EmbeddingTable.from_dense(Ep, sparse=False, adagrad=True)
    tr = TwoTowerInBatch(ts, tp, B, lr=0.05, tower_lr=1e-3, loss="softmax", scale=4.0, seed=3, tower_matmul="tf32")
    ps = {k: v.cpu().numpy().copy() for k, v in tr.scene_tower.p.items()}
    pp = {k: v.cpu().numpy().copy() for k, v in tr.product_tower.p.items()}
    mk = lambda p: dict(count=0, mu={k: np.zeros_like(v) for k, v in p.items()}, nu={k: np.zeros_like(v) for k, v in p.items()})
    s_ids, p_ids = synth.pair_batches(V, V, B, 1, 9)
    before = torch.backends.cuda.matmul.allow_tf32
    got = float(tr.step(torch.from_numpy(s_ids[0]).cuda(), torch.from_numpy(p_ids[0]).cuda()).item())
    assert torch.backends.cuda.matmul.allow_tf32 == before
    want = oib.two_tower_step(Es.copy(), np.full_like(Es, 0.1), Ep.copy(), np.full_like(Ep, 0.1), ps, pp, mk(ps), mk(pp),
                              s_ids[0], p_ids[0], 0.05, 1e-3, "softmax", 1.0, 4.0)
    assert abs(got - want) <= 2e-3 * max(1.0, abs(want)), (got, want)


@pytest.mark.parametrize("D", [32, 128, 256])
def test_segment_sum_rows_long_and_short_segments(D):
    """esr_segment_sum_rows_f32 (per-row gradient sums of the in-batch trainers, the owner-side merge of the NCCL
    exchange): one id owning a third of the slots (the block-cooperative path at D >= 128), pairs, singletons; twice
    the same launch gives the same bits (fixed summation tree)."""
    import ctypes as C
    from esrecsys_b200 import _lib as L
    from esrecsys_b200 import engine
    n, V = 6000, 4000
    rng = np.random.default_rng(D)
    keys = rng.integers(0, V, size=n).astype(np.int32)
    keys[rng.permutation(n)[:2000]] = 77                      # a 2000-slot segment
    keys[rng.permutation(n)[:70]] = 1234                      # one just above the long-segment threshold
    g = rng.standard_normal((n, D)).astype(np.float32)
    gb = rng.standard_normal(n).astype(np.float32)
    plan = engine.IndexPlan(n, V, with_partner=False).build(torch.from_numpy(keys).cuda())
    U = int(plan.n_uniq.item())
    g_d, gb_d = torch.from_numpy(g).cuda(), torch.from_numpy(gb).cuda()
    outs = []
    for _ in range(2):
        out = torch.full((n, D), np.nan, dtype=torch.float32, device="cuda")
        outb = torch.full((n,), np.nan, dtype=torch.float32, device="cuda")
        L.check(L.lib().esr_segment_sum_rows_f32(C.byref(plan.s), D, L.ptr(g_d), L.ptr(gb_d), L.ptr(out), L.ptr(outb),
                                                 L.stream_ptr()), "esr_segment_sum_rows_f32")
        outs.append((out[:U].cpu().numpy(), outb[:U].cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    uniq, inv = np.unique(keys, return_inverse=True)
    want = np.zeros((len(uniq), D), np.float64)
    wantb = np.zeros(len(uniq), np.float64)
    np.add.at(want, inv, g.astype(np.float64))
    np.add.at(wantb, inv, gb.astype(np.float64))
    assert U == len(uniq)
    np.testing.assert_allclose(outs[0][0], want, rtol=1e-5, atol=2e-5 * np.sqrt(2000))
    np.testing.assert_allclose(outs[0][1], wantb, rtol=1e-5, atol=2e-5 * np.sqrt(2000))
