"""Multi-rank parity of the peer-memory sharded path on ONE GPU: N virtual ranks (tests/virtual_peers.py) drive the
same libesr kernels PeerShardedGloveTrainer uses, with every "peer" pointer local.

* integer kernels (esr_peer_pull_ids_i32 / esr_peer_resolve_i32 / esr_peer_emit_plan_i32) bit-exact against
  oracle/index.py for N in {2, 3, 4, 8};
* the whole sharded step against the single-table oracle on the rank-major concatenated batch, 1e-5.

First run on a B200 in round 2 (12 passed); part of the default -m gpu suite since."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _batches(V, B_loc, n, steps, seed):
    from esrecsys_b200 import synth
    return synth.glove_batches(V, B_loc * n, steps, seed)          # global batches, rank r takes columns [r B, (r+1) B)


@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_peer_integer_kernels_bit_exact(n):
    from oracle import index as oidx
    from virtual_peers import VirtualPeerGlove
    V, D, B = 3001, 64, 512
    vp = VirtualPeerGlove(V, D, B, n)
    ids, _ = _batches(V, B, n, 1, 7)
    per_rank = [torch.from_numpy(np.ascontiguousarray(ids[0][:, r * B:(r + 1) * B])) for r in range(n)]
    vp.plan_phase(per_rank)
    vp.resolve_phase()
    torch.cuda.synchronize()
    uniqs, invs = [], []
    for r, k in enumerate(vp.ranks):
        U = int(k.plan.n_uniq.item())
        uniq = k.plan.uniq[:U].cpu().numpy()
        assert np.array_equal(uniq, np.unique(per_rank[r].numpy()))
        c, _, sl, order = oidx.route_plan(uniq, n)
        assert np.array_equal(k.counts[:n].cpu().numpy(), c)
        assert np.array_equal(k.send_local[:U].cpu().numpy(), sl)
        assert np.array_equal(k.order[:U].cpu().numpy(), order)
        inv = np.empty(U, np.int32)
        inv[order] = np.arange(U, dtype=np.int32)
        assert np.array_equal(k.inv_order[:U].cpu().numpy(), inv)
        uniqs.append(uniq)
        invs.append(inv)
    counts = np.stack([k.counts[:n].cpu().numpy() for k in vp.ranks])
    sls = [k.send_local.cpu().numpy() for k in vp.ranks]
    for me, k in enumerate(vp.ranks):
        recv, meta, smap = oidx.peer_pull_ids(counts, sls, me, vp.map_stride)
        total = int(meta[3 * n])
        got_meta = k.src_meta.cpu().numpy()
        assert np.array_equal(got_meta[:3 * n + 1], meta[:3 * n + 1])
        assert np.array_equal(k.recv_ids[:total].cpu().numpy(), recv)
        assert np.array_equal(k.slot_map.cpu().numpy(), smap)
        desc, own = oidx.peer_resolve(n, recv, meta, smap)
        assert np.array_equal(k.desc[:total * n].cpu().numpy().reshape(total, n), desc)
        n_own = int(got_meta[3 * n + 1])
        assert n_own == int(own.sum())
        recs = k.desc[vp.inbox_cap * n: vp.inbox_cap * n + 2 * n_own].cpu().numpy().reshape(n_own, 2)   # {k | multi << 31, x}
        own_k = recs[:, 0] & 0x7fffffff
        order_k = np.argsort(own_k)
        assert np.array_equal(own_k[order_k], np.flatnonzero(own))
        assert np.array_equal(recs[order_k, 1], recv[own])                          # owner-local row of every record
        assert np.array_equal(recs[order_k, 0] < 0, (desc[own] >= 0).sum(axis=1) > 1)   # several-sources flag
        em = oidx.peer_emit_map(counts, me, uniqs[me], invs[me], n)
        assert np.array_equal(k.emit_map[:len(em)].cpu().numpy(), em)
        assert int(k.err.item()) == 0


@pytest.mark.parametrize("bias_mode", ["reference_broadcast", "per_pair"])
@pytest.mark.parametrize("n,V,D,B_loc", [(2, 5000, 64, 1024), (3, 300, 128, 512), (4, 5000, 128, 512), (8, 2000, 64, 256)])
def test_virtual_sharded_step_matches_single_table_oracle(n, V, D, B_loc, bias_mode):
    from esrecsys_b200 import synth
    from oracle import glove as og
    from oracle import optim as oopt
    from virtual_peers import VirtualPeerGlove
    steps = 3
    E, _ = synth.init_glove_tables(V, D, 0)
    b = (np.random.default_rng(5).standard_normal(V) * 0.05).astype(np.float32)
    ids, counts = _batches(V, B_loc, n, steps, 1)
    vp = VirtualPeerGlove(V, D, B_loc, n, lr=0.05, bias_mode=bias_mode)
    vp.load_dense(E, b)
    Eo, bo = E.copy(), b.copy()
    aE, ab = np.full_like(Eo, oopt.ADAGRAD_INIT_ACC), np.full_like(bo, oopt.ADAGRAD_INIT_ACC)
    for s in range(steps):
        per_ids = [torch.from_numpy(np.ascontiguousarray(ids[s][:, r * B_loc:(r + 1) * B_loc])) for r in range(n)]
        per_cnt = [torch.from_numpy(np.ascontiguousarray(counts[s][r * B_loc:(r + 1) * B_loc])) for r in range(n)]
        loss = vp.step(per_ids, per_cnt)
        oloss = og.step_adagrad(Eo, bo, aE, ab, ids[s, 0], ids[s, 1], counts[s], 0.05, bias_mode)
        np.testing.assert_allclose(float(loss.item()), oloss, rtol=2e-5, atol=1e-5)
        assert all(int(k.err.item()) == 0 for k in vp.ranks)
    Eg, bg = vp.gather_dense()
    np.testing.assert_allclose(Eg.cpu().numpy(), Eo, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(bg.cpu().numpy(), bo, rtol=1e-5, atol=1e-5)
    for r, k in enumerate(vp.ranks):                                # Adagrad slots too (owner-side merge order is fixed)
        np.testing.assert_allclose(k.shard.acc.cpu().numpy(), aE[r::n], rtol=1e-5, atol=1e-7)
        assert bool((k.slot_map == -1).all()), "slot_map not restored"


# ---- owner-computes pair routing (OwnerRoutedGloveTrainer; tests/virtual_peers.VirtualOwnerRoutedGlove) ------------------
def test_padded_plan_with_device_slot_count_equals_plain_plan():
    """EsrPlan.n_valid: a step over a fixed-capacity slot array whose padding sorts to the end must equal the step over
    exactly the real pairs -- rows and accumulators bit for bit (same chunking => same summation tree)."""
    from esrecsys_b200 import engine as eng, synth
    V, D, m, cap = 20000, 128, 3000, 4096
    E, _ = synth.init_glove_tables(V, D, 2)
    b = (np.random.default_rng(9).standard_normal(V) * 0.05).astype(np.float32)
    ids, counts = synth.glove_batches(V, m, 2, 11)
    outs = []
    for padded in (False, True):
        t = eng.EmbeddingTable.from_dense(E, b, sparse=True)
        if padded:
            nv = torch.zeros(1, dtype=torch.int32, device="cuda")
            plan = eng.IndexPlan(2 * cap, V + 1, n_valid=nv)
            step = eng.GloveStep(t, cap, chunk=32, B_global=m)
        else:
            plan = eng.IndexPlan(2 * m, V)
            step = eng.GloveStep(t, m, chunk=32)
        losses = []
        for k in range(2):
            if padded:
                keys = torch.full((2 * cap,), V, dtype=torch.int32)
                keys[:m] = torch.from_numpy(ids[k, 0])
                keys[cap:cap + m] = torch.from_numpy(ids[k, 1])
                cnt = torch.zeros(cap)
                cnt[:m] = torch.from_numpy(counts[k])
                nv.fill_(2 * m)
                plan.build(keys.cuda())
                losses.append(float(step.run(plan, cnt.cuda())[eng.L.SC_LOSS].item()))
                assert int(plan.n_uniq.item()) == np.unique(ids[k]).size
            else:
                plan.build(torch.from_numpy(ids[k].reshape(-1)).cuda())
                losses.append(float(step.run(plan, torch.from_numpy(counts[k]).cuda())[eng.L.SC_LOSS].item()))
        outs.append((t.dense().cpu().numpy(), t.acc.cpu().numpy(), t.bias.cpu().numpy(), np.array(losses)))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(outs[0][3], outs[1][3], rtol=1e-6)


@pytest.mark.parametrize("n", [2, 3, 8])
def test_pair_routing_kernels_bit_exact(n):
    from oracle import index as oidx
    from virtual_peers import VirtualOwnerRoutedGlove
    V, D, B = 3001, 64, 700                                      # B is not a multiple of the routing tile
    vp = VirtualOwnerRoutedGlove(V, D, B, n)
    ids, counts = _batches(V, B, n, 1, 3)
    per_ids = [np.ascontiguousarray(ids[0][:, r * B:(r + 1) * B]) for r in range(n)]
    per_cnt = [np.ascontiguousarray(counts[0][r * B:(r + 1) * B]) for r in range(n)]
    vp.route_phase([torch.from_numpy(x) for x in per_ids], [torch.from_numpy(x) for x in per_cnt])
    torch.cuda.synchronize()
    routed = [oidx.route_pairs(per_ids[r], per_cnt[r], n) for r in range(n)]
    for o, k in enumerate(vp.ranks):
        regions = [routed[s][0][o] for s in range(n)]
        assert k.pin_counts[:n].cpu().tolist() == [int(routed[s][1][o]) for s in range(n)]
        rec = k.pin_rec.cpu().numpy()                               # [n][B][4] = {i, j, count bits, 0}
        for s in range(n):
            c = regions[s][0].size
            assert np.array_equal(rec[s, :c, 0], regions[s][0]) and np.array_equal(rec[s, :c, 1], regions[s][1])
            assert np.array_equal(rec[s, :c, 2].view(np.uint32), regions[s][2].view(np.uint32))
        keys, cnt, nv, over = oidx.collect_pairs(regions, vp.B_cap, V)
        assert not over and int(k.n_valid.item()) == nv and int(k.err.item()) == 0
        assert np.array_equal(k.keys.cpu().numpy(), keys)
        assert np.array_equal(k.cnt_l.cpu().numpy().view(np.uint32), cnt.view(np.uint32))
        assert k.my_counts[:n].cpu().tolist() == routed[o][1].tolist()


def test_pair_routing_overflow_is_flagged():
    from virtual_peers import VirtualOwnerRoutedGlove
    V, D, B, n = 4000, 64, 512, 2
    vp = VirtualOwnerRoutedGlove(V, D, B, n, pair_cap=600)
    ids = [torch.stack([torch.arange(2, 2 + 2 * B, 2, dtype=torch.int32), torch.ones(B, dtype=torch.int32)]) for _ in range(n)]
    vp.route_phase(ids, [torch.ones(B) for _ in range(n)])        # every i is even: all 1024 pairs go to rank 0
    torch.cuda.synchronize()
    assert int(vp.ranks[0].err.item()) & 2 and int(vp.ranks[0].n_valid.item()) == 1200
    assert int(vp.ranks[1].err.item()) == 0 and int(vp.ranks[1].n_valid.item()) == 0


@pytest.mark.parametrize("bias_mode", ["reference_broadcast", "per_pair"])
@pytest.mark.parametrize("n,V,D,B_loc", [(2, 5000, 64, 1024), (3, 300, 128, 512), (4, 5000, 128, 512), (8, 2000, 64, 256),
                                         (8, 50000, 128, 4096)])
def test_virtual_owner_routed_step_matches_single_table_oracle(n, V, D, B_loc, bias_mode):
    """One global step of the owner-routed sharded trainer == the single-table step on the concatenated batch
    (wikipedia/train_cooccurence.py:71-101 with the north-star Adagrad rule), for 3 consecutive steps."""
    from esrecsys_b200 import synth
    from oracle import glove as og
    from oracle import optim as oopt
    from virtual_peers import VirtualOwnerRoutedGlove
    steps = 3
    E, _ = synth.init_glove_tables(V, D, 0)
    b = (np.random.default_rng(5).standard_normal(V) * 0.05).astype(np.float32)
    ids, counts = _batches(V, B_loc, n, steps, 1)
    vp = VirtualOwnerRoutedGlove(V, D, B_loc, n, lr=0.05, bias_mode=bias_mode)
    vp.load_dense(E, b)
    Eo, bo = E.copy(), b.copy()
    aE, ab = np.full_like(Eo, oopt.ADAGRAD_INIT_ACC), np.full_like(bo, oopt.ADAGRAD_INIT_ACC)
    for s in range(steps):
        per_ids = [torch.from_numpy(np.ascontiguousarray(ids[s][:, r * B_loc:(r + 1) * B_loc])) for r in range(n)]
        per_cnt = [torch.from_numpy(np.ascontiguousarray(counts[s][r * B_loc:(r + 1) * B_loc])) for r in range(n)]
        loss = vp.step(per_ids, per_cnt)
        oloss = og.step_adagrad(Eo, bo, aE, ab, ids[s, 0], ids[s, 1], counts[s], 0.05, bias_mode)
        np.testing.assert_allclose(float(loss.item()), oloss, rtol=2e-5, atol=1e-5)
        assert all(int(k.err.item()) == 0 for k in vp.ranks)
        assert sum(int(k.n_valid.item()) for k in vp.ranks) == 2 * n * B_loc           # every pair processed exactly once
    Eg, bg = vp.gather_dense()
    np.testing.assert_allclose(Eg.cpu().numpy(), Eo, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(bg.cpu().numpy(), bo, rtol=1e-5, atol=1e-5)
    for r, k in enumerate(vp.ranks):
        np.testing.assert_allclose(k.shard.acc.cpu().numpy(), aE[r::n], rtol=1e-5, atol=1e-7)
        assert bool((k.slot_map == -1).all()), "slot_map not restored"
