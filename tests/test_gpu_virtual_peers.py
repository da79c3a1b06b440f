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
        own_list = k.desc[vp.inbox_cap * n: vp.inbox_cap * n + n_own].cpu().numpy()
        assert np.array_equal(np.sort(own_list), np.flatnonzero(own))
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
