"""GPU parity of the row-sharded GloVe path: n ranks (NCCL, one process per GPU) must reproduce the
single-table oracle on the concatenated global batch -- losses and the re-assembled table within 1e-5.
Runs with as many ranks as there are GPUs (capped at 4); on a 1-GPU box it still runs the whole sharded
code path with world_size 1."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, V, D, B_loc, steps, bias_mode, q, peer=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from esrecsys_b200 import synth
        from esrecsys_b200.sharded import OwnerRoutedGloveTrainer, PeerShardedGloveTrainer, ShardedGloveTrainer
        from oracle import glove as og
        from oracle import optim as oopt
        E, b = synth.init_glove_tables(V, D, 0)
        b = (np.random.default_rng(5).standard_normal(V) * 0.05).astype(np.float32)
        ids, counts = synth.glove_batches(V, B_loc * world, steps, 1)     # global batches
        # experimental switches of the peer path: libesr peer all-reduce / barrier kernels; id phases on a side stream
        kw = {"fast": {"fast_sync": True}, "overlap": {"overlap_ids": True},
              "fast_overlap": {"fast_sync": True, "overlap_ids": True}}.get(peer, {})
        if peer == "routed":    # owner-computes pair routing (bench default for N > 1); CUDA graphs from step 4 on
            tr = OwnerRoutedGloveTrainer(V, D, B_loc, lr=0.05, bias_mode=bias_mode)
        else:
            tr = (PeerShardedGloveTrainer if peer else ShardedGloveTrainer)(V, D, B_loc, lr=0.05, bias_mode=bias_mode, **kw)
        tr.load_dense(E, b)
        Eo, bo = E.copy(), b.copy()
        aE, ab = np.full_like(Eo, oopt.ADAGRAD_INIT_ACC), np.full_like(bo, oopt.ADAGRAD_INIT_ACC)
        lo, hi = rank * B_loc, (rank + 1) * B_loc
        for k in range(steps):
            loss = tr.step(torch.from_numpy(np.ascontiguousarray(ids[k][:, lo:hi])), torch.from_numpy(counts[k][lo:hi]))
            oloss = og.step_adagrad(Eo, bo, aE, ab, ids[k, 0], ids[k, 1], counts[k], 0.05, bias_mode)
            got = tr.loss_value() if hasattr(tr, "loss_value") else float(loss.item())     # routed trainer: own streams
            np.testing.assert_allclose(got, oloss, rtol=2e-5, atol=1e-5)
        Eg, bg = tr.gather_dense()
        np.testing.assert_allclose(Eg.cpu().numpy(), Eo, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(bg.cpu().numpy(), bo, rtol=1e-5, atol=1e-5)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("peer", [False, True, "fast_overlap", "routed"],
                         ids=["nccl_a2a", "peer_memory", "peer_fast_sync_overlap_ids", "owner_routed_graphs"])
@pytest.mark.parametrize("bias_mode", ["reference_broadcast", "per_pair"])
@pytest.mark.parametrize("V,D,B_loc", [(5000, 64, 1024), (300, 128, 512)])
def test_sharded_matches_single_table_oracle(V, D, B_loc, bias_mode, peer):
    world = max(1, min(4, torch.cuda.device_count()))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200) + (V % 7) + (50 if peer else 0)
    steps = 7 if peer == "routed" else 3        # the routed trainer captures its CUDA graphs before step 4
    procs = [ctx.Process(target=_worker, args=(r, world, port, V, D, B_loc, steps, bias_mode, q, peer)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res


def test_route_plan_bit_exact():
    from esrecsys_b200.sharded import LibesrOps
    from oracle import index as oidx
    rng = np.random.default_rng(0)
    for n_ranks in (1, 2, 3, 8):
        uniq = np.unique(rng.integers(0, 100000, size=5000)).astype(np.int32)
        U = len(uniq)
        cap = 8192
        u = torch.zeros(cap, dtype=torch.int32, device="cuda")
        u[:U] = torch.from_numpy(uniq).cuda()
        ops = LibesrOps(torch.device("cuda"))
        order, send_local, counts = ops.route_plan(u, torch.tensor([U], dtype=torch.int32, device="cuda"), n_ranks)
        oc, _, osl, oo = oidx.route_plan(uniq, n_ranks)
        assert np.array_equal(counts.cpu().numpy(), oc)
        assert np.array_equal(order[:U].cpu().numpy(), oo)
        assert np.array_equal(send_local[:U].cpu().numpy(), osl)


def _inbatch_worker(rank, world, port, V, D, B_loc, steps, kind, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from esrecsys_b200 import synth
        from esrecsys_b200.inbatch import ShardedSharedTableInBatch
        from oracle import inbatch as oib
        rng = np.random.default_rng(11)
        E = (rng.standard_normal((V, D)) / D ** 0.25).astype(np.float32)
        qs, ks = synth.pair_batches(V, V, B_loc * world, steps, 5)      # global batches, rank-major slices
        tr = ShardedSharedTableInBatch(V, D, B_loc, lr=0.05, loss=kind)
        tr.load_dense(E)
        Eo, acc = E.copy(), np.full_like(E, 0.1)
        lo, hi = rank * B_loc, (rank + 1) * B_loc
        for s in range(steps):
            ids = torch.from_numpy(np.stack([qs[s][lo:hi], ks[s][lo:hi]])).cuda()
            got = float(tr.step(ids).item())
            want = oib.shared_table_step(Eo, acc, qs[s], ks[s], 0.05, kind)
            assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (s, got, want)
        Eg = tr.gather_dense().cpu().numpy()
        bad = np.abs(Eg - Eo) > 1e-5 + 1e-5 * np.abs(Eo)
        assert bad.mean() < 2e-3 and np.abs(Eg - Eo).max() < 1e-3, (bad.mean(), np.abs(Eg - Eo).max())
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["hinge", "softmax"])
def test_sharded_inbatch_matches_single_table_oracle(kind):
    """configs[2] sharded: n ranks with B_local pairs each == one in-batch step over the concatenated batch."""
    world = max(1, min(4, torch.cuda.device_count()))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40) + (0 if kind == "hinge" else 1)
    procs = [ctx.Process(target=_inbatch_worker, args=(r, world, port, 4000, 128, 256, 3, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res


def _two_tower_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from esrecsys_b200 import synth
        from esrecsys_b200.inbatch import ShardedTwoTowerInBatch
        from oracle import inbatch as oib
        V, D, B_loc, steps = 3000, 64, 128, 3
        rng = np.random.default_rng(12)
        Es = (rng.standard_normal((V, D)) / D ** 0.5).astype(np.float32)
        Ep = (rng.standard_normal((V, D)) / D ** 0.5).astype(np.float32)
        tr = ShardedTwoTowerInBatch(V, V, D, B_loc, lr=0.05, tower_lr=1e-3, loss="softmax", scale=4.0, seed=3)
        tr.load_dense(Es, Ep)
        ps = {k: v.cpu().numpy().copy() for k, v in tr.scene_tower.p.items()}
        pp = {k: v.cpu().numpy().copy() for k, v in tr.product_tower.p.items()}
        mk = lambda p: dict(count=0, mu={k: np.zeros_like(v) for k, v in p.items()}, nu={k: np.zeros_like(v) for k, v in p.items()})
        os_, op_ = mk(ps), mk(pp)
        Eso, Epo = Es.copy(), Ep.copy()
        accs, accp = np.full_like(Es, 0.1), np.full_like(Ep, 0.1)
        s_ids, p_ids = synth.pair_batches(V, V, B_loc * world, steps, 9)
        lo, hi = rank * B_loc, (rank + 1) * B_loc
        for s in range(steps):
            got = float(tr.step(torch.from_numpy(s_ids[s][lo:hi]).cuda(), torch.from_numpy(p_ids[s][lo:hi]).cuda()).item())
            want = oib.two_tower_step(Eso, accs, Epo, accp, ps, pp, os_, op_, s_ids[s], p_ids[s], 0.05, 1e-3, "softmax", 1.0, 4.0)
            assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (s, got, want)
        Egs, Egp = tr.gather_dense()
        np.testing.assert_allclose(Egs.cpu().numpy(), Eso, rtol=1e-3, atol=2e-4)
        np.testing.assert_allclose(Egp.cpu().numpy(), Epo, rtol=1e-3, atol=2e-4)
        # optax.adam moves a parameter by ~lr * sign(g) in the first steps whatever |g| is, so an element whose gradient
        # cancels to ~0 (bias sums) can differ by up to 2 * lr per step between two fp32 summation orders (here: the
        # all-reduce of per-rank partial sums vs one global sum).  Weights: tight; biases: bounded by that envelope and
        # mostly equal.
        for k in ps:
            for got, want in ((tr.scene_tower.p[k].cpu().numpy(), ps[k]), (tr.product_tower.p[k].cpu().numpy(), pp[k])):
                if k.startswith("W"):
                    bad = np.abs(got - want) > 2e-4 + 1e-3 * np.abs(want)
                    assert bad.mean() < 5e-3, (k, bad.mean())
                assert np.abs(got - want).max() <= 2.2 * steps * 1e-3, (k, np.abs(got - want).max())
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_two_tower_matches_oracle():
    """configs[3] sharded: row-sharded id tables, replicated MLP towers (all-reduced grads), global in-batch softmax."""
    world = max(1, min(4, torch.cuda.device_count()))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29990 + (os.getpid() % 9)
    procs = [ctx.Process(target=_two_tower_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res
