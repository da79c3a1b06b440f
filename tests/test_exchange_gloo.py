"""World-size-2 gloo test (CPU) of the row-sharded exchange choreography in esrecsys_b200/sharded.py.

The per-rank compute is supplied by NumPy ops that follow oracle/index.py, so what is pinned here is
the HOST logic: owner bucketing order, count/id/row all-to-alls, split sizes, reassembly in unique
order, and the gradient return + cross-rank duplicate merge."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Shard:
    def __init__(self, rows, bias):
        self.rows0, self.bias = rows, bias
        self.V, self.D = rows.shape
        self.acc = torch.full_like(rows, 0.1)
        self.bias_acc = torch.full_like(bias, 0.1)


class NumpyOps:
    """Same interface as sharded.LibesrOps, on CPU tensors, via oracle.index / oracle.optim."""

    def route_plan(self, uniq, n_uniq, n_ranks):
        from oracle import index as oidx
        U = int(n_uniq.item())
        counts, _, send_local, order = oidx.route_plan(uniq[:U].numpy(), n_ranks)
        cap = uniq.numel()
        o = torch.zeros(cap, dtype=torch.int32)
        s = torch.zeros(cap, dtype=torch.int32)
        o[:U] = torch.from_numpy(order)
        s[:U] = torch.from_numpy(send_local)
        return o, s, torch.from_numpy(counts.astype(np.int32))

    def gather_rows(self, shard, ids):
        return shard.rows0[ids.long()]

    def gather_scalar(self, src, ids):
        return src[ids.long()]

    def permute_rows(self, src, idx, n, scatter, out):
        i = idx[:n].long()
        if scatter:
            out[i] = src[:n]
        else:
            out[:n] = src[i]
        return out

    def owner_update(self, shard, recv_ids, recv_g, recv_gb, lr, eps):
        from oracle import index as oidx
        from oracle import optim as oopt
        ids = recv_ids.numpy()
        sk, perm = oidx.sort_slots(ids)
        uniq, off = oidx.segments(sk)
        g = np.add.reduceat(recv_g.numpy()[perm], off[:-1], axis=0) if len(uniq) else np.zeros((0, shard.D), np.float32)
        gb = np.add.reduceat(recv_gb.numpy()[perm], off[:-1]) if len(uniq) else np.zeros(0, np.float32)
        E, a = oopt.adagrad_update(shard.rows0.numpy()[uniq], g.astype(np.float32), shard.acc.numpy()[uniq], lr, eps)
        shard.rows0[uniq] = torch.from_numpy(E)
        shard.acc[uniq] = torch.from_numpy(a)
        bb, ab = oopt.adagrad_update(shard.bias.numpy()[uniq], gb.astype(np.float32), shard.bias_acc.numpy()[uniq], lr, eps)
        shard.bias[uniq] = torch.from_numpy(bb)
        shard.bias_acc[uniq] = torch.from_numpy(ab)


def _worker(rank, world, port, V, D, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from esrecsys_b200.sharded import RowExchange, shard_rows
        from oracle import optim as oopt
        rng = np.random.default_rng(0)                       # same table on every rank
        E = rng.standard_normal((V, D)).astype(np.float32)
        b = rng.standard_normal(V).astype(np.float32)
        mine = np.arange(rank, V, world)
        assert len(mine) == shard_rows(V, rank, world)
        shard = _Shard(torch.from_numpy(E[mine].copy()), torch.from_numpy(b[mine].copy()))
        xchg = RowExchange(NumpyOps(), shard)
        # each rank asks for its own sorted unique rows (overlapping between ranks; rank 1 asks for few)
        r2 = np.random.default_rng(100 + rank)
        uniq = np.unique(r2.integers(0, V, size=40 if rank == 0 else 7)).astype(np.int32)
        U = len(uniq)
        cap = 64
        u_t = torch.zeros(cap, dtype=torch.int32)
        u_t[:U] = torch.from_numpy(uniq)
        rows, bias = xchg.fetch(u_t, torch.tensor([U], dtype=torch.int32))
        assert np.array_equal(rows[:U].numpy(), E[uniq]), "fetched rows are not table[uniq]"
        assert np.array_equal(bias[:U].numpy(), b[uniq])
        # gradients: g = row id + 1000*rank in every column (so the merge is checkable exactly)
        dE = torch.zeros(cap, D)
        db = torch.zeros(cap)
        dE[:U] = torch.from_numpy((uniq[:, None] * 0.001 + rank).astype(np.float32)).expand(U, D)
        db[:U] = torch.from_numpy((uniq * 0.01 + rank).astype(np.float32))
        xchg.push(dE, db, lr=0.05)
        # expected: every rank's request list is known to everyone (deterministic seeds)
        gsum = np.zeros((V, D), np.float32)
        gbsum = np.zeros(V, np.float32)
        for r in range(world):
            rr = np.random.default_rng(100 + r)
            uq = np.unique(rr.integers(0, V, size=40 if r == 0 else 7)).astype(np.int32)
            gsum[uq] += (uq[:, None] * 0.001 + r).astype(np.float32)
            gbsum[uq] += (uq * 0.01 + r).astype(np.float32)
        touched = np.flatnonzero(gbsum != 0) if False else np.unique(np.concatenate(
            [np.unique(np.random.default_rng(100 + r).integers(0, V, size=40 if r == 0 else 7)) for r in range(world)]))
        Eexp, bexp = E.copy(), b.copy()
        Eexp[touched], _ = oopt.adagrad_update(E[touched], gsum[touched], np.full((len(touched), D), 0.1, np.float32), 0.05)
        bexp[touched], _ = oopt.adagrad_update(b[touched], gbsum[touched], np.full(len(touched), 0.1, np.float32), 0.05)
        np.testing.assert_allclose(shard.rows0.numpy(), Eexp[mine], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(shard.bias.numpy(), bexp[mine], rtol=1e-6, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_row_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, 101, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


def _inbatch_worker(rank, world, port, kind, q):
    """The sharded in-batch decomposition (esrecsys_b200/inbatch.py, ShardedSharedTableInBatch) with gloo collectives:
    every rank scores its B_local queries against the all-gathered items (diag_off = rank * B_local, b_norm = global B);
    concatenated dQ, reduce-scattered dK and all-reduced loss must equal the single global step of the oracle."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import inbatch as oib
        B, D = 24, 16
        Bg = B * world
        rng = np.random.default_rng(3)
        Q = oib.bf16_round(rng.standard_normal((Bg, D)).astype(np.float32))
        K = oib.bf16_round(rng.standard_normal((Bg, D)).astype(np.float32))
        lo, hi = rank * B, (rank + 1) * B
        K_all = torch.empty(Bg, D)
        dist.all_gather_into_tensor(K_all, torch.from_numpy(K[lo:hi].copy()))
        assert np.array_equal(K_all.numpy(), K)
        fn = oib.hinge if kind == "hinge" else oib.softmax
        kw = dict(margin=0.5) if kind == "hinge" else {}
        loss, dQ, dK_part, _ = fn(Q[lo:hi], K_all.numpy(), off=rank * B, scale=0.25, b_norm=Bg, **kw)
        dK_loc = torch.empty(B, D)
        dist.reduce_scatter_tensor(dK_loc, torch.from_numpy(dK_part))
        lt = torch.tensor([float(loss)], dtype=torch.float64)
        dist.all_reduce(lt)
        gl, gdQ, gdK, _ = fn(Q, K, off=0, scale=0.25, b_norm=Bg, **kw)
        assert abs(lt.item() - gl) < 1e-5 * max(1.0, abs(gl))
        np.testing.assert_allclose(dQ, gdQ[lo:hi], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(dK_loc.numpy(), gdK[lo:hi], rtol=1e-5, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("kind", ["hinge", "softmax"])
def test_sharded_inbatch_decomposition_gloo(kind):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + (os.getpid() % 100) + (0 if kind == "hinge" else 1)
    procs = [ctx.Process(target=_inbatch_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


def _routed_worker(rank, world, port, bias_mode, q):
    """The owner-computes decomposition of the GloVe step (esrecsys_b200/sharded.py OwnerRoutedGloveTrainer; kernels
    esr_peer_route_pairs_i32 / esr_peer_collect_pairs_i32 / esr_peer_gather_remote_f32 / esr_peer_apply_parts_f32) with gloo
    collectives in place of the NVLink peer stores: pairs go to the owner of row i, the owner fetches the partner rows it
    does not own, the batch sums are all-reduced (B = global batch), gradients of foreign rows return to their owners and
    are merged there.  Every shard must then equal the oracle's single global step on the concatenated batch."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from esrecsys_b200.sharded import pair_capacity
        from oracle import glove as og
        from oracle import index as oidx
        from oracle import optim as oopt
        V, D, B, lr = 211, 8, 96, 0.05
        rng = np.random.default_rng(5)                                   # same table and batches on every rank
        E = (0.1 * rng.standard_normal((V, D))).astype(np.float32)
        b = (0.1 * rng.standard_normal(V)).astype(np.float32)
        ids = [np.minimum(rng.zipf(1.3, size=(2, B)) - 1, V - 1).astype(np.int32) for _ in range(world)]
        cnt = [rng.uniform(1, 200, size=B).astype(np.float32) for _ in range(world)]
        mine = np.arange(rank, V, world)
        Es, bs_ = E[mine].copy(), b[mine].copy()
        accE, accb = np.full_like(Es, 0.1), np.full_like(bs_, 0.1)

        # 1. route: stable partition by owner(i); "peer store" = exchange of the per-owner regions
        per_owner, send_counts = oidx.route_pairs(ids[rank], cnt[rank], world)
        inbox = [None] * world
        dist.all_gather_object(inbox, per_owner)
        regions = [inbox[s][rank] for s in range(world)]                  # source-major
        cap = pair_capacity(B, world)
        keys, x, n_valid, over = oidx.collect_pairs(regions, cap, V)
        assert not over
        m = n_valid // 2
        i, j, x = keys[:m], keys[cap:cap + m], x[:m]
        assert np.all(i % world == rank)
        exp = oidx.routed_batches(ids, cnt, world)[rank]
        assert np.array_equal(i, exp[0]) and np.array_equal(j, exp[1]) and np.array_equal(x, exp[2])

        # 2. fetch the partner rows this rank does not own
        need = np.unique(j[j % world != rank])
        asks = [None] * world
        dist.all_gather_object(asks, [need[need % world == o] for o in range(world)])
        reply = [(Es[asks[s][rank] // world], bs_[asks[s][rank] // world]) for s in range(world)]
        got = [None] * world
        dist.all_gather_object(got, reply)
        Eloc = np.zeros((V, D), np.float32)
        bloc = np.zeros(V, np.float32)
        Eloc[mine], bloc[mine] = Es, bs_
        for o in range(world):
            if o != rank:
                want = need[need % world == o]
                Eloc[want], bloc[want] = got[o][rank]

        # 3. batch sums over ALL ranks (the three-float all-reduce of the step), then the closed-form gradients
        dot = np.einsum("cd,cd->c", Eloc[i], Eloc[j])
        bsum = bloc[i] + bloc[j]
        w, t = og.weight_fn(x), og.log_target_fn(x)
        res = t - dot
        part = torch.tensor([w.sum(), (w * res).sum(), (w * res * res).sum(), bsum.sum(), (bsum * bsum).sum(),
                             (w * (res - bsum) ** 2).sum()], dtype=torch.float64)
        dist.all_reduce(part)
        S0, S1, S2, sb, sb2, pp = part.tolist()
        Bg = float(B * world)
        if bias_mode == "reference_broadcast":
            loss = (S2 - 2 * (sb / Bg) * S1 + (sb2 / Bg) * S0) / Bg
            g = -(2 / Bg) * w * (res - sb / Bg)
            h = -(2 / (Bg * Bg)) * (S1 - bsum * S0)
        else:
            loss = pp / Bg
            g = -(2 / Bg) * w * (res - bsum)
            h = g
        dE = np.zeros((V, D), np.float64)
        db = np.zeros(V, np.float64)
        np.add.at(dE, i, g[:, None] * Eloc[j])
        np.add.at(dE, j, g[:, None] * Eloc[i])
        np.add.at(db, i, h)
        np.add.at(db, j, h)

        # 4. gradients of foreign rows return to the owner; the owner merges all sources and applies Adagrad
        touched = np.unique(np.concatenate([i, j]))
        outs = [None] * world
        dist.all_gather_object(outs, [(touched[touched % world == o], dE[touched[touched % world == o]],
                                       db[touched[touched % world == o]]) for o in range(world)])
        gE = np.zeros((V, D), np.float64)
        gb = np.zeros(V, np.float64)
        for s in range(world):
            rows, a, c = outs[s][rank]
            gE[rows] += a
            gb[rows] += c
            assert np.all(rows % world == rank)
        upd = np.unique(np.concatenate([outs[s][rank][0] for s in range(world)]))
        loc = upd // world
        Es[loc], accE[loc] = oopt.adagrad_update(Es[loc], gE[upd].astype(np.float32), accE[loc], lr)
        bs_[loc], accb[loc] = oopt.adagrad_update(bs_[loc], gb[upd].astype(np.float32), accb[loc], lr)

        # oracle: ONE global step on the concatenated batch
        Eg, bg = E.copy(), b.copy()
        aE, ab = np.full_like(Eg, 0.1), np.full_like(bg, 0.1)
        gi = np.concatenate([a[0] for a in ids])
        gj = np.concatenate([a[1] for a in ids])
        gx = np.concatenate(cnt)
        gl = og.step_adagrad(Eg, bg, aE, ab, gi, gj, gx, lr, bias_mode)
        assert abs(loss - gl) < 2e-5 * max(1.0, abs(gl)), (loss, gl)
        np.testing.assert_allclose(Es, Eg[mine], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(bs_, bg[mine], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(accE, aE[mine], rtol=2e-5, atol=2e-7)
        untouched = np.setdiff1d(mine, np.unique(np.concatenate([gi, gj])))
        assert np.array_equal(Es[untouched // world], E[untouched])       # rows nobody touched are bit-unchanged
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,bias_mode", [(2, "reference_broadcast"), (2, "per_pair"), (3, "reference_broadcast")])
def test_owner_routed_glove_decomposition_gloo(world, bias_mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 100) + world + (7 if bias_mode == "per_pair" else 0)
    procs = [ctx.Process(target=_routed_worker, args=(r, world, port, bias_mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res


def _topk_worker(rank, world, port, q):
    """Retrieval over a row-sharded table (esrecsys_b200/sharded.py sharded_table_topk): local top-k of every shard (here
    NumPy on the shard, on the GPU the fused scan kernel) -> all-gather -> merge; equals the oracle's top-k over the
    whole table in BOTH tie conventions, planted exact ties across shards included."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from esrecsys_b200.sharded import sharded_query_rows, sharded_table_topk
        from oracle import glove as og
        V, D, k = 103, 8, 10
        rng = np.random.default_rng(4)
        E = rng.integers(-3, 4, size=(V, D)).astype(np.float32)          # small integers: many exactly equal scores
        E[50] = E[7]
        E[51] = E[7]                                                      # rows 7, 50, 51 tie for every query
        tokens = np.array([7, 13, 99], np.int32)

        class Shard:
            pass
        Shard.V = len(range(rank, V, world))
        Shard.rows = torch.from_numpy(E[rank::world].copy())

        def local_topk(shard, queries, kl, high_first):
            sc = shard.rows.numpy() @ queries.numpy().T                   # (V_loc, T)
            idx = np.arange(shard.V)
            out_i, out_v = [], []
            for t in range(sc.shape[1]):
                order = np.lexsort((-idx if high_first else idx, -sc[:, t]))[:kl]
                out_i.append(order)
                out_v.append(sc[order, t])
            return torch.from_numpy(np.array(out_v, np.float32)), torch.from_numpy(np.array(out_i, np.int32))

        queries = sharded_query_rows(Shard.rows, tokens, rank, world)
        assert np.array_equal(queries.numpy(), E[tokens])
        for high_first in (True, False):
            val, rows = sharded_table_topk(Shard, queries, k, rank, world, ties_high_index_first=high_first,
                                           local_topk=local_topk)
            sc = E @ E[tokens].T
            for t in range(len(tokens)):
                want = np.lexsort((-np.arange(V) if high_first else np.arange(V), -sc[:, t]))[:k]
                assert rows[t].tolist() == want.tolist(), (high_first, t, rows[t].tolist(), want.tolist())
                assert np.array_equal(val[t].numpy(), sc[want, t])
            if high_first:                                                # the convention dump_knn reads (oracle.glove.top_k)
                top, _ = og.top_k(E, tokens, k)
                assert np.array_equal(rows.numpy(), top)
        # k larger than a shard: padding never wins
        val, rows = sharded_table_topk(Shard, queries, 60, rank, world, local_topk=local_topk)
        assert rows.min() >= 0 and torch.isfinite(val).all() and all(len(set(r.tolist())) == 60 for r in rows)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_table_topk_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 100) + world
    procs = [ctx.Process(target=_topk_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res
