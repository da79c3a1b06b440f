"""In-batch-negative trainers of BASELINE.json configs[2] and configs[3] on libesr.

The reference scores explicit negatives (spotify/train_spotify.py:77-106: 64 sampled tracks per
playlist; pinterest/train_shop_the_look.py:93-109: one pre-sampled negative product per pair); the
north star replaces them by the other items of the batch:

* ``SharedTableInBatch``  -- configs[2]: ONE id-embedding table (2M x 128), a batch of (query id, item id)
  pairs, step = gather rows -> B x B scores + loss + dQ/dK on the tensor cores
  (``esr_inbatch_fwd_bwd_bf16``) -> per-row segment sum of the 2B gradient rows -> sparse Adagrad.
* ``TwoTowerInBatch``     -- configs[3]: scene-id and product-id tables (x 256) each followed by a 2-layer
  MLP tower ``Linear -> ReLU -> Linear`` (the north-star substitute for the CNN towers of
  pinterest/models.py:23-46), same scorer; tables take sparse Adagrad, tower weights ``optax.adam``
  (pinterest/train_shop_the_look.py:175).  The tower GEMMs are plain library GEMMs (cuBLAS through torch).

Contract: oracle/inbatch.py (``shared_table_step`` / ``two_tower_step``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import engine


class _RowUpdater:
    """Gradient rows (n, D) keyed by row id -> deterministic per-row sum -> sparse Adagrad in place."""

    def __init__(self, table: engine.EmbeddingTable, n_slots: int, lr: float, eps: float = 1e-7):
        self.table, self.lr, self.eps = table, float(lr), float(eps)
        self.plan = engine.IndexPlan(n_slots, table.V, table.device, with_partner=False, sort="wide")
        self.gsum = torch.empty(n_slots, table.D, dtype=torch.float32, device=table.device)

    def apply(self, ids_i32, grads, stream=None):
        self.plan.build(ids_i32, stream)
        L.check(L.lib().esr_segment_sum_rows_f32(C.byref(self.plan.s), self.table.D, L.ptr(grads), None, L.ptr(self.gsum), None,
                                                 L.stream_ptr(stream)), "esr_segment_sum_rows_f32")
        engine.sparse_adagrad(self.table, self.plan.uniq, self.plan.n_uniq, self.gsum, None, self.lr, self.eps, stream)


class _GraphedStep:
    """One trainer step captured into a CUDA graph (static shapes, static buffers): the eager step is 15-40 small launches
    -- libesr kernels, cub sort passes and, for the two-tower model, torch GEMM / elementwise ops -- and at B = 4096-8192
    the host, not the GPU, sets its pace (measured: 0.84 ms eager vs the 0.08 ms the scorer needs at configs[3]).  Inputs are
    copied into static tensors, the graph is replayed, the loss is a static device scalar."""

    def __init__(self, fn, example_inputs, warmup=3):
        self.static_in = [x.clone() for x in example_inputs]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):                    # lazy initialisation (cuBLAS handles, function attributes) outside capture
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class SharedTableInBatch:
    """configs[2]: Spotify-style skip-gram over one table with in-batch negatives."""

    def __init__(self, table: engine.EmbeddingTable, B: int, lr: float = 0.05, loss: str = "hinge", margin: float = 1.0,
                 scale: float = 1.0):
        assert not table.sparse and table.acc is not None, "single-buffer table with an Adagrad slot (sparse=False, adagrad=True)"
        self.table, self.B = table, int(B)
        dev = table.device
        self.scorer = engine.InBatchScorer(B, table.D, loss=loss, margin=margin, scale=scale, device=dev)
        self.X = torch.empty(2 * self.B, table.D, dtype=torch.float32, device=dev)      # [Q ; K] gathered rows
        self.dX = torch.empty(2 * self.B, table.D, dtype=torch.float32, device=dev)     # [dQ ; dK]
        self.scorer.dQ, self.scorer.dK = self.dX[:self.B], self.dX[self.B:]
        self.upd = _RowUpdater(table, 2 * self.B, lr)

    def step(self, ids, stream=None):
        """ids: int32 CUDA (2, B) = [query ids ; item ids].  Returns the device loss scalar."""
        flat = ids.reshape(-1)
        assert flat.dtype == torch.int32 and flat.is_cuda and flat.numel() == 2 * self.B
        L.check(L.lib().esr_table_gather_f32(C.byref(self.table.struct()), L.ptr(flat), 2 * self.B, L.ptr(self.X),
                                             L.stream_ptr(stream)), "esr_table_gather_f32")
        loss, _, _ = self.scorer.run(self.X[:self.B], self.X[self.B:], stream)
        self.upd.apply(flat, self.dX, stream)
        return loss

    def graphed(self, example_ids):
        """``step`` as a CUDA graph: ``g = tr.graphed(ids); loss = g(ids)``.  The three warm-up calls and the capture itself
        do not train (the table, its accumulator and the trainer state are restored)."""
        snap = (self.table.rows0.clone(), self.table.acc.clone())
        g = _GraphedStep(lambda ids: self.step(ids), [example_ids])
        self.table.rows0.copy_(snap[0])
        self.table.acc.copy_(snap[1])
        return g


class _MatmulPrecision:
    """fp32 (IEEE, cuBLAS SIMT kernels: exact parity with the NumPy oracle) or tf32 (tensor cores, 10-bit mantissa inputs --
    what XLA's DEFAULT precision does with a float32 Dense on an NVIDIA GPU) for the GEMMs issued inside the block."""

    def __init__(self, mode):
        assert mode in ("fp32", "tf32"), mode
        self.tf32 = mode == "tf32"

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


class MLPTower:
    """``Linear(D, H) -> ReLU -> Linear(H, O)`` with explicit backward; weights updated by ``optax.adam``
    semantics through ``esr_dense_adam_f32``.  The six GEMMs of a step are library calls (cuBLAS through torch): measured
    share of the configs[3] step in profiles/r2_summary.md section 4."""

    def __init__(self, D, H, O, gen, device, lr=1e-3, matmul="fp32"):
        self.precision = _MatmulPrecision(matmul)
        mk = lambda i, o: (torch.randn(i, o, generator=gen) / np.sqrt(i)).to(device)
        init = {"W1": mk(D, H), "b1": torch.zeros(H, device=device), "W2": mk(H, O), "b2": torch.zeros(O, device=device)}
        # parameters, gradients and the two Adam moments are views of four flat buffers (16-byte aligned pieces), so the
        # optimiser is ONE esr_dense_adam_f32 launch per tower instead of one per tensor
        off, total = {}, 0
        for k, v in init.items():
            off[k] = total
            total += (v.numel() + 3) // 4 * 4
        flat = lambda: torch.zeros(total, dtype=torch.float32, device=device)
        self.flat = {"p": flat(), "g": flat(), "mu": flat(), "nu": flat()}
        view = lambda buf, k: buf[off[k]:off[k] + init[k].numel()].view(init[k].shape)
        self.p = {k: view(self.flat["p"], k) for k in init}
        self.g = {k: view(self.flat["g"], k) for k in init}
        self.mu = {k: view(self.flat["mu"], k) for k in init}
        self.nu = {k: view(self.flat["nu"], k) for k in init}
        for k, v in init.items():
            self.p[k].copy_(v)
        self.ones = None
        self.count, self.lr = 0, float(lr)

    def forward(self, x):
        self.x = x
        with self.precision:
            self.h = torch.relu(torch.addmm(self.p["b1"], x, self.p["W1"]))
            return torch.addmm(self.p["b2"], self.h, self.p["W2"])

    def backward(self, dy):
        g = self.g
        with self.precision:
            if self.ones is None or self.ones.shape[1] != dy.shape[0]:
                self.ones = torch.ones(1, dy.shape[0], dtype=torch.float32, device=dy.device)
            torch.mm(self.h.t(), dy, out=g["W2"])
            torch.mm(self.ones, dy, out=g["b2"].view(1, -1))          # column sums as a GEMV (torch.sum: 16 us each)
            dh = torch.mm(dy, self.p["W2"].t())
            dh.mul_(self.h > 0)
            torch.mm(self.x.t(), dh, out=g["W1"])
            torch.mm(self.ones, dh, out=g["b1"].view(1, -1))
            return torch.mm(dh, self.p["W1"].t())

    def update(self):
        self.count += 1
        f = self.flat
        engine.dense_adam(f["p"], f["g"], f["mu"], f["nu"], self.lr, self.count)


class TwoTowerInBatch:
    """configs[3]: scene / product id tables + MLP towers, B x B in-batch negatives."""

    def __init__(self, scene_table, product_table, B, hidden=None, out=None, lr=0.05, tower_lr=1e-3, loss="softmax",
                 margin=1.0, scale=1.0, seed=0, tower_matmul="fp32"):
        """``tower_matmul``: "fp32" (default; parity with the oracle at 1e-5) or "tf32" (tensor-core GEMMs in the towers)."""
        for t in (scene_table, product_table):
            assert not t.sparse and t.acc is not None
        self.ts, self.tp, self.B = scene_table, product_table, int(B)
        dev = scene_table.device
        D = scene_table.D
        H = int(hidden or D)
        O = int(out or D)
        gen = torch.Generator(device="cpu").manual_seed(seed)
        self.scene_tower = MLPTower(D, H, O, gen, dev, tower_lr, tower_matmul)
        self.product_tower = MLPTower(D, H, O, gen, dev, tower_lr, tower_matmul)
        self.scorer = engine.InBatchScorer(B, O, loss=loss, margin=margin, scale=scale, device=dev)
        self.xs = torch.empty(self.B, D, dtype=torch.float32, device=dev)
        self.xp = torch.empty(self.B, D, dtype=torch.float32, device=dev)
        self.us = _RowUpdater(scene_table, self.B, lr)
        self.up = _RowUpdater(product_table, self.B, lr)

    def step(self, scene_ids, product_ids):
        """int32 CUDA (B,) each.  Returns the device loss scalar."""
        self.ts.gather(scene_ids, out=self.xs)
        self.tp.gather(product_ids, out=self.xp)
        q = self.scene_tower.forward(self.xs)
        k = self.product_tower.forward(self.xp)
        loss, dq, dk = self.scorer.run(q.contiguous(), k.contiguous())
        dxs = self.scene_tower.backward(dq)
        dxp = self.product_tower.backward(dk)
        self.scene_tower.update()
        self.product_tower.update()
        self.us.apply(scene_ids, dxs.contiguous())
        self.up.apply(product_ids, dxp.contiguous())
        return loss

    def graphed(self, example_scene_ids, example_product_ids):
        """``step`` as a CUDA graph (tables, accumulators, tower weights and Adam moments are restored after the warm-up
        and capture calls).  Adam's bias correction takes the step count from the host, so the captured graph freezes it:
        valid for benchmarking and for runs past the first few hundred steps (1 - beta^t -> 1), not a bit-exact replacement
        of the eager ``step`` at small t."""
        tensors = [self.ts.rows0, self.ts.acc, self.tp.rows0, self.tp.acc]
        for tw in (self.scene_tower, self.product_tower):
            tensors += list(tw.p.values()) + list(tw.mu.values()) + list(tw.nu.values())
        snap = [t.clone() for t in tensors]
        counts = (self.scene_tower.count, self.product_tower.count)
        g = _GraphedStep(lambda a, b: self.step(a, b), [example_scene_ids, example_product_ids])
        for t, v in zip(tensors, snap):
            t.copy_(v)
        self.scene_tower.count, self.product_tower.count = counts
        return g


class ShardedSharedTableInBatch:
    """configs[2] across the GPUs of one box ("table row-sharded 1 -> 8 B200 with index all-to-all"): the table is
    sharded cyclically (owner = row % n, as the GloVe path), every rank holds B_local (query, item) pairs, and the
    negatives of a query are the items of ALL ranks:

      1. index plan of the local 2*B_local ids; unique rows fetched from their owners      RowExchange.fetch (ids / rows all-to-all)
      2. Q, K_local = fetched rows in batch order                                          esr_permute_rows_f32
      3. all-gather of K over the ranks                                                    NCCL
      4. scores of the local queries against every item, loss, dQ, partial dK              esr_inbatch_fwd_bwd_bf16
         (Bq = B_local, Bk = n * B_local, diag_off = rank * B_local, b_norm = n * B_local)
      5. reduce-scatter of dK, all-reduce of the loss                                       NCCL
      6. per-row sum of the 2*B_local gradient rows, gradient all-to-all, owners' Adagrad   RowExchange.push

    One global step equals ``oracle.inbatch.shared_table_step`` on the rank-major concatenation of the batches."""

    def __init__(self, V, D, B_local, lr=0.05, loss="hinge", margin=1.0, scale=1.0, group=None, device=None):
        import torch.distributed as dist
        from .sharded import LibesrOps, RowExchange, shard_rows
        L.require_cuda()
        self.dist, self.group = dist, group
        self.n, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.V, self.D, self.B, self.lr = int(V), int(D), int(B_local), float(lr)
        self.shard = engine.EmbeddingTable(shard_rows(V, self.rank, self.n), D, self.dev, sparse=False, adagrad=True)
        self.ops = LibesrOps(self.dev)
        self.xchg = RowExchange(self.ops, self.shard, group)
        n_slots = 2 * self.B
        self.plan = engine.IndexPlan(n_slots, V, self.dev, with_partner=False, sort="wide")
        Bg = self.B * self.n
        self.scorer = engine.InBatchScorer(self.B, D, Bk=Bg, loss=loss, diag_off=self.rank * self.B, margin=margin,
                                           scale=scale, b_norm=Bg, device=self.dev)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.X = torch.empty(n_slots, D, **f32)
        self.K_all = torch.empty(Bg, D, **f32)
        self.dX = torch.empty(n_slots, D, **f32)
        self.scorer.dQ = self.dX[:self.B]
        self.gsum = torch.empty(n_slots, D, **f32)
        self.zeros_b = torch.zeros(n_slots, **f32)
        self.remap = torch.empty(n_slots, dtype=torch.int32, device=self.dev)

    def load_dense(self, E):
        idx = torch.arange(self.rank, self.V, self.n)
        self.shard.rows0.copy_(torch.as_tensor(E)[idx].to(self.dev))

    def gather_dense(self):
        E = torch.zeros(self.V, self.D, device=self.dev)
        E[torch.arange(self.rank, self.V, self.n, device=self.dev)] = self.shard.rows0
        self.dist.all_reduce(E, group=self.group)
        return E

    def step(self, ids):
        """ids: int32 (2, B_local) global rows [queries ; items].  Returns the GLOBAL loss (device scalar)."""
        dist, B, D = self.dist, self.B, self.D
        flat = ids.to(self.dev, non_blocking=True).reshape(-1).contiguous()
        plan = self.plan.build(flat)
        rows, _ = self.xchg.fetch(plan.uniq, plan.n_uniq)
        plan.remap_ids(self.remap)
        self.ops.permute_rows(rows, self.remap, 2 * B, False, self.X)           # X[s] = row of slot s
        dist.all_gather_into_tensor(self.K_all, self.X[B:], group=self.group)
        loss, _, dK = self.scorer.run(self.X[:B], self.K_all)
        dist.reduce_scatter_tensor(self.dX[B:], dK, group=self.group)
        loss = loss.clone()
        dist.all_reduce(loss, group=self.group)
        L.check(L.lib().esr_segment_sum_rows_f32(C.byref(plan.s), D, L.ptr(self.dX), None, L.ptr(self.gsum), None,
                                                 L.stream_ptr()), "esr_segment_sum_rows_f32")
        self.xchg.push(self.gsum, self.zeros_b, self.lr)
        return loss[0]


class ShardedTwoTowerInBatch:
    """configs[3] across the GPUs of one box: scene-id and product-id tables row-sharded cyclically, the two MLP
    towers replicated (their gradients all-reduced -- SURVEY.md 8(e)), B_local pairs per rank, in-batch negatives over
    the items of ALL ranks (all-gather of the product-tower outputs, reduce-scatter of their gradients).  One global
    step equals ``oracle.inbatch.two_tower_step`` on the rank-major concatenation of the batches."""

    def __init__(self, Vs, Vp, D, B_local, hidden=None, out=None, lr=0.05, tower_lr=1e-3, loss="softmax", margin=1.0,
                 scale=1.0, seed=0, group=None, device=None):
        import torch.distributed as dist
        from .sharded import LibesrOps, RowExchange, shard_rows
        L.require_cuda()
        self.dist, self.group = dist, group
        self.n, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.Vs, self.Vp, self.D, self.B, self.lr = int(Vs), int(Vp), int(D), int(B_local), float(lr)
        H, O = int(hidden or D), int(out or D)
        mk = lambda V: engine.EmbeddingTable(shard_rows(V, self.rank, self.n), D, self.dev, sparse=False, adagrad=True)
        self.shard_s, self.shard_p = mk(Vs), mk(Vp)
        self.ops = LibesrOps(self.dev)
        self.xs, self.xp = RowExchange(self.ops, self.shard_s, group), RowExchange(self.ops, self.shard_p, group)
        self.plan_s = engine.IndexPlan(self.B, Vs, self.dev, with_partner=False, sort="wide")
        self.plan_p = engine.IndexPlan(self.B, Vp, self.dev, with_partner=False, sort="wide")
        gen = torch.Generator(device="cpu").manual_seed(seed)           # same seed on every rank: replicated towers
        self.scene_tower = MLPTower(D, H, O, gen, self.dev, tower_lr)
        self.product_tower = MLPTower(D, H, O, gen, self.dev, tower_lr)
        Bg = self.B * self.n
        self.scorer = engine.InBatchScorer(self.B, O, Bk=Bg, loss=loss, diag_off=self.rank * self.B, margin=margin,
                                           scale=scale, b_norm=Bg, device=self.dev)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.Xs, self.Xp = torch.empty(self.B, D, **f32), torch.empty(self.B, D, **f32)
        self.K_all = torch.empty(Bg, O, **f32)
        self.dk_loc = torch.empty(self.B, O, **f32)
        self.gs, self.gp = torch.empty(self.B, D, **f32), torch.empty(self.B, D, **f32)
        self.zeros_b = torch.zeros(self.B, **f32)
        self.remap = torch.empty(self.B, dtype=torch.int32, device=self.dev)

    def load_dense(self, Es, Ep):
        self.shard_s.rows0.copy_(torch.as_tensor(Es)[torch.arange(self.rank, self.Vs, self.n)].to(self.dev))
        self.shard_p.rows0.copy_(torch.as_tensor(Ep)[torch.arange(self.rank, self.Vp, self.n)].to(self.dev))

    def gather_dense(self):
        out = []
        for V, sh in ((self.Vs, self.shard_s), (self.Vp, self.shard_p)):
            E = torch.zeros(V, self.D, device=self.dev)
            E[torch.arange(self.rank, V, self.n, device=self.dev)] = sh.rows0
            self.dist.all_reduce(E, group=self.group)
            out.append(E)
        return out

    def _fetch(self, plan, xchg, ids, X):
        plan.build(ids)
        rows, _ = xchg.fetch(plan.uniq, plan.n_uniq)
        plan.remap_ids(self.remap)
        self.ops.permute_rows(rows, self.remap, self.B, False, X)

    def _push(self, plan, xchg, dX, gsum):
        L.check(L.lib().esr_segment_sum_rows_f32(C.byref(plan.s), self.D, L.ptr(dX), None, L.ptr(gsum), None, L.stream_ptr()),
                "esr_segment_sum_rows_f32")
        xchg.push(gsum, self.zeros_b, self.lr)

    def step(self, scene_ids, product_ids):
        """int32 (B_local,) global ids each.  Returns the GLOBAL loss (device scalar)."""
        dist = self.dist
        s_ids = scene_ids.to(self.dev, torch.int32).contiguous()
        p_ids = product_ids.to(self.dev, torch.int32).contiguous()
        self._fetch(self.plan_s, self.xs, s_ids, self.Xs)
        self._fetch(self.plan_p, self.xp, p_ids, self.Xp)
        q = self.scene_tower.forward(self.Xs)
        k = self.product_tower.forward(self.Xp)
        dist.all_gather_into_tensor(self.K_all, k.contiguous(), group=self.group)
        loss, dq, dK = self.scorer.run(q.contiguous(), self.K_all)
        dist.reduce_scatter_tensor(self.dk_loc, dK, group=self.group)
        loss = loss.clone()
        dist.all_reduce(loss, group=self.group)
        dxs = self.scene_tower.backward(dq)
        dxp = self.product_tower.backward(self.dk_loc)
        for tower in (self.scene_tower, self.product_tower):       # replicated dense params: sum of the ranks' gradients
            dist.all_reduce(tower.flat["g"], group=self.group)     # (one flat buffer per tower: one collective)
            tower.update()
        self._push(self.plan_s, self.xs, dxs.contiguous(), self.gs)
        self._push(self.plan_p, self.xp, dxp.contiguous(), self.gp)
        return loss[0]
