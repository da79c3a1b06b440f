"""Row-sharded GloVe training across the GPUs of one box (BASELINE.json north star; the reference
itself is single-device -- SURVEY.md 0, 8(e)).

One process per GPU.  The table is sharded CYCLICALLY (owner = row % n, local = row // n; rows are
frequency ranks, wikipedia/make_dictionary.py:113-116).  A step on rank r, for its B_local pairs:

  1. index plan of the local batch (sort / unique / segments)             libesr
  2. bucket the unique rows by owner                                      libesr  esr_route_plan_i32
  3. count all-to-all, id all-to-all                                      NCCL
  4. owners gather the requested rows (+ biases)                          libesr
  5. row all-to-all back; scatter into the compact table in unique order  NCCL + libesr
  6. the GloVe step on the compact table in EMIT_GRADS mode, with the batch sums all-reduced
     (3 floats before the row pass, 2 after) so every rank uses the GLOBAL mean(bs), S0, S1
  7. gradient all-to-all to the owners                                    NCCL
  8. owners merge duplicates across source ranks (sorted, deterministic) and apply sparse Adagrad

The exchange choreography (``RowExchange``) is device-agnostic given an ``ops`` object; the product
uses ``LibesrOps`` (CUDA only -- it raises without a GPU).  tests/ drives the same choreography on CPU
with gloo and NumPy ops to pin the routing.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from .engine import EmbeddingTable, GloveStep, IndexPlan


def shard_rows(V, rank, n):
    return (V - rank + n - 1) // n


class LibesrOps:
    """The CUDA implementation of the per-rank pieces (every method enqueues on the current stream)."""

    def __init__(self, device):
        L.require_cuda()
        self.dev = device
        self._route_ws = None
        self._plans = {}

    def route_plan(self, uniq, n_uniq, n_ranks, out=None):
        cap = uniq.numel()
        i32 = dict(dtype=torch.int32, device=self.dev)
        inv_order = None
        if out is not None:
            order, send_local, send_counts = out[:3]
            inv_order = out[3] if len(out) > 3 else None
        else:
            order = torch.empty(cap, **i32)
            send_local = torch.empty(cap, **i32)
            send_counts = torch.zeros(n_ranks, **i32)
        need = int(L.lib().esr_route_workspace_bytes(cap))
        if self._route_ws is None or self._route_ws.numel() < need:
            self._route_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        L.check(L.lib().esr_route_plan_i32(L.ptr(uniq), L.ptr(n_uniq), cap, n_ranks, L.ptr(order), L.ptr(send_local),
                                           L.ptr(send_counts), L.ptr(inv_order), L.ptr(self._route_ws), self._route_ws.numel(),
                                           L.stream_ptr()), "esr_route_plan_i32")
        return order, send_local, send_counts

    def gather_rows(self, table: EmbeddingTable, ids):
        return table.gather(ids)

    def gather_scalar(self, src, ids):
        out = torch.empty(ids.numel(), dtype=torch.float32, device=self.dev)
        L.check(L.lib().esr_gather_scalar_f32(L.ptr(src), L.ptr(ids), ids.numel(), L.ptr(out), L.stream_ptr()),
                "esr_gather_scalar_f32")
        return out

    def permute_rows(self, src, idx, n, scatter, out):
        """scatter: out[idx[k]] = src[k]; else out[k] = src[idx[k]]; k < n."""
        D = src.shape[1]
        L.check(L.lib().esr_permute_rows_f32(L.ptr(src), L.ptr(idx), None, int(n), D, 1 if scatter else 0, L.ptr(out),
                                             L.stream_ptr()), "esr_permute_rows_f32")
        return out

    def owner_update(self, shard: EmbeddingTable, recv_ids, recv_g, recv_gb, lr, eps):
        """Merge the gradients received for owner-local rows (duplicates across source ranks) and apply
        optax.adagrad to the shard (App. A.5)."""
        R = recv_ids.numel()
        if R == 0:
            return
        cap = 1 << max(10, (R - 1).bit_length())
        plan = self._plans.get(cap)
        if plan is None:
            plan = IndexPlan(cap, shard.V, self.dev, with_partner=False)
            self._plans[cap] = plan
        plan.build(recv_ids)
        D = shard.D
        gsum = torch.empty(R, D, dtype=torch.float32, device=self.dev)
        gbsum = torch.empty(R, dtype=torch.float32, device=self.dev)
        L.check(L.lib().esr_segment_sum_rows_f32(C.byref(plan.s), D, L.ptr(recv_g), L.ptr(recv_gb), L.ptr(gsum),
                                                 L.ptr(gbsum), L.stream_ptr()), "esr_segment_sum_rows_f32")
        L.check(L.lib().esr_sparse_adagrad_f32(C.byref(shard.struct()), L.ptr(plan.uniq), L.ptr(plan.n_uniq), R,
                                               L.ptr(gsum), L.ptr(gbsum), lr, eps, L.stream_ptr()), "esr_sparse_adagrad_f32")


class RowExchange:
    """fetch(): unique global rows -> their current values from the owners; push(): per-row gradients
    -> owners, merged and applied.  ``ops`` supplies the per-rank compute; ``shard`` is this rank's table."""

    def __init__(self, ops, shard, group=None):
        self.ops, self.shard, self.group = ops, shard, group
        self.n = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.state = None

    def _a2a(self, send, out_splits, in_splits, trailing=()):
        out = send.new_empty((sum(out_splits),) + tuple(trailing))
        dist.all_to_all_single(out, send, out_splits, in_splits, group=self.group)
        return out

    def fetch(self, uniq, n_uniq_dev):
        """uniq: sorted unique global rows (capacity-sized; first U valid).  Returns (rows[U,D], bias[U])
        in the order of ``uniq``."""
        ops = self.ops
        order, send_local, send_counts = ops.route_plan(uniq, n_uniq_dev, self.n)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc = [int(x) for x in send_counts.cpu().tolist()]       # host split sizes (synchronises)
        rc = [int(x) for x in recv_counts.cpu().tolist()]
        U = sum(sc)
        recv_ids = self._a2a(send_local[:U].contiguous(), rc, sc)
        rows_out = ops.gather_rows(self.shard, recv_ids)
        bias_out = ops.gather_scalar(self.shard.bias, recv_ids)
        D = self.shard.D
        got_rows = self._a2a(rows_out.reshape(-1, D), sc, rc, (D,))
        got_bias = self._a2a(bias_out, sc, rc)
        rows = got_rows.new_empty(max(U, 1), D)
        ops.permute_rows(got_rows, order, U, True, rows)           # rows[order[k]] = got_rows[k]
        bias = got_bias.new_zeros(max(U, 1))
        bias[order[:U].long()] = got_bias
        self.state = (order, sc, rc, recv_ids, U)
        return rows, bias

    def push(self, dE, db, lr, eps=1e-7):
        """dE[U,D], db[U] in the order of ``uniq`` -> owners."""
        order, sc, rc, recv_ids, U = self.state
        ops = self.ops
        D = self.shard.D
        send_g = dE.new_empty(max(U, 1), D)
        ops.permute_rows(dE, order, U, False, send_g)              # send_g[k] = dE[order[k]]
        send_gb = db[order[:U].long()]
        recv_g = self._a2a(send_g[:U], rc, sc, (D,))
        recv_gb = self._a2a(send_gb.contiguous(), rc, sc)
        ops.owner_update(self.shard, recv_ids, recv_g, recv_gb, lr, eps)


class ShardedGloveTrainer:
    # libesr kernels one step launches (NCCL excluded): plan 5 (digit histogram, 3 sort passes, head pass), route 5, owner
    # gathers 2, row permutes 3, compact 2, prep / rows / combine / finish 4, owner plan 5 + segment sum 1 + Adagrad 2
    LAUNCHES_PER_STEP = 29

    def __init__(self, V, D, B_local, lr=0.05, bias_mode="reference_broadcast", group=None, device=None, chunk=0):
        L.require_cuda()
        self.group = group
        self.n = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.V, self.D, self.B, self.lr = int(V), int(D), int(B_local), float(lr)
        self.shard = EmbeddingTable(shard_rows(V, self.rank, self.n), D, self.dev, sparse=False, adagrad=True)
        self.ops = LibesrOps(self.dev)
        self.xchg = RowExchange(self.ops, self.shard, group)
        n_slots = 2 * self.B
        self.plan = IndexPlan(n_slots, V, self.dev)
        # compact table of fetched rows + the plan re-expressed in unique-row indices
        self.compact = EmbeddingTable(n_slots, D, self.dev, sparse=False, adagrad=False)
        self.cplan = IndexPlan(n_slots, n_slots, self.dev)
        self.scratch = torch.empty(n_slots, dtype=torch.int32, device=self.dev)
        self.step_fn = GloveStep(self.compact, self.B, lr=lr, bias_mode=bias_mode, chunk=chunk, emit_grads=True,
                                 B_global=self.B * self.n)
        self.loss = None

    def load_dense(self, E, b):
        """Scatter a dense (V,D) table / (V,) bias (host or device) into the shards."""
        idx = torch.arange(self.rank, self.V, self.n)
        self.shard.rows0.copy_(torch.as_tensor(E)[idx].to(self.dev))
        self.shard.bias.copy_(torch.as_tensor(b).reshape(-1)[idx].to(self.dev))

    def step(self, ids, counts):
        """ids: int32 (2, B_local) global rows; counts: f32 (B_local,).  Returns the GLOBAL loss (device scalar)."""
        ids = ids.to(self.dev, non_blocking=True).reshape(-1).contiguous()
        counts = counts.to(self.dev, non_blocking=True)
        plan, cplan = self.plan, self.cplan
        plan.build(ids)
        rows, bias = self.xchg.fetch(plan.uniq, plan.n_uniq)
        U = rows.shape[0]
        self.compact.rows0[:U].copy_(rows)
        self.compact.bias[:U].copy_(bias)
        L.check(L.lib().esr_plan_compact_i32(C.byref(plan.s), L.ptr(cplan.sorted_keys), L.ptr(cplan.partner),
                                             L.ptr(cplan.uniq), L.ptr(self.scratch), L.stream_ptr()), "esr_plan_compact_i32")
        # the compact plan shares perm / useg / seg_off / n_uniq with the original
        cs = cplan.s
        cs.n_slots = plan.n_slots
        cs.perm, cs.useg, cs.seg_off, cs.n_uniq = plan.s.perm, plan.s.useg, plan.s.seg_off, plan.s.n_uniq
        st = self.step_fn
        st.prep(cplan, counts)
        dist.all_reduce(st.scalars[0:3], group=self.group)
        st.rows(cplan)
        dist.all_reduce(st.scalars[3:5], group=self.group)
        st.finish(cplan)
        self.xchg.push(st.dE, st.db, self.lr)
        self.loss = st.scalars[L.SC_LOSS].clone()
        return self.loss

    def gather_dense(self):
        """All shards -> dense (V,D) table and (V,) bias on every rank (test / checkpoint helper)."""
        E = torch.zeros(self.V, self.D, device=self.dev)
        b = torch.zeros(self.V, device=self.dev)
        idx = torch.arange(self.rank, self.V, self.n, device=self.dev)
        E[idx] = self.shard.rows0
        b[idx] = self.shard.bias
        dist.all_reduce(E, group=self.group)
        dist.all_reduce(b, group=self.group)
        return E, b


class PeerShardedGloveTrainer:
    """Same step as ShardedGloveTrainer, but the exchanges are libesr kernels over NVLink peer memory
    (csrc/peer_ops.cu) on buffers from torch's symmetric-memory rendezvous: no NCCL all-to-all, no id
    exchange, no host-known sizes, no host synchronisation inside the step.

      side stream : index plan + route plan (published in symmetric memory, double-buffered) + compact plan of batch t+1
      main stream : peer gather of the unique rows (NVLink loads) -> prep -> all-reduce(3 floats) -> owners pull the
                    id lists and resolve who sends what -> emit plan -> row pass, whose gradient rows are STORED
                    STRAIGHT INTO THE OWNERS' INBOXES over NVLink -> all-reduce(2) -> finish -> barrier
                    -> owners merge their inbox locally, Adagrad -> barrier
    ``overlap_ids=True`` (EXPERIMENTAL, off): pull / resolve / emit plan move to a second side stream behind a barrier at
    the top of the step, which also replaces the barrier at its end.
    NOTE (round 2): this trainer runs its main half on the CALLER's stream and makes the side stream wait for it, so the
    two halves do not overlap (the owner-routed trainer below fixed that for itself: 385 -> 303 us per step at 2 GPUs);
    kept as the round-1 reference point.
    ``graphs=True`` captures both halves into CUDA graphs per parity after two eager steps (the eager step is ~25 host
    calls and partly host-bound).  EXPERIMENTAL and off by default: in round 1 the replay of the captured
    NCCL all-reduce + symmetric-memory barrier sequence dead-locked in the 2-GPU bench (killed by its timeout).
    """

    DEPTH = 2
    # libesr kernels one step launches (NCCL all-reduces and symmetric-memory barriers excluded): plan 5 (digit histogram,
    # 3 sort passes, head pass), route 5, compact 2, gather, prep, pull, resolve, emit plan, rows, combine, finish, merge,
    # clear map
    LAUNCHES_PER_STEP = 22

    def __init__(self, V, D, B_local, lr=0.05, bias_mode="reference_broadcast", group=None, device=None, chunk=0,
                 graphs=False, fast_sync=False, overlap_ids=False, impl="auto"):
        import torch.distributed._symmetric_memory as symm_mem
        L.require_cuda()
        self.group = group if group is not None else dist.group.WORLD
        self.n = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.n > 8:
            raise ValueError("at most 8 ranks (one NVSwitch domain; ESR_MAX_PEERS)")
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.V, self.D, self.B, self.lr = int(V), int(D), int(B_local), float(lr)
        n_slots = 2 * self.B
        gname = self.group.group_name
        self._hdls = []

        def symm(shape, dtype):
            t = symm_mem.empty(shape, dtype=dtype, device=self.dev)
            h = symm_mem.rendezvous(t, gname)
            self._hdls.append(h)
            ptrs = (C.c_void_p * 8)(*[int(p) for p in h.buffer_ptrs])
            return t, ptrs

        i32 = dict(dtype=torch.int32, device=self.dev)
        V_loc = shard_rows(V, self.rank, self.n)
        V_max = shard_rows(V, 0, self.n)
        rows, self.p_rows = symm((V_max, D), torch.float32)
        bias, self.p_bias = symm((V_max,), torch.float32)
        rows.zero_()
        bias.zero_()
        self.shard = EmbeddingTable.wrap(rows[:V_loc], bias=bias[:V_loc],
                                         acc=torch.full((V_loc, D), 0.1, device=self.dev),
                                         bias_acc=torch.full((V_loc,), 0.1, device=self.dev))
        # route plan published per step parity: peers read set k of step t while set 1-k is rebuilt for t+1
        self.pub = []
        for _ in range(self.DEPTH):
            counts, p_counts = symm((16,), torch.int32)
            send_local, p_send_local = symm((n_slots,), torch.int32)
            self.pub.append(dict(counts=counts, p_counts=p_counts, send_local=send_local, p_send_local=p_send_local,
                                 order=torch.empty(n_slots, **i32), inv_order=torch.empty(n_slots, **i32)))
        # gradient inbox: the sources' row passes scatter their rows for my shard straight into it (NVLink stores)
        # every source can name up to n_slots rows of ONE owner (ids congruent mod n), so the worst case is n * n_slots rows;
        # only the part a step touches costs bandwidth.  k_peer_emit_plan still flags an overflow in ``err`` (check()).
        self.inbox_cap = n_slots * self.n
        self.inbox_dE, self.p_inbox_dE = symm((self.inbox_cap, D), torch.float32)
        self.inbox_db, self.p_inbox_db = symm((self.inbox_cap,), torch.float32)
        self.emit_map = torch.zeros(n_slots, **i32)
        self.err = torch.zeros(1, **i32)
        self.ops = LibesrOps(self.dev)
        self.plans = [IndexPlan(n_slots, V, self.dev) for _ in range(self.DEPTH)]
        self.compact = EmbeddingTable(n_slots, D, self.dev, sparse=False, adagrad=False)
        # the plan re-expressed in unique-row indices depends on the ids only: built on the side stream, per parity;
        # perm / useg / seg_off / n_uniq are shared with the original plan
        self.cplans = [IndexPlan(n_slots, n_slots, self.dev) for _ in range(self.DEPTH)]
        for cp, pl in zip(self.cplans, self.plans):
            cp.s.n_slots = n_slots
            cp.s.perm, cp.s.useg, cp.s.seg_off, cp.s.n_uniq = pl.s.perm, pl.s.useg, pl.s.seg_off, pl.s.n_uniq
        self.cplan = self.cplans[0]
        self.scratch = torch.empty(n_slots, **i32)
        # full persistent grid: leaving CTA slots free (as the single-GPU trainer does for its plan stream) measured
        # 2 % slower here -- the id kernels of this path are short and run between the phases anyway
        self.step_fn = GloveStep(self.compact, self.B, lr=lr, bias_mode=bias_mode, chunk=chunk, emit_grads=True,
                                 B_global=self.B * self.n, dE=self.inbox_dE, db=self.inbox_db, impl=impl)
        cfg = self.step_fn.cfg
        cfg.emit_map = L.ptr(self.emit_map)
        cfg.emit_peers_dE = C.cast(self.p_inbox_dE, C.c_void_p)
        cfg.emit_peers_db = C.cast(self.p_inbox_db, C.c_void_p)
        cfg.n_emit_peers = self.n
        self.recv_cap = self.inbox_cap
        self.recv_ids = torch.empty(self.recv_cap, **i32)
        self.src_meta = torch.zeros(3 * 8 + 4, **i32)
        self.map_stride = V_max
        self.slot_map = torch.full((self.n, V_max), -1, **i32)
        self.desc = torch.empty(self.recv_cap * (self.n + 2), **i32)   # per entry: n source rows; then 8-byte owner records
        self.s_side = torch.cuda.Stream(self.dev)
        self.ev_plan = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.ev_done = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self._keep = [None] * self.DEPTH
        # static staging of the batch: the step's kernels (and, once captured, its CUDA graphs) always read these
        self.st_ids = [torch.ones(n_slots, **i32) for _ in range(self.DEPTH)]
        self.st_counts = [torch.ones(self.B, dtype=torch.float32, device=self.dev) for _ in range(self.DEPTH)]
        self.use_graphs = bool(graphs)
        self.g_plan = [None] * self.DEPTH
        self.g_step = [None] * self.DEPTH
        self.t = 0
        self.loss = None
        self.loss_log = torch.zeros(4096, dtype=torch.float32, device=self.dev)
        # fast_sync (EXPERIMENTAL, off): the two batch-sum all-reduces and the two phase barriers of a step as libesr
        # peer-memory kernels (esr_peer_allreduce_f32) instead of NCCL + the symmetric-memory barrier -- one NVLink
        # round trip each, and no library collective left inside the step, so the whole step can be graph-captured
        self.fast_sync = bool(fast_sync)
        if self.fast_sync:
            words = int(L.lib().esr_peer_sync_bytes()) // 4
            self.sync, self.p_sync = symm((words,), torch.int32)
            self.sync.zero_()
            self.sync_seq = torch.zeros(1, **i32)
        # overlap_ids (EXPERIMENTAL, off): the owner-side id pull / resolve and the source-side emit plan depend on the
        # published route plans only, so they run on a second side stream next to gather + prep instead of between the
        # first all-reduce and the row pass; a barrier at the top of the step (all plans published, all updates of the
        # previous step applied) replaces the one at its end
        self.overlap_ids = bool(overlap_ids)
        self.s_ids = torch.cuda.Stream(self.dev)
        self.ev_top = torch.cuda.Event()
        self.ev_ids = torch.cuda.Event()
        torch.cuda.current_stream().synchronize()
        self._hdls[0].barrier()

    def barrier(self):
        if self.fast_sync:
            L.check(L.lib().esr_peer_allreduce_f32(self.p_sync, self.n, self.rank, None, None, 0, L.ptr(self.sync_seq),
                                                   L.stream_ptr()), "esr_peer_allreduce_f32")
        else:
            self._hdls[0].barrier()

    def _all_reduce(self, view):
        """Sum of a few floats over the ranks, in place."""
        if self.fast_sync:
            L.check(L.lib().esr_peer_allreduce_f32(self.p_sync, self.n, self.rank, L.ptr(view), L.ptr(view), view.numel(),
                                                   L.ptr(self.sync_seq), L.stream_ptr()), "esr_peer_allreduce_f32")
        else:
            dist.all_reduce(view, group=self.group)

    # -- the two halves of a step; every call enqueues on the CURRENT stream ------------------------------------
    def _plan_body(self, k):
        """Everything that depends on the ids only (side stream): index plan, owner routing (published for the
        peers), the plan re-expressed in unique-row indices."""
        plan, pub, cplan = self.plans[k], self.pub[k], self.cplans[k]
        plan.build(self.st_ids[k])
        self.ops.route_plan(plan.uniq, plan.n_uniq, self.n, out=(pub["order"], pub["send_local"], pub["counts"],
                                                                 pub["inv_order"]))
        cplan.s.n_slots = plan.n_slots
        L.check(L.lib().esr_plan_compact_i32(C.byref(plan.s), L.ptr(cplan.sorted_keys), L.ptr(cplan.partner),
                                             L.ptr(cplan.uniq), L.ptr(self.scratch), L.stream_ptr()), "esr_plan_compact_i32")

    def _ids_body(self, k, sp):
        """Owner side: which rows the sources send me and where each source's gradient for a row lands in my inbox;
        source side: where my gradient rows go.  Depends on the published route plans only."""
        lib, n = L.lib(), self.n
        plan, pub = self.plans[k], self.pub[k]
        L.check(lib.esr_peer_pull_ids_i32(pub["p_counts"], pub["p_send_local"], n, self.rank, self.recv_cap,
                                          L.ptr(self.recv_ids), L.ptr(self.src_meta), L.ptr(self.slot_map), self.map_stride,
                                          sp), "esr_peer_pull_ids_i32")
        L.check(lib.esr_peer_resolve_i32(n, L.ptr(self.recv_ids), L.ptr(self.src_meta), L.ptr(self.slot_map),
                                         self.map_stride, L.ptr(self.desc), self.recv_cap, sp), "esr_peer_resolve_i32")
        L.check(lib.esr_peer_emit_plan_i32(pub["p_counts"], n, self.rank, L.ptr(plan.uniq), L.ptr(plan.n_uniq),
                                           plan.capacity, L.ptr(pub["inv_order"]), self.inbox_cap, L.ptr(self.emit_map),
                                           L.ptr(self.err), sp), "esr_peer_emit_plan_i32")

    def _step_body(self, k):
        lib, n = L.lib(), self.n
        plan, cplan, st = self.plans[k], self.cplans[k], self.step_fn
        sp = L.stream_ptr()
        if self.overlap_ids:
            main = torch.cuda.current_stream(self.dev)
            self.barrier()                                      # every route plan of this step is published and every
            self.ev_top.record(main)                            # owner has applied the previous step's updates
            self.s_ids.wait_event(self.ev_top)
            with torch.cuda.stream(self.s_ids):
                self._ids_body(k, L.stream_ptr())
                self.ev_ids.record(self.s_ids)
        L.check(lib.esr_peer_gather_f32(self.p_rows, self.p_bias, n, L.ptr(plan.uniq), L.ptr(plan.n_uniq), plan.capacity,
                                        self.D, L.ptr(self.compact.rows0), L.ptr(self.compact.bias), sp), "esr_peer_gather_f32")
        st.prep(cplan, self.st_counts[k])
        self._all_reduce(st.scalars[0:3])                       # also orders: all fetches done, all route plans published
        if self.overlap_ids:
            torch.cuda.current_stream(self.dev).wait_event(self.ev_ids)
        else:
            self._ids_body(k, sp)
        st.rows(cplan)                                          # gradient rows go straight to the owners' inboxes
        self._all_reduce(st.scalars[3:5])
        st.finish(cplan)
        self.barrier()                                          # every rank's gradients have landed in the inboxes
        L.check(lib.esr_peer_apply_adagrad_f32(C.byref(self.shard.struct()), L.ptr(self.inbox_dE), L.ptr(self.inbox_db), n,
                                               L.ptr(self.recv_ids), L.ptr(self.src_meta), L.ptr(self.slot_map),
                                               self.map_stride, L.ptr(self.desc), self.recv_cap, self.lr, 1e-7, sp),
                "esr_peer_apply_adagrad_f32")
        if not self.overlap_ids:
            self.barrier()                                      # every owner has applied its updates

    def _capture(self):
        """Capture the two halves per parity into CUDA graphs (the eager step is ~25 host calls: at 0.5 ms per step the
        host was the bound).  Collective: every rank captures the same sequence.  Falls back to eager launches if the
        runtime refuses to capture a collective."""
        try:
            torch.cuda.synchronize(self.dev)
            cap = torch.cuda.Stream(self.dev)
            for k in range(self.DEPTH):
                gp = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gp, stream=cap, capture_error_mode="thread_local"):
                    self._plan_body(k)
                gs = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gs, stream=cap, capture_error_mode="thread_local"):
                    self._step_body(k)
                self.g_plan[k], self.g_step[k] = gp, gs
            torch.cuda.synchronize(self.dev)
        except Exception as e:  # pragma: no cover
            import warnings
            warnings.warn("sharded step: CUDA-graph capture failed (%s); staying on eager launches" % e)
            self.g_plan = [None] * self.DEPTH
            self.g_step = [None] * self.DEPTH
            self.use_graphs = False

    def load_dense(self, E, b):
        idx = torch.arange(self.rank, self.V, self.n)
        self.shard.rows0.copy_(torch.as_tensor(E)[idx].to(self.dev))
        self.shard.bias.copy_(torch.as_tensor(b).reshape(-1)[idx].to(self.dev))
        torch.cuda.current_stream().synchronize()
        self.barrier()

    def check(self):
        """Synchronise and raise if any step overflowed a gradient inbox (the device flag of esr_peer_emit_plan_i32):
        the overflowing rows were clamped, so the tables can no longer be trusted."""
        torch.cuda.synchronize(self.dev)
        if int(self.err.item()) != 0:
            raise RuntimeError("PeerShardedGloveTrainer: gradient inbox overflow (inbox_cap=%d rows); training state is invalid"
                               % self.inbox_cap)

    def synchronize(self):
        self.check()

    def step(self, ids, counts):
        """ids: int32 (2, B_local) global rows (host pinned or device); counts: f32 (B_local,).
        Enqueues one step; returns the GLOBAL loss as a device scalar."""
        k = self.t % self.DEPTH
        main = torch.cuda.current_stream(self.dev)
        side = self.s_side
        if self.use_graphs and self.g_step[0] is None and self.t == 2 * self.DEPTH:
            self._capture()                    # after two eager steps per parity (lazy NCCL / allocator state is warm)
        # ---- side stream: stage the batch, then everything that depends on the ids only ----
        side.wait_stream(main)                 # inputs may have been produced on the caller's stream
        side.wait_event(self.ev_done[k])       # peers are done with publish set k (barrier of step t-2 passed)
        with torch.cuda.stream(side):
            self.st_ids[k].copy_(ids.reshape(-1), non_blocking=True)
            self.st_counts[k].copy_(counts, non_blocking=True)
            self._keep[k] = (ids, counts)
            if self.g_plan[k] is not None:
                self.g_plan[k].replay()
            else:
                self._plan_body(k)
            self.ev_plan[k].record(side)
        main.wait_event(self.ev_plan[k])
        # ---- main stream ----
        if self.g_step[k] is not None:
            self.g_step[k].replay()
        else:
            self._step_body(k)
        self.ev_done[k].record(main)
        slot = self.t % self.loss_log.numel()
        self.loss_log[slot: slot + 1].copy_(self.step_fn.scalars[L.SC_LOSS: L.SC_LOSS + 1], non_blocking=True)
        self.loss = self.loss_log[slot]
        self.t += 1
        return self.loss

    def gather_dense(self):
        self.check()
        E = torch.zeros(self.V, self.D, device=self.dev)
        b = torch.zeros(self.V, device=self.dev)
        idx = torch.arange(self.rank, self.V, self.n, device=self.dev)
        E[idx] = self.shard.rows0
        b[idx] = self.shard.bias
        dist.all_reduce(E, group=self.group)
        dist.all_reduce(b, group=self.group)
        return E, b


# ---- retrieval over a row-sharded table ("sharded: local top-k -> all-gather -> merge") ---------------
def merge_topk(cand_val, cand_idx, k, ties_high_index_first=False):
    """Top-k of per-query candidate lists.  ``cand_val`` f32 (T, M) (-inf = empty), ``cand_idx`` int64 (T, M) global rows
    (-1 = empty).  Equal scores are ordered by GLOBAL row: lower row first (``jax.lax.top_k``, spotify/train_spotify.py:
    121-124) or higher row first (the tail of the stable ascending ``argsort`` dump_knn reads, wikipedia/
    train_cooccurence.py:91-97,121-125).  Returns ``(values (T,k), rows (T,k))``."""
    order = torch.argsort(cand_idx, dim=1, descending=bool(ties_high_index_first), stable=True)
    ci, cv = torch.gather(cand_idx, 1, order), torch.gather(cand_val, 1, order)
    best = torch.argsort(cv, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(cv, 1, best), torch.gather(ci, 1, best)


def sharded_query_rows(shard_rows, rows, rank, n, group=None):
    """Rows ``rows`` (global ids, identical on every rank) of a cyclically sharded table, on every rank: each owner
    contributes the rows it holds, one all-reduce.  ``shard_rows``: this rank's (V_local, D) rows (row r lives at
    r // n on rank r % n)."""
    rows = torch.as_tensor(rows, device=shard_rows.device).long()
    out = torch.zeros(rows.numel(), shard_rows.shape[1], dtype=shard_rows.dtype, device=shard_rows.device)
    mine = (rows % n) == rank
    out[mine] = shard_rows[rows[mine] // n]
    dist.all_reduce(out, group=group)
    return out


def sharded_table_topk(shard_table, queries, k, rank, n, group=None, ties_high_index_first=False, local_topk=None):
    """dump_knn / find_top_k over a row-sharded table (owner = row % n): every rank scans ITS rows once with the fused
    score + top-k kernel (esr_topk_scan_f32 through engine.table_topk), the n local lists are all-gathered and merged in
    the reference's tie order (a row of the global top-k is beaten by fewer than k rows, hence by fewer than k rows of
    its own shard: it is in its shard's list).  ``queries`` (T, D) must be the same on every rank
    (``sharded_query_rows``).  Returns ``(values (T,k), global rows (T,k))``, identical on every rank."""
    from . import engine
    k = int(k)
    kl = min(k, int(shard_table.V))
    val, idx = (local_topk or engine.table_topk)(shard_table, queries, kl, ties_high_index_first)
    T = val.shape[0]
    pv = torch.full((T, k), float("-inf"), dtype=torch.float32, device=val.device)
    pi = torch.full((T, k), -1, dtype=torch.int64, device=val.device)
    pv[:, :kl] = val
    pi[:, :kl] = idx.long() * n + rank
    all_v = torch.empty(n * T, k, dtype=torch.float32, device=val.device)     # rank-major concatenation along dim 0
    all_i = torch.empty(n * T, k, dtype=torch.int64, device=val.device)
    dist.all_gather_into_tensor(all_v, pv.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, pi.contiguous(), group=group)
    return merge_topk(all_v.view(n, T, k).permute(1, 0, 2).reshape(T, n * k),
                      all_i.view(n, T, k).permute(1, 0, 2).reshape(T, n * k), k, ties_high_index_first)


def pair_capacity(B_local, n):
    """Pairs an owner must be able to take per step: its expected share is B_local (the global batch is n * B_local and
    ownership is cyclic over frequency-ranked ids), the margin covers the binomial spread (sigma ~ sqrt(B_local)) many
    times over; data that concentrates row i on one owner overflows and is reported by check()."""
    if n == 1:
        return int(B_local)
    cap = int(B_local) + max(int(B_local) // 16, 512)
    return (cap + 255) // 256 * 256


class OwnerRoutedGloveTrainer:
    """Row-sharded GloVe step with OWNER-COMPUTES pair routing (csrc/peer_ops.cu): every pair (i, j, x) of the global
    batch is processed by the rank that owns row i, so only the unique PARTNER rows j cross NVLink (and only their
    gradients travel back) -- 4.3x fewer NVLink bytes per direction than PeerShardedGloveTrainer on the bench stream at
    8 ranks.  Everything else is the peer-memory machinery of PeerShardedGloveTrainer: cyclic ownership, compact table
    of the step's unique rows, the row pass storing gradient rows straight into the owners' inboxes, owner-side merge in
    source order + Adagrad; no NCCL inside the step (libesr all-reduce / barrier kernels), both halves CUDA-graphed.

      side stream : stage the batch -> route pairs to their owners (peer stores, stable) -> barrier (side sequence)
                    -> collect my pairs -> index plan (device-side slot count) -> route plan (published) -> plan in
                    unified-table addresses
      main stream : barrier -> { gather stream: remote rows over NVLink | ids stream: owners pull the id lists, resolve,
                    emit plan | main: remote biases -> prep -> all-reduce(3) } -> row pass (gradients -> inboxes) -> all-reduce(2) -> { main: owner merge of the rows
                    + Adagrad | ids stream: finish -> barrier -> owner bias merge }

    One global step equals the single-table step on the concatenated batch (tests: N virtual ranks on one GPU against
    oracle.glove.step_adagrad; bench.py --gpus N runs the same check before its timed region)."""

    DEPTH = 2
    # libesr kernels per step: route pairs 3, collect 1, plan 5 (digit histogram, 3 sort passes at V <= 2^24, head pass),
    # route plan 3, address plan 2, gather 2 (rows, biases), prep, pull, resolve, emit plan, rows, combine, finish,
    # merge rows, merge bias, clear map, 2 all-reduces + 3 barriers
    LAUNCHES_PER_STEP = 31

    def __init__(self, V, D, B_local, lr=0.05, bias_mode="reference_broadcast", group=None, device=None, chunk=0,
                 graphs=True, pair_cap=None, impl="auto", row_blocks=None):
        import torch.distributed._symmetric_memory as symm_mem
        L.require_cuda()
        self.group = group if group is not None else dist.group.WORLD
        self.n = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.n > 8:
            raise ValueError("at most 8 ranks (one NVSwitch domain; ESR_MAX_PEERS)")
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.V, self.D, self.B, self.lr = int(V), int(D), int(B_local), float(lr)
        self.B_cap = int(pair_cap) if pair_cap is not None else pair_capacity(self.B, self.n)
        n, B, B_cap = self.n, self.B, self.B_cap
        n_slots = 2 * B_cap
        gname = self.group.group_name
        self._hdls = []

        def symm(shape, dtype):
            t = symm_mem.empty(shape, dtype=dtype, device=self.dev)
            h = symm_mem.rendezvous(t, gname)
            self._hdls.append(h)
            return t, (C.c_void_p * 8)(*[int(p) for p in h.buffer_ptrs])

        i32 = dict(dtype=torch.int32, device=self.dev)
        V_loc, V_max = shard_rows(V, self.rank, n), shard_rows(V, 0, n)
        # ONE allocation per rank: [shard rows (V_max) ; fetch region (n_slots)].  The row pass addresses a unique row of the
        # step either where it lives in the shard (rows this rank owns: every i, 1/n of the j) or in the fetch region
        # (rows pulled from their owners), so local rows are never copied (esr_plan_compact_owner_i32).
        rows, self.p_rows = symm((V_max + n_slots, D), torch.float32)
        bias, self.p_bias = symm((V_max + n_slots,), torch.float32)
        rows.zero_()
        bias.zero_()
        self.V_max = V_max
        self.shard = EmbeddingTable.wrap(rows[:V_loc], bias=bias[:V_loc], acc=torch.full((V_loc, D), 0.1, device=self.dev),
                                         bias_acc=torch.full((V_loc,), 0.1, device=self.dev))
        self.unified = EmbeddingTable.wrap(rows, bias=bias)
        self.fetch_rows, self.fetch_bias = rows[V_max:], bias[V_max:]
        self.pub, self.pin = [], []
        for _ in range(self.DEPTH):
            counts, p_counts = symm((16,), torch.int32)
            send_local, p_send_local = symm((n_slots,), torch.int32)
            self.pub.append(dict(counts=counts, p_counts=p_counts, send_local=send_local, p_send_local=p_send_local,
                                 order=torch.empty(n_slots, **i32), inv_order=torch.empty(n_slots, **i32)))
            rec, p_rec = symm((n, B, 4), torch.int32)           # pair inbox: 16-byte records {i, j, count bits, 0} per source
            cc, p_cc = symm((16,), torch.int32)
            cc.zero_()
            self.pin.append(dict(rec=rec, p_rec=p_rec, counts=cc, p_counts=p_cc))
        self.inbox_cap = n_slots * n
        self.inbox_dE, self.p_inbox_dE = symm((self.inbox_cap, D), torch.float32)
        self.inbox_db, self.p_inbox_db = symm((self.inbox_cap,), torch.float32)
        words = int(L.lib().esr_peer_sync_bytes()) // 4
        # three independent synchronisation sequences (one per stream that synchronises across ranks): 0 main stream
        # (barriers + all-reduces), 1 side stream (pairs routed), 2 ids stream (bias gradients landed)
        self.syncs = []
        for _ in range(3):
            blk, p_blk = symm((words,), torch.int32)
            blk.zero_()
            self.syncs.append((blk, p_blk, torch.zeros(1, **i32)))
        self.emit_map = torch.zeros(n_slots, **i32)
        self.err = torch.zeros(1, **i32)
        self.ops = LibesrOps(self.dev)
        self.my_counts = torch.zeros(16, **i32)
        self.route_ws = torch.empty(int(L.lib().esr_peer_route_pairs_workspace_bytes(B)), dtype=torch.uint8, device=self.dev)
        # per parity: the pairs I own this step (flat [i ; j] keys + counts + device-side slot count) and their plans
        self.keys = [torch.full((n_slots,), V, **i32) for _ in range(self.DEPTH)]
        self.cnt_l = [torch.zeros(B_cap, dtype=torch.float32, device=self.dev) for _ in range(self.DEPTH)]
        self.n_valid = [torch.zeros(1, **i32) for _ in range(self.DEPTH)]
        # pad key = V; libesr's wide sort: SMs idle in the exchange phases while the plan is built (2 GPUs: 265 vs 270 us/step)
        self.plans = [IndexPlan(n_slots, V + 1, self.dev, n_valid=self.n_valid[k], sort="wide") for k in range(self.DEPTH)]
        self.cplans = [IndexPlan(n_slots, V_max + n_slots, self.dev, n_valid=self.n_valid[k]) for k in range(self.DEPTH)]
        for cp, pl in zip(self.cplans, self.plans):
            cp.s.n_slots = n_slots
            cp.s.perm, cp.s.useg, cp.s.seg_off, cp.s.n_uniq = pl.s.perm, pl.s.useg, pl.s.seg_off, pl.s.n_uniq
        self.scratch = torch.empty(n_slots, **i32)
        if row_blocks is None:
            # as GloveTrainer: the persistent row pass leaves 1/9 of its CTA slots to the side stream (pair routing + plan of
            # the next batch: ~190 us of kernels that otherwise only run in the gaps of the main chain)
            sm = C.c_int(0)
            L.check(L.lib().esr_device_info(C.byref(sm), None, None), "esr_device_info")
            row_blocks = max(1, (2 * sm.value * 8) // 9)
        self.step_fn = GloveStep(self.unified, B_cap, lr=lr, bias_mode=bias_mode, chunk=chunk, emit_grads=True,
                                 B_global=B * n, dE=self.inbox_dE, db=self.inbox_db, impl=impl, row_blocks=row_blocks)
        cfg = self.step_fn.cfg
        cfg.emit_map = L.ptr(self.emit_map)
        cfg.emit_peers_dE = C.cast(self.p_inbox_dE, C.c_void_p)
        cfg.emit_peers_db = C.cast(self.p_inbox_db, C.c_void_p)
        cfg.n_emit_peers = n
        # the finish kernel logs every step's (global) loss itself: device log + pinned host mirror (a 4-byte posted PCIe
        # write per step) -- no copy node between the step graphs (a D2H copy there cost 30 us per step at 2 GPUs)
        self.loss_log = torch.zeros(4096, dtype=torch.float32, device=self.dev)
        self.loss_step = torch.zeros(1, **i32)
        self.loss_host = torch.zeros(4096, dtype=torch.float32).pin_memory()
        cfg.loss_log, cfg.loss_step, cfg.loss_log_len = L.ptr(self.loss_log), L.ptr(self.loss_step), 4096
        cfg.loss_host = self.loss_host.data_ptr()
        self.recv_cap = self.inbox_cap
        self.recv_ids = torch.empty(self.recv_cap, **i32)
        self.src_meta = torch.zeros(3 * 8 + 4, **i32)
        self.map_stride = V_max
        self.slot_map = torch.full((n, V_max), -1, **i32)
        self.desc = torch.empty(self.recv_cap * (n + 2), **i32)   # per entry: n source rows; then 8-byte owner records
        self.s_main = torch.cuda.Stream(self.dev)
        self.s_side = torch.cuda.Stream(self.dev)
        self.s_ids = torch.cuda.Stream(self.dev)
        self.s_copy = torch.cuda.Stream(self.dev)         # host batches: uploaded beside the previous step's routing + plan
        self.ev_up = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.ev_plan = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.ev_done = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.s_gather = torch.cuda.Stream(self.dev)
        self.ev_top, self.ev_ids = torch.cuda.Event(), torch.cuda.Event()
        self.ev_gather, self.ev_rows, self.ev_bias = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._keep = [None] * self.DEPTH
        self.st_ids = [torch.ones(2 * B, **i32) for _ in range(self.DEPTH)]
        self.st_counts = [torch.ones(B, dtype=torch.float32, device=self.dev) for _ in range(self.DEPTH)]
        self.use_graphs = bool(graphs)
        self.g_plan = [None] * self.DEPTH
        self.g_step = [None] * self.DEPTH
        self.t = 0
        self.loss = None
        torch.cuda.current_stream().synchronize()
        self._hdls[0].barrier()

    # -- synchronisation: libesr kernels over symmetric memory (one NVLink round trip, CUDA-graph capturable) ---------
    def _sync(self, view=None, side=0):
        _, p, seq = self.syncs[int(side)]
        buf = L.ptr(view) if view is not None else None
        L.check(L.lib().esr_peer_allreduce_f32(p, self.n, self.rank, buf, buf, view.numel() if view is not None else 0,
                                               L.ptr(seq), L.stream_ptr()), "esr_peer_allreduce_f32")

    def barrier(self):
        self._sync()

    # -- the two halves of a step; every call enqueues on the CURRENT stream ------------------------------------
    def _plan_body(self, k):
        """Ids only (side stream): route my pairs to their owners, collect the pairs I own, index plan, owner routing of
        the unique rows (published for the peers), the plan re-expressed in unique-row indices."""
        lib, n = L.lib(), self.n
        pin, pub, plan, cplan = self.pin[k], self.pub[k], self.plans[k], self.cplans[k]
        sp = L.stream_ptr()
        L.check(lib.esr_peer_route_pairs_i32(L.ptr(self.st_ids[k]), L.ptr(self.st_counts[k]), self.B, n, self.rank,
                                             pin["p_rec"], pin["p_counts"], L.ptr(self.my_counts),
                                             L.ptr(self.route_ws), self.route_ws.numel(), sp), "esr_peer_route_pairs_i32")
        self._sync(side=True)                                   # every source's pairs for me have landed
        L.check(lib.esr_peer_collect_pairs_i32(L.ptr(pin["rec"]), L.ptr(pin["counts"]), n, self.B,
                                               self.B_cap, self.V, L.ptr(self.keys[k]), L.ptr(self.cnt_l[k]),
                                               L.ptr(self.n_valid[k]), L.ptr(self.err), sp), "esr_peer_collect_pairs_i32")
        plan.build(self.keys[k])
        cplan.s.n_slots = plan.n_slots
        # owner routing of the unique rows (published for the peers) and the plan in unified-table addresses: they need the
        # plan, not the table.  (With both halves of the step truly overlapped the main chain is the longer one, so they
        # sit here; while the halves were serialised by mistake their placement made no difference.)
        self.ops.route_plan(plan.uniq, plan.n_uniq, n, out=(pub["order"], pub["send_local"], pub["counts"], pub["inv_order"]))
        L.check(lib.esr_plan_compact_owner_i32(C.byref(plan.s), n, self.rank, self.V_max, L.ptr(cplan.sorted_keys),
                                               L.ptr(cplan.partner), L.ptr(self.scratch), sp), "esr_plan_compact_owner_i32")

    def _ids_body(self, k, sp):
        lib, n = L.lib(), self.n
        plan, pub = self.plans[k], self.pub[k]
        L.check(lib.esr_peer_pull_ids_i32(pub["p_counts"], pub["p_send_local"], n, self.rank, self.recv_cap,
                                          L.ptr(self.recv_ids), L.ptr(self.src_meta), L.ptr(self.slot_map), self.map_stride,
                                          sp), "esr_peer_pull_ids_i32")
        L.check(lib.esr_peer_resolve_i32(n, L.ptr(self.recv_ids), L.ptr(self.src_meta), L.ptr(self.slot_map),
                                         self.map_stride, L.ptr(self.desc), self.recv_cap, sp), "esr_peer_resolve_i32")
        L.check(lib.esr_peer_emit_plan_i32(pub["p_counts"], n, self.rank, L.ptr(plan.uniq), L.ptr(plan.n_uniq),
                                           plan.capacity, L.ptr(pub["inv_order"]), self.inbox_cap, L.ptr(self.emit_map),
                                           L.ptr(self.err), sp), "esr_peer_emit_plan_i32")

    def _step_body(self, k):
        """Main-stream half of a step.  After the top barrier three things run side by side -- the NVLink fetch of the remote
        rows (gather stream), the owner-side id bookkeeping (ids stream), and bias fetch -> prep -> all-reduce (main) -- and
        after the row pass the owner merge of the embedding rows starts behind the S1/S2 all-reduce (which is also the
        "every rank's gradient rows have landed" barrier) while finish -> barrier -> bias merge run on the ids stream."""
        lib, n = L.lib(), self.n
        plan, cplan, st, pub = self.plans[k], self.cplans[k], self.step_fn, self.pub[k]
        main = torch.cuda.current_stream(self.dev)
        sp = L.stream_ptr()

        def gather(parts, stream_ptr):
            L.check(lib.esr_peer_gather_remote_f32(self.p_rows, self.p_bias, n, self.rank, L.ptr(plan.uniq), L.ptr(pub["order"]),
                                                   L.ptr(pub["counts"]), plan.capacity, self.D, L.ptr(self.fetch_rows),
                                                   L.ptr(self.fetch_bias), parts, stream_ptr), "esr_peer_gather_remote_f32")

        def apply(parts, stream_ptr):
            L.check(lib.esr_peer_apply_parts_f32(C.byref(self.shard.struct()), L.ptr(self.inbox_dE), L.ptr(self.inbox_db), n,
                                                 L.ptr(self.recv_ids), L.ptr(self.src_meta), L.ptr(self.slot_map),
                                                 self.map_stride, L.ptr(self.desc), self.recv_cap, self.lr, 1e-7, parts,
                                                 stream_ptr), "esr_peer_apply_parts_f32")

        self.barrier()                      # every route plan of this step is published; every owner applied step t-1
        self.ev_top.record(main)
        self.s_ids.wait_event(self.ev_top)
        self.s_gather.wait_event(self.ev_top)
        with torch.cuda.stream(self.s_ids):
            self._ids_body(k, L.stream_ptr())
            self.ev_ids.record(self.s_ids)
        with torch.cuda.stream(self.s_gather):
            gather(1, L.stream_ptr())       # remote embedding rows -> fetch region (NVLink loads)
            self.ev_gather.record(self.s_gather)
        gather(2, sp)                       # remote biases (4 bytes per row): all that prep needs
        st.prep(cplan, self.cnt_l[k])
        self._sync(st.scalars[0:3])         # global sum(bs), sum(bs^2), S0
        main.wait_event(self.ev_ids)
        main.wait_event(self.ev_gather)
        st.rows(cplan)                      # gradient rows go straight to the owners' inboxes
        self._sync(st.scalars[3:5])         # global S1, S2 -- and: every rank's row pass is done, its gradient rows landed
        self.ev_rows.record(main)
        self.s_ids.wait_event(self.ev_rows)
        with torch.cuda.stream(self.s_ids):
            st.finish(cplan, stream=self.s_ids)                 # bias gradients -> the owners' inboxes
            self._sync(side=2)                                  # every rank's bias gradients have landed
            apply(2, L.stream_ptr())                            # owner bias merge + Adagrad, slot_map restored
            self.ev_bias.record(self.s_ids)
        apply(1, sp)                        # owner merge of the embedding rows in source order + Adagrad
        main.wait_event(self.ev_bias)

    def _capture(self):
        """Both halves per parity into CUDA graphs.  Collective: every rank captures the same sequence."""
        try:
            torch.cuda.synchronize(self.dev)
            cap = torch.cuda.Stream(self.dev)
            for k in range(self.DEPTH):
                gp = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gp, stream=cap, capture_error_mode="thread_local"):
                    self._plan_body(k)
                gs = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gs, stream=cap, capture_error_mode="thread_local"):
                    self._step_body(k)
                self.g_plan[k], self.g_step[k] = gp, gs
            torch.cuda.synchronize(self.dev)
        except Exception as e:  # pragma: no cover
            import warnings
            warnings.warn("owner-routed step: CUDA-graph capture failed (%s); staying on eager launches" % e)
            self.g_plan = [None] * self.DEPTH
            self.g_step = [None] * self.DEPTH
            self.use_graphs = False

    def load_dense(self, E, b):
        idx = torch.arange(self.rank, self.V, self.n)
        self.shard.rows0.copy_(torch.as_tensor(E)[idx].to(self.dev))
        self.shard.bias.copy_(torch.as_tensor(b).reshape(-1)[idx].to(self.dev))
        torch.cuda.current_stream().synchronize()
        self._hdls[0].barrier()

    def step(self, ids, counts):
        """ids: int32 (2, B_local) global rows (host pinned or device) -- this rank's share of the global batch, in the
        layout of wikipedia/cooccurrence_matrix.py:103-114; counts: f32 (B_local,).  Enqueues one step on the trainer's
        own streams; returns the GLOBAL loss as a device scalar that is valid once the trainer's main stream has reached
        it (``loss_value()``, ``read_loss_to()``, ``synchronize()``)."""
        k = self.t % self.DEPTH
        cur = torch.cuda.current_stream(self.dev)
        main, side = self.s_main, self.s_side
        if self.use_graphs and self.g_step[0] is None and self.t == 2 * self.DEPTH:
            self._capture()                    # after two eager steps per parity (lazy module / allocator state is warm)
        # The trainer runs on its OWN streams.  (Round 1 and the first round-2 version ran the main half on the caller's
        # stream and made the side stream wait for that stream "for the inputs" -- which also made the routing + plan of
        # batch t+1 wait for the whole step t: the two halves never overlapped, step = main + side.)
        on_device = torch.is_tensor(ids) and ids.is_cuda
        if on_device:
            side.wait_stream(cur)              # device inputs produced on the caller's stream
        side.wait_event(self.ev_done[k])       # step t-2 is done with parity k's buffers on this rank
        if not on_device:
            # host batch: its upload runs on a stream of its own, next to routing + plan of the previous step instead of in
            # front of this step's routing (measured at 2 GPUs: no difference, 300 us per e2e step either way -- the side
            # chain is not the critical path there; what cost 30 us per step was the loss copy between the step graphs)
            self.s_copy.wait_event(self.ev_done[k])
            with torch.cuda.stream(self.s_copy):
                self.st_ids[k].copy_(ids.reshape(-1), non_blocking=True)
                self.st_counts[k].copy_(counts, non_blocking=True)
                self.ev_up[k].record(self.s_copy)
            side.wait_event(self.ev_up[k])
        with torch.cuda.stream(side):
            if on_device:
                self.st_ids[k].copy_(ids.reshape(-1), non_blocking=True)
                self.st_counts[k].copy_(counts, non_blocking=True)
            self._keep[k] = (ids, counts)
            if self.g_plan[k] is not None:
                self.g_plan[k].replay()
            else:
                self._plan_body(k)
            self.ev_plan[k].record(side)
        main.wait_event(self.ev_plan[k])
        with torch.cuda.stream(main):
            if self.g_step[k] is not None:
                self.g_step[k].replay()
            else:
                self._step_body(k)
            self.ev_done[k].record(main)
        self.loss = self.loss_log[self.t % self.loss_log.numel()]     # written by this step's finish kernel
        self.t += 1
        return self.loss

    def loss_value(self):
        """The last step's global loss as a Python float (waits for the trainer's main stream)."""
        self.s_main.synchronize()
        return float(self.loss.item())

    def host_loss(self, step):
        """Loss of step ``step`` (0-based) from the pinned host mirror the finish kernel writes -- valid once that step
        has run (``synchronize()``, or simply later: the mirror keeps the last 4096 steps)."""
        return float(self.loss_host[step % self.loss_host.numel()])

    def read_loss_to(self, pinned_slot):
        """Asynchronous device->host copy of the last step's loss into a pinned 1-element tensor, ordered on the trainer's
        main stream.  (Every step's loss also arrives in ``loss_host`` without any copy: ``host_loss``.)"""
        with torch.cuda.stream(self.s_main):
            pinned_slot.copy_(self.loss.reshape(1), non_blocking=True)

    def check(self):
        """Synchronise and raise if a step overflowed the pair capacity (bit 1) or a gradient inbox (bit 0)."""
        torch.cuda.synchronize(self.dev)
        e = int(self.err.item())
        if e:
            raise RuntimeError("OwnerRoutedGloveTrainer: %s; training state is invalid" % " and ".join(
                m for b, m in ((2, "an owner received more than pair_cap=%d pairs in a step" % self.B_cap),
                               (1, "a gradient inbox overflowed (inbox_cap=%d rows)" % self.inbox_cap)) if e & b))

    def synchronize(self):
        self.check()

    def gather_dense(self):
        self.check()
        E = torch.zeros(self.V, self.D, device=self.dev)
        b = torch.zeros(self.V, device=self.dev)
        idx = torch.arange(self.rank, self.V, self.n, device=self.dev)
        E[idx] = self.shard.rows0
        b[idx] = self.shard.bias
        dist.all_reduce(E, group=self.group)
        dist.all_reduce(b, group=self.group)
        return E, b
