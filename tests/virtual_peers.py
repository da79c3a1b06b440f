"""TEST INFRASTRUCTURE: the peer-memory sharded GloVe step (esrecsys_b200/sharded.py, PeerShardedGloveTrainer) with
N VIRTUAL ranks on ONE GPU.

Every libesr entry point of the peer path takes an explicit array of per-rank device pointers (include/esr.h,
"Row-sharded table over NVLink peer memory").  On one device those pointers can all be local, so the whole multi-rank
choreography -- route plans, peer gather, id pull / resolve, emit map, the row pass scattering gradient rows into the
owners' inboxes, owner-side merge + Adagrad -- runs through the SAME kernels, phase by phase over the virtual ranks
on one stream (stream order replaces the device barriers; the few-float all-reduces are summed in rank order).
A 1-GPU box can then check multi-rank parity (tests/test_gpu_virtual_peers.py) and develop the sharded step without
paying for N GPUs.  Mirrors PeerShardedGloveTrainer._plan_body / _step_body line by line; not a product path.
"""
from __future__ import annotations

import ctypes as C

import torch

from esrecsys_b200 import _lib as L
from esrecsys_b200.engine import EmbeddingTable, GloveStep, IndexPlan
from esrecsys_b200.sharded import LibesrOps, pair_capacity, shard_rows


def _ptr_array(tensors):
    return (C.c_void_p * 8)(*[int(t.data_ptr()) for t in tensors])


class _Rank:
    pass


class VirtualPeerGlove:
    def __init__(self, V, D, B_local, n_ranks, lr=0.05, bias_mode="reference_broadcast", device="cuda", B_cap=None):
        """``B_cap``: owner-computes routing (VirtualOwnerRoutedGlove) -- the step runs over a padded slot array of capacity
        2 * B_cap whose real length is only known on the device (EsrPlan.n_valid)."""
        L.require_cuda()
        assert 1 <= n_ranks <= 8
        self.V, self.D, self.B, self.n, self.lr = int(V), int(D), int(B_local), int(n_ranks), float(lr)
        self.dev = torch.device(device)
        self.B_cap = int(B_cap) if B_cap is not None else None
        n, n_slots = self.n, 2 * (self.B_cap if B_cap is not None else self.B)
        i32 = dict(dtype=torch.int32, device=self.dev)
        V_max = shard_rows(V, 0, n)
        self.inbox_cap = n_slots * n
        self.ops = LibesrOps(self.dev)
        self.ranks = []
        for r in range(n):
            k = _Rank()
            V_loc = shard_rows(V, r, n)
            # owner-routed: [shard ; fetch region] in one allocation (local rows are addressed in place, never copied)
            extra = n_slots if B_cap is not None else 0
            k.rows = torch.zeros(V_max + extra, D, device=self.dev)
            k.bias = torch.zeros(V_max + extra, device=self.dev)
            k.shard = EmbeddingTable.wrap(k.rows[:V_loc], bias=k.bias[:V_loc], acc=torch.full((V_loc, D), 0.1, device=self.dev),
                                          bias_acc=torch.full((V_loc,), 0.1, device=self.dev))
            k.counts = torch.zeros(16, **i32)
            k.send_local = torch.zeros(n_slots, **i32)
            k.order = torch.zeros(n_slots, **i32)
            k.inv_order = torch.zeros(n_slots, **i32)
            k.inbox_dE = torch.zeros(self.inbox_cap, D, device=self.dev)
            k.inbox_db = torch.zeros(self.inbox_cap, device=self.dev)
            k.emit_map = torch.zeros(n_slots, **i32)
            k.err = torch.zeros(1, **i32)
            k.n_valid = torch.zeros(1, **i32) if B_cap is not None else None
            k.plan = IndexPlan(n_slots, V + (1 if B_cap is not None else 0), self.dev, n_valid=k.n_valid)   # pad key = V
            k.compact = (EmbeddingTable.wrap(k.rows, bias=k.bias) if B_cap is not None
                         else EmbeddingTable(n_slots, D, self.dev, sparse=False, adagrad=False))
            k.cplan = IndexPlan(n_slots, V_max + n_slots, self.dev, n_valid=k.n_valid)
            cs, ps = k.cplan.s, k.plan.s
            cs.n_slots = n_slots
            cs.perm, cs.useg, cs.seg_off, cs.n_uniq = ps.perm, ps.useg, ps.seg_off, ps.n_uniq
            k.scratch = torch.empty(n_slots, **i32)
            k.recv_ids = torch.zeros(self.inbox_cap, **i32)
            k.src_meta = torch.zeros(3 * 8 + 4, **i32)
            k.slot_map = torch.full((n, V_max), -1, **i32)
            k.desc = torch.zeros(self.inbox_cap * (n + 2), **i32)   # per entry: n source rows; then 8-byte owner records
            self.ranks.append(k)
        self.map_stride = V_max
        self.p_rows = _ptr_array([k.rows for k in self.ranks])
        self.p_bias = _ptr_array([k.bias for k in self.ranks])
        self.p_counts = _ptr_array([k.counts for k in self.ranks])
        self.p_send_local = _ptr_array([k.send_local for k in self.ranks])
        self.p_inbox_dE = _ptr_array([k.inbox_dE for k in self.ranks])
        self.p_inbox_db = _ptr_array([k.inbox_db for k in self.ranks])
        for k in self.ranks:
            k.step_fn = GloveStep(k.compact, self.B_cap if B_cap is not None else self.B, lr=lr, bias_mode=bias_mode,
                                  emit_grads=True, B_global=self.B * n, dE=k.inbox_dE, db=k.inbox_db)
            cfg = k.step_fn.cfg
            cfg.emit_map = L.ptr(k.emit_map)
            cfg.emit_peers_dE = C.cast(self.p_inbox_dE, C.c_void_p)
            cfg.emit_peers_db = C.cast(self.p_inbox_db, C.c_void_p)
            cfg.n_emit_peers = n
        self.loss = None

    # -- table in / out ------------------------------------------------------------------------------------------
    def load_dense(self, E, b):
        E, b = torch.as_tensor(E), torch.as_tensor(b).reshape(-1)
        for r, k in enumerate(self.ranks):
            idx = torch.arange(r, self.V, self.n)
            k.shard.rows0.copy_(E[idx].to(self.dev))
            k.shard.bias.copy_(b[idx].to(self.dev))

    def gather_dense(self):
        E = torch.zeros(self.V, self.D, device=self.dev)
        b = torch.zeros(self.V, device=self.dev)
        for r, k in enumerate(self.ranks):
            idx = torch.arange(r, self.V, self.n, device=self.dev)
            E[idx] = k.shard.rows0
            b[idx] = k.shard.bias
        return E, b

    # -- phases (each loops over the virtual ranks; one stream, so a finished loop is a passed barrier) ---------------
    def plan_phase(self, ids):
        lib = L.lib()
        for r, k in enumerate(self.ranks):
            k.plan.build(ids[r].to(self.dev).reshape(-1).contiguous())
            self.ops.route_plan(k.plan.uniq, k.plan.n_uniq, self.n, out=(k.order, k.send_local, k.counts, k.inv_order))
            k.cplan.s.n_slots = k.plan.n_slots
            L.check(lib.esr_plan_compact_i32(C.byref(k.plan.s), L.ptr(k.cplan.sorted_keys), L.ptr(k.cplan.partner),
                                             L.ptr(k.cplan.uniq), L.ptr(k.scratch), L.stream_ptr()), "esr_plan_compact_i32")

    def _all_reduce(self, lo, hi):
        tot = self.ranks[0].step_fn.scalars[lo:hi].clone()
        for k in self.ranks[1:]:
            tot += k.step_fn.scalars[lo:hi]                      # rank order, like esr_peer_allreduce_f32
        for k in self.ranks:
            k.step_fn.scalars[lo:hi] = tot

    def fetch_phase(self, counts):
        lib, n, sp = L.lib(), self.n, L.stream_ptr()
        for r, k in enumerate(self.ranks):
            L.check(lib.esr_peer_gather_f32(self.p_rows, self.p_bias, n, L.ptr(k.plan.uniq), L.ptr(k.plan.n_uniq),
                                            k.plan.capacity, self.D, L.ptr(k.compact.rows0), L.ptr(k.compact.bias), sp),
                    "esr_peer_gather_f32")
            k.counts_dev = counts[r].to(self.dev).contiguous()
            k.step_fn.prep(k.cplan, k.counts_dev)
        self._all_reduce(0, 3)

    def resolve_phase(self):
        lib, n, sp = L.lib(), self.n, L.stream_ptr()
        for r, k in enumerate(self.ranks):
            L.check(lib.esr_peer_pull_ids_i32(self.p_counts, self.p_send_local, n, r, self.inbox_cap, L.ptr(k.recv_ids),
                                              L.ptr(k.src_meta), L.ptr(k.slot_map), self.map_stride, sp), "esr_peer_pull_ids_i32")
            L.check(lib.esr_peer_resolve_i32(n, L.ptr(k.recv_ids), L.ptr(k.src_meta), L.ptr(k.slot_map), self.map_stride,
                                             L.ptr(k.desc), self.inbox_cap, sp), "esr_peer_resolve_i32")
            L.check(lib.esr_peer_emit_plan_i32(self.p_counts, n, r, L.ptr(k.plan.uniq), L.ptr(k.plan.n_uniq), k.plan.capacity,
                                               L.ptr(k.inv_order), self.inbox_cap, L.ptr(k.emit_map), L.ptr(k.err), sp),
                    "esr_peer_emit_plan_i32")

    def rows_phase(self):
        for k in self.ranks:
            k.step_fn.rows(k.cplan)                              # gradient rows -> the owners' inboxes
        self._all_reduce(3, 5)
        for k in self.ranks:
            k.step_fn.finish(k.cplan)

    def apply_phase(self):
        lib, n, sp = L.lib(), self.n, L.stream_ptr()
        for k in self.ranks:
            L.check(lib.esr_peer_apply_adagrad_f32(C.byref(k.shard.struct()), L.ptr(k.inbox_dE), L.ptr(k.inbox_db), n,
                                                   L.ptr(k.recv_ids), L.ptr(k.src_meta), L.ptr(k.slot_map), self.map_stride,
                                                   L.ptr(k.desc), self.inbox_cap, self.lr, 1e-7, sp), "esr_peer_apply_adagrad_f32")

    def step(self, ids, counts):
        """ids[r]: int32 (2, B_local) global rows of virtual rank r; counts[r]: f32 (B_local,).  Returns the GLOBAL loss."""
        self.plan_phase(ids)
        self.fetch_phase(counts)
        self.resolve_phase()
        self.rows_phase()
        self.apply_phase()
        self.loss = self.ranks[0].step_fn.scalars[L.SC_LOSS].clone()
        return self.loss


class VirtualOwnerRoutedGlove(VirtualPeerGlove):
    """OwnerRoutedGloveTrainer (esrecsys_b200/sharded.py) with N virtual ranks on one GPU: the pairs are first routed to
    the rank owning row i (esr_peer_route_pairs_i32 -> esr_peer_collect_pairs_i32), then the step runs exactly as in
    VirtualPeerGlove over each rank's RECEIVED pairs (padded capacity, device-side slot count)."""

    def __init__(self, V, D, B_local, n_ranks, lr=0.05, bias_mode="reference_broadcast", device="cuda", pair_cap=None):
        cap = int(pair_cap) if pair_cap is not None else pair_capacity(B_local, n_ranks)
        super().__init__(V, D, B_local, n_ranks, lr=lr, bias_mode=bias_mode, device=device, B_cap=cap)
        n, B = self.n, self.B
        i32 = dict(dtype=torch.int32, device=self.dev)
        for k in self.ranks:
            k.pin_rec = torch.zeros(n, B, 4, **i32)              # 16-byte records {i, j, count bits, 0} per source
            k.pin_counts = torch.zeros(16, **i32)
            k.my_counts = torch.zeros(16, **i32)
            k.keys = torch.full((2 * cap,), V, **i32)
            k.cnt_l = torch.zeros(cap, dtype=torch.float32, device=self.dev)
            k.route_ws = torch.empty(int(L.lib().esr_peer_route_pairs_workspace_bytes(B)), dtype=torch.uint8, device=self.dev)
        self.p_pin_rec = _ptr_array([k.pin_rec for k in self.ranks])
        self.p_pin_counts = _ptr_array([k.pin_counts for k in self.ranks])

    def route_phase(self, ids, counts):
        lib, n, sp = L.lib(), self.n, L.stream_ptr()
        for r, k in enumerate(self.ranks):
            k.ids_dev = ids[r].to(self.dev).reshape(-1).contiguous()
            k.cnt_dev = counts[r].to(self.dev).contiguous()
            L.check(lib.esr_peer_route_pairs_i32(L.ptr(k.ids_dev), L.ptr(k.cnt_dev), self.B, n, r, self.p_pin_rec,
                                                 self.p_pin_counts, L.ptr(k.my_counts), L.ptr(k.route_ws),
                                                 k.route_ws.numel(), sp), "esr_peer_route_pairs_i32")
        for k in self.ranks:                                      # (side-stream barrier here in the product)
            L.check(lib.esr_peer_collect_pairs_i32(L.ptr(k.pin_rec), L.ptr(k.pin_counts), n, self.B,
                                                   self.B_cap, self.V, L.ptr(k.keys), L.ptr(k.cnt_l), L.ptr(k.n_valid),
                                                   L.ptr(k.err), sp), "esr_peer_collect_pairs_i32")

    def plan_phase(self, ids=None):
        lib = L.lib()
        for k in self.ranks:
            k.plan.build(k.keys)
            self.ops.route_plan(k.plan.uniq, k.plan.n_uniq, self.n, out=(k.order, k.send_local, k.counts, k.inv_order))
            k.cplan.s.n_slots = k.plan.n_slots
        for r, k in enumerate(self.ranks):
            L.check(lib.esr_plan_compact_owner_i32(C.byref(k.plan.s), self.n, r, self.map_stride, L.ptr(k.cplan.sorted_keys),
                                                   L.ptr(k.cplan.partner), L.ptr(k.scratch), L.stream_ptr()),
                    "esr_plan_compact_owner_i32")

    def fetch_phase(self, counts):
        lib, n, sp = L.lib(), self.n, L.stream_ptr()
        V_max = self.map_stride
        for r, k in enumerate(self.ranks):
            for parts in (2, 1):                                 # biases, then rows: the product runs them on two streams
                L.check(lib.esr_peer_gather_remote_f32(self.p_rows, self.p_bias, n, r, L.ptr(k.plan.uniq), L.ptr(k.order),
                                                       L.ptr(k.counts), k.plan.capacity, self.D, L.ptr(k.rows[V_max:]),
                                                       L.ptr(k.bias[V_max:]), parts, sp), "esr_peer_gather_remote_f32")
            k.step_fn.prep(k.cplan, counts[r])
        self._all_reduce(0, 3)

    def apply_phase(self):
        lib, n, sp = L.lib(), self.n, L.stream_ptr()
        for parts in (1, 2):                                     # embedding rows, then biases (two streams in the product)
            for k in self.ranks:
                L.check(lib.esr_peer_apply_parts_f32(C.byref(k.shard.struct()), L.ptr(k.inbox_dE), L.ptr(k.inbox_db), n,
                                                     L.ptr(k.recv_ids), L.ptr(k.src_meta), L.ptr(k.slot_map), self.map_stride,
                                                     L.ptr(k.desc), self.inbox_cap, self.lr, 1e-7, parts, sp),
                        "esr_peer_apply_parts_f32")

    def step(self, ids, counts):
        self.route_phase(ids, counts)
        self.plan_phase()
        self.fetch_phase([k.cnt_l for k in self.ranks])
        self.resolve_phase()
        self.rows_phase()
        self.apply_phase()
        self.loss = self.ranks[0].step_fn.scalars[L.SC_LOSS].clone()
        return self.loss
