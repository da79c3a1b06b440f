"""Stream/graph orchestration of the fused GloVe step: the host loop of
``train_epoch`` (wikipedia/train_cooccurence.py:103-112) without a tracing compiler.

Per step the reference does ``next(train_it)`` -> ``apply_model`` -> ``update_model``.  Here a step is
two CUDA graphs on two streams:

* side stream : stage the batch (H2D from pinned memory, or D2D) and build its index plan
                (sort / unique / segments -- depends only on the ids),
* main stream : prep -> rows (+combine) -> finish on the table.

Plans and staging buffers are double-buffered, so the plan of batch t+1 is built while batch t
trains; events order the two streams.  Everything a step launches is inside the timed region of
bench.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .engine import EmbeddingTable, GloveStep, IndexPlan


class GloveTrainer:
    def __init__(self, table: EmbeddingTable, B, lr=0.05, bias_mode="reference_broadcast", chunk=0, impl="auto",
                 graphs=True, depth=2, loss_log=4096, row_blocks=None, priorities=False):
        L.require_cuda()
        if row_blocks is None:
            # The persistent row pass fills every SM with 2 CTAs, which leaves no registers for the plan
            # kernels of batch t+1 on the side stream (measured on B200: 170 us/step, plan serialised behind
            # the row pass).  Leaving ~11 % of the CTA slots free lets them co-run: 155-161 us/step.
            sm = C.c_int(0)
            L.check(L.lib().esr_device_info(C.byref(sm), None, None), "esr_device_info")
            row_blocks = max(1, (2 * sm.value * 8) // 9)
        self.table = table
        self.B = int(B)
        self.dev = table.device
        self.depth = int(depth)
        self.step_fn = GloveStep(table, B, lr=lr, bias_mode=bias_mode, chunk=chunk, impl=impl, row_blocks=row_blocks)
        self.plans = [IndexPlan(2 * self.B, table.V, self.dev) for _ in range(self.depth)]
        self.ids = [torch.zeros(2 * self.B, dtype=torch.int32, device=self.dev) for _ in range(self.depth)]
        self.counts = [torch.ones(self.B, dtype=torch.float32, device=self.dev) for _ in range(self.depth)]
        # priorities (EXPERIMENTAL, off): the step's short kernels (prep / combine / finish) on a high-priority stream, so
        # their CTAs are placed ahead of the pending radix-sort CTAs of the next batch's plan instead of queueing behind them
        self.s_main = torch.cuda.Stream(self.dev, priority=-1 if priorities else 0)
        self.s_side = torch.cuda.Stream(self.dev)
        self.ev_plan = [torch.cuda.Event() for _ in range(self.depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(self.depth)]
        self.loss_log = torch.zeros(loss_log, dtype=torch.float32, device=self.dev)
        self.loss_host = torch.zeros(loss_log, dtype=torch.float32).pin_memory()
        self.t = 0
        self.kernels_per_step = None
        self.g_plan = [None] * self.depth
        self.g_step = [None] * self.depth
        self.use_graphs = bool(graphs)
        if self.use_graphs:
            self._capture()

    # kernels one step launches (libesr only; the radix sort is cub code compiled into libesr)
    LAUNCHES_PLAN = 8    # iota, cub histogram + <=4 onesweep passes (key_bits<=32), head count, scan, head write
    LAUNCHES_STEP = 4    # prep (+ fused reduction), rows, combine, finish

    def _plan_body(self, k):
        self.plans[k].build(self.ids[k])

    def _step_body(self, k):
        self.step_fn.run(self.plans[k], self.counts[k])

    def _capture(self):
        # warm up once outside capture (function attributes, lazy module load), on valid ids
        for k in range(self.depth):
            self.ids[k].zero_()
            self.ids[k][: self.B] = 1
        snap = self._snapshot()
        torch.cuda.synchronize(self.dev)
        for k in range(self.depth):
            with torch.cuda.stream(self.s_side):
                self._plan_body(k)
            self.s_side.synchronize()
            with torch.cuda.stream(self.s_main):
                self._step_body(k)
            self.s_main.synchronize()
        for k in range(self.depth):
            gp = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gp, stream=self.s_side):
                self._plan_body(k)
            gs = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gs, stream=self.s_main):
                self._step_body(k)
            self.g_plan[k], self.g_step[k] = gp, gs
        self._restore(snap)
        torch.cuda.synchronize(self.dev)

    def _snapshot(self):
        """The warm-up steps touch rows 0 and 1: save and restore them so capture leaves no trace."""
        t = self.table
        rows = torch.tensor([0, min(1, t.V - 1)], device=self.dev)
        return rows, [x[rows].clone() if x is not None else None
                      for x in (t.rows0, t.rows1, t.ver, t.acc, t.bias, t.bias_acc)]

    def _restore(self, snap):
        rows, vals = snap
        t = self.table
        for x, v in zip((t.rows0, t.rows1, t.ver, t.acc, t.bias, t.bias_acc), vals):
            if x is not None:
                x[rows] = v

    def submit(self, ids, counts):
        """Enqueue one training step.  ``ids``: int32 (2,B) -- the batch layout of
        wikipedia/cooccurrence_matrix.py:103-114 -- and ``counts``: f32 (B,); pinned host tensors
        (copied H2D asynchronously) or device tensors.  Returns the step number."""
        k = self.t % self.depth
        side, main = self.s_side, self.s_main
        cur = torch.cuda.current_stream(self.dev)
        side.wait_stream(cur)                 # inputs produced on the caller's stream
        side.wait_event(self.ev_done[k])      # staging/plan buffers k are free again
        with torch.cuda.stream(side):
            self.ids[k].copy_(ids.reshape(-1), non_blocking=True)
            self.counts[k].copy_(counts, non_blocking=True)
            if self.use_graphs:
                self.g_plan[k].replay()
            else:
                self._plan_body(k)
            self.ev_plan[k].record(side)
        main.wait_event(self.ev_plan[k])
        with torch.cuda.stream(main):
            if self.use_graphs:
                self.g_step[k].replay()
            else:
                self._step_body(k)
            slot = self.t % self.loss_log.numel()
            self.loss_log[slot: slot + 1].copy_(self.step_fn.scalars[L.SC_LOSS: L.SC_LOSS + 1], non_blocking=True)
            self.ev_done[k].record(main)
        self.t += 1
        return self.t - 1

    def read_loss(self, step):
        """Device->host read of one step's loss (asynchronous copy into pinned memory on the main stream)."""
        slot = step % self.loss_log.numel()
        with torch.cuda.stream(self.s_main):
            self.loss_host[slot: slot + 1].copy_(self.loss_log[slot: slot + 1], non_blocking=True)
        return slot

    def synchronize(self):
        self.s_side.synchronize()
        self.s_main.synchronize()

    def losses(self, first, last):
        """Losses of steps [first, last) as NumPy (synchronises)."""
        self.synchronize()
        n = self.loss_log.numel()
        idx = torch.arange(first, last, device=self.dev) % n
        return self.loss_log[idx].cpu().numpy()

    @property
    def launches_per_step(self):
        return self.LAUNCHES_PLAN + self.LAUNCHES_STEP
