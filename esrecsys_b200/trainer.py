"""Stream/graph orchestration of the fused GloVe step: the host loop of
``train_epoch`` (wikipedia/train_cooccurence.py:103-112) without a tracing compiler.

Per step the reference does ``next(train_it)`` -> ``apply_model`` -> ``update_model``.  Here a step is
two CUDA graphs on three streams, driven by libesr's native pipeline object (csrc/pipeline.cu: ONE C call
per step -- the torch-level loop of round 1 spent 170-200 us of host time per step, more than the GPU needs):

* copy stream : stage the batch (H2D from pinned memory on a copy engine, or D2D),
* side stream : build its index plan (sort / unique / segments -- depends only on the ids),
* main stream : prep -> rows (+combine) -> finish on the table.

Staging buffers and plans are triple-buffered: batch t+2 uploads while the plan of batch t+1 is built and
batch t trains, so a step from pinned host memory costs what a step from device memory does; events order
the streams.  Everything a step launches is inside the timed region of bench.py.  ``graphs=False`` keeps an
eager torch-stream version of the same choreography (debugging).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .engine import EmbeddingTable, GloveStep, IndexPlan


class GloveTrainer:
    def __init__(self, table: EmbeddingTable, B, lr=0.05, bias_mode="reference_broadcast", chunk=0, impl="auto",
                 graphs=True, depth=3, loss_log=4096, row_blocks=None, priorities=False, validate_ids="first"):
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
        if not 1 <= self.depth <= 4:
            raise ValueError("depth must be 1..4")
        self.step_fn = GloveStep(table, B, lr=lr, bias_mode=bias_mode, chunk=chunk, impl=impl, row_blocks=row_blocks)
        self.loss_log = torch.zeros(loss_log, dtype=torch.float32, device=self.dev)
        self.loss_step = torch.zeros(1, dtype=torch.int32, device=self.dev)
        cfg = self.step_fn.cfg           # the finish kernel logs every step's loss itself (device-side slot counter)
        cfg.loss_log, cfg.loss_step, cfg.loss_log_len = L.ptr(self.loss_log), L.ptr(self.loss_step), int(loss_log)
        self.loss_host = torch.zeros(loss_log, dtype=torch.float32).pin_memory()
        cfg.loss_host = self.loss_host.data_ptr()      # pinned host mirror: every step's loss arrives as a 4-byte PCIe write
        self.plans = [IndexPlan(2 * self.B, table.V, self.dev) for _ in range(self.depth)]
        # staging: ids and counts of a parity share one allocation, so an adjacent host batch uploads in ONE copy
        self._stage = [torch.zeros(3 * self.B, dtype=torch.int32, device=self.dev) for _ in range(self.depth)]
        self.ids = [b[: 2 * self.B] for b in self._stage]
        self.counts = [b[2 * self.B:].view(torch.float32) for b in self._stage]
        for c in self.counts:
            c.fill_(1.0)
        self.t = 0
        # The plan sorts ceil(log2 V) key bits and the row pass uses ids as raw row offsets, so an id outside [0, V) would
        # corrupt memory where XLA clamps / drops it.  "first": the first batch is validated on the
        # device (no host sync; the flag is read at the first synchronize()); "always": every batch, with a host sync
        # per step (debug); "never": the caller guarantees the range.
        if validate_ids not in ("first", "always", "never"):
            raise ValueError("validate_ids must be 'first', 'always' or 'never'")
        self.validate_ids = validate_ids
        self.n_bad = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._bad_pending = False
        self._keep = []
        self.use_graphs = bool(graphs)
        self.pipe = None
        if self.use_graphs:
            # native runtime: it owns the three streams (priorities: the step's stream above the plan stream, experimental)
            h = C.c_void_p()
            with torch.cuda.device(self.dev):
                L.check(L.lib().esr_pipeline_create(self.depth, 1 if priorities else 0, C.byref(h)), "esr_pipeline_create")
            self.pipe = h
            sc, ss, sm_ = C.c_void_p(), C.c_void_p(), C.c_void_p()
            L.check(L.lib().esr_pipeline_streams(self.pipe, C.byref(sc), C.byref(ss), C.byref(sm_)), "esr_pipeline_streams")
            self.s_copy = torch.cuda.ExternalStream(sc.value, device=self.dev)
            self.s_side = torch.cuda.ExternalStream(ss.value, device=self.dev)
            self.s_main = torch.cuda.ExternalStream(sm_.value, device=self.dev)
            ids_p = (C.c_void_p * self.depth)(*[x.data_ptr() for x in self.ids])
            cnt_p = (C.c_void_p * self.depth)(*[x.data_ptr() for x in self.counts])
            L.check(L.lib().esr_pipeline_set_buffers(self.pipe, ids_p, cnt_p, self.ids[0].numel() * 4, self.counts[0].numel() * 4,
                                                     self.step_fn.scalars.data_ptr() + 4 * L.SC_LOSS, self.loss_log.data_ptr(),
                                                     self.loss_log.numel(), None),
                    "esr_pipeline_set_buffers")
            self.g_plan = [True] * self.depth      # the graphs live in the native object
            self.g_step = [True] * self.depth
            self._capture()
        else:
            self.s_main = torch.cuda.Stream(self.dev, priority=-1 if priorities else 0)
            self.s_side = torch.cuda.Stream(self.dev)
            self.s_copy = torch.cuda.Stream(self.dev)
            self.ev_copy = [torch.cuda.Event() for _ in range(self.depth)]
            self.ev_plan = [torch.cuda.Event() for _ in range(self.depth)]
            self.ev_done = [torch.cuda.Event() for _ in range(self.depth)]
            self.g_plan = [None] * self.depth
            self.g_step = [None] * self.depth

    def __del__(self):
        try:
            if getattr(self, "pipe", None):
                L.lib().esr_pipeline_destroy(self.pipe)
                self.pipe = None
        except Exception:  # pragma: no cover  (interpreter shutdown)
            pass

    # kernels one step launches (libesr only)
    LAUNCHES_STEP = 4    # prep (+ fused reduction and work-list clear), rows, combine, finish

    @property
    def LAUNCHES_PLAN(self):
        """digit histogram, one sort pass per 8 key bits (3 at V = 1M, 4 at V = 100M), head pass"""
        return 2 + (self.plans[0].key_bits + 7) // 8

    def _plan_body(self, k, stream=None):
        self.plans[k].build(self.ids[k], stream=stream)

    def _step_body(self, k, stream=None):
        self.step_fn.run(self.plans[k], self.counts[k], stream=stream)

    def _capture(self):
        # warm up once outside capture (function attributes, lazy module load), on valid ids
        lib = L.lib()
        for k in range(self.depth):
            self.ids[k].zero_()
            self.ids[k][: self.B] = 1
        snap = self._snapshot()
        torch.cuda.synchronize(self.dev)
        for k in range(self.depth):
            self._plan_body(k, self.s_side)
            self.s_side.synchronize()
            self._step_body(k, self.s_main)
            self.s_main.synchronize()
        for k in range(self.depth):
            L.check(lib.esr_pipeline_capture_begin(self.pipe, 0), "esr_pipeline_capture_begin")
            try:
                self._plan_body(k, self.s_side)
            finally:
                L.check(lib.esr_pipeline_capture_end(self.pipe, 0, k), "esr_pipeline_capture_end")
            L.check(lib.esr_pipeline_capture_begin(self.pipe, 1), "esr_pipeline_capture_begin")
            try:
                self._step_body(k, self.s_main)
            finally:
                L.check(lib.esr_pipeline_capture_end(self.pipe, 1, k), "esr_pipeline_capture_end")
        self._restore(snap)
        self.loss_step.zero_()            # the warm-up / capture steps advanced the device-side step counter
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

    def _validate(self, k, side_stream):
        """Count the ids of staging buffer k outside [0, V) on the side stream (after its plan: the buffer is intact)."""
        L.check(L.lib().esr_check_ids_i32(L.ptr(self.ids[k]), self.ids[k].numel(), self.table.V, L.ptr(self.n_bad),
                                          L.stream_ptr(side_stream)), "esr_check_ids_i32")
        if self.validate_ids == "always":
            side_stream.synchronize()
            self._raise_if_bad()
        else:
            self._bad_pending = True

    def submit(self, ids, counts, read_loss=False):
        """Enqueue one training step.  ``ids``: int32 (2,B) -- the batch layout of
        wikipedia/cooccurrence_matrix.py:103-114 -- and ``counts``: f32 (B,); pinned host tensors
        (copied H2D asynchronously) or device tensors.  ``read_loss``: also copy the step's loss to the pinned host
        log (``loss_host[step % len]``, asynchronous).  Returns the step number."""
        k = self.t % self.depth
        want_check = self.validate_ids == "always" or (self.validate_ids == "first" and self.t == 0)
        if self.pipe is not None:
            if ids.dtype != torch.int32 or counts.dtype != torch.float32 or ids.numel() != 2 * self.B or \
                    counts.numel() != self.B or not ids.is_contiguous() or not counts.is_contiguous():
                raise ValueError("submit: ids must be contiguous int32 (2,%d), counts contiguous float32 (%d,)" % (self.B, self.B))
            if not ids.is_cuda and not ids.is_pinned():
                ids, counts = ids.pin_memory(), counts.pin_memory()      # pageable memory would serialise the copy
            self._keep.append((ids, counts))                             # alive while the asynchronous copy may still read them
            if len(self._keep) > 2 * self.depth + 2:
                self._keep.pop(0)
            flags = (1 if read_loss else 0) | (2 if ids.is_cuda else 0)  # device inputs: order behind the caller's stream
            if ids.untyped_storage().data_ptr() == counts.untyped_storage().data_ptr():
                flags |= 4                                               # one allocation (pinned_batch()): single upload
            if not ids.is_cuda:
                flags |= 8                                               # pinned host memory: staged by a kernel over PCIe
            caller = torch.cuda.current_stream(self.dev).cuda_stream if ids.is_cuda else None
            L.check(L.lib().esr_pipeline_submit(self.pipe, ids.data_ptr(), counts.data_ptr(), caller, flags, None),
                    "esr_pipeline_submit")
            if want_check:
                self._validate(k, self.s_side)
            self.t += 1
            return self.t - 1
        # eager torch-stream version of the same choreography (graphs=False)
        side, main, copy = self.s_side, self.s_main, self.s_copy
        cur = torch.cuda.current_stream(self.dev)
        copy.wait_stream(cur)                 # inputs produced on the caller's stream
        copy.wait_event(self.ev_done[k])      # staging/plan buffers k are free again (step t - depth is done)
        with torch.cuda.stream(copy):
            self.ids[k].copy_(ids.reshape(-1), non_blocking=True)
            self.counts[k].copy_(counts, non_blocking=True)
            self.ev_copy[k].record(copy)
        side.wait_event(self.ev_copy[k])
        with torch.cuda.stream(side):
            self._plan_body(k)
            if want_check:
                self._validate(k, side)
            self.ev_plan[k].record(side)
        main.wait_event(self.ev_plan[k])
        with torch.cuda.stream(main):
            self._step_body(k)
            slot = self.t % self.loss_log.numel()
            if read_loss:
                self.loss_host[slot: slot + 1].copy_(self.loss_log[slot: slot + 1], non_blocking=True)
            self.ev_done[k].record(main)
        self.t += 1
        return self.t - 1

    def pinned_batch(self):
        """A pinned host batch in the layout of wikipedia/cooccurrence_matrix.py:103-114 -- ``(ids int32 (2,B), counts f32
        (B,))`` -- whose two arrays are adjacent in ONE pinned block, so ``submit`` uploads it with a single copy.  For
        loaders to fill in place."""
        buf = torch.empty(3 * self.B, dtype=torch.int32).pin_memory()
        return buf[: 2 * self.B].view(2, self.B), buf[2 * self.B:].view(torch.float32)

    def wait_staged(self, step):
        """Host-side wait (any thread) until the batch of step ``step`` has been staged on the device, i.e. its host
        memory may be rewritten: the hand-shake of ``wikipedia.input_pipeline.PinnedBatchLoader``."""
        import time
        while self.t <= step:                 # not submitted yet
            time.sleep(20e-6)
        if self.pipe is not None:
            L.check(L.lib().esr_pipeline_wait_staged(self.pipe, int(step)), "esr_pipeline_wait_staged")
        else:
            self.ev_copy[step % self.depth].synchronize()

    def read_loss(self, step):
        """Device->host read of one step's loss (asynchronous copy into pinned memory on the main stream); prefer
        ``submit(..., read_loss=True)``, which folds it into the step's call."""
        slot = step % self.loss_log.numel()
        with torch.cuda.stream(self.s_main):
            self.loss_host[slot: slot + 1].copy_(self.loss_log[slot: slot + 1], non_blocking=True)
        return slot

    def _raise_if_bad(self):
        bad = int(self.n_bad.item())
        if bad:
            raise ValueError("GloveTrainer: %d ids outside [0, %d) in a submitted batch; the table may be corrupted "
                             "(XLA would clamp the gather and drop the scatter -- wikipedia/models.py:31-34)" % (bad, self.table.V))

    def synchronize(self):
        if self.pipe is not None:
            L.check(L.lib().esr_pipeline_sync(self.pipe), "esr_pipeline_sync")
        else:
            self.s_copy.synchronize()
            self.s_side.synchronize()
            self.s_main.synchronize()
        if self._bad_pending:
            self._bad_pending = False
            self._raise_if_bad()

    def losses(self, first, last):
        """Losses of steps [first, last) as NumPy (synchronises)."""
        self.synchronize()
        n = self.loss_log.numel()
        idx = torch.arange(first, last, device=self.dev) % n
        return self.loss_log[idx].cpu().numpy()

    @property
    def launches_per_step(self):
        return self.LAUNCHES_PLAN + self.LAUNCHES_STEP
