"""Host-side objects around the C ABI: the versioned embedding table, the per-batch index plan
and the fused GloVe step.  torch owns every device buffer and stream; all arithmetic happens in
``libesr.so`` (see include/esr.h).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L


def _dev(device):
    L.require_cuda()
    return torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())


class EmbeddingTable:
    """``nn.Embed`` parameter ``{'embedding': f32[V, D]}`` (wikipedia/models.py:16-19,
    spotify/models.py:30-31) plus its Adagrad slot, laid out for the batch-synchronous sparse update:
    two row buffers + a per-row version byte (include/esr.h, EsrTable).

    ``sparse=False`` keeps a single buffer (dense Adam / SGD-momentum parity modes, or the compact
    fetched-row table of the row-sharded path).
    """

    def __init__(self, V, D, device=None, sparse=True, with_bias=True, init_acc=0.1, adagrad=None):
        self.device = _dev(device)
        if D % 4:
            raise ValueError("D must be a multiple of 4 (16-byte rows)")
        self.V, self.D, self.sparse = int(V), int(D), bool(sparse)
        dev = self.device
        self.rows0 = torch.zeros(V, D, dtype=torch.float32, device=dev)
        self.rows1 = torch.zeros(V, D, dtype=torch.float32, device=dev) if sparse else None
        self.ver = torch.zeros(V, dtype=torch.uint8, device=dev) if sparse else None
        adagrad = sparse if adagrad is None else bool(adagrad)   # single-buffer shards still carry the Adagrad slot
        self.acc = torch.full((V, D), float(init_acc), dtype=torch.float32, device=dev) if adagrad else None
        self.bias = torch.zeros(V, dtype=torch.float32, device=dev) if with_bias else None
        self.bias_acc = (torch.full((V,), float(init_acc), dtype=torch.float32, device=dev)
                         if (with_bias and adagrad) else None)
        self._struct = None

    # -- construction ----------------------------------------------------------------------
    @classmethod
    def from_dense(cls, E, bias=None, device=None, sparse=True, init_acc=0.1, adagrad=None):
        E = torch.as_tensor(E)
        t = cls(E.shape[0], E.shape[1], device, sparse, bias is not None, init_acc, adagrad)
        t.rows0.copy_(E.to(torch.float32))
        if bias is not None:
            t.bias.copy_(torch.as_tensor(bias).to(torch.float32).reshape(-1))
        return t

    @classmethod
    def wrap(cls, rows, bias=None, acc=None, bias_acc=None):
        """Zero-copy single-buffer view over existing (V,D) / (V,) CUDA tensors (reference-style params)."""
        t = cls.__new__(cls)
        assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous() and rows.shape[1] % 4 == 0
        t.device = rows.device
        t.V, t.D, t.sparse = rows.shape[0], rows.shape[1], False
        t.rows0, t.rows1, t.ver = rows, None, None
        t.acc = acc
        t.bias = bias.reshape(-1) if bias is not None else None
        t.bias_acc = bias_acc.reshape(-1) if bias_acc is not None else None
        t._struct = None
        return t

    def struct(self) -> L.EsrTable:
        if self._struct is None:
            s = L.EsrTable()
            s.struct_size = C.sizeof(L.EsrTable)
            s.D, s.V = self.D, self.V
            s.rows[0] = L.ptr(self.rows0)
            s.rows[1] = L.ptr(self.rows1)
            s.ver = L.ptr(self.ver)
            s.acc = L.ptr(self.acc)
            s.bias = L.ptr(self.bias)
            s.bias_acc = L.ptr(self.bias_acc)
            self._struct = s
        return self._struct

    # -- reads -----------------------------------------------------------------------------
    def gather(self, ids, out=None):
        """``jnp.take(embedding, ids, axis=0)`` (nn.Embed.__call__)."""
        ids = ids.to(device=self.device, dtype=torch.int32).contiguous()
        n = ids.numel()
        if out is None:
            out = torch.empty(n, self.D, dtype=torch.float32, device=self.device)
        L.check(L.lib().esr_table_gather_f32(C.byref(self.struct()), L.ptr(ids), n, L.ptr(out), L.stream_ptr()),
                "esr_table_gather_f32")
        return out.view(*ids.shape, self.D)

    def dense(self):
        """The dense ``f32[V, D]`` array ``state.params[..]['embedding']`` holds in the reference."""
        if not self.sparse:
            return self.rows0
        out = torch.empty(self.V, self.D, dtype=torch.float32, device=self.device)
        L.check(L.lib().esr_table_export_f32(C.byref(self.struct()), L.ptr(out), L.stream_ptr()), "esr_table_export_f32")
        return out

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in
                   (self.rows0, self.rows1, self.ver, self.acc, self.bias, self.bias_acc) if t is not None)


class IndexPlan:
    """Device buffers + descriptor of one batch's index plan (EsrPlan)."""

    SORTS = {"auto": L.ESR_SORT_AUTO, "wide": L.ESR_SORT_WIDE, "library": L.ESR_SORT_LIBRARY}

    def __init__(self, n_slots, V, device=None, with_partner=True, n_valid=None, sort="auto"):
        """``n_valid``: optional device int32 scalar -- only the first ``n_valid`` SORTED slots are real (the row-sharded
        path pads its fixed-capacity slot array with a key larger than every row id; EsrPlan.n_valid).
        ``sort``: "auto" / "wide" = libesr's own radix sort (cub's single-tile kernel up to 6144 slots), "library" = cub
        (the measurement control); EsrPlan.sort_impl."""
        self.device = _dev(device)
        n = int(n_slots)
        self.n_slots = n
        self.capacity = n
        dev = self.device
        i32 = dict(dtype=torch.int32, device=dev)
        cap = max(n, 1)
        self.sorted_keys = torch.empty(cap, **i32)
        self.perm = torch.empty(cap, **i32)
        self.partner = torch.empty(cap, **i32) if with_partner else None
        self.useg = torch.empty(cap, **i32)
        self.uniq = torch.empty(cap, **i32)
        self.seg_off = torch.empty(cap + 1, **i32)
        self.n_uniq = torch.zeros(1, **i32)
        self.ws_bytes = int(L.lib().esr_plan_workspace_bytes(n))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.key_bits = max(1, int(math.ceil(math.log2(max(int(V), 2)))))
        s = L.EsrPlan()
        s.struct_size = C.sizeof(L.EsrPlan)
        s.key_bits = self.key_bits
        s.n_slots = n
        s.keys = None
        s.sorted_keys, s.perm = L.ptr(self.sorted_keys), L.ptr(self.perm)
        s.partner = L.ptr(self.partner)
        s.useg, s.uniq, s.seg_off, s.n_uniq = (L.ptr(self.useg), L.ptr(self.uniq), L.ptr(self.seg_off),
                                               L.ptr(self.n_uniq))
        self.n_valid = n_valid
        s.n_valid = L.ptr(n_valid) if n_valid is not None else None
        s.sort_impl = self.SORTS[sort]
        self.s = s
        self._keys = None

    def build(self, keys, stream=None):
        """keys: int32 device tensor (for GloVe the flat (2,B) batch).  Fewer keys than the capacity
        the plan was created with are allowed (owner-side plans of the sharded path vary per step)."""
        assert keys.dtype == torch.int32 and keys.is_cuda and keys.is_contiguous()
        assert keys.numel() <= self.capacity
        self._keys = keys  # keep alive until the next build
        self.n_slots = keys.numel()
        self.s.n_slots = self.n_slots
        self.s.keys = L.ptr(keys)
        L.check(L.lib().esr_plan_build_i32(C.byref(self.s), L.ptr(self.ws), self.ws_bytes, L.stream_ptr(stream)),
                "esr_plan_build_i32")
        return self

    def remap_ids(self, out=None, stream=None):
        if out is None:
            out = torch.empty(self.n_slots, dtype=torch.int32, device=self.device)
        L.check(L.lib().esr_plan_remap_ids_i32(C.byref(self.s), L.ptr(out), L.stream_ptr(stream)), "esr_plan_remap_ids_i32")
        return out

    def host_view(self):
        """(sorted_keys, perm, uniq, seg_off) as NumPy, trimmed to n_uniq (synchronises)."""
        U = int(self.n_uniq.item())
        n = self.n_slots
        return (self.sorted_keys[:n].cpu().numpy(), self.perm[:n].cpu().numpy(), self.uniq[:U].cpu().numpy(),
                self.seg_off[:U + 1].cpu().numpy())


class GloveStep:
    """apply_model + update_model of wikipedia/train_cooccurence.py:71-101 as three stream-ordered
    phases over an IndexPlan (include/esr.h: esr_glove_prep/rows/finish)."""

    def __init__(self, table: EmbeddingTable, B, lr=0.05, bias_mode="reference_broadcast", eps=1e-7,
                 x_max=100.0, alpha=0.75, chunk=0, emit_grads=False, B_global=None, impl="auto", dE=None, db=None,
                 row_blocks=0, variant=0):
        self.table = table
        self.B = int(B)
        dev = table.device
        cfg = L.EsrGloveCfg()
        cfg.struct_size = C.sizeof(L.EsrGloveCfg)
        cfg.bias_mode = L.BIAS_MODES[bias_mode]
        cfg.rows_mode = L.ROWS_EMIT_GRADS if emit_grads else L.ROWS_UPDATE
        if impl == "fifo":          # the group row pass with bulk-copy FIFO staging (A/B candidate for the default)
            impl, variant = "auto", 2
        if impl == "endfirst":      # A/B probe: the round-1 work order of the persistent row pass (cold end of the stream first)
            impl, variant = "auto", 7
        cfg.impl = impl if isinstance(impl, int) else {"auto": L.IMPL_AUTO, "ldg": L.IMPL_LDG, "tma": L.IMPL_TMA}[impl]
        cfg.B = self.B
        cfg.B_global = int(B_global if B_global is not None else B)
        cfg.lr, cfg.eps, cfg.x_max, cfg.alpha = lr, eps, x_max, alpha
        cfg.chunk = chunk
        cfg.row_blocks = int(row_blocks)
        cfg.reserved = int(variant)     # row-pass staging A/B: 0 cp.async (default), 1 registers, 2 bulk-copy FIFO
        self.cfg = cfg
        self.emit = bool(emit_grads)
        self.ws_bytes = int(L.lib().esr_glove_workspace_bytes(self.B, table.D, chunk))
        self.ws = torch.empty(max(self.ws_bytes, 256), dtype=torch.uint8, device=dev)
        self.scalars = torch.zeros(L.GLOVE_NSCAL, dtype=torch.float32, device=dev)
        n = max(2 * self.B, 1)
        # EMIT outputs; the caller may supply them (e.g. symmetric memory the owners read over NVLink)
        self.dE = (dE if dE is not None else torch.empty(n, table.D, dtype=torch.float32, device=dev)) if emit_grads else None
        self.db = (db if db is not None else torch.empty(n, dtype=torch.float32, device=dev)) if emit_grads else None

    def _args(self, plan):
        return C.byref(self.table.struct()), C.byref(plan.s), C.byref(self.cfg)

    def prep(self, plan, counts, stream=None):
        t, p, c = self._args(plan)
        L.check(L.lib().esr_glove_prep_f32(t, p, L.ptr(counts), c, L.ptr(self.scalars), L.ptr(self.ws), self.ws_bytes,
                                           L.stream_ptr(stream)), "esr_glove_prep_f32")

    def rows(self, plan, stream=None):
        t, p, c = self._args(plan)
        L.check(L.lib().esr_glove_rows_f32(t, p, c, L.ptr(self.scalars), L.ptr(self.dE), L.ptr(self.ws), self.ws_bytes,
                                           L.stream_ptr(stream)), "esr_glove_rows_f32")

    def rows_main(self, plan, stream=None):
        t, p, c = self._args(plan)
        L.check(L.lib().esr_glove_rows_main_f32(t, p, c, L.ptr(self.scalars), L.ptr(self.dE), L.ptr(self.ws),
                                                self.ws_bytes, L.stream_ptr(stream)), "esr_glove_rows_main_f32")

    def rows_combine(self, plan, stream=None):
        t, p, c = self._args(plan)
        L.check(L.lib().esr_glove_rows_combine_f32(t, p, c, L.ptr(self.scalars), L.ptr(self.dE), L.ptr(self.ws),
                                                   self.ws_bytes, L.stream_ptr(stream)), "esr_glove_rows_combine_f32")

    def finish(self, plan, stream=None):
        t, p, c = self._args(plan)
        L.check(L.lib().esr_glove_finish_f32(t, p, c, L.ptr(self.scalars), L.ptr(self.db), L.ptr(self.ws), self.ws_bytes,
                                             L.stream_ptr(stream)), "esr_glove_finish_f32")

    def run(self, plan, counts, stream=None):
        """All three phases.  Returns the device scalars block (loss at index SC_LOSS)."""
        assert counts.dtype == torch.float32 and counts.is_cuda and counts.numel() == self.B
        t, p, c = self._args(plan)
        L.check(L.lib().esr_glove_step_f32(t, p, L.ptr(counts), c, L.ptr(self.scalars), L.ptr(self.dE), L.ptr(self.db),
                                           L.ptr(self.ws), self.ws_bytes, L.stream_ptr(stream)), "esr_glove_step_f32")
        return self.scalars


def sparse_adagrad(table: EmbeddingTable, uniq, n_uniq, g, gb, lr, eps=1e-7, stream=None):
    cap = uniq.numel()
    L.check(L.lib().esr_sparse_adagrad_f32(C.byref(table.struct()), L.ptr(uniq), L.ptr(n_uniq), cap, L.ptr(g), L.ptr(gb),
                                           lr, eps, L.stream_ptr(stream)), "esr_sparse_adagrad_f32")


def scatter_rows(dst, uniq, n_uniq, g, accumulate=False, stream=None):
    D = dst.shape[1] if dst.dim() == 2 else 1
    L.check(L.lib().esr_scatter_rows_f32(L.ptr(dst), D, L.ptr(uniq), L.ptr(n_uniq), uniq.numel(), L.ptr(g),
                                         1 if accumulate else 0, L.stream_ptr(stream)), "esr_scatter_rows_f32")


def dense_adam(p, g, mu, nu, lr, count, b1=0.9, b2=0.999, eps=1e-8, stream=None):
    L.check(L.lib().esr_dense_adam_f32(L.ptr(p), L.ptr(g), L.ptr(mu), L.ptr(nu), p.numel(), lr, b1, b2, eps, int(count),
                                       L.stream_ptr(stream)), "esr_dense_adam_f32")


def dense_sgdm(p, g, trace, lr, momentum, stream=None):
    L.check(L.lib().esr_dense_sgdm_f32(L.ptr(p), L.ptr(g), L.ptr(trace), p.numel(), lr, momentum, L.stream_ptr(stream)),
            "esr_dense_sgdm_f32")


def check_ids(ids, V):
    """Debug validator: number of ids outside [0, V) (synchronises)."""
    ids = ids.contiguous()
    n_bad = torch.zeros(1, dtype=torch.int32, device=ids.device)
    L.check(L.lib().esr_check_ids_i32(L.ptr(ids), ids.numel(), int(V), L.ptr(n_bad), L.stream_ptr()), "esr_check_ids_i32")
    return int(n_bad.item())


def rowwise_dot(x, y):
    """jax.vmap(jnp.dot) over rows (wikipedia/models.py:35-36); sum(a*b, -1) (pinterest/models.py:67-72)."""
    x, y = x.contiguous(), y.contiguous()
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    L.check(L.lib().esr_rowwise_dot_f32(L.ptr(x), L.ptr(y), x.shape[0], x.shape[1], L.ptr(out), L.stream_ptr()),
            "esr_rowwise_dot_f32")
    return out


def score_all(table: EmbeddingTable, queries):
    """(V, T) scores of every table row against T query vectors (Glove.score_all)."""
    queries = queries.contiguous()
    T = queries.shape[0]
    out = torch.empty(table.V, T, dtype=torch.float32, device=table.device)
    L.check(L.lib().esr_score_all_f32(C.byref(table.struct()), L.ptr(queries), T, L.ptr(out), L.stream_ptr()),
            "esr_score_all_f32")
    return out


def sample_uniform(seed, step, n, hi, device=None, out=None):
    """n uniform int32 in [0, hi) on the device (counter-based, reproducible from (seed, step))."""
    dev = _dev(device)
    if out is None:
        out = torch.empty(int(n), dtype=torch.int32, device=dev)
    L.check(L.lib().esr_sample_uniform_i32(int(seed), int(step), int(n), int(hi), L.ptr(out), L.stream_ptr()),
            "esr_sample_uniform_i32")
    return out


def sort_cols(scores, k=None, descending=False, want_values=False):
    """Stable per-column ranking of a (V, T) score matrix -> int32 (k, T) row indices (and values).
    ``descending=False, k=None`` is ``jnp.argsort(scores, axis=0)`` (find_knn); ``descending=True`` is
    ``jax.lax.top_k`` per column."""
    assert scores.is_cuda and scores.dtype == torch.float32 and scores.dim() == 2 and scores.is_contiguous()
    V, T = scores.shape
    k = V if k is None else int(k)
    ws_bytes = int(L.lib().esr_sort_cols_workspace_bytes(V))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=scores.device)
    idx = torch.empty(k, T, dtype=torch.int32, device=scores.device)
    val = torch.empty(k, T, dtype=torch.float32, device=scores.device) if want_values else None
    L.check(L.lib().esr_sort_cols_f32(L.ptr(scores), V, T, 1 if descending else 0, k, L.ptr(idx), L.ptr(val), L.ptr(ws),
                                      ws_bytes, L.stream_ptr()), "esr_sort_cols_f32")
    return (idx, val) if want_values else idx


def topk_scan(rows_a, queries, k, *, rows_a1=None, ver=None, rows_b=None, idx_a=None, idx_b=None, mod_a=0,
              max_over_queries=False, ctx_a=None, ctx_b=None, boost=0.1, ties_high_index_first=False, stream=None):
    """Fused retrieval (``esr_topk_scan_f32``): one pass over the candidate rows, running top-k per list in shared memory --
    no (N, T) score matrix, no sort of N keys.  Returns ``(values, indices)`` of shape ``(lists, k)``, best first; ``lists``
    is T, or 1 with ``max_over_queries`` (the eval_step affinity of spotify/models.py:78-80).  See include/esr.h."""
    dev = rows_a.device
    queries = queries.contiguous()
    T = queries.shape[0]
    N = int(idx_a.numel() if idx_a is not None else (idx_b.numel() if idx_b is not None else rows_a.shape[0]))
    cfg = L.EsrTopkCfg()
    cfg.struct_size = C.sizeof(L.EsrTopkCfg)
    cfg.T = T
    cfg.rows_a, cfg.rows_a1, cfg.ver = L.ptr(rows_a), L.ptr(rows_a1), L.ptr(ver)
    cfg.rows_b, cfg.idx_a, cfg.idx_b = L.ptr(rows_b), L.ptr(idx_a), L.ptr(idx_b)
    cfg.queries, cfg.ctx_a, cfg.ctx_b = L.ptr(queries), L.ptr(ctx_a), L.ptr(ctx_b)
    cfg.N = N
    cfg.Da = rows_a.shape[1]
    cfg.Db = rows_b.shape[1] if rows_b is not None else 0
    cfg.mod_a = int(mod_a)
    cfg.max_over_queries = 1 if max_over_queries else 0
    cfg.n_ctx_a = int(ctx_a.numel()) if ctx_a is not None else 0
    cfg.n_ctx_b = int(ctx_b.numel()) if ctx_b is not None else 0
    cfg.boost = float(boost)
    cfg.k = int(k)
    cfg.ties_high_index_first = 1 if ties_high_index_first else 0
    assert queries.shape[1] == cfg.Da + cfg.Db
    lists = 1 if max_over_queries else T
    ws_bytes = int(L.lib().esr_topk_workspace_bytes(N, cfg.Da + cfg.Db, T, cfg.max_over_queries, cfg.k))
    if ws_bytes == 0:
        raise L.EsrError("esr_topk_scan_f32: unsupported shape (N=%d, D=%d, T=%d, k=%d)" % (N, cfg.Da + cfg.Db, T, k))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    idx = torch.empty(lists, int(k), dtype=torch.int32, device=dev)
    val = torch.empty(lists, int(k), dtype=torch.float32, device=dev)
    L.check(L.lib().esr_topk_scan_f32(C.byref(cfg), L.ptr(idx), L.ptr(val), L.ptr(ws), ws_bytes, L.stream_ptr(stream)),
            "esr_topk_scan_f32")
    return val, idx


def table_topk(table: EmbeddingTable, queries, k, ties_high_index_first=False):
    """Top-k rows of a table per query vector: ``(values (T,k), indices (T,k))`` -- dump_knn / find_top_k without the
    (V,T) score matrix."""
    t = table
    return topk_scan(t.rows0, queries, k, rows_a1=t.rows1, ver=t.ver, ties_high_index_first=ties_high_index_first)


def top_k(scores_1d, k):
    """``jax.lax.top_k(scores, k)`` of a 1-D score vector -> (values, indices)."""
    idx, val = sort_cols(scores_1d.reshape(-1, 1).contiguous(), k, True, True)
    return val[:, 0], idx[:, 0]


class InBatchScorer:
    """B x B in-batch-negative scoring + loss + gradients on the tensor cores
    (``esr_inbatch_fwd_bwd_bf16``; csrc/inbatch_scores.cu).  Generalises the triplet scoring of
    pinterest/models.py:67-72 + pinterest/train_shop_the_look.py:99-104 to in-batch negatives
    (BASELINE.json configs[2], configs[3]).  The positive of query ``i`` is item ``i + diag_off``."""

    def __init__(self, Bq, D, Bk=None, loss="hinge", diag_off=0, margin=1.0, scale=1.0, b_norm=None, splits=0,
                 chunk_rows=0, device=None):
        self.device = _dev(device)
        cfg = L.EsrInbatchCfg()
        cfg.struct_size = C.sizeof(L.EsrInbatchCfg)
        cfg.loss_kind = {"hinge": L.LOSS_HINGE, "softmax": L.LOSS_SOFTMAX}[loss]
        cfg.Bq, cfg.Bk = int(Bq), int(Bk if Bk is not None else Bq)
        cfg.diag_off, cfg.D, cfg.splits = int(diag_off), int(D), int(splits)
        cfg.margin, cfg.scale = float(margin), float(scale)
        cfg.b_norm = float(b_norm if b_norm is not None else Bq)
        cfg.chunk_rows = int(chunk_rows)
        self.cfg = cfg
        self.Bq, self.Bk, self.D = cfg.Bq, cfg.Bk, cfg.D
        self.ws_bytes = int(L.lib().esr_inbatch_workspace_bytes(C.byref(cfg)))
        if self.ws_bytes == 0:
            raise L.EsrError("esr_inbatch_workspace_bytes: unsupported shape (D must be 64, 128, 192 or 256)")
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.dQ = torch.empty(self.Bq, self.D, dtype=torch.float32, device=self.device)
        self.dK = torch.empty(self.Bk, self.D, dtype=torch.float32, device=self.device)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)

    def run(self, Q, K, stream=None):
        """Q (Bq, D), K (Bk, D) fp32 CUDA.  Returns (loss[1], dQ, dK) device tensors (reused across calls)."""
        assert Q.is_cuda and K.is_cuda and Q.dtype == torch.float32 and K.dtype == torch.float32
        assert Q.is_contiguous() and K.is_contiguous() and tuple(Q.shape) == (self.Bq, self.D) and tuple(K.shape) == (self.Bk, self.D)
        L.check(L.lib().esr_inbatch_fwd_bwd_bf16(L.ptr(Q), L.ptr(K), C.byref(self.cfg), L.ptr(self.dQ), L.ptr(self.dK),
                                                 L.ptr(self.loss), L.ptr(self.ws), self.ws_bytes, L.stream_ptr(stream)),
                "esr_inbatch_fwd_bwd_bf16")
        return self.loss, self.dQ, self.dK

    def debug_views(self):
        """(G bf16 [rows of the LAST chunk, Bk], diag f32 [Bq], cnt f32 [R, Bq], lse f32 [Bq] (natural log)) out of
        the workspace (tests; pass chunk_rows >= Bq to keep the whole dL/dS matrix in one chunk)."""
        o = (C.c_int64 * 10)()
        L.check(L.lib().esr_inbatch_ws_layout(C.byref(self.cfg), o), "esr_inbatch_ws_layout")
        g_off, ldG, d_off, c_off, R, l_off, n_chunks, _, Bc = (int(o[k]) for k in range(9))
        rows = self.Bq - (n_chunks - 1) * Bc
        G = self.ws[g_off:g_off + rows * ldG * 2].view(torch.bfloat16).view(rows, ldG)[:, :self.Bk]
        diag = self.ws[d_off:d_off + 4 * self.Bq].view(torch.float32)
        cnt = self.ws[c_off:c_off + 4 * R * self.Bq].view(torch.float32).view(R, self.Bq)
        lse = self.ws[l_off:l_off + 4 * self.Bq].view(torch.float32) * 0.6931471805599453
        return G, diag, cnt, lse

    def plan(self):
        o = (C.c_int64 * 10)()
        L.check(L.lib().esr_inbatch_ws_layout(C.byref(self.cfg), o), "esr_inbatch_ws_layout")
        return dict(n_chunks=int(o[6]), chunk_rows=int(o[8]), R=int(o[4]), Sq=int(o[9]) // 100, Sk=int(o[9]) % 100)
