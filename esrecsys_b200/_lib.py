"""ctypes binding of ``libesr.so`` -- the C ABI declared in ``include/esr.h``.

This is the same stub a maintainer of the reference would add to call the library from the
reference's Python trainers.  There is NO fallback: if the shared library
is missing, or a compute entry point is called without a CUDA device, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libesr.so")

ESR_OK, ESR_EINVAL, ESR_EWORKSPACE, ESR_ECUDA, ESR_ENOTSUP, ESR_ENOMEM = 0, -1, -2, -3, -4, -5
OPT_ADAGRAD, OPT_ADAM, OPT_SGDM = 0, 1, 2
BIAS_REFERENCE_BROADCAST, BIAS_PER_PAIR = 0, 1
ROWS_UPDATE, ROWS_EMIT_GRADS = 0, 1
IMPL_AUTO, IMPL_LDG, IMPL_TMA = 0, 1, 2
LOSS_HINGE, LOSS_SOFTMAX = 0, 1
SC_SUM_BS, SC_SUM_BS2, SC_S0, SC_S1, SC_S2, SC_LOSS, GLOVE_NSCAL = 0, 1, 2, 3, 4, 5, 8

BIAS_MODES = {"reference_broadcast": BIAS_REFERENCE_BROADCAST, "per_pair": BIAS_PER_PAIR}


class EsrError(RuntimeError):
    pass


class EsrTable(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("D", C.c_int32), ("V", C.c_int64),
                ("rows", C.c_void_p * 2), ("ver", C.c_void_p), ("acc", C.c_void_p),
                ("bias", C.c_void_p), ("bias_acc", C.c_void_p)]


class EsrPlan(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("key_bits", C.c_int32), ("n_slots", C.c_int64),
                ("keys", C.c_void_p), ("sorted_keys", C.c_void_p), ("perm", C.c_void_p),
                ("partner", C.c_void_p), ("useg", C.c_void_p), ("uniq", C.c_void_p),
                ("seg_off", C.c_void_p), ("n_uniq", C.c_void_p), ("n_valid", C.c_void_p),
                ("sort_impl", C.c_int32), ("reserved", C.c_int32)]


ESR_SORT_AUTO, ESR_SORT_WIDE, ESR_SORT_LIBRARY = 0, 1, 2


class EsrTopkCfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("T", C.c_int32), ("rows_a", C.c_void_p), ("rows_a1", C.c_void_p),
                ("ver", C.c_void_p), ("rows_b", C.c_void_p), ("idx_a", C.c_void_p), ("idx_b", C.c_void_p),
                ("queries", C.c_void_p), ("ctx_a", C.c_void_p), ("ctx_b", C.c_void_p), ("N", C.c_int64),
                ("Da", C.c_int32), ("Db", C.c_int32), ("mod_a", C.c_int32), ("max_over_queries", C.c_int32),
                ("n_ctx_a", C.c_int32), ("n_ctx_b", C.c_int32), ("boost", C.c_float), ("k", C.c_int32),
                ("ties_high_index_first", C.c_int32), ("reserved", C.c_int32)]


class EsrGloveCfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("bias_mode", C.c_int32), ("rows_mode", C.c_int32),
                ("impl", C.c_int32), ("B", C.c_int64), ("B_global", C.c_int64),
                ("lr", C.c_float), ("eps", C.c_float), ("x_max", C.c_float), ("alpha", C.c_float),
                ("chunk", C.c_int32), ("reserved", C.c_int32), ("emit_map", C.c_void_p),
                ("emit_peers_dE", C.c_void_p), ("emit_peers_db", C.c_void_p), ("n_emit_peers", C.c_int32),
                ("row_blocks", C.c_int32), ("loss_log", C.c_void_p), ("loss_step", C.c_void_p),
                ("loss_log_len", C.c_int32), ("reserved2", C.c_int32), ("loss_host", C.c_void_p)]


class EsrInbatchCfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("loss_kind", C.c_int32), ("Bq", C.c_int64), ("Bk", C.c_int64),
                ("diag_off", C.c_int64), ("D", C.c_int32), ("splits", C.c_int32), ("margin", C.c_float),
                ("scale", C.c_float), ("b_norm", C.c_float), ("chunk_rows", C.c_int32)]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "esr_version": (C.c_int, []),
    "esr_strerror": (C.c_char_p, [C.c_int]),
    "esr_last_cuda_error": (C.c_char_p, []),
    "esr_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "esr_table_gather_f32": (C.c_int, [C.POINTER(EsrTable), _P, C.c_int64, _P, _P]),
    "esr_table_export_f32": (C.c_int, [C.POINTER(EsrTable), _P, _P]),
    "esr_rowwise_dot_f32": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P]),
    "esr_score_all_f32": (C.c_int, [C.POINTER(EsrTable), _P, C.c_int32, _P, _P]),
    "esr_sort_cols_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "esr_sort_cols_f32": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    "esr_sample_uniform_i32": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, _P, _P]),
    "esr_check_ids_i32": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "esr_plan_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "esr_plan_build_i32": (C.c_int, [C.POINTER(EsrPlan), _P, C.c_size_t, _P]),
    "esr_plan_remap_ids_i32": (C.c_int, [C.POINTER(EsrPlan), _P, _P]),
    "esr_glove_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "esr_glove_prep_f32": (C.c_int, [C.POINTER(EsrTable), C.POINTER(EsrPlan), _P, C.POINTER(EsrGloveCfg), _P, _P,
                                     C.c_size_t, _P]),
    "esr_glove_rows_f32": (C.c_int, [C.POINTER(EsrTable), C.POINTER(EsrPlan), C.POINTER(EsrGloveCfg), _P, _P, _P,
                                     C.c_size_t, _P]),
    "esr_glove_rows_main_f32": (C.c_int, [C.POINTER(EsrTable), C.POINTER(EsrPlan), C.POINTER(EsrGloveCfg), _P, _P, _P,
                                          C.c_size_t, _P]),
    "esr_glove_rows_combine_f32": (C.c_int, [C.POINTER(EsrTable), C.POINTER(EsrPlan), C.POINTER(EsrGloveCfg), _P, _P, _P,
                                             C.c_size_t, _P]),
    "esr_glove_finish_f32": (C.c_int, [C.POINTER(EsrTable), C.POINTER(EsrPlan), C.POINTER(EsrGloveCfg), _P, _P, _P,
                                       C.c_size_t, _P]),
    "esr_glove_step_f32": (C.c_int, [C.POINTER(EsrTable), C.POINTER(EsrPlan), _P, C.POINTER(EsrGloveCfg), _P, _P, _P,
                                     _P, C.c_size_t, _P]),
    "esr_sparse_adagrad_f32": (C.c_int, [C.POINTER(EsrTable), _P, _P, C.c_int64, _P, _P, C.c_float, C.c_float, _P]),
    "esr_scatter_rows_f32": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int64, _P, C.c_int32, _P]),
    "esr_dense_adam_f32": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_int64, _P]),
    "esr_dense_sgdm_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_float, _P]),
    "esr_stl_triplet_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P, _P]),
    "esr_spotify_fwd_bwd_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          _P, _P, _P, _P, _P, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "esr_peer_gather_f32": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_int64, C.c_int32, _P, _P, _P]),
    "esr_peer_emit_plan_i32": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, C.c_int64, _P, C.c_int64, _P, _P, _P]),
    "esr_peer_pull_ids_i32": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, C.c_int64, _P]),
    "esr_peer_merge_adagrad_f32": (C.c_int, [C.POINTER(EsrTable), _P, _P, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64,
                                             C.c_float, C.c_float, _P]),
    "esr_peer_sync_bytes": (C.c_size_t, []),
    "esr_peer_allreduce_f32": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P]),
    "esr_peer_resolve_i32": (C.c_int, [C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64, _P]),
    "esr_pipeline_create": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "esr_pipeline_streams": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "esr_pipeline_set_buffers": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_size_t, _P, _P, C.c_int64, _P]),
    "esr_pipeline_capture_begin": (C.c_int, [_P, C.c_int32]),
    "esr_pipeline_capture_end": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "esr_pipeline_submit": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.POINTER(C.c_int64)]),
    "esr_pipeline_trace": (C.c_int, [_P, C.c_int32]),
    "esr_pipeline_trace_read": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "esr_pipeline_wait_staged": (C.c_int, [_P, C.c_int64]),
    "esr_pipeline_sync": (C.c_int, [_P]),
    "esr_pipeline_destroy": (C.c_int, [_P]),
    "esr_topk_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "esr_topk_scan_f32": (C.c_int, [C.POINTER(EsrTopkCfg), _P, _P, _P, C.c_size_t, _P]),
    "esr_peer_gather_remote_f32": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, C.c_int32, _P, _P, C.c_int32, _P]),
    "esr_peer_apply_parts_f32": (C.c_int, [C.POINTER(EsrTable), _P, _P, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64,
                                           C.c_float, C.c_float, C.c_int32, _P]),
    "esr_plan_compact_owner_i32": (C.c_int, [C.POINTER(EsrPlan), C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "esr_peer_route_pairs_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "esr_peer_route_pairs_i32": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_size_t, _P]),
    "esr_peer_collect_pairs_i32": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]),
    "esr_peer_apply_adagrad_f32": (C.c_int, [C.POINTER(EsrTable), _P, _P, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64,
                                             C.c_float, C.c_float, _P]),
    "esr_route_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "esr_route_plan_i32": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "esr_plan_compact_i32": (C.c_int, [C.POINTER(EsrPlan), _P, _P, _P, _P, _P]),
    "esr_gather_scalar_f32": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "esr_permute_rows_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "esr_segment_sum_rows_f32": (C.c_int, [C.POINTER(EsrPlan), C.c_int32, _P, _P, _P, _P, _P]),
    "esr_decode_cooccur_b64": (C.c_int64, [C.c_char_p, C.c_size_t, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64),
                                           C.POINTER(C.c_size_t)]),
    "esr_decode_tfrecord_int64": (C.c_int64, [_P, C.c_size_t, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(_P),
                                              C.POINTER(C.c_int64), C.POINTER(_P), C.c_int64, C.POINTER(C.c_size_t)]),
    "esr_host_shuffle_gather": (C.c_int, [_P, _P, _P, C.c_int64, C.c_uint64, C.c_int64, C.c_int64, _P, _P, _P, C.c_int32]),
    "esr_inbatch_workspace_bytes": (C.c_size_t, [C.POINTER(EsrInbatchCfg)]),
    "esr_inbatch_fwd_bwd_bf16": (C.c_int, [_P, _P, C.POINTER(EsrInbatchCfg), _P, _P, _P, _P, C.c_size_t, _P]),
    "esr_inbatch_ws_layout": (C.c_int, [C.POINTER(EsrInbatchCfg), C.POINTER(C.c_int64)]),
}

_lib = None


def declared_symbols():
    """Every entry point this binding knows (tests compare it with include/esr.h)."""
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    """Loads libesr.so (once).  Raises EsrError when it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EsrError("libesr.so is missing (%s): run `python -m esrecsys_b200.build` or "
                       "__graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    h = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(h, name)   # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = h
    return h


def check(rc: int, what: str = "") -> None:
    if rc == ESR_OK:
        return
    h = lib()
    msg = h.esr_strerror(rc).decode()
    if rc == ESR_ECUDA:
        msg += ": " + h.esr_last_cuda_error().decode()
    raise EsrError("%s failed: %s" % (what or "libesr call", msg))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise EsrError("esrecsys_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(stream=None) -> int:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
