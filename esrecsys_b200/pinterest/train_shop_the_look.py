"""Hot-path functions of pinterest/train_shop_the_look.py with the same names and argument meaning:
``train_step`` (:93-109) and ``eval_step`` (:111-122), on the ID-embedding towers of ``STLModel`` (the CNN towers
are out of scope).

``train_step`` = tower lookups (``esr_table_gather_f32``) -> fused triplet forward + backward
(``esr_stl_triplet_f32``) -> per-row gradient sums over the sorted id plan (``esr_segment_sum_rows_f32``; the VJP of
the lookup is a scatter-add) -> the reference's optimizer through ``TrainState.apply_gradients``
(``optax.adam``, train_shop_the_look.py:175).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L
from .. import engine
from ..train_state import RowGrads, TrainState
from .models import STLModel, triplet_loss_and_grads


def _inner(params):
    return params["params"] if "params" in params else params


def _row_grads(table, ids, g):
    """Sum the per-example gradient rows ``g`` by table row (scatter-add over duplicates, fixed order)."""
    V, D = table.shape
    plan = engine.IndexPlan(ids.numel(), V, table.device, with_partner=False, sort="wide").build(ids)
    gsum = torch.empty(ids.numel(), D, device=table.device)
    L.check(L.lib().esr_segment_sum_rows_f32(C.byref(plan.s), D, L.ptr(g), None, L.ptr(gsum), None, L.stream_ptr()),
            "esr_segment_sum_rows_f32")
    return RowGrads(V, plan.uniq, plan.n_uniq, gsum)


def loss_and_grads(model: STLModel, params, scene, pos_product, neg_product, regularization, batch_size):
    """``jax.value_and_grad(loss_fn)`` of train_step (:94-107): (loss, grads pytree shaped like ``params``)."""
    p = _inner(params)
    dev = p["scene_cnn"]["embedding"].device
    ids = [torch.as_tensor(v, device=dev).to(torch.int32).reshape(-1).contiguous() for v in (scene, pos_product, neg_product)]
    _, _, s, pp, pn = model(p, *ids, True)
    loss, ds, dp, dn, _, _ = triplet_loss_and_grads(s, pp, pn, regularization, batch_size)
    grads = {"scene_cnn": {"embedding": _row_grads(p["scene_cnn"]["embedding"], ids[0], ds.contiguous())},
             "product_cnn": {"embedding": _row_grads(p["product_cnn"]["embedding"], torch.cat([ids[1], ids[2]]),
                                                     torch.cat([dp, dn]).contiguous())}}
    return loss, ({"params": grads} if "params" in params else grads)


def train_step(state: TrainState, model: STLModel, scene, pos_product, neg_product, regularization, batch_size):
    """train_shop_the_look.py:93-109: returns (new_state, loss)."""
    loss, grads = loss_and_grads(model, state.params, scene, pos_product, neg_product, regularization, batch_size)
    return state.apply_gradients(grads=grads), loss


def eval_step(state: TrainState, model: STLModel, scene, pos_product, neg_product):
    """train_shop_the_look.py:111-122: the triplet hinge at the fixed margin, no regulariser, no division."""
    p = _inner(state.params)
    dev = p["scene_cnn"]["embedding"].device
    ids = [torch.as_tensor(v, device=dev).to(torch.int32).reshape(-1).contiguous() for v in (scene, pos_product, neg_product)]
    _, _, s, pp, pn = model(p, *ids, True)
    return triplet_loss_and_grads(s, pp, pn, 0.0, 1.0)[0]
