"""Hot-path functions of wikipedia/train_cooccurence.py with the same names and argument meaning:
``apply_model`` (:71-89), ``update_model`` (:99-101), ``find_knn`` (:91-97), ``train_epoch`` (:103-112).

``apply_model`` returns per-row gradients (``RowGrads``) instead of the dense ``(V,D)`` pytree; the
reference-exact optimizer (dense ``optax.adam``) is applied by ``update_model`` through
``TrainState.apply_gradients``.  The fused fast path (sparse Adagrad, no gradient materialised) is
``esrecsys_b200.trainer.GloveTrainer``.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import engine
from ..train_state import RowGrads, TrainState
from .models import Glove

_cache = {}


def _workspace(table, B, bias_mode):
    key = (table.rows0.data_ptr(), table.bias.data_ptr(), table.V, table.D, B, bias_mode)
    if key not in _cache:
        _cache.clear()
        _cache[key] = (engine.IndexPlan(2 * B, table.V, table.device),
                       engine.GloveStep(table, B, bias_mode=bias_mode, emit_grads=True))
    return _cache[key]


def apply_model(state: TrainState, inputs, target, bias_mode="reference_broadcast"):
    """Computes the gradients and loss for a single batch (train_cooccurence.py:71-89)."""
    params = state.params
    E = params["_token_embedding"]["embedding"]
    table = engine.EmbeddingTable.wrap(E, params["_bias"]["embedding"])
    ids = torch.as_tensor(np.asarray(inputs) if not torch.is_tensor(inputs) else inputs).to(table.device, torch.int32)
    ids = ids.reshape(-1).contiguous()
    B = ids.numel() // 2
    target = torch.as_tensor(target).to(table.device, torch.float32).contiguous()
    fresh = (table.rows0.data_ptr(), table.bias.data_ptr(), table.V, table.D, B, bias_mode) not in _cache
    plan, step = _workspace(table, B, bias_mode)
    if fresh and engine.check_ids(ids, table.V):
        # XLA clamps an out-of-range gather and drops the scatter (wikipedia/models.py:31-34); here ids are raw row
        # offsets, so the first batch of every (table, B) workspace is validated instead of corrupting memory
        raise ValueError("apply_model: ids outside [0, %d)" % table.V)
    plan.build(ids)
    sc = step.run(plan, target)
    loss = sc[L.SC_LOSS].clone()
    V = table.V
    grads = {"_token_embedding": {"embedding": RowGrads(V, plan.uniq.clone(), plan.n_uniq.clone(), step.dE.clone())},
             "_bias": {"embedding": RowGrads(V, plan.uniq.clone(), plan.n_uniq.clone(), step.db.clone().reshape(-1, 1))}}
    return grads, loss


def update_model(state: TrainState, grads):
    """train_cooccurence.py:99-101."""
    return state.apply_gradients(grads=grads)


def find_knn(model: Glove, params, token):
    """train_cooccurence.py:91-97: scores (V,T) and the ascending argsort over axis 0."""
    scores = model.apply({"params": params}, token, method=Glove.score_all)
    indices = engine.sort_cols(scores)          # stable ascending argsort over axis 0, as jnp.argsort
    return scores, indices


def dump_knn(model: Glove, params, tokens, token_dictionary=None, k=10):
    """train_cooccurence.py:114-126: the k nearest neighbours of each query token.  The reference reads them from the
    tail of the ascending stable argsort of the (V, T) score matrix (ties therefore come out HIGHER index first); here
    one fused table scan keeps the running top-k per query (``esr_topk_scan_f32``: no score matrix, no sort of V keys).
    Returns ``[(query, [(neighbour, score), ...]), ...]`` (the reference logs them)."""
    t = model.table(params)
    tok = torch.as_tensor(np.asarray(tokens) if not torch.is_tensor(tokens) else tokens).to(t.device, torch.int32).reshape(-1)
    val, idx = engine.table_topk(t, t.gather(tok), k, ties_high_index_first=True)          # (T, k) best first
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    # ids reach the dictionary as NumPy scalars, as the reference's jax scalars do (its ``is 0`` test never fires for them)
    name = (lambda t: token_dictionary.get_token_from_embedding_index(np.int32(t))) if token_dictionary is not None else int
    out = []
    for i, token in enumerate(np.asarray(tok.cpu()).reshape(-1)):
        out.append((name(int(token)), [(name(int(idx[i, j])), float(val[i, j])) for j in range(k)]))
    return out


def train_epoch(state, steps_per_epoch, train_it):
    """train_cooccurence.py:103-112."""
    epoch_loss = []
    for _ in range(steps_per_epoch):
        inputs, targets = next(train_it)
        grads, loss = apply_model(state, inputs, targets)
        state = update_model(state, grads)
        epoch_loss.append(loss)
    return state, float(torch.stack(epoch_loss).mean().item())


def save_state(state, step, checkpoint_dir):
    """train_cooccurence.py:129-134: ``checkpoint-%05d.flax`` holding ``flax.serialization.to_bytes(state)``."""
    import os
    from .. import checkpoint
    filename = os.path.join(checkpoint_dir, "checkpoint-%05d.flax" % step)
    with open(filename, "wb") as f:
        f.write(checkpoint.to_bytes(state))
    return filename


def resume_state(state, resume_checkpoint):
    """train_cooccurence.py:173-177: ``flax.serialization.from_bytes(state, contents)``."""
    from .. import checkpoint
    with open(resume_checkpoint, "rb") as f:
        return checkpoint.from_bytes(state, f.read())
