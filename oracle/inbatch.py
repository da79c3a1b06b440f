"""Oracle for the in-batch-negative score kernel (TEST INFRASTRUCTURE, not product code).

Contract of ``esr_inbatch_fwd_bwd_bf16`` (include/esr.h, csrc/inbatch_scores.cu), in NumPy.

There is NO reference counterpart: the reference only scores explicit triplets
(pinterest/models.py:67-72 row-wise dot, pinterest/train_shop_the_look.py:99-104 hinge(1 + neg - pos);
spotify/train_spotify.py:91-94 affinity hinges).  The B x B in-batch form is the north-star
generalisation, so this file is the definition and **parity is
unpinned**; it is validated against torch float64 autograd in tests/test_oracle_inbatch.py and
collapses to ``oracle.stl.inbatch_hinge`` / ``inbatch_softmax`` for square batches, offset 0,
scale = margin = 1 on bf16-representable inputs.

Numerics of the contract (what "bf16 tensor-core contraction" means here):
  * Q and K are rounded to bfloat16 (round-to-nearest-even) first; every product below is then exact
    in fp32 and only the summation order differs between this file and the tensor cores;
  * scores are accumulated in fp32 and never rounded;
  * hinge: dL/dS is an exact {0,1} mask (scaled in fp32 afterwards);
  * softmax: the probabilities are rounded to bfloat16 before the two backward contractions;
  * gradients are straight-through to the fp32 inputs.
"""
from __future__ import annotations

import numpy as np


def bf16_round(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).reshape(x.shape)


def scores(Q, K):
    """S = bf16(Q) bf16(K)^T with fp32 accumulation (float64 here, then cast: every product is exact)."""
    Qh, Kh = bf16_round(Q), bf16_round(K)
    return (Qh.astype(np.float64) @ Kh.astype(np.float64).T).astype(np.float32), Qh, Kh


def _pos_index(Bq, Bk, off):
    pj = np.arange(Bq) + off
    ok = (pj >= 0) & (pj < Bk)
    return pj, ok


def hinge_mask(S, off, margin=1.0, scale=1.0):
    """mask[i, j] = [j != pos(i)] * [margin + scale*S_ij - scale*S_i,pos(i) > 0] and the margins h."""
    Bq, Bk = S.shape
    pj, ok = _pos_index(Bq, Bk, off)
    sii = np.where(ok, np.float32(scale) * S[np.arange(Bq), np.clip(pj, 0, Bk - 1)], np.float32(0)).astype(np.float32)
    h = (np.float32(margin) + np.float32(scale) * S - sii[:, None]).astype(np.float32)
    mask = h > 0
    mask[np.arange(Bq)[ok], pj[ok]] = False
    return mask, h


def hinge_backward(mask, Qh, Kh, off, scale=1.0, b_norm=None):
    """dQ, dK for a GIVEN mask (lets a test feed the kernel's own mask back: the backward is then exact
    up to fp32 summation order)."""
    Bq, Bk = mask.shape
    b_norm = float(Bq if b_norm is None else b_norm)
    pj, ok = _pos_index(Bq, Bk, off)
    m = mask.astype(np.float64)
    cnt = m.sum(axis=1)
    dQ = m @ Kh.astype(np.float64)
    dK = m.T @ Qh.astype(np.float64)
    dQ[ok] -= cnt[ok, None] * Kh[pj[ok]].astype(np.float64)
    np.subtract.at(dK, pj[ok], cnt[ok, None] * Qh[ok].astype(np.float64))
    c = scale / b_norm
    return (dQ * c).astype(np.float32), (dK * c).astype(np.float32)


def hinge(Q, K, off=0, margin=1.0, scale=1.0, b_norm=None):
    """Returns (loss, dQ, dK, mask)."""
    S, Qh, Kh = scores(Q, K)
    b_norm = float(Q.shape[0] if b_norm is None else b_norm)
    mask, h = hinge_mask(S, off, margin, scale)
    loss = np.float32(np.sum(np.where(mask, h, 0).astype(np.float64)) / b_norm)
    dQ, dK = hinge_backward(mask, Qh, Kh, off, scale, b_norm)
    return loss, dQ, dK, mask


def softmax_probs(S, scale=1.0):
    """(P fp32, lse) with P = exp(scale*S - lse) row-wise."""
    Z = (np.float32(scale) * S).astype(np.float64)
    mx = Z.max(axis=1, keepdims=True)
    se = np.exp(Z - mx).sum(axis=1, keepdims=True)
    lse = (mx + np.log(se))[:, 0]
    return np.exp(Z - lse[:, None]).astype(np.float32), lse.astype(np.float32)


def softmax_backward(P_bf16, Qh, Kh, off, scale=1.0, b_norm=None):
    """dQ, dK for GIVEN (bf16-rounded) probabilities."""
    Bq, Bk = P_bf16.shape
    b_norm = float(Bq if b_norm is None else b_norm)
    pj, ok = _pos_index(Bq, Bk, off)
    p = P_bf16.astype(np.float64)
    dQ = p @ Kh.astype(np.float64)
    dK = p.T @ Qh.astype(np.float64)
    dQ[ok] -= Kh[pj[ok]].astype(np.float64)
    np.subtract.at(dK, pj[ok], Qh[ok].astype(np.float64))
    c = scale / b_norm
    return (dQ * c).astype(np.float32), (dK * c).astype(np.float32)


def softmax(Q, K, off=0, scale=1.0, b_norm=None):
    """Returns (loss, dQ, dK, P_bf16)."""
    S, Qh, Kh = scores(Q, K)
    Bq, Bk = S.shape
    b_norm = float(Bq if b_norm is None else b_norm)
    P, lse = softmax_probs(S, scale)
    pj, ok = _pos_index(Bq, Bk, off)
    pos = np.where(ok, np.float32(scale) * S[np.arange(Bq), np.clip(pj, 0, Bk - 1)], np.float32(0))
    loss = np.float32(np.sum(lse.astype(np.float64) - pos.astype(np.float64)) / b_norm)
    Pb = bf16_round(P)
    dQ, dK = softmax_backward(Pb, Qh, Kh, off, scale, b_norm)
    return loss, dQ, dK, Pb


# --------------------------------------------------------------------------------------------------
# Whole training steps of the in-batch configs (contract of esrecsys_b200/inbatch.py)
# --------------------------------------------------------------------------------------------------
def _loss_and_grads(kind, Q, K, margin, scale):
    if kind == "hinge":
        loss, dQ, dK, _ = hinge(Q, K, 0, margin, scale)
    else:
        loss, dQ, dK, _ = softmax(Q, K, 0, scale)
    return loss, dQ, dK


def _adagrad_rows(E, acc, ids, grads, lr, eps=1e-7):
    """Batch-synchronous sparse Adagrad: gradients of duplicate ids are summed first (what XLA's
    scatter-add VJP of jnp.take gives -- SURVEY.md App. A.6), then optax.adagrad on the touched rows."""
    uniq, inv = np.unique(ids, return_inverse=True)
    g = np.zeros((uniq.size, E.shape[1]), np.float64)
    np.add.at(g, inv, grads.astype(np.float64))
    g = g.astype(np.float32)
    a = acc[uniq] + g * g
    E[uniq] = E[uniq] - np.float32(lr) * g * (np.float32(1.0) / np.sqrt(a + np.float32(eps)))
    acc[uniq] = a


def shared_table_step(E, acc, q_ids, k_ids, lr, kind="hinge", margin=1.0, scale=1.0):
    """configs[2]: one table, (query, item) id pairs, in-batch negatives.  Updates E, acc in place; returns the loss."""
    loss, dQ, dK = _loss_and_grads(kind, E[q_ids], E[k_ids], margin, scale)
    _adagrad_rows(E, acc, np.concatenate([q_ids, k_ids]), np.concatenate([dQ, dK]), lr)
    return loss


def mlp_forward(x, p):
    h = np.maximum(x @ p["W1"] + p["b1"], 0).astype(np.float32)
    return (h @ p["W2"] + p["b2"]).astype(np.float32), h


def mlp_backward(x, h, dy, p):
    g = {"W2": h.T @ dy, "b2": dy.sum(0)}
    dh = (dy @ p["W2"].T) * (h > 0)
    g["W1"] = x.T @ dh
    g["b1"] = dh.sum(0)
    return (dh @ p["W1"].T).astype(np.float32), {k: v.astype(np.float32) for k, v in g.items()}


def two_tower_step(Es, accs, Ep, accp, ps, pp, opt_s, opt_p, s_ids, p_ids, lr, tower_lr, kind="softmax", margin=1.0,
                   scale=1.0):
    """configs[3]: id tables -> 2-layer MLP towers -> in-batch loss; tables sparse Adagrad, towers optax.adam
    (``opt_*``: dict(count, mu, nu) of dicts).  In place; returns the loss."""
    from . import optim as oo
    xs, xp = Es[s_ids], Ep[p_ids]
    q, hs = mlp_forward(xs, ps)
    k, hp = mlp_forward(xp, pp)
    loss, dq, dk = _loss_and_grads(kind, q, k, margin, scale)
    dxs, gs = mlp_backward(xs, hs, dq, ps)
    dxp, gp = mlp_backward(xp, hp, dk, pp)
    for p, g, opt in ((ps, gs, opt_s), (pp, gp, opt_p)):
        for name in p:
            p[name], opt["mu"][name], opt["nu"][name], _ = oo.adam_update(p[name], g[name], opt["mu"][name], opt["nu"][name],
                                                                          opt["count"], tower_lr)
        opt["count"] += 1
    _adagrad_rows(Es, accs, s_ids, dxs, lr)
    _adagrad_rows(Ep, accp, p_ids, dxp, lr)
    return loss
