"""Multi-threaded CPU restatement of the GloVe step with torch CPU ops (TEST INFRASTRUCTURE).

Same algorithm as ``oracle.glove`` (which stays the parity reference), written with torch CPU
tensors so the baseline legs of ``bench.py`` can use every host thread: ``index_select`` for the
``jnp.take`` gathers (wikipedia/models.py:31-34), the closed-form ``value_and_grad`` of
``glove_loss`` (wikipedia/train_cooccurence.py:76-87), ``index_add_`` for the
scatter-add VJP, and

* ``step_adam_dense``    -- what the reference itself executes: a dense ``(V, D)`` gradient and
  ``optax.adam`` over every row (wikipedia/train_cooccurence.py:99-101, :171),
* ``step_adagrad_sparse`` -- the north-star rule on the touched rows only (same algorithm as the GPU path).

Parity unpinned (no reference golden vectors exist); tests/test_oracle_glove_torch.py checks both
against ``oracle.glove`` / ``oracle.optim``.  Only tests/ and bench.py's baseline legs import this.
"""
from __future__ import annotations

import torch

X_MAX = 100.0
ALPHA = 0.75
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8
ADAGRAD_EPS = 1e-7


def loss_and_slot_grads(E, b, i, j, x, bias_mode="reference_broadcast"):
    """Returns (loss, g[B], h[B], ei, ej): dL/d dot_c, dL/d bs_r and the gathered rows."""
    B = i.shape[0]
    ei = E.index_select(0, i)
    ej = E.index_select(0, j)
    dot = (ei * ej).sum(1)
    bs = b.index_select(0, i) + b.index_select(0, j)
    w = torch.clamp(x / X_MAX, max=1.0).pow(ALPHA)
    t = torch.log10(1.0 + x)
    res = t - dot
    if bias_mode == "reference_broadcast":
        S0, S1, S2 = w.sum(), (w * res).sum(), (w * res * res).sum()
        mbs, mbs2 = bs.mean(), (bs * bs).mean()
        loss = (S2 - 2.0 * mbs * S1 + mbs2 * S0) / B
        g = (-2.0 / B) * w * (res - mbs)
        h = (-2.0 / (B * B)) * (S1 - bs * S0)
    else:
        r2 = res - bs
        loss = (w * r2 * r2).sum() / B
        g = (-2.0 / B) * w * r2
        h = g
    return loss, g, h, ei, ej


def dense_grads(E, b, i, j, x, bias_mode="reference_broadcast"):
    loss, g, h, ei, ej = loss_and_slot_grads(E, b, i, j, x, bias_mode)
    dE = torch.zeros_like(E)
    dE.index_add_(0, i, g[:, None] * ej)
    dE.index_add_(0, j, g[:, None] * ei)
    db = torch.zeros_like(b)
    db.index_add_(0, i, h)
    db.index_add_(0, j, h)
    return loss, dE, db


def adam_(p, g, mu, nu, count, lr):
    mu.mul_(ADAM_B1).add_(g, alpha=1.0 - ADAM_B1)
    nu.mul_(ADAM_B2).addcmul_(g, g, value=1.0 - ADAM_B2)
    c1 = 1.0 - ADAM_B1 ** count
    c2 = 1.0 - ADAM_B2 ** count
    p.sub_(lr * (mu / c1) / ((nu / c2).sqrt() + ADAM_EPS))


def step_adam_dense(E, b, st, i, j, x, lr, bias_mode="reference_broadcast"):
    """apply_model + update_model exactly as the reference runs them.  ``st`` = dict(count, muE, nuE, mub, nub)."""
    loss, dE, db = dense_grads(E, b, i, j, x, bias_mode)
    st["count"] += 1
    adam_(E, dE, st["muE"], st["nuE"], st["count"], lr)
    adam_(b, db, st["mub"], st["nub"], st["count"], lr)
    return float(loss)


def step_adagrad_sparse(E, b, accE, accb, i, j, x, lr, bias_mode="reference_broadcast", eps=ADAGRAD_EPS):
    loss, g, h, ei, ej = loss_and_slot_grads(E, b, i, j, x, bias_mode)
    keys = torch.cat([i, j])
    uniq, inv = torch.unique(keys, return_inverse=True)
    dE = torch.zeros(uniq.shape[0], E.shape[1], dtype=E.dtype)
    dE.index_add_(0, inv, torch.cat([g[:, None] * ej, g[:, None] * ei]))
    db = torch.zeros(uniq.shape[0], dtype=E.dtype)
    db.index_add_(0, inv, torch.cat([h, h]))
    a = accE.index_select(0, uniq) + dE * dE
    accE.index_copy_(0, uniq, a)
    E.index_add_(0, uniq, -lr * dE * torch.rsqrt(a + eps))
    ab = accb.index_select(0, uniq) + db * db
    accb.index_copy_(0, uniq, ab)
    b.index_add_(0, uniq, -lr * db * torch.rsqrt(ab + eps))
    return float(loss)
