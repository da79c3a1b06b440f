"""Oracle for the GloVe trainer's hot path (TEST INFRASTRUCTURE).

Restates, in NumPy:

* ``Glove.__call__``  wikipedia/models.py:21-38  -- shared token table for both
  roles (:16-19,:31-34), row-wise dot (:35-36), and the ``(B,) + (B,1) + (B,1)``
  broadcast that makes the output ``(B,B)`` (:37): ``out[r,c] = dot[c] + b[i_r] + b[j_r]``.
* ``glove_loss``      wikipedia/train_cooccurence.py:76-84 -- ``w = min(1, x/100)^0.75``
  (:79-81), ``t = log10(1+x)`` (:82), ``mean((t - out)^2 * w)`` over the B*B cells (:83).
* ``jax.value_and_grad`` (:86-87): the VJP of ``jnp.take`` is a scatter-add, restated
  here as a stable-sorted segment sum (oracle.index).
* ``update_model``    wikipedia/train_cooccurence.py:99-101 + ``optax.adam`` (:171), and
  the north-star sparse Adagrad rule (oracle.optim).
* ``Glove.score_all`` / ``find_knn``  wikipedia/models.py:40-55, train_cooccurence.py:91-97.

``bias_mode="reference_broadcast"`` is what the reference computes;
``bias_mode="per_pair"`` is textbook GloVe.

PINNING: jax/flax (jax 0.3.25, flax 0.5.2 -- wikipedia/requirements.txt:18-20) are not
installable here and the reference holds no golden vectors.  The closed form below is
pinned against (1) the reference's own models.py / train_cooccurence.py executed on the
jax/flax/optax stand-in of tests/golden/refshim (tests/golden/ref_glove_*.npz, 1e-12),
(2) the literal (B,B) evaluation in this file and (3) torch float64 autograd of that
literal forward (tests/test_oracle_glove.py).  Real-jax parity stays unpinned.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import index as oidx
from . import optim as oopt

X_MAX = 100.0   # wikipedia/train_cooccurence.py:80
ALPHA = 0.75    # wikipedia/train_cooccurence.py:81


def weight_fn(x):
    """train_cooccurence.py:79-81."""
    x = np.asarray(x)
    return np.power(np.minimum(np.ones_like(x), x / x.dtype.type(X_MAX)), x.dtype.type(ALPHA))


def log_target_fn(x):
    """train_cooccurence.py:82."""
    x = np.asarray(x)
    return np.log10(x.dtype.type(1.0) + x)


def forward_literal(E, b, i, j):
    """Glove.__call__ evaluated literally, including the (B,B) broadcast (models.py:30-38)."""
    e1 = E[i]
    e2 = E[j]
    b1 = b[i].reshape(-1, 1)
    b2 = b[j].reshape(-1, 1)
    dot = np.einsum("cd,cd->c", e1, e2)
    return dot + b1 + b2          # (B,) + (B,1) + (B,1) -> (B,B)


def loss_literal(E, b, i, j, x):
    """glove_loss evaluated literally on the (B,B) prediction (train_cooccurence.py:76-84)."""
    pred = forward_literal(E, b, i, j)
    w = weight_fn(x)
    t = log_target_fn(x)
    return np.mean(np.square(t - pred) * w)


@dataclass
class GloveGrads:
    loss: float
    dot: np.ndarray        # (B,)
    g: np.ndarray          # (B,) dL/d dot_c
    h: np.ndarray          # (B,) dL/d bs_r   (bs_r = b[i_r] + b[j_r])
    sorted_keys: np.ndarray
    perm: np.ndarray
    uniq: np.ndarray       # (U,)
    seg_off: np.ndarray    # (U+1,)
    dE: np.ndarray         # (U, D) per-unique-row embedding gradient
    db: np.ndarray         # (U,)   per-unique-row bias gradient
    S0: float = 0.0
    S1: float = 0.0
    S2: float = 0.0
    mean_bs: float = 0.0


def loss_and_grads(E, b, i, j, x, bias_mode="reference_broadcast") -> GloveGrads:
    """Closed form of value_and_grad(glove_loss).

    Gradients are returned per unique touched row, accumulated in stable
    sorted-slot order (the summation order the CUDA path uses).
    """
    dt = E.dtype
    i = np.asarray(i, np.int32)
    j = np.asarray(j, np.int32)
    B = i.shape[0]
    ei = E[i]
    ej = E[j]
    dot = np.einsum("cd,cd->c", ei, ej).astype(dt)
    bs = (b[i] + b[j]).astype(dt)
    w = weight_fn(x.astype(dt))
    t = log_target_fn(x.astype(dt))
    res = (t - dot).astype(dt)
    S0 = w.sum(dtype=dt)
    S1 = (w * res).sum(dtype=dt)
    S2 = (w * res * res).sum(dtype=dt)
    Bf = dt.type(B)
    if bias_mode == "reference_broadcast":
        mbs = bs.sum(dtype=dt) / Bf
        mbs2 = (bs * bs).sum(dtype=dt) / Bf
        loss = (S2 - dt.type(2.0) * mbs * S1 + mbs2 * S0) / Bf
        g = (-(dt.type(2.0) / Bf) * w * (res - mbs)).astype(dt)
        h = (-(dt.type(2.0) / (Bf * Bf)) * (S1 - bs * S0)).astype(dt)
    elif bias_mode == "per_pair":
        mbs = bs.sum(dtype=dt) / Bf
        r2 = (res - bs).astype(dt)
        loss = (w * r2 * r2).sum(dtype=dt) / Bf
        g = (-(dt.type(2.0) / Bf) * w * r2).astype(dt)
        h = g.copy()
    else:
        raise ValueError(bias_mode)

    keys = oidx.slot_keys(i, j)
    sk, perm = oidx.sort_slots(keys)
    uniq, seg_off = oidx.segments(sk)
    pair = perm % B                       # pair index of each sorted slot
    side_j = perm >= B                    # slot is the j role -> partner is i
    partner = np.where(side_j, i[pair], j[pair])
    contrib = (g[pair][:, None] * E[partner]).astype(dt)     # (2B, D)
    hcon = h[pair]
    U = uniq.shape[0]
    dE = np.zeros((U, E.shape[1]), dt)
    db = np.zeros(U, dt)
    useg = oidx.slot_segment_index(sk)
    # in-order accumulation per segment (np.add.at applies updates sequentially)
    np.add.at(dE, useg, contrib)
    np.add.at(db, useg, hcon)
    return GloveGrads(loss=float(loss), dot=dot, g=g, h=h, sorted_keys=sk, perm=perm,
                      uniq=uniq, seg_off=seg_off, dE=dE, db=db,
                      S0=float(S0), S1=float(S1), S2=float(S2), mean_bs=float(mbs))


def dense_grads(V, gr: GloveGrads, D):
    """The dense pytree jax.value_and_grad returns (train_cooccurence.py:86-87)."""
    dE = np.zeros((V, D), gr.dE.dtype)
    db = np.zeros(V, gr.db.dtype)
    dE[gr.uniq] = gr.dE
    db[gr.uniq] = gr.db
    return dE, db


def step_adagrad(E, b, accE, accb, i, j, x, lr, bias_mode="reference_broadcast",
                 eps=oopt.ADAGRAD_EPS):
    """One batch-synchronous sparse Adagrad step (north-star rule). In place. Returns loss."""
    gr = loss_and_grads(E, b, i, j, x, bias_mode)
    u = gr.uniq
    E[u], accE[u] = oopt.adagrad_update(E[u], gr.dE, accE[u], lr, eps)
    b[u], accb[u] = oopt.adagrad_update(b[u], gr.db, accb[u], lr, eps)
    return gr.loss


def step_adam(E, b, st, i, j, x, lr, bias_mode="reference_broadcast"):
    """The reference's own rule: dense Adam over every row (train_cooccurence.py:99-101,171).

    ``st`` = dict(count, muE, nuE, mub, nub).  In place.  Returns loss.
    """
    gr = loss_and_grads(E, b, i, j, x, bias_mode)
    dE, db = dense_grads(E.shape[0], gr, E.shape[1])
    c = st["count"]
    E[...], st["muE"], st["nuE"], _ = oopt.adam_update(E, dE, st["muE"], st["nuE"], c, lr)
    b[...], st["mub"], st["nub"], st["count"] = oopt.adam_update(b, db, st["mub"], st["nub"], c, lr)
    return gr.loss


def step_sgdm(E, b, st, i, j, x, lr, momentum, bias_mode="reference_broadcast"):
    """Dense SGD + momentum trace (optax.sgd) applied to the GloVe tables."""
    gr = loss_and_grads(E, b, i, j, x, bias_mode)
    dE, db = dense_grads(E.shape[0], gr, E.shape[1])
    E[...], st["trE"] = oopt.sgdm_update(E, dE, st["trE"], lr, momentum)
    b[...], st["trb"] = oopt.sgdm_update(b, db, st["trb"], lr, momentum)
    return gr.loss


def score_all(E, tokens):
    """Glove.score_all (models.py:40-55): scores[v, t] = E[v] . E[tokens[t]]  -> (V, T)."""
    return (E @ E[np.asarray(tokens)].T).astype(E.dtype)


def find_knn(E, tokens):
    """find_knn (train_cooccurence.py:91-97): ascending stable argsort over axis 0."""
    scores = score_all(E, tokens)
    return scores, np.argsort(scores, axis=0, kind="stable").astype(np.int32)


def top_k(E, tokens, k):
    """What dump_knn reads (train_cooccurence.py:121-125): the last k of the argsort, reversed."""
    scores, idx = find_knn(E, tokens)
    top = idx[::-1][:k].T.copy()                       # (T, k), best first
    return top, np.take_along_axis(scores.T, top, axis=1)
