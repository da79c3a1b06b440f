"""Oracle for the shop-the-look two-tower scoring + loss (TEST INFRASTRUCTURE).

Restates, in NumPy:

* ``STLModel.__call__`` scoring  pinterest/models.py:63-74 -- row-wise dots
  ``pos = sum(scene*pos_prod, -1)``, ``neg = sum(scene*neg_prod, -1)`` (:67-72).
* ``train_step`` loss            pinterest/train_shop_the_look.py:93-109 --
  ``(sum relu(1 + neg - pos) + reg * sum_b sum_e relu(||e_b|| - 1)) / batch_size`` (:99-104)
  and its gradient wrt the three embedding matrices (``jax.value_and_grad`` :106-107).
* ``eval_step``                  pinterest/train_shop_the_look.py:111-122.
* ``find_top_k``                 pinterest/make_recommendations.py:49-65.

The CNN towers (pinterest/models.py:23-46) are out of scope; the north star
substitutes ID-embedding + 2-layer MLP towers, restated in ``mlp_tower*`` below
(our own definition -- SURVEY.md D5 -- so "parity" there is against this file and
torch autograd only).

The in-batch generalisations (``inbatch_hinge``, ``inbatch_softmax``; SURVEY.md
App. A.4) have NO reference counterpart at all.

PINNING: jax 0.3.25 / flax 0.5.2 / optax 0.1.2 (pinterest/requirements.txt:5-8) are not
installable here; no reference golden vectors exist.  Pinned (1e-12) against the reference's own
pinterest/models.py (STLModel, CNN towers swapped for ID towers) + train_shop_the_look.py +
make_recommendations.py executed on the jax/flax/optax stand-in of tests/golden/refshim
(tests/golden/ref_stl.npz, tests/test_ref_golden.py) and against torch float64 autograd
(tests/test_oracle_stl.py).  Real-jax parity stays unpinned.
"""
from __future__ import annotations

import numpy as np


def scores(scene, pos, neg):
    """pinterest/models.py:67-72."""
    return np.sum(scene * pos, axis=-1), np.sum(scene * neg, axis=-1)


def triplet_loss(scene, pos, neg, regularization, batch_size):
    """pinterest/train_shop_the_look.py:99-104."""
    dt = scene.dtype
    ps, ns = scores(scene, pos, neg)
    relu = lambda v: np.maximum(v, dt.type(0))
    trip = np.sum(relu(dt.type(1.0) + ns - ps))

    def reg_fn(e):
        return relu(np.sqrt(np.sum(np.square(e), axis=-1)) - dt.type(1.0))

    reg = np.sum(reg_fn(scene) + reg_fn(pos) + reg_fn(neg))
    return dt.type((trip + dt.type(regularization) * reg) / dt.type(batch_size))


def triplet_loss_and_grads(scene, pos, neg, regularization, batch_size):
    """Loss and d loss / d (scene, pos, neg), relu'(0) = 0."""
    dt = scene.dtype
    loss = triplet_loss(scene, pos, neg, regularization, batch_size)
    ps, ns = scores(scene, pos, neg)
    a = ((dt.type(1.0) + ns - ps) > 0).astype(dt)[:, None]
    inv = dt.type(1.0 / batch_size)

    def reg_grad(e):
        n = np.sqrt(np.sum(np.square(e), axis=-1, keepdims=True))
        act = (n - dt.type(1.0)) > 0
        return np.where(act, e / np.where(n > 0, n, 1), 0).astype(dt)

    r = dt.type(regularization)
    ds = (a * (neg - pos) + r * reg_grad(scene)) * inv
    dp = (-a * scene + r * reg_grad(pos)) * inv
    dn = (a * scene + r * reg_grad(neg)) * inv
    return loss, ds.astype(dt), dp.astype(dt), dn.astype(dt)


def eval_loss(scene, pos, neg):
    """pinterest/train_shop_the_look.py:111-122 (fixed margin, no regulariser, no division)."""
    dt = scene.dtype
    ps, ns = scores(scene, pos, neg)
    return dt.type(np.sum(np.maximum(dt.type(1.0) + ns - ps, dt.type(0))))


def find_top_k(scene_embedding, product_embeddings, k):
    """pinterest/make_recommendations.py:49-65: scores = sum(scene*products, -1); lax.top_k.

    Ties broken by lower index first (lax.top_k).
    """
    sc = np.sum(scene_embedding * product_embeddings, axis=-1)
    order = np.lexsort((np.arange(sc.shape[0]), -sc.astype(np.float64)))[:k]
    return sc[order], order.astype(np.int32)


# --------------------------------------------------------------------------
# North-star substitutions (no reference counterpart)
# --------------------------------------------------------------------------

def mlp_tower(x, W1, b1, W2, b2):
    """ID-embedding rows -> Linear -> ReLU -> Linear."""
    h = np.maximum(x @ W1 + b1, 0)
    return h @ W2 + b2


def inbatch_hinge(Q, K):
    """L = (1/B) sum_i sum_{j != i} relu(1 + S_ij - S_ii),  S = Q K^T.

    Returns (loss, dQ, dK).
    """
    dt = Q.dtype
    B = Q.shape[0]
    S = Q @ K.T
    d = np.diag(S)[:, None]
    act = ((dt.type(1.0) + S - d) > 0)
    np.fill_diagonal(act, False)
    loss = np.sum(np.where(act, dt.type(1.0) + S - d, 0)) / dt.type(B)
    dS = act.astype(dt) / dt.type(B)
    dS[np.arange(B), np.arange(B)] = -act.sum(axis=1).astype(dt) / dt.type(B)
    return dt.type(loss), (dS @ K).astype(dt), (dS.T @ Q).astype(dt)


def inbatch_softmax(Q, K):
    """L = (1/B) sum_i [logsumexp_j S_ij - S_ii];  dS = (softmax_row(S) - I)/B."""
    dt = Q.dtype
    B = Q.shape[0]
    S = Q @ K.T
    mx = S.max(axis=1, keepdims=True)
    ex = np.exp(S - mx)
    se = ex.sum(axis=1, keepdims=True)
    lse = (np.log(se) + mx)[:, 0]
    loss = np.sum(lse - np.diag(S)) / dt.type(B)
    dS = ex / se
    dS[np.arange(B), np.arange(B)] -= dt.type(1.0)
    dS = dS / dt.type(B)
    return dt.type(loss), (dS @ K).astype(dt), (dS.T @ Q).astype(dt)
