"""Oracle for the Spotify playlist trainer's hot path (TEST INFRASTRUCTURE).

Restates, in NumPy:

* ``SpotifyModel.get_embeddings``  spotify/models.py:33-46 -- ``concat(album_embed[album % 100000],
  artist_embed[artist])``.
* ``SpotifyModel.__call__``        spotify/models.py:48-91 -- max-over-context affinity (:74,:78),
  0.1 ``isin`` boosts on the RAW (un-modded) album ids and artist ids (:75-76,:79-80),
  per-row L2 norms (:82-83), self-affinity grams with a row flip (:85-87).
* ``train_step`` loss              spotify/train_spotify.py:91-105 and its gradient
  (``jax.value_and_grad`` :108-109) -- derived by hand with JAX's VJP conventions:
  ``max``/``min`` split the cotangent equally over ties, ``relu'(0) = 0``.
* ``apply_gradients`` with ``optax.sgd(lr, momentum)`` spotify/train_spotify.py:110,238-241 -- DENSE
  (every row's trace decays and moves every step).
* ``eval_step`` scoring + top-k    spotify/train_spotify.py:113-131.

PINNING: jax 0.4.10 / flax 0.6.9 / optax 0.1.5 (spotify/requirements.txt:17,30,53) are not
installable here; no reference golden vectors exist.  Pinned (1e-12) against the reference's
own spotify/models.py + train_spotify.py executed on the jax/flax/optax stand-in of
tests/golden/refshim (tests/golden/ref_spotify.npz, tests/test_ref_golden.py) and against torch
float64 autograd of a literal forward transcription (tests/test_oracle_spotify.py).  The jax
VJP / optax rules themselves are restated there, so real-jax parity stays unpinned.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import optim as oopt

MAX_ALBUMS = 100000      # spotify/models.py:29
NUM_ARTISTS = 295861     # spotify/models.py:31


def get_embeddings(A, R, album, artist):
    """spotify/models.py:33-46."""
    return np.concatenate([A[np.mod(album, A.shape[0])], R[artist]], axis=-1)


def forward(A, R, album_ctx, artist_ctx, next_album, next_artist, neg_album, neg_artist):
    """SpotifyModel.__call__ (spotify/models.py:48-91). Track ids are unused by the model."""
    dt = A.dtype
    ctx = get_embeddings(A, R, album_ctx, artist_ctx)
    nxt = get_embeddings(A, R, next_album, next_artist)
    neg = get_embeddings(A, R, neg_album, neg_artist)
    pos_aff = np.max(nxt @ ctx.T, axis=-1)
    pos_aff = pos_aff + dt.type(0.1) * np.isin(next_album, album_ctx)
    pos_aff = pos_aff + dt.type(0.1) * np.isin(next_artist, artist_ctx)
    neg_aff = np.max(neg @ ctx.T, axis=-1)
    neg_aff = neg_aff + dt.type(0.1) * np.isin(neg_album, album_ctx)
    neg_aff = neg_aff + dt.type(0.1) * np.isin(neg_artist, artist_ctx)
    allemb = np.concatenate([ctx, nxt, neg], axis=-2)
    l2 = np.sqrt(np.sum(np.square(allemb), axis=-1))
    g_ctx = np.flip(ctx, axis=-2) @ ctx.T
    g_nxt = np.flip(nxt, axis=-2) @ nxt.T
    g_neg = np.flip(neg, axis=-2) @ neg.T
    return (pos_aff.astype(dt), neg_aff.astype(dt), g_ctx.astype(dt), g_nxt.astype(dt),
            g_neg.astype(dt), l2.astype(dt))


def loss_from_outputs(out, regularization):
    """spotify/train_spotify.py:91-105."""
    pos, neg, g_ctx, g_nxt, g_neg, l2 = out
    dt = pos.dtype
    relu = lambda v: np.maximum(v, dt.type(0))
    mean_trip = relu(dt.type(1.0) + np.mean(neg) - np.mean(pos))
    ext_trip = relu(dt.type(1.0) + np.max(neg) - np.min(pos))
    ctx_l = np.mean(relu(dt.type(0.5) - g_ctx))
    nxt_l = np.mean(relu(dt.type(0.5) - g_nxt))
    neg_l = np.mean(relu(g_neg))
    reg_l = np.sum(relu(l2 - dt.type(regularization)))
    return dt.type(ext_trip + mean_trip + reg_l + ctx_l + nxt_l + neg_l)


@dataclass
class SpotifyGrads:
    loss: float
    dX: np.ndarray           # (5+m+o, 2F) gradient wrt the stacked [ctx; next; neg] embeddings
    album_rows: np.ndarray   # (5+m+o,) modded album row of every stacked embedding
    artist_rows: np.ndarray  # (5+m+o,)


def _max_cotangent(S, dvec):
    """VJP of max over the last axis with equal tie-splitting (JAX reduce_max rule)."""
    mx = S.max(axis=-1, keepdims=True)
    ind = (S == mx).astype(S.dtype)
    return dvec[:, None] * ind / ind.sum(axis=-1, keepdims=True)


def _gram_grad(X, scale_sign, thresh_fn):
    """Gradient of mean(hinge(flip(X) @ X.T)) wrt X.

    ``G[a,b] = X[n-1-a] . X[b]``; ``Q = dL/dG``; dX[b] += sum_a Q[a,b] X[n-1-a];
    dX[n-1-a] += sum_b Q[a,b] X[b].
    """
    n = X.shape[0]
    Xf = X[::-1]
    G = Xf @ X.T
    Q = (thresh_fn(G).astype(X.dtype) * X.dtype.type(scale_sign / (n * n)))
    d = Q.T @ Xf                      # contribution to X[b]
    d = d + (Q @ X)[::-1]             # contribution to X[n-1-a]
    return d


def loss_and_grads(A, R, album_ctx, artist_ctx, next_album, next_artist, neg_album, neg_artist,
                   regularization) -> SpotifyGrads:
    dt = A.dtype
    out = forward(A, R, album_ctx, artist_ctx, next_album, next_artist, neg_album, neg_artist)
    pos, neg, g_ctx, g_nxt, g_neg, l2 = out
    loss = loss_from_outputs(out, regularization)
    ctx = get_embeddings(A, R, album_ctx, artist_ctx)
    nxt = get_embeddings(A, R, next_album, next_artist)
    ngx = get_embeddings(A, R, neg_album, neg_artist)
    m, o = nxt.shape[0], ngx.shape[0]
    nc = ctx.shape[0]

    dpos = np.zeros(m, dt)
    dneg = np.zeros(o, dt)
    if 1.0 + np.mean(neg) - np.mean(pos) > 0:                 # mean triplet hinge (:91-93)
        dneg += dt.type(1.0 / o)
        dpos -= dt.type(1.0 / m)
    if 1.0 + np.max(neg) - np.min(pos) > 0:                   # extremal hinge (:95-97)
        imax = (neg == neg.max()).astype(dt)
        imin = (pos == pos.min()).astype(dt)
        dneg += imax / imax.sum()
        dpos -= imin / imin.sum()
    dS_pos = _max_cotangent(nxt @ ctx.T, dpos)                # (m, nc)
    dS_neg = _max_cotangent(ngx @ ctx.T, dneg)                # (o, nc)
    d_ctx = dS_pos.T @ nxt + dS_neg.T @ ngx
    d_nxt = dS_pos @ ctx
    d_neg = dS_neg @ ctx
    d_ctx = d_ctx + _gram_grad(ctx, -1.0, lambda G: (0.5 - G) > 0)   # :99
    d_nxt = d_nxt + _gram_grad(nxt, -1.0, lambda G: (0.5 - G) > 0)   # :100
    d_neg = d_neg + _gram_grad(ngx, +1.0, lambda G: G > 0)           # :101
    X = np.concatenate([ctx, nxt, ngx], axis=0)
    dX = np.concatenate([d_ctx, d_nxt, d_neg], axis=0).astype(dt)
    act = (l2 - dt.type(regularization)) > 0                   # :103
    dX = dX + (act[:, None] * X / np.where(l2 > 0, l2, 1)[:, None]).astype(dt)
    album_rows = np.mod(np.concatenate([album_ctx, next_album, neg_album]), A.shape[0]).astype(np.int64)
    artist_rows = np.concatenate([artist_ctx, next_artist, neg_artist]).astype(np.int64)
    assert dX.shape[0] == nc + m + o
    return SpotifyGrads(float(loss), dX.astype(dt), album_rows, artist_rows)


def dense_grads(A, R, gr: SpotifyGrads):
    F = A.shape[1]
    dA = np.zeros_like(A)
    dR = np.zeros_like(R)
    np.add.at(dA, gr.album_rows, gr.dX[:, :F])
    np.add.at(dR, gr.artist_rows, gr.dX[:, F:])
    return dA, dR


def train_step(A, R, trA, trR, x, regularization, lr, momentum):
    """train_step (spotify/train_spotify.py:77-111) with optax.sgd(lr, momentum). In place; returns loss."""
    gr = loss_and_grads(A, R, x["album_context"], x["artist_context"], x["next_album"],
                        x["next_artist"], x["neg_album"], x["neg_artist"], regularization)
    dA, dR = dense_grads(A, R, gr)
    A[...], trA[...] = oopt.sgdm_update(A, dA, trA, lr, momentum)
    R[...], trR[...] = oopt.sgdm_update(R, dR, trR, lr, momentum)
    return gr.loss


def eval_scores(A, R, album_ctx, artist_ctx, all_albums, all_artists):
    """The ``result[1]`` of eval_step (spotify/train_spotify.py:114-119): affinity of every track."""
    dt = A.dtype
    ctx = get_embeddings(A, R, album_ctx, artist_ctx)
    cand = get_embeddings(A, R, all_albums, all_artists)
    aff = np.max(cand @ ctx.T, axis=-1)
    aff = aff + dt.type(0.1) * np.isin(all_albums, album_ctx)
    aff = aff + dt.type(0.1) * np.isin(all_artists, artist_ctx)
    return aff.astype(dt)


def eval_step(A, R, y, all_tracks, all_albums, all_artists, k=500):
    """eval_step (spotify/train_spotify.py:113-131): top-k recall for tracks and artists.

    ``jax.lax.top_k`` returns the k largest, ties broken by LOWER index first.
    """
    aff = eval_scores(A, R, y["album_context"], y["artist_context"], all_albums, all_artists)
    order = np.lexsort((np.arange(aff.shape[0]), -aff.astype(np.float64)))[:k]
    top_tracks = all_tracks[order]
    top_artists = all_artists[order]
    t = np.sum(np.isin(top_tracks, y["next_track"])).astype(np.float32)
    a = np.sum(np.isin(top_artists, y["next_artist"])).astype(np.float32)
    return np.stack([t / y["next_track"].shape[0], a / y["next_artist"].shape[0]]), order
