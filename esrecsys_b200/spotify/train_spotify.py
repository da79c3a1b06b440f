"""Hot-path functions of spotify/train_spotify.py with the same names and argument meaning:
``train_step`` (:77-111), ``eval_step`` (:113-131), ``sample_negative`` (:139-150).

``train_step`` runs the fused forward + backward of the reference loss for a pack of playlists
(``SpotifyModel.loss_and_grads``) and applies the reference's optimizer through ``TrainState``;
``eval_step`` scores every track against the 5 context rows on the device and ranks them with libesr.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine
from ..train_state import TrainState
from .models import SpotifyModel


def train_step(state: TrainState, model: SpotifyModel, examples, regularization=10.0):
    """train_spotify.py:77-111 for a list of playlist dicts (the reference passes one): returns (state, loss[P])."""
    loss, grads = model.loss_and_grads(state.params, examples, regularization)
    return state.apply_gradients(grads=grads), loss


def eval_scores(model: SpotifyModel, params, y, all_albums, all_artists):
    """``result[1]`` of eval_step (:114-119): affinity of every track of the corpus to the playlist context,
    ``max_k(track . ctx_k) + 0.1 isin(album, album_ctx) + 0.1 isin(artist, artist_ctx)`` (spotify/models.py:78-80)."""
    dev = params["album_embed"]["embedding"].device
    ctx = model.get_embeddings(params, y["album_context"], y["artist_context"])            # (5, 2F)
    cand = model.get_embeddings(params, all_albums, all_artists)                           # (N, 2F)
    scores = engine.score_all(engine.EmbeddingTable.wrap(cand.contiguous()), ctx.contiguous())    # (N, 5), one table scan
    aff = scores.max(dim=1).values
    alb = torch.as_tensor(np.asarray(all_albums)).to(dev)
    art = torch.as_tensor(np.asarray(all_artists)).to(dev)
    aff = aff + 0.1 * torch.isin(alb, torch.as_tensor(np.asarray(y["album_context"])).to(dev)).float()
    aff = aff + 0.1 * torch.isin(art, torch.as_tensor(np.asarray(y["artist_context"])).to(dev)).float()
    return aff


def eval_topk(model: SpotifyModel, params, y, all_albums, all_artists, k=500):
    """``jax.lax.top_k(result[1], k)`` of eval_step (:114-120) as ONE fused pass over the corpus: every track's row is
    gathered on the fly (``concat(album_embed[album % 100000], artist_embed[artist])``, spotify/models.py:33-46), scored
    against the 5 context rows, max-reduced, boosted by the two raw-id ``isin`` terms (:78-80) and kept only if it beats
    the running top-k -- no (N, 2F) candidate matrix, no (N, 5) score matrix, no sort of N keys.
    ``all_albums`` / ``all_artists`` may be device int32 tensors (kept resident across the 1000s of eval calls)."""
    A, R = params["album_embed"]["embedding"], params["artist_embed"]["embedding"]
    dev = A.device

    def ids(x):
        return x if torch.is_tensor(x) and x.device == dev and x.dtype == torch.int32 else \
            torch.as_tensor(np.asarray(x)).to(dev, torch.int32).contiguous()
    ctx = model.get_embeddings(params, y["album_context"], y["artist_context"]).contiguous()      # (5, 2F)
    val, idx = engine.topk_scan(A, ctx, k, rows_b=R, idx_a=ids(all_albums), idx_b=ids(all_artists), mod_a=A.shape[0],
                                max_over_queries=True, ctx_a=ids(y["album_context"]), ctx_b=ids(y["artist_context"]),
                                boost=0.1)
    return val[0], idx[0]


def eval_step(model: SpotifyModel, params, y, all_tracks, all_albums, all_artists, k=500):
    """train_spotify.py:113-131: recall of the next tracks / artists among the top-k tracks.  Returns
    (metrics f32[2], top_k_indices)."""
    _, top = eval_topk(model, params, y, all_albums, all_artists, k)                        # jax.lax.top_k
    dev = top.device
    top_l = top.long()
    top_tracks = torch.as_tensor(np.asarray(all_tracks)).to(dev)[top_l]
    top_artists = torch.as_tensor(np.asarray(all_artists)).to(dev)[top_l]
    nt = torch.as_tensor(np.asarray(y["next_track"])).to(dev)
    na = torch.as_tensor(np.asarray(y["next_artist"])).to(dev)
    t = torch.isin(top_tracks, nt).sum().float() / nt.numel()
    a = torch.isin(top_artists, na).sum().float() / na.numel()
    return torch.stack([t, a]), top


def sample_negative(x, generator: torch.Generator, num_negatives, all_tracks, all_albums, all_artists):
    """train_spotify.py:139-150: uniform negatives, ``randint(0, N - 1)`` -- the upper bound is EXCLUSIVE in
    jax.random.randint, so the last track is never drawn (kept).  The stream is torch's, not threefry: parity
    harnesses pass the negatives in as inputs."""
    n = len(all_tracks)
    idx = torch.randint(0, n - 1, (num_negatives,), generator=generator).numpy()
    out = dict(x)
    out["neg_track"] = np.asarray(all_tracks)[idx]
    out["neg_album"] = np.asarray(all_albums)[idx]
    out["neg_artist"] = np.asarray(all_artists)[idx]
    return out


def sample_negative_device(x, seed, step, num_negatives, all_tracks_d, all_albums_d, all_artists_d):
    """sample_negative with the draw and the three gathers on the GPU: ``all_*_d`` are device
    tensors of the corpus arrays; returns ``x`` plus device ``neg_track / neg_album / neg_artist``."""
    n = all_tracks_d.numel()
    idx = engine.sample_uniform(seed, step, num_negatives, n - 1, all_tracks_d.device).long()
    out = dict(x)
    out["neg_track"], out["neg_album"], out["neg_artist"] = all_tracks_d[idx], all_albums_d[idx], all_artists_d[idx]
    return out
