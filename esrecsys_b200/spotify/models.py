"""``SpotifyModel`` with the class surface of spotify/models.py:23-91 on libesr.

Same constructor field (``feature_size``), same tables (``album_embed`` (100000,F) hashed by mod,
``artist_embed`` (295861,F) -- models.py:27-31), same param tree
(``{'params': {'album_embed': {'embedding'}, 'artist_embed': {'embedding'}}}``), ``get_embeddings`` and
``__call__`` with the 9 positional id arrays and the 6-tuple return.  ``loss_and_grads`` is the fused
forward+backward of train_step's loss (spotify/train_spotify.py:77-109) for a PACK of playlists.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import engine
from ..train_state import RowGrads

MAX_ALBUMS = 100000     # models.py:29
NUM_ARTISTS = 295861    # models.py:31


def _i32(x, dev):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(dev).to(torch.int32).reshape(-1).contiguous()


class SpotifyModel:
    """Spotify model that takes a context and predicts the next tracks."""

    def __init__(self, feature_size: int, max_albums: int = MAX_ALBUMS, num_artists: int = NUM_ARTISTS):
        self.feature_size = int(feature_size)
        self.max_albums, self.num_artists = int(max_albums), int(num_artists)

    def init(self, key, *ids, device=None):
        L.require_cuda()
        dev = torch.device(device if device is not None else "cuda")
        gen = key if isinstance(key, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(key))
        F = self.feature_size
        mk = lambda n: (torch.randn(n, F, generator=gen) / np.sqrt(F)).to(dev)
        return {"params": {"album_embed": {"embedding": mk(self.max_albums)},
                           "artist_embed": {"embedding": mk(self.num_artists)}}}

    def apply(self, variables, *ids, method=None):
        params = variables["params"]
        if method is not None:
            return getattr(self, getattr(method, "__name__", method))(params, *ids)
        return self(params, *ids)

    def get_embeddings(self, params, album, artist):
        """models.py:33-46: concat(album_embed[album % max_albums], artist_embed[artist])."""
        A = engine.EmbeddingTable.wrap(params["album_embed"]["embedding"])
        R = engine.EmbeddingTable.wrap(params["artist_embed"]["embedding"])
        album = _i32(album, A.device)
        artist = _i32(artist, A.device)
        return torch.cat([A.gather(torch.remainder(album, self.max_albums)), R.gather(artist)], dim=-1)

    def pack(self, params, examples, regularization=10.0, want_forward=False):
        """Fused forward + backward for a list of playlists (dicts with the keys of
        spotify/input_pipeline.py:23-30 plus neg_*).  Returns dict(loss[P], dXa, dXr, album_rows,
        artist_rows, row_base[P+1], pos_aff, neg_aff, l2)."""
        A = params["album_embed"]["embedding"]
        R = params["artist_embed"]["embedding"]
        dev = A.device
        P = len(examples)
        nc = len(examples[0]["album_context"])
        o = len(examples[0]["neg_album"])
        ms = [len(x["next_album"]) for x in examples]
        off = np.zeros(P + 1, np.int32)
        off[1:] = np.cumsum(ms)
        cat = lambda k: _i32(np.concatenate([np.asarray(x[k]).reshape(-1) for x in examples]), dev)
        T = P * (nc + o) + int(off[-1])
        F = self.feature_size
        out = dict(loss=torch.empty(P, device=dev), dXa=torch.empty(T, F, device=dev), dXr=torch.empty(T, F, device=dev),
                   album_rows=torch.empty(T, dtype=torch.int32, device=dev),
                   artist_rows=torch.empty(T, dtype=torch.int32, device=dev),
                   row_base=np.arange(P + 1) * (nc + o) + off, next_off=off)
        fw = want_forward
        out["pos_aff"] = torch.empty(int(off[-1]), device=dev) if fw else None
        out["neg_aff"] = torch.empty(P * o, device=dev) if fw else None
        out["l2"] = torch.empty(T, device=dev) if fw else None
        ids = {k: cat(k) for k in ("album_context", "artist_context", "next_album", "next_artist", "neg_album",
                                   "neg_artist")}      # keep the device id arrays alive across the launch
        ids["off"] = torch.from_numpy(off).to(dev)
        out["_ids"] = ids
        L.check(L.lib().esr_spotify_fwd_bwd_f32(
            L.ptr(A), A.shape[0], L.ptr(R), F, P, nc, o, int(max(ms)), L.ptr(ids["album_context"]),
            L.ptr(ids["artist_context"]), L.ptr(ids["next_album"]), L.ptr(ids["next_artist"]),
            L.ptr(ids["off"]), L.ptr(ids["neg_album"]), L.ptr(ids["neg_artist"]), float(regularization),
            L.ptr(out["loss"]), L.ptr(out["dXa"]), L.ptr(out["dXr"]), L.ptr(out["album_rows"]), L.ptr(out["artist_rows"]),
            L.ptr(out["pos_aff"]), L.ptr(out["neg_aff"]), L.ptr(out["l2"]), L.stream_ptr()), "esr_spotify_fwd_bwd_f32")
        return out

    def __call__(self, params, track_context, album_context, artist_context, next_track, next_album, next_artist,
                 neg_track, neg_album, neg_artist):
        """models.py:48-91.  Track ids are unused by the model (as in the reference)."""
        x = dict(album_context=album_context, artist_context=artist_context, next_album=next_album,
                 next_artist=next_artist, neg_album=neg_album, neg_artist=neg_artist)
        r = self.pack(params, [x], want_forward=True)
        ctx = self.get_embeddings(params, album_context, artist_context)
        nxt = self.get_embeddings(params, next_album, next_artist)
        neg = self.get_embeddings(params, neg_album, neg_artist)
        # the three small self-affinity grams are only returned for inspection here (the training loss
        # consumes them inside the fused kernel); library GEMM
        g_ctx = torch.flip(ctx, [0]) @ ctx.T
        g_nxt = torch.flip(nxt, [0]) @ nxt.T
        g_neg = torch.flip(neg, [0]) @ neg.T
        return r["pos_aff"], r["neg_aff"], g_ctx, g_nxt, g_neg, r["l2"]

    def loss_and_grads(self, params, examples, regularization=10.0):
        """value_and_grad of train_step's loss summed over the pack -> (loss[P], grads pytree of RowGrads)."""
        r = self.pack(params, examples, regularization)
        dev = r["loss"].device
        grads = {}
        for name, rows, g, V in (("album_embed", r["album_rows"], r["dXa"], self.max_albums),
                                 ("artist_embed", r["artist_rows"], r["dXr"], self.num_artists)):
            T = rows.numel()
            plan = engine.IndexPlan(T, V, dev, with_partner=False, sort="wide").build(rows)
            gsum = torch.empty(T, self.feature_size, device=dev)
            import ctypes as C
            L.check(L.lib().esr_segment_sum_rows_f32(C.byref(plan.s), self.feature_size, L.ptr(g), None, L.ptr(gsum), None,
                                                     L.stream_ptr()), "esr_segment_sum_rows_f32")
            grads[name] = {"embedding": RowGrads(V, plan.uniq, plan.n_uniq, gsum)}
        return r["loss"], grads
