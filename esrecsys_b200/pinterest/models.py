"""``STLModel`` scoring surface of pinterest/models.py:48-74 on libesr.

The reference's towers are CNNs over 512x512 JPEGs (pinterest/models.py:23-46): OUT OF SCOPE (SURVEY.md D5,
8(a) a16).  Per the north star the towers here are ID-embedding tables (scene ids / product ids -> rows);
``__call__`` keeps the reference's signature and return tuple
``(pos_score, neg_score, scene_embed, pos_product_embed, neg_product_embed)``.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import engine


class STLModel:
    """Shop the look model that takes in a scene and item and computes a score for them."""

    def __init__(self, output_size: int, num_scenes: int = 1024, num_products: int = 1024):
        self.output_size = int(output_size)
        self.num_scenes, self.num_products = int(num_scenes), int(num_products)

    def init(self, key, scene=None, pos_product=None, neg_product=None, device=None):
        L.require_cuda()
        dev = torch.device(device if device is not None else "cuda")
        gen = key if isinstance(key, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(key))
        mk = lambda n: (torch.randn(n, self.output_size, generator=gen) / np.sqrt(self.output_size)).to(dev)
        return {"params": {"scene_cnn": {"embedding": mk(self.num_scenes)},
                           "product_cnn": {"embedding": mk(self.num_products)}}}

    def _tower(self, params, name, ids):
        t = engine.EmbeddingTable.wrap(params[name]["embedding"])
        return t.gather(torch.as_tensor(ids, device=t.device).to(torch.int32))

    def get_scene_embed(self, params, scene):
        return self._tower(params, "scene_cnn", scene)           # models.py:57-58

    def get_product_embed(self, params, product):
        return self._tower(params, "product_cnn", product)       # models.py:60-61

    def apply(self, variables, *args, method=None, **kw):
        params = variables["params"]
        if method is not None:
            return getattr(self, getattr(method, "__name__", method))(params, *args)
        return self(params, *args)

    def __call__(self, params, scene, pos_product, neg_product, train: bool = True):
        scene_embed = self.get_scene_embed(params, scene)
        pos_product_embed = self.get_product_embed(params, pos_product)
        neg_product_embed = self.get_product_embed(params, neg_product)
        pos_score = engine.rowwise_dot(scene_embed, pos_product_embed)      # models.py:67-68
        neg_score = engine.rowwise_dot(scene_embed, neg_product_embed)      # models.py:71-72
        return pos_score, neg_score, scene_embed, pos_product_embed, neg_product_embed


def triplet_loss_and_grads(scene_embed, pos_embed, neg_embed, regularization, batch_size):
    """train_step's loss (pinterest/train_shop_the_look.py:99-104) and its gradient wrt the three
    embedding matrices, one fused kernel.  Returns (loss, d_scene, d_pos, d_neg, pos_score, neg_score)."""
    import ctypes as C  # noqa: F401
    s, p, n = (x.contiguous() for x in (scene_embed, pos_embed, neg_embed))
    B, D = s.shape
    dev = s.device
    ds, dp, dn = torch.empty_like(s), torch.empty_like(p), torch.empty_like(n)
    ps = torch.empty(B, device=dev)
    ns = torch.empty(B, device=dev)
    loss = torch.zeros(1, device=dev)
    ws = torch.empty(max(B, 1), device=dev)
    L.check(L.lib().esr_stl_triplet_f32(L.ptr(s), L.ptr(p), L.ptr(n), B, D, float(regularization), float(batch_size),
                                        L.ptr(ds), L.ptr(dp), L.ptr(dn), L.ptr(ps), L.ptr(ns), L.ptr(loss), L.ptr(ws),
                                        L.stream_ptr()), "esr_stl_triplet_f32")
    return loss[0], ds, dp, dn, ps, ns
