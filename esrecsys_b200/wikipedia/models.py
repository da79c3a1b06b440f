"""``Glove`` with the class surface of wikipedia/models.py:8-55, executing on libesr.

Same constructor fields and defaults (``num_embeddings=1024, features=64`` -- models.py:12-13), same
param tree (``{'params': {'_token_embedding': {'embedding': (V,D)}, '_bias': {'embedding': (V,1)}}}`` --
flax names submodules after the attribute, models.py:16-19), ``init`` / ``apply`` / ``score_all``.
Parameters are float32 torch CUDA tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L
from .. import engine


class Glove:
    """A simple embedding model based on gloVe (wikipedia/models.py:8-11)."""

    def __init__(self, num_embeddings: int = 1024, features: int = 64):
        self.num_embeddings = int(num_embeddings)
        self.features = int(features)

    # -- flax-like functional surface ---------------------------------------------------------
    def init(self, key, x=None, device=None):
        """``model.init(key, x)`` (train_cooccurence.py:170).  ``key``: int seed or torch.Generator.
        nn.Embed default init: N(0, 1/features) (variance_scaling fan_in, out_axis=0); bias zeros (models.py:19)."""
        L.require_cuda()
        dev = torch.device(device if device is not None else "cuda")
        gen = key if isinstance(key, torch.Generator) else torch.Generator(device="cpu").manual_seed(int(key))
        E = torch.randn(self.num_embeddings, self.features, generator=gen, dtype=torch.float32) / np.sqrt(self.features)
        return {"params": {"_token_embedding": {"embedding": E.to(dev)},
                           "_bias": {"embedding": torch.zeros(self.num_embeddings, 1, dtype=torch.float32, device=dev)}}}

    def apply(self, variables, inputs, method=None):
        params = variables["params"]
        if method is not None:
            name = getattr(method, "__name__", method)
            if name == "score_all":
                return self.score_all(params, inputs)
            if name != "__call__":
                raise AttributeError(name)
        return self(params, inputs)

    def table(self, params):
        return engine.EmbeddingTable.wrap(params["_token_embedding"]["embedding"], params["_bias"]["embedding"])

    def __call__(self, params, inputs):
        """models.py:21-38: ``dot + bias1 + bias2`` with the reference's ``(B,)+(B,1)+(B,1) -> (B,B)`` broadcast."""
        t = self.table(params)
        token1, token2 = inputs[0], inputs[1]
        token1 = torch.as_tensor(token1, device=t.device).to(torch.int32)
        token2 = torch.as_tensor(token2, device=t.device).to(torch.int32)
        embed1 = t.gather(token1)
        embed2 = t.gather(token2)
        bias1 = params["_bias"]["embedding"][token1.long()]          # (B,1) index plumbing
        bias2 = params["_bias"]["embedding"][token2.long()]
        dot = engine.rowwise_dot(embed1, embed2)                     # (B,)
        return dot + bias1 + bias2

    def score_all(self, params, token):
        """models.py:40-55: score of ``token`` (int array (T,)) vs all tokens -> (V, T)."""
        t = self.table(params)
        token = torch.as_tensor(token, device=t.device).to(torch.int32).reshape(-1)
        return engine.score_all(t, t.gather(token))
