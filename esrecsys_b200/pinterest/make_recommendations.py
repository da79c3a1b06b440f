"""``find_top_k`` of pinterest/make_recommendations.py:49-65 on libesr: scores of one scene embedding against
every product embedding (one scan of the product table) and ``jax.lax.top_k``."""
from __future__ import annotations

import torch

from .. import engine


def find_top_k(scene_embedding, product_embeddings, k):
    """Returns (scores[k], indices[k]), best first, ties by lower index."""
    p = torch.as_tensor(product_embeddings).to("cuda", torch.float32).contiguous()
    s = torch.as_tensor(scene_embedding).to("cuda", torch.float32).reshape(1, -1).contiguous()
    scores = engine.score_all(engine.EmbeddingTable.wrap(p), s)          # (N, 1)
    return engine.top_k(scores[:, 0].contiguous(), k)
