"""``find_top_k`` of pinterest/make_recommendations.py:49-65 on libesr: one fused scan of the product table scores every
product against the scene embedding and keeps the running top-k (``esr_topk_scan_f32``; ``jax.lax.top_k`` order)."""
from __future__ import annotations

import torch

from .. import engine


def find_top_k(scene_embedding, product_embeddings, k):
    """Returns (scores[k], indices[k]), best first, ties by lower index."""
    p = torch.as_tensor(product_embeddings).to("cuda", torch.float32).contiguous()
    s = torch.as_tensor(scene_embedding).to("cuda", torch.float32).reshape(1, -1).contiguous()
    val, idx = engine.topk_scan(p, s, k)                                 # (1, k): no (N, 1) score vector, no sort of N keys
    return val[0], idx[0]
