"""``flax.training.train_state.TrainState`` shim (wikipedia/train_cooccurence.py:172,
spotify/train_spotify.py:242-243): ``step``, ``apply_fn``, ``params``, ``tx``, ``opt_state``,
``create()``, ``apply_gradients()``.  Parameters are torch CUDA tensors updated IN PLACE by libesr
kernels (the reference rebinds a new pytree; nothing on the hot path depends on that).

Gradients are ``RowGrads`` (rows that received gradient + their values) instead of the dense pytree
``jax.value_and_grad`` returns; ``RowGrads.dense()`` materialises the reference's dense array.
"""
from __future__ import annotations

import torch

from . import engine
from . import optim as O


class RowGrads:
    """Gradient of one embedding table: rows ``uniq[:n]`` got ``g[:n]``; every other row is zero."""

    def __init__(self, V, uniq, n_uniq, g):
        self.V, self.uniq, self.n_uniq, self.g = int(V), uniq, n_uniq, g

    def dense(self, out=None):
        """The dense array of the reference's grads pytree (zero rows included)."""
        shape = (self.V,) + tuple(self.g.shape[1:])
        if out is None:
            out = torch.zeros(shape, dtype=torch.float32, device=self.g.device)
        else:
            out.zero_()
        g2 = self.g if self.g.dim() == 2 else self.g.reshape(-1, 1)
        o2 = out if out.dim() == 2 else out.reshape(-1, 1)
        if g2.shape[1] % 4 == 0:
            engine.scatter_rows(o2, self.uniq, self.n_uniq, g2)
        else:   # (V,1) bias tables: width-1 rows are below the 16-byte row kernels; plain index copy
            n = int(self.n_uniq.item())
            o2[self.uniq[:n].long()] = g2[:n]
        return out


class TrainState:
    def __init__(self, step, apply_fn, params, tx, opt_state):
        self.step, self.apply_fn, self.params, self.tx, self.opt_state = step, apply_fn, params, tx, opt_state

    @classmethod
    def create(cls, *, apply_fn, params, tx):
        leaves = _leaves(params)
        if tx.kind == "adam":
            opt = {"count": 0, "mu": {k: torch.zeros_like(v) for k, v in leaves.items()},
                   "nu": {k: torch.zeros_like(v) for k, v in leaves.items()}}
        elif tx.kind == "sgd":
            opt = {"trace": {k: torch.zeros_like(v) for k, v in leaves.items()}}
        elif tx.kind == "adagrad":
            opt = {"acc": {k: torch.full_like(v, tx.initial_accumulator_value) for k, v in leaves.items()}}
        else:
            raise ValueError(tx)
        return cls(0, apply_fn, params, tx, opt)

    def apply_gradients(self, *, grads):
        """``grads``: dict leaf-path -> RowGrads | dense tensor, same tree as ``params``."""
        leaves = _leaves(self.params)
        gl = _leaves(grads)
        tx = self.tx
        if tx.kind == "adam":
            self.opt_state["count"] += 1
        for k, p in leaves.items():
            g = gl[k]
            gd = g.dense() if isinstance(g, RowGrads) else g
            if tx.kind == "adam":
                engine.dense_adam(p, gd, self.opt_state["mu"][k], self.opt_state["nu"][k], tx.learning_rate,
                                  self.opt_state["count"], tx.b1, tx.b2, tx.eps)
            elif tx.kind == "sgd":
                engine.dense_sgdm(p, gd, self.opt_state["trace"][k], tx.learning_rate, tx.momentum)
            else:
                _adagrad_rows(p, self.opt_state["acc"][k], g, tx)
        self.step += 1
        return self


def _adagrad_rows(p, acc, g, tx):
    """optax.adagrad restricted to the rows with gradient (identical to the dense rule: rows with g == 0 do
    not move)."""
    if not isinstance(g, RowGrads) or p.dim() != 2 or p.shape[1] % 4:
        gd = g.dense() if isinstance(g, RowGrads) else g
        acc.add_(gd * gd)          # (V,1) bias / dense fallback: tiny elementwise plumbing
        p.sub_(tx.learning_rate * gd * torch.where(acc > 0, torch.rsqrt(acc + tx.eps), torch.zeros_like(acc)))
        return
    t = engine.EmbeddingTable.wrap(p, acc=acc)
    engine.sparse_adagrad(t, g.uniq, g.n_uniq, g.g, None, tx.learning_rate, tx.eps)


def _leaves(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        path = prefix + "/" + k if prefix else k
        if isinstance(v, dict):
            out.update(_leaves(v, path))
        else:
            out[path] = v
    return out
