"""Initializers: shapes and distributions as flax documents them; the random STREAM is numpy's, not threefry."""
import torch as _t

from jax.numpy import FLOAT as _FLOAT


def zeros(key, shape, dtype=None):
    return _t.zeros(shape, dtype=_FLOAT)


def default_embed_init(key, shape, dtype=None):
    # variance_scaling(1.0, 'fan_in', 'normal', out_axis=0): fan_in = shape[1]  ->  N(0, 1/features)
    return _t.tensor(key.standard_normal(shape) / (shape[1] ** 0.5), dtype=_FLOAT)
