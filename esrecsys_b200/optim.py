"""Optimizer descriptors with optax's constructor names and defaults.

``adam(lr)``  -- optax.adam: wikipedia/train_cooccurence.py:171, pinterest/train_shop_the_look.py:175
``sgd(lr, momentum)`` -- optax.sgd: spotify/train_spotify.py:238-241
``adagrad(lr)`` -- optax.adagrad (north-star sparse rule; the reference never calls it)

They only carry hyper-parameters; the arithmetic is in libesr (esr_dense_adam_f32, esr_dense_sgdm_f32,
esr_sparse_adagrad_f32 / the fused row pass).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Adam:
    learning_rate: float
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    kind: str = "adam"


@dataclass(frozen=True)
class Sgd:
    learning_rate: float
    momentum: float = 0.0
    kind: str = "sgd"


@dataclass(frozen=True)
class Adagrad:
    learning_rate: float
    initial_accumulator_value: float = 0.1
    eps: float = 1e-7
    kind: str = "adagrad"


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8):
    return Adam(float(learning_rate), b1, b2, eps)


def sgd(learning_rate, momentum=0.0):
    return Sgd(float(learning_rate), float(momentum or 0.0))


def adagrad(learning_rate, initial_accumulator_value=0.1, eps=1e-7):
    return Adagrad(float(learning_rate), initial_accumulator_value, eps)
