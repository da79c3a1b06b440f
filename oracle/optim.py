"""Oracle for the optimizer rules (TEST INFRASTRUCTURE).

The rules live in optax, a third-party dependency that is NOT vendored in
/root/reference and NOT importable here:

* optax 0.1.2  (wikipedia/requirements.txt:21, pinterest/requirements.txt:8)
* optax 0.1.5  (spotify/requirements.txt:53)

Call sites in the reference: ``optax.adam(lr)`` wikipedia/train_cooccurence.py:171
and pinterest/train_shop_the_look.py:175; ``optax.sgd(lr, momentum)``
spotify/train_spotify.py:238-241; ``TrainState.apply_gradients``
wikipedia/train_cooccurence.py:101, spotify/train_spotify.py:110.
``optax.adagrad`` is the north-star sparse rule (BASELINE.json); the reference
never calls it.

Restated from the published optax algorithms.  Parity with
real optax is unpinned (not installable here, no reference golden vectors):
tests/test_oracle_optim.py checks these against torch.optim (Adam / SGD momentum /
Adagrad), which implement the same published rules, and tests/test_ref_golden.py
against three-step trajectories driven by the reference's own update_model /
train_step (whose optax calls resolve to an independent restatement under
tests/golden/refshim/optax).
"""
from __future__ import annotations

import numpy as np

ADAM_B1 = 0.9
ADAM_B2 = 0.999
ADAM_EPS = 1e-8
ADAGRAD_INIT_ACC = 0.1
ADAGRAD_EPS = 1e-7


def adam_update(p, g, mu, nu, count, lr, b1=ADAM_B1, b2=ADAM_B2, eps=ADAM_EPS):
    """optax.adam: scale_by_adam(b1,b2,eps,eps_root=0) -> scale(-lr).  Dense.

    ``count`` is the step count BEFORE this update.  Returns (p, mu, nu, count+1).
    """
    dt = p.dtype
    count = count + 1
    mu = (b1 * mu + (1.0 - b1) * g).astype(dt)
    nu = (b2 * nu + (1.0 - b2) * (g * g)).astype(dt)
    c1 = dt.type(1.0 - b1 ** count)
    c2 = dt.type(1.0 - b2 ** count)
    mhat = mu / c1
    vhat = nu / c2
    p = (p - dt.type(lr) * mhat / (np.sqrt(vhat) + dt.type(eps))).astype(dt)
    return p, mu, nu, count


def sgdm_update(p, g, trace, lr, momentum):
    """optax.sgd(lr, momentum): trace(decay=momentum, nesterov=False) -> scale(-lr)."""
    dt = p.dtype
    trace = (g + dt.type(momentum) * trace).astype(dt)
    p = (p - dt.type(lr) * trace).astype(dt)
    return p, trace


def adagrad_update(p, g, acc, lr, eps=ADAGRAD_EPS):
    """optax.adagrad(lr, initial_accumulator_value=0.1, eps=1e-7).

    ``acc += g^2; p -= lr * g * rsqrt(acc + eps)`` (0 where acc == 0).  Rows with
    g == 0 do not move, so applying it to the touched rows only equals the dense
    rule exactly.
    """
    dt = p.dtype
    acc = (acc + g * g).astype(dt)
    inv = np.where(acc > 0, 1.0 / np.sqrt(acc + dt.type(eps)), 0.0).astype(dt)
    p = (p - dt.type(lr) * g * inv).astype(dt)
    return p, acc
