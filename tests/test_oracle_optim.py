"""Pins oracle.optim against torch.optim (same published rules as optax)."""
import numpy as np
import torch

from oracle import optim as oopt


def _run_torch(opt_cls, p0, grads, **kw):
    p = torch.tensor(p0.copy(), requires_grad=True)
    opt = opt_cls([p], **kw)
    for g in grads:
        opt.zero_grad()
        p.grad = torch.tensor(g)
        opt.step()
    return p.detach().numpy()


def test_adam():
    rng = np.random.default_rng(0)
    p = rng.standard_normal((5, 3)); grads = [rng.standard_normal((5, 3)) for _ in range(4)]
    q, mu, nu, c = p.copy(), np.zeros_like(p), np.zeros_like(p), 0
    for g in grads:
        q, mu, nu, c = oopt.adam_update(q, g, mu, nu, c, 1e-2)
    ref = _run_torch(torch.optim.Adam, p, grads, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    assert np.abs(q - ref).max() < 1e-9 and c == 4


def test_adam_zero_grad_rows_still_move():
    p = np.ones((2, 2)); mu = np.zeros_like(p); nu = np.zeros_like(p)
    g = np.array([[1.0, 1.0], [0.0, 0.0]])
    p1, mu, nu, c = oopt.adam_update(p, g, mu, nu, 0, 1e-2)
    p2, mu, nu, c = oopt.adam_update(p1, np.zeros_like(p), mu, nu, c, 1e-2)
    assert (p2[0] != p1[0]).all() and (p2[1] == p1[1]).all()


def test_sgd_momentum():
    rng = np.random.default_rng(1)
    p = rng.standard_normal(7); grads = [rng.standard_normal(7) for _ in range(5)]
    q, tr = p.copy(), np.zeros_like(p)
    for g in grads:
        q, tr = oopt.sgdm_update(q, g, tr, 1e-3, 0.98)
    ref = _run_torch(torch.optim.SGD, p, grads, lr=1e-3, momentum=0.98)
    assert np.abs(q - ref).max() < 1e-14


def test_adagrad():
    rng = np.random.default_rng(2)
    p = rng.standard_normal(9); grads = [rng.standard_normal(9) for _ in range(5)]
    q, acc = p.copy(), np.full_like(p, 0.1)
    for g in grads:
        q, acc = oopt.adagrad_update(q, g, acc, 0.05)
    # torch: p -= lr * g / (sqrt(acc) + eps); optax: p -= lr * g / sqrt(acc + eps). Differ by O(eps).
    ref = _run_torch(torch.optim.Adagrad, p, grads, lr=0.05, initial_accumulator_value=0.1, eps=0.0)
    assert np.abs(q - ref).max() < 1e-6
    z, acc0 = oopt.adagrad_update(np.ones(3), np.zeros(3), np.zeros(3), 0.1)
    assert (z == 1).all() and (acc0 == 0).all()
