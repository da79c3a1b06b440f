"""Pins oracle.glove: closed form == literal (B,B) evaluation == torch float64 autograd.

The reference has no tests or golden vectors and jax cannot be
imported here, so this independent second derivation is the pin ("parity unpinned").
"""
import numpy as np
import pytest
import torch

from esrecsys_b200 import synth
from oracle import glove as og
from oracle import optim as oopt


def _case(V=50, D=8, B=16, seed=0, dtype=np.float64, zero_bias=False):
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((V, D)).astype(dtype) / np.sqrt(D).astype(dtype)
    b = (np.zeros(V) if zero_bias else rng.standard_normal(V) * 0.1).astype(dtype)
    ids, counts = synth.glove_batches(V, B, 1, seed)
    return E, b, ids[0, 0], ids[0, 1], counts[0].astype(dtype)


def _torch_literal(E, b, i, j, x):
    """Literal transcription of wikipedia/models.py:30-38 + train_cooccurence.py:76-84 in torch f64."""
    Et = torch.tensor(E, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(b.reshape(-1, 1), dtype=torch.float64, requires_grad=True)
    it = torch.tensor(i, dtype=torch.long)
    jt = torch.tensor(j, dtype=torch.long)
    xt = torch.tensor(x, dtype=torch.float64)
    e1, e2 = Et[it], Et[jt]
    b1, b2 = bt[it], bt[jt]                     # (B,1)
    dot = (e1 * e2).sum(-1)                     # (B,)
    out = dot + b1 + b2                         # (B,B)
    w = torch.minimum(torch.ones_like(xt), xt / 100.0) ** 0.75
    lt = torch.log10(1.0 + xt)
    loss = torch.mean(torch.square(lt - out) * w)
    loss.backward()
    return loss.item(), Et.grad.numpy(), bt.grad.numpy()[:, 0]


def test_micro_hand_computed():
    # B=2, D=2: everything by hand.
    E = np.array([[0., 0.], [1., 2.], [3., -1.], [0.5, 0.5]])
    b = np.array([0., 0.1, -0.2, 0.3])
    i = np.array([2, 3], np.int32)
    j = np.array([1, 2], np.int32)
    x = np.array([9.0, 999.0])
    out = og.forward_literal(E, b, i, j)
    dot = np.array([3 * 1 + -1 * 2, 0.5 * 3 + 0.5 * -1])            # [1, 1]
    bs = np.array([b[2] + b[1], b[3] + b[2]])                        # [-0.1, 0.1]
    assert np.allclose(out, dot[None, :] + bs[:, None])
    w = np.array([0.09 ** 0.75, 1.0])
    t = np.array([1.0, 3.0])
    expect = np.mean((t[None, :] - out) ** 2 * w[None, :])
    assert np.isclose(og.loss_literal(E, b, i, j, x), expect)
    assert np.isclose(og.loss_and_grads(E, b, i, j, x).loss, expect)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_closed_form_vs_literal_and_autograd(seed):
    E, b, i, j, x = _case(seed=seed)
    gr = og.loss_and_grads(E, b, i, j, x)
    assert np.isclose(gr.loss, og.loss_literal(E, b, i, j, x), rtol=1e-12)
    tl, tE, tb = _torch_literal(E, b, i, j, x)
    assert np.isclose(gr.loss, tl, rtol=1e-12)
    dE, db = og.dense_grads(E.shape[0], gr, E.shape[1])
    assert np.abs(dE - tE).max() < 1e-13
    assert np.abs(db - tb).max() < 1e-13


def test_duplicate_heavy_and_same_row_both_roles():
    # tiny vocabulary: every row is hit many times and in both roles
    E, b, i, j, x = _case(V=5, D=4, B=64, seed=3)
    gr = og.loss_and_grads(E, b, i, j, x)
    _, tE, tb = _torch_literal(E, b, i, j, x)
    dE, db = og.dense_grads(5, gr, 4)
    assert np.abs(dE - tE).max() < 1e-12 and np.abs(db - tb).max() < 1e-12
    assert gr.seg_off[-1] == 128 and gr.uniq.size <= 4        # row 0 (mask) never drawn


def test_per_pair_mode_autograd():
    E, b, i, j, x = _case(seed=4)
    gr = og.loss_and_grads(E, b, i, j, x, bias_mode="per_pair")
    Et = torch.tensor(E, requires_grad=True)
    bt = torch.tensor(b, requires_grad=True)
    it, jt = torch.tensor(i, dtype=torch.long), torch.tensor(j, dtype=torch.long)
    xt = torch.tensor(x)
    pred = (Et[it] * Et[jt]).sum(-1) + bt[it] + bt[jt]
    w = torch.minimum(torch.ones_like(xt), xt / 100.0) ** 0.75
    loss = torch.mean((torch.log10(1 + xt) - pred) ** 2 * w)
    loss.backward()
    dE, db = og.dense_grads(E.shape[0], gr, E.shape[1])
    assert np.isclose(gr.loss, loss.item(), rtol=1e-12)
    assert np.abs(dE - Et.grad.numpy()).max() < 1e-13 and np.abs(db - bt.grad.numpy()).max() < 1e-13


def test_modes_coincide_when_bias_sums_equal():
    E, b, i, j, x = _case(seed=5, zero_bias=True)
    a = og.loss_and_grads(E, b, i, j, x, "reference_broadcast")
    c = og.loss_and_grads(E, b, i, j, x, "per_pair")
    assert np.isclose(a.loss, c.loss, rtol=1e-12) and np.abs(a.dE - c.dE).max() < 1e-13


def test_batch_permutation_invariance():
    E, b, i, j, x = _case(seed=6)
    p = np.random.default_rng(0).permutation(i.size)
    a = og.loss_and_grads(E, b, i, j, x)
    c = og.loss_and_grads(E, b, i[p], j[p], x[p])
    assert np.isclose(a.loss, c.loss, rtol=1e-12)
    assert (a.uniq == c.uniq).all() and np.abs(a.dE - c.dE).max() < 1e-12


def test_fp32_closed_form_close_to_fp64():
    E, b, i, j, x = _case(V=2000, D=64, B=512, seed=7)
    a = og.loss_and_grads(E, b, i, j, x)
    c = og.loss_and_grads(E.astype(np.float32), b.astype(np.float32), i, j, x.astype(np.float32))
    assert abs(a.loss - c.loss) < 1e-5 * max(1, abs(a.loss))
    assert np.abs(a.dE - c.dE).max() < 1e-6


def test_adam_step_matches_torch_adam():
    E, b, i, j, x = _case(seed=8)
    E0, b0 = E.copy(), b.copy()
    st = dict(count=0, muE=np.zeros_like(E), nuE=np.zeros_like(E), mub=np.zeros_like(b), nub=np.zeros_like(b))
    Et = torch.tensor(E0, requires_grad=True)
    bt = torch.tensor(b0, requires_grad=True)
    opt = torch.optim.Adam([Et, bt], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(3):
        og.step_adam(E, b, st, i, j, x, 1e-3)
        opt.zero_grad()
        _, gE, gb = _torch_literal(Et.detach().numpy(), bt.detach().numpy(), i, j, x)
        Et.grad, bt.grad = torch.tensor(gE), torch.tensor(gb)
        opt.step()
    # torch divides by (sqrt(v)/sqrt(c2) + eps), optax by (sqrt(v/c2) + eps): identical up to rounding
    assert np.abs(E - Et.detach().numpy()).max() < 1e-10
    assert np.abs(b - bt.detach().numpy()).max() < 1e-10
    assert st["count"] == 3


def test_adagrad_sparse_equals_dense():
    E, b, i, j, x = _case(seed=9)
    E1, b1 = E.copy(), b.copy()
    aE, ab = np.full_like(E, 0.1), np.full_like(b, 0.1)
    og.step_adagrad(E1, b1, aE, ab, i, j, x, 0.05)
    gr = og.loss_and_grads(E, b, i, j, x)
    dE, db = og.dense_grads(E.shape[0], gr, E.shape[1])
    E2, _ = oopt.adagrad_update(E, dE, np.full_like(E, 0.1), 0.05)
    assert np.abs(E1 - E2).max() == 0.0
    untouched = np.setdiff1d(np.arange(E.shape[0]), gr.uniq)
    assert (E1[untouched] == E[untouched]).all()


def test_score_all_and_knn_order():
    E, b, i, j, x = _case(V=30, seed=10)
    toks = np.array([3, 7], np.int32)
    scores, idx = og.find_knn(E, toks)
    assert scores.shape == (30, 2) and idx.shape == (30, 2)
    top, ts = og.top_k(E, toks, 5)
    for t in range(2):
        s = E @ E[toks[t]]
        assert np.allclose(np.sort(s)[::-1][:5], ts[t])
        assert (idx[-1, t] == top[t, 0])
