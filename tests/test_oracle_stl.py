"""Pins oracle.stl against torch float64 autograd."""
import numpy as np
import pytest
import torch

from oracle import stl as ostl


def _emb(B=16, D=32, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((B, D)) * scale / np.sqrt(D) * 1.5 for _ in range(3)]


def _torch_triplet(s, p, n, reg, bs):
    st, pt, nt = [torch.tensor(a, requires_grad=True) for a in (s, p, n)]
    pos = (st * pt).sum(-1)
    neg = (st * nt).sum(-1)
    trip = torch.relu(1.0 + neg - pos).sum()
    rf = lambda e: torch.relu(torch.sqrt((e ** 2).sum(-1)) - 1.0)
    loss = (trip + reg * (rf(st) + rf(pt) + rf(nt)).sum()) / bs
    loss.backward()
    return loss.item(), st.grad.numpy(), pt.grad.numpy(), nt.grad.numpy()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_triplet_loss_and_grads(seed):
    s, p, n = _emb(seed=seed)
    loss, ds, dp, dn = ostl.triplet_loss_and_grads(s, p, n, 0.1, 16)
    tl, ts, tp, tn = _torch_triplet(s, p, n, 0.1, 16)
    assert np.isclose(loss, tl, rtol=1e-12)
    for a, b in ((ds, ts), (dp, tp), (dn, tn)):
        assert np.abs(a - b).max() < 1e-14


def test_eval_loss_and_topk():
    s, p, n = _emb(seed=3)
    ps, ns = ostl.scores(s, p, n)
    assert np.isclose(ostl.eval_loss(s, p, n), np.maximum(1 + ns - ps, 0).sum())
    sc, idx = ostl.find_top_k(s[0], p, 5)
    full = (s[0] * p).sum(-1)
    assert np.allclose(sc, np.sort(full)[::-1][:5]) and np.allclose(full[idx], sc)


@pytest.mark.parametrize("fn,tfn", [("inbatch_hinge", "hinge"), ("inbatch_softmax", "softmax")])
def test_inbatch_losses_vs_autograd(fn, tfn):
    rng = np.random.default_rng(4)
    Q = rng.standard_normal((12, 8))
    K = rng.standard_normal((12, 8))
    loss, dQ, dK = getattr(ostl, fn)(Q, K)
    Qt, Kt = torch.tensor(Q, requires_grad=True), torch.tensor(K, requires_grad=True)
    S = Qt @ Kt.T
    B = 12
    if tfn == "hinge":
        d = torch.diag(S)[:, None]
        M = torch.relu(1.0 + S - d) * (1 - torch.eye(B, dtype=torch.float64))
        tl = M.sum() / B
    else:
        tl = (torch.logsumexp(S, 1) - torch.diag(S)).sum() / B
    tl.backward()
    assert np.isclose(loss, tl.item(), rtol=1e-12)
    assert np.abs(dQ - Qt.grad.numpy()).max() < 1e-13 and np.abs(dK - Kt.grad.numpy()).max() < 1e-13


def test_mlp_tower_shape():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4, 6))
    out = ostl.mlp_tower(x, rng.standard_normal((6, 6)), np.zeros(6), rng.standard_normal((6, 3)), np.zeros(3))
    assert out.shape == (4, 3)
