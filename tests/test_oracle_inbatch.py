"""oracle/inbatch.py against torch float64 autograd and against oracle.stl's square-batch forms."""
import numpy as np
import pytest
import torch

from oracle import inbatch as oib
from oracle import stl as ostl


def test_bf16_round_matches_torch():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 10.0 ** rng.integers(-6, 6, 4096),
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 3.3895314e38], np.float32)])
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(oib.bf16_round(x), want)


@pytest.mark.parametrize("Bq,Bk,off", [(12, 12, 0), (8, 20, 5)])
@pytest.mark.parametrize("kind", ["hinge", "softmax"])
def test_inbatch_vs_autograd(kind, Bq, Bk, off):
    rng = np.random.default_rng(1)
    Q = oib.bf16_round(rng.standard_normal((Bq, 16)).astype(np.float32))
    K = oib.bf16_round(rng.standard_normal((Bk, 16)).astype(np.float32))
    margin, scale, bn = 0.7, 0.5, 37.0
    Qt, Kt = torch.tensor(Q, dtype=torch.float64, requires_grad=True), torch.tensor(K, dtype=torch.float64, requires_grad=True)
    S = scale * (Qt @ Kt.T)
    pos = S[torch.arange(Bq), torch.arange(Bq) + off]
    if kind == "hinge":
        notpos = torch.ones(Bq, Bk, dtype=torch.float64)
        notpos[torch.arange(Bq), torch.arange(Bq) + off] = 0
        tl = (torch.relu(margin + S - pos[:, None]) * notpos).sum() / bn
        loss, dQ, dK, _ = oib.hinge(Q, K, off, margin, scale, bn)
        tol = 1e-6
    else:
        tl = (torch.logsumexp(S, 1) - pos).sum() / bn
        loss, dQ, dK, _ = oib.softmax(Q, K, off, scale, bn)
        tol = 2.0 ** -8   # the contract rounds the probabilities to bf16
    tl.backward()
    assert abs(loss - tl.item()) < 1e-5 * max(1, abs(tl.item()))
    assert np.abs(dQ - Qt.grad.numpy()).max() <= tol * max(1e-3, np.abs(Qt.grad.numpy()).max())
    assert np.abs(dK - Kt.grad.numpy()).max() <= tol * max(1e-3, np.abs(Kt.grad.numpy()).max())


def test_collapses_to_square_forms():
    rng = np.random.default_rng(2)
    Q = oib.bf16_round(rng.standard_normal((24, 8)).astype(np.float32))
    K = oib.bf16_round(rng.standard_normal((24, 8)).astype(np.float32))
    l1, dQ1, dK1, _ = oib.hinge(Q, K)
    l2, dQ2, dK2 = ostl.inbatch_hinge(Q.astype(np.float64), K.astype(np.float64))
    assert abs(l1 - l2) < 1e-5 and np.abs(dQ1 - dQ2).max() < 1e-6 and np.abs(dK1 - dK2).max() < 1e-6
    l1, *_ = oib.softmax(Q, K)
    l2, *_ = ostl.inbatch_softmax(Q.astype(np.float64), K.astype(np.float64))
    assert abs(l1 - l2) < 1e-5
