"""Golden vectors (tests/golden/*.npz, generator tests/golden/make_golden.py): a float64 torch-autograd
transcription of the reference's forward + loss, statement by statement.  CPU tests pin the oracle to them;
GPU tests pin the CUDA path (through the C ABI) to the same files.  fp32 paths: 1e-5 abs + rel."""
import os

import numpy as np
import pytest
import torch

from oracle import glove as og
from oracle import optim as oopt
from oracle import spotify as osp
from oracle import stl as ostl

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GLOVE = ["glove_V60_D8_B32.npz", "glove_V500_D64_B256.npz"]


def _load(name):
    return dict(np.load(os.path.join(G, name)))


def test_fixtures_present_and_regenerable():
    for f in GLOVE + ["stl_B16_D32.npz", "spotify_m7.npz", "make_golden.py"]:
        assert os.path.exists(os.path.join(G, f)), f


@pytest.mark.parametrize("name", GLOVE)
def test_oracle_glove_vs_golden(name):
    g = _load(name)
    E, b = g["E"].astype(np.float64), g["b"].astype(np.float64)
    gr = og.loss_and_grads(E, b, g["i"], g["j"], g["x"].astype(np.float64))
    assert abs(gr.loss - g["loss"]) < 1e-12
    dE, db = og.dense_grads(E.shape[0], gr, E.shape[1])
    assert np.abs(dE - g["dE"]).max() < 1e-13 and np.abs(db - g["db"]).max() < 1e-13
    # fp32 oracle (what the GPU is compared with) stays within the fp32 budget of the f64 golden
    gr32 = og.loss_and_grads(g["E"], g["b"], g["i"], g["j"], g["x"])
    dE32, _ = og.dense_grads(E.shape[0], gr32, E.shape[1])
    np.testing.assert_allclose(dE32, g["dE"], rtol=1e-4, atol=1e-6)
    # optimizer rules
    Ea, ba = E.copy(), b.copy()
    accE, accb = np.full_like(Ea, 0.1), np.full_like(ba, 0.1)
    og.step_adagrad(Ea, ba, accE, accb, g["i"], g["j"], g["x"].astype(np.float64), 0.05)
    assert np.abs(Ea - g["adagrad_E"]).max() < 1e-12 and np.abs(accE - g["adagrad_acc_E"]).max() < 1e-12
    assert np.abs(ba - g["adagrad_b"]).max() < 1e-12
    Em, bm = E.copy(), b.copy()
    st = dict(count=0, muE=np.zeros_like(Em), nuE=np.zeros_like(Em), mub=np.zeros_like(bm), nub=np.zeros_like(bm))
    og.step_adam(Em, bm, st, g["i"], g["j"], g["x"].astype(np.float64), 1e-3)
    assert np.abs(Em - g["adam_E"]).max() < 1e-10 and np.abs(bm - g["adam_b"]).max() < 1e-10


def test_oracle_stl_vs_golden():
    g = _load("stl_B16_D32.npz")
    s, p, n = (g[k].astype(np.float64) for k in ("scene", "pos", "neg"))
    loss, ds, dp, dn = ostl.triplet_loss_and_grads(s, p, n, 0.1, 16)
    assert abs(loss - g["loss"]) < 1e-12
    for got, k in ((ds, "d_scene"), (dp, "d_pos"), (dn, "d_neg")):
        assert np.abs(got - g[k]).max() < 1e-13


def test_oracle_spotify_vs_golden():
    g = _load("spotify_m7.npz")
    A, R = g["A"].astype(np.float64), g["R"].astype(np.float64)
    args = [g[k] for k in ("album_context", "artist_context", "next_album", "next_artist", "neg_album", "neg_artist")]
    out = osp.forward(A, R, *args)
    assert np.abs(out[0] - g["pos_aff"]).max() < 1e-12 and np.abs(out[1] - g["neg_aff"]).max() < 1e-12
    assert np.abs(out[5] - g["l2"]).max() < 1e-12
    gr = osp.loss_and_grads(A, R, *args, float(g["reg"]))
    assert abs(gr.loss - g["loss"]) < 1e-12
    dA, dR = osp.dense_grads(A, R, gr)
    assert np.abs(dA - g["dA"]).max() < 1e-12 and np.abs(dR - g["dR"]).max() < 1e-12


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", GLOVE)
def test_cuda_glove_vs_golden(name):
    from esrecsys_b200 import _lib as L, engine
    g = _load(name)
    V, D = g["E"].shape
    B = g["i"].shape[0]
    ids = torch.from_numpy(np.stack([g["i"], g["j"]]).astype(np.int32)).cuda().reshape(-1)
    counts = torch.from_numpy(g["x"]).cuda()
    # gradients (EMIT mode) against the golden dense gradient
    t = engine.EmbeddingTable.from_dense(g["E"], g["b"], sparse=False, adagrad=False)
    plan = engine.IndexPlan(2 * B, V).build(ids)
    st = engine.GloveStep(t, B, emit_grads=True)
    sc = st.run(plan, counts)
    torch.cuda.synchronize()
    U = int(plan.n_uniq.item())
    uniq = plan.uniq[:U].cpu().numpy()
    assert abs(float(sc[L.SC_LOSS].item()) - g["loss"]) <= 1e-5 * max(1.0, abs(g["loss"]))
    np.testing.assert_allclose(st.dE[:U].cpu().numpy(), g["dE"][uniq], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(st.db[:U].cpu().numpy(), g["db"][uniq], rtol=1e-5, atol=1e-5)
    untouched = np.setdiff1d(np.arange(V), uniq)
    assert np.all(g["dE"][untouched] == 0)
    # fused sparse Adagrad step against the golden post-update table
    t2 = engine.EmbeddingTable.from_dense(g["E"], g["b"], sparse=True)
    engine.GloveStep(t2, B, lr=0.05).run(engine.IndexPlan(2 * B, V).build(ids), counts)
    np.testing.assert_allclose(t2.dense().cpu().numpy(), g["adagrad_E"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(t2.bias.cpu().numpy(), g["adagrad_b"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(t2.acc.cpu().numpy(), g["adagrad_acc_E"], rtol=1e-5, atol=1e-5)
    # dense Adam (the reference's own rule) from the golden dense gradients
    E = torch.from_numpy(g["E"]).cuda()
    dE = torch.from_numpy(g["dE"].astype(np.float32)).cuda()
    mu, nu = torch.zeros_like(E), torch.zeros_like(E)
    engine.dense_adam(E, dE, mu, nu, 1e-3, 1)
    np.testing.assert_allclose(E.cpu().numpy(), g["adam_E"], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_cuda_stl_vs_golden():
    from esrecsys_b200.pinterest.models import triplet_loss_and_grads
    g = _load("stl_B16_D32.npz")
    loss, ds, dp, dn, ps, ns = triplet_loss_and_grads(*(torch.from_numpy(g[k]).cuda() for k in ("scene", "pos", "neg")),
                                                      0.1, 16)
    assert abs(float(loss.item()) - g["loss"]) <= 1e-5 * max(1.0, abs(g["loss"]))
    np.testing.assert_allclose(ps.cpu().numpy(), g["pos_score"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ns.cpu().numpy(), g["neg_score"], rtol=1e-5, atol=1e-5)
    for got, k in ((ds, "d_scene"), (dp, "d_pos"), (dn, "d_neg")):
        np.testing.assert_allclose(got.cpu().numpy(), g[k], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_cuda_spotify_vs_golden():
    from esrecsys_b200.spotify.models import SpotifyModel
    g = _load("spotify_m7.npz")
    VA, F = g["A"].shape
    model = SpotifyModel(feature_size=F, max_albums=VA, num_artists=g["R"].shape[0])
    params = {"album_embed": {"embedding": torch.from_numpy(g["A"]).cuda()},
              "artist_embed": {"embedding": torch.from_numpy(g["R"]).cuda()}}
    x = {k: g[k] for k in ("album_context", "artist_context", "next_album", "next_artist", "neg_album", "neg_artist")}
    loss, grads = model.loss_and_grads(params, [x], regularization=float(g["reg"]))
    assert abs(float(loss[0].item()) - g["loss"]) <= 2e-5 * max(1.0, abs(g["loss"]))
    np.testing.assert_allclose(grads["album_embed"]["embedding"].dense().cpu().numpy(), g["dA"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(grads["artist_embed"]["embedding"].dense().cpu().numpy(), g["dR"], rtol=1e-5, atol=1e-5)
