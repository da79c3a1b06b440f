"""Vectors produced by EXECUTING the reference's own files (tests/golden/ref_*.npz, generator
tests/golden/make_ref_golden.py: /root/reference/{wikipedia,spotify,pinterest}/*.py imported unmodified on top of the
torch-float64 jax/flax/optax stand-in under tests/golden/refshim).

CPU tests pin the oracle (oracle/*.py, run in float64) to them at ~1e-12; GPU tests pin the CUDA path, called through
the reference-named host mirrors and the C ABI, to the same files at the north star's fp32 tolerance (1e-5)."""
import os

import numpy as np
import pytest
import torch

from oracle import glove as og
from oracle import optim as oopt
from oracle import spotify as osp
from oracle import stl as ostl

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GLOVE = ["ref_glove_V60_D8_B32.npz", "ref_glove_V500_D64_B256.npz"]
SP_KEYS = ("track_context", "album_context", "artist_context", "next_track", "next_album", "next_artist",
           "neg_track", "neg_album", "neg_artist")


def _load(name):
    return dict(np.load(os.path.join(G, name)))


def _spotify_tables(g, dtype):
    F = int(g["F"])
    A = np.zeros((osp.MAX_ALBUMS, F), dtype)
    R = np.zeros((osp.NUM_ARTISTS, F), dtype)
    A[g["arows"]], R[g["rrows"]] = g["avals"], g["rvals"]
    return A, R


def _spotify_x(g, s):
    return {k: g["x%d_%s" % (s, k)] for k in SP_KEYS if "x%d_%s" % (s, k) in g}


def test_ref_fixtures_present():
    for f in GLOVE + ["ref_spotify.npz", "ref_stl.npz", "make_ref_golden.py", "refshim/README.md"]:
        assert os.path.exists(os.path.join(G, f)), f


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("name", GLOVE)
def test_oracle_glove_vs_reference_run(name):
    g = _load(name)
    E0, b0 = g["E"].astype(np.float64), g["b"].astype(np.float64)
    ids, x = g["ids"], g["x"].astype(np.float64)
    V, D = E0.shape
    steps = ids.shape[0]
    if "forward0" in g:                                           # Glove.__call__, the (B,B) broadcast included
        assert np.abs(og.forward_literal(E0, b0, ids[0, 0], ids[0, 1]) - g["forward0"]).max() < 1e-13
    gr = og.loss_and_grads(E0, b0, ids[0, 0], ids[0, 1], x[0])
    dE, db = og.dense_grads(V, gr, D)
    assert abs(gr.loss - g["loss_adam"][0]) < 1e-13
    assert np.abs(dE - g["dE0"]).max() < 1e-14 and np.abs(db - g["db0"]).max() < 1e-14
    # apply_model + update_model with optax.adam, three steps (count / bias correction advance)
    E, b = E0.copy(), b0.copy()
    st = dict(count=0, muE=np.zeros_like(E), nuE=np.zeros_like(E), mub=np.zeros_like(b), nub=np.zeros_like(b))
    losses = [og.step_adam(E, b, st, ids[s, 0], ids[s, 1], x[s], 1e-3) for s in range(steps)]
    assert np.abs(np.array(losses) - g["loss_adam"]).max() < 1e-12
    assert np.abs(E - g["E_adam"]).max() < 1e-11 and np.abs(b - g["b_adam"]).max() < 1e-11
    assert np.abs(st["muE"] - g["adam_mu_E"]).max() < 1e-14 and np.abs(st["nuE"] - g["adam_nu_E"]).max() < 1e-14
    assert abs(np.mean(losses) - g["epoch_loss"]) < 1e-12         # train_epoch's np.mean(epoch_loss)
    # the north-star optimizer through the reference's update_model
    E, b = E0.copy(), b0.copy()
    accE, accb = np.full_like(E, 0.1), np.full_like(b, 0.1)
    losses = [og.step_adagrad(E, b, accE, accb, ids[s, 0], ids[s, 1], x[s], 0.05) for s in range(steps)]
    assert np.abs(np.array(losses) - g["loss_adagrad"]).max() < 1e-12
    assert np.abs(E - g["E_adagrad"]).max() < 1e-12 and np.abs(b - g["b_adagrad"]).max() < 1e-12
    assert np.abs(accE - g["adagrad_acc_E"]).max() < 1e-13 and np.abs(accb - g["adagrad_acc_b"]).max() < 1e-13
    E, b = E0.copy(), b0.copy()
    st = dict(trE=np.zeros_like(E), trb=np.zeros_like(b))
    losses = [og.step_sgdm(E, b, st, ids[s, 0], ids[s, 1], x[s], 0.05, 0.9) for s in range(steps)]
    assert np.abs(np.array(losses) - g["loss_sgdm"]).max() < 1e-12
    assert np.abs(E - g["E_sgdm"]).max() < 1e-12 and np.abs(b - g["b_sgdm"]).max() < 1e-12
    # find_knn on the trained table: scores and the full stable ascending argsort
    sc, idx = og.find_knn(g["E_adam"], g["tokens"])
    assert np.abs(sc - g["knn_scores"]).max() < 1e-13
    assert np.array_equal(idx, g["knn_indices"])


def test_oracle_dump_knn_lines_vs_reference_run():
    """dump_knn's log lines (train_cooccurence.py:114-126) rebuilt from the oracle's top-k and the host mirror of the
    token dictionary (pure Python, no compute)."""
    from esrecsys_b200.wikipedia.token_dictionary import TokenDictionary
    g = _load(GLOVE[0])
    td = TokenDictionary(os.path.join(G, "token.tstat.pb.b64.bz2"))
    top, sc = og.top_k(g["E_adam"], g["tokens"], 10)
    name = td.get_token_from_embedding_index
    lines = ["Nearest neighbors for %s: %s" % (name(int(tok)), " ".join(
        "%s:%f" % (name(int(top[t, k])), sc[t, k]) for k in range(10))) for t, tok in enumerate(g["tokens"])]
    assert lines == [str(v) for v in g["dump_knn_lines"]]


def test_oracle_spotify_vs_reference_run():
    g = _load("ref_spotify.npz")
    A, R = _spotify_tables(g, np.float64)
    reg, lr, mom = float(g["reg"]), float(g["lr"]), float(g["momentum"])
    x0 = _spotify_x(g, 0)
    out = osp.forward(A, R, x0["album_context"], x0["artist_context"], x0["next_album"], x0["next_artist"],
                      x0["neg_album"], x0["neg_artist"])
    for got, k in zip(out, ("pos_affinity", "neg_affinity", "context_self", "next_self", "neg_self", "l2")):
        assert np.abs(got - g["fwd0_" + k]).max() < 1e-12, k
    gr = osp.loss_and_grads(A, R, x0["album_context"], x0["artist_context"], x0["next_album"], x0["next_artist"],
                            x0["neg_album"], x0["neg_artist"], reg)
    dA, dR = osp.dense_grads(A, R, gr)
    assert np.abs(dA[g["arows"]] - g["dA0"]).max() < 1e-12 and np.abs(dR[g["rrows"]] - g["dR0"]).max() < 1e-12
    mask = np.ones(A.shape[0], bool)
    mask[g["arows"]] = False
    assert not dA[mask].any()
    trA, trR = np.zeros_like(A), np.zeros_like(R)
    losses = [osp.train_step(A, R, trA, trR, _spotify_x(g, s), reg, lr, mom) for s in range(3)]
    assert np.abs(np.array(losses) - g["losses"]).max() < 1e-11
    assert np.abs(A[g["arows"]] - g["A_rows_final"]).max() < 1e-12
    assert np.abs(R[g["rrows"]] - g["R_rows_final"]).max() < 1e-12
    assert np.abs(trA[g["arows"]] - g["A_trace_final"]).max() < 1e-12
    assert np.abs(trR[g["rrows"]] - g["R_trace_final"]).max() < 1e-12
    y = _spotify_x(g, 3)
    aff = osp.eval_scores(A, R, y["album_context"], y["artist_context"], g["all_albums"], g["all_artists"])
    assert np.abs(aff - g["eval_affinity"]).max() < 1e-12
    metrics, order = osp.eval_step(A, R, y, g["all_tracks"], g["all_albums"], g["all_artists"])
    assert np.array_equal(order.astype(np.int32), g["eval_top500"])
    assert np.abs(metrics - g["eval_metrics"]).max() < 1e-7


def _stl_table_grads(S, P, scene, pos, neg, reg, B):
    s, p, n = S[scene], P[pos], P[neg]
    loss, ds, dp, dn = ostl.triplet_loss_and_grads(s, p, n, reg, B)
    dS, dP = np.zeros_like(S), np.zeros_like(P)
    np.add.at(dS, scene, ds)                                      # VJP of the tower lookup = scatter-add
    np.add.at(dP, pos, dp)
    np.add.at(dP, neg, dn)
    return loss, dS, dP


def test_oracle_stl_vs_reference_run():
    g = _load("ref_stl.npz")
    S, P = g["S"].astype(np.float64), g["P"].astype(np.float64)
    scene, pos, neg = g["scene"], g["pos"], g["neg"]
    reg, lr, B = float(g["reg"]), float(g["lr"]), scene.shape[1]
    ps, ns = ostl.scores(S[scene[0]], P[pos[0]], P[neg[0]])
    assert np.abs(ps - g["fwd0_pos_score"]).max() < 1e-13 and np.abs(ns - g["fwd0_neg_score"]).max() < 1e-13
    assert np.array_equal(S[scene[0]], g["fwd0_scene_embed"]) and np.array_equal(P[neg[0]], g["fwd0_neg_embed"])
    assert np.array_equal(S[scene[0]], g["scene_embed_method"]) and np.array_equal(P[pos[0]], g["product_embed_method"])
    _, dS, dP = _stl_table_grads(S, P, scene[0], pos[0], neg[0], reg, B)
    assert np.abs(dS - g["dS0"]).max() < 1e-13 and np.abs(dP - g["dP0"]).max() < 1e-13
    muS, nuS, muP, nuP, count = np.zeros_like(S), np.zeros_like(S), np.zeros_like(P), np.zeros_like(P), 0
    for s in range(scene.shape[0]):
        assert abs(ostl.eval_loss(S[scene[s]], P[pos[s]], P[neg[s]]) - g["eval_losses"][s]) < 1e-12
        loss, dS, dP = _stl_table_grads(S, P, scene[s], pos[s], neg[s], reg, B)
        assert abs(loss - g["losses"][s]) < 1e-12
        S, muS, nuS, _ = oopt.adam_update(S, dS, muS, nuS, count, lr)
        P, muP, nuP, count = oopt.adam_update(P, dP, muP, nuP, count, lr)
    assert np.abs(S - g["S_final"]).max() < 1e-11 and np.abs(P - g["P_final"]).max() < 1e-11
    sc, idx = ostl.find_top_k(g["S"][7].astype(np.float64), g["P"].astype(np.float64), 10)
    assert np.array_equal(idx, g["topk_indices"]) and np.abs(sc - g["topk_scores"]).max() < 1e-13


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)
RTOL, ATOL = 1e-5, 1e-5


def _t(a, dtype=np.float32):
    return torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).cuda()


def _c(t):
    return t.detach().cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GLOVE)
def test_cuda_glove_vs_reference_run(name):
    """apply_model / update_model / train_epoch / find_knn / dump_knn of the host mirror (libesr kernels underneath)
    against what the reference's own functions returned for the same inputs."""
    from esrecsys_b200 import engine
    from esrecsys_b200 import optim as O
    from esrecsys_b200.train_state import TrainState
    from esrecsys_b200.wikipedia.models import Glove
    from esrecsys_b200.wikipedia.token_dictionary import TokenDictionary
    from esrecsys_b200.wikipedia.train_cooccurence import apply_model, dump_knn, find_knn, train_epoch, update_model
    g = _load(name)
    V, D = g["E"].shape
    ids, x = g["ids"], g["x"]
    steps, B = ids.shape[0], ids.shape[2]
    model = Glove(num_embeddings=V, features=D)

    def fresh():
        return {"_token_embedding": {"embedding": _t(g["E"])}, "_bias": {"embedding": _t(g["b"].reshape(V, 1))}}

    if "forward0" in g:
        out = model.apply({"params": fresh()}, ids[0])
        np.testing.assert_allclose(_c(out), g["forward0"], rtol=RTOL, atol=ATOL)
    for tag, tx in (("adam", O.adam(1e-3)), ("adagrad", O.adagrad(0.05)), ("sgdm", O.sgd(0.05, momentum=0.9))):
        state = TrainState.create(apply_fn=model.apply, params=fresh(), tx=tx)
        for s in range(steps):
            grads, loss = apply_model(state, ids[s], x[s])
            if s == 0:
                np.testing.assert_allclose(_c(grads["_token_embedding"]["embedding"].dense()), g["dE0"], rtol=RTOL, atol=ATOL)
                np.testing.assert_allclose(_c(grads["_bias"]["embedding"].dense()).reshape(-1), g["db0"], rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(float(loss), g["loss_" + tag][s], rtol=2e-5, atol=ATOL)
            state = update_model(state, grads)
        assert state.step == steps
        np.testing.assert_allclose(_c(state.params["_token_embedding"]["embedding"]), g["E_" + tag], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(_c(state.params["_bias"]["embedding"]).reshape(-1), g["b_" + tag], rtol=RTOL, atol=ATOL)
    # the fused step (sparse Adagrad inside the row pass) against the reference's update_model(optax.adagrad) run
    table = engine.EmbeddingTable.from_dense(g["E"], g["b"], sparse=True)
    step = engine.GloveStep(table, B, lr=0.05)
    plan = engine.IndexPlan(2 * B, V)
    for s in range(steps):
        plan.build(torch.from_numpy(ids[s].reshape(-1)).cuda())
        sc = step.run(plan, torch.from_numpy(x[s]).cuda())
        np.testing.assert_allclose(float(sc[5].item()), g["loss_adagrad"][s], rtol=2e-5, atol=ATOL)
    np.testing.assert_allclose(_c(table.dense()), g["E_adagrad"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_c(table.bias), g["b_adagrad"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_c(table.acc), g["adagrad_acc_E"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_c(table.bias_acc), g["adagrad_acc_b"], rtol=RTOL, atol=ATOL)
    # train_epoch: mean loss of the epoch + the state it leaves
    state = TrainState.create(apply_fn=model.apply, params=fresh(), tx=O.adam(1e-3))
    state, train_loss = train_epoch(state, steps, iter([(ids[s], x[s]) for s in range(steps)]))
    np.testing.assert_allclose(train_loss, g["epoch_loss"], rtol=2e-5, atol=ATOL)
    np.testing.assert_allclose(_c(state.params["_token_embedding"]["embedding"]), g["E_adam"], rtol=RTOL, atol=ATOL)
    # find_knn / dump_knn from the reference-trained table (so ranking differences are not training differences)
    params = {"_token_embedding": {"embedding": _t(g["E_adam"])}, "_bias": {"embedding": _t(g["b_adam"].reshape(V, 1))}}
    scores, idx = find_knn(model, params, g["tokens"])
    np.testing.assert_allclose(_c(scores), g["knn_scores"], rtol=RTOL, atol=ATOL)
    got, want = _c(idx), g["knn_indices"]
    assert got.shape == want.shape and np.array_equal(np.sort(got, axis=0), np.sort(want, axis=0))   # a permutation per query
    ref_sorted = np.take_along_axis(g["knn_scores"], want.astype(np.int64), axis=0)
    sep = np.ones(want.shape, bool)                                     # positions whose f64 score is separated from both neighbours
    gap = np.abs(np.diff(ref_sorted, axis=0)) > 1e-5
    sep[1:] &= gap
    sep[:-1] &= gap
    assert np.array_equal(got[sep], want[sep])
    if "dump_knn_lines" in g:
        td = TokenDictionary(os.path.join(G, "token.tstat.pb.b64.bz2"))
        knn = dump_knn(model, params, g["tokens"], td, k=10)
        for (query, nbrs), line in zip(knn, g["dump_knn_lines"]):
            head, body = str(line).split(": ", 1)
            assert head == "Nearest neighbors for %s" % query
            want_items = body.split(" ")
            got_items = " ".join("%s:%f" % (w, sc) for w, sc in nbrs).split(" ")
            assert len(got_items) == len(want_items)
            for a, b_ in zip(got_items, want_items):                     # "word:score" (or the halves of "MINHASH n:score")
                if ":" in a:
                    assert a.rsplit(":", 1)[0] == b_.rsplit(":", 1)[0]
                    assert abs(float(a.rsplit(":", 1)[1]) - float(b_.rsplit(":", 1)[1])) <= 2e-5
                else:
                    assert a == b_


@pytest.mark.gpu
def test_cuda_spotify_vs_reference_run():
    """SpotifyModel.apply, train_step (optax.sgd momentum) and eval_step of the host mirror against the reference's run."""
    from esrecsys_b200 import optim as O
    from esrecsys_b200.spotify.models import SpotifyModel
    from esrecsys_b200.spotify.train_spotify import eval_scores, eval_step, train_step
    from esrecsys_b200.train_state import TrainState
    g = _load("ref_spotify.npz")
    A, R = _spotify_tables(g, np.float32)
    reg, lr, mom = float(g["reg"]), float(g["lr"]), float(g["momentum"])
    model = SpotifyModel(feature_size=int(g["F"]))                      # reference table heights (100000, 295861)
    params = {"album_embed": {"embedding": _t(A)}, "artist_embed": {"embedding": _t(R)}}
    x0 = _spotify_x(g, 0)
    out = model.apply({"params": params}, *[x0[k] for k in SP_KEYS])
    for got, k in zip(out, ("pos_affinity", "neg_affinity", "context_self", "next_self", "neg_self", "l2")):
        np.testing.assert_allclose(_c(got), g["fwd0_" + k], rtol=RTOL, atol=2e-5, err_msg=k)
    loss, grads = model.loss_and_grads(params, [x0], regularization=reg)
    np.testing.assert_allclose(float(loss[0]), g["losses"][0], rtol=2e-5, atol=ATOL)
    dA = _c(grads["album_embed"]["embedding"].dense())
    dR = _c(grads["artist_embed"]["embedding"].dense())
    np.testing.assert_allclose(dA[g["arows"]], g["dA0"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(dR[g["rrows"]], g["dR0"], rtol=RTOL, atol=ATOL)
    mask = np.ones(A.shape[0], bool)
    mask[g["arows"]] = False
    assert not dA[mask].any()
    state = TrainState.create(apply_fn=model.apply, params=params, tx=O.sgd(lr, momentum=mom))
    for s in range(3):
        state, loss = train_step(state, model, [_spotify_x(g, s)], reg)
        np.testing.assert_allclose(float(loss[0]), g["losses"][s], rtol=2e-5, atol=ATOL)
    assert state.step == 3
    np.testing.assert_allclose(_c(state.params["album_embed"]["embedding"])[g["arows"]], g["A_rows_final"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_c(state.params["artist_embed"]["embedding"])[g["rrows"]], g["R_rows_final"], rtol=RTOL, atol=ATOL)
    # eval_step from the reference-trained tables
    A3, R3 = A.copy(), R.copy()
    A3[g["arows"]], R3[g["rrows"]] = g["A_rows_final"], g["R_rows_final"]
    p3 = {"album_embed": {"embedding": _t(A3)}, "artist_embed": {"embedding": _t(R3)}}
    y = _spotify_x(g, 3)
    aff = eval_scores(model, p3, y, g["all_albums"], g["all_artists"])
    np.testing.assert_allclose(_c(aff), g["eval_affinity"], rtol=RTOL, atol=2e-5)
    metrics, top = eval_step(model, p3, y, g["all_tracks"], g["all_albums"], g["all_artists"], k=500)
    got, want = _c(top).astype(np.int64), g["eval_top500"].astype(np.int64)
    ref_sorted = g["eval_affinity"][want]
    gap = np.abs(np.diff(ref_sorted)) > 1e-4
    sep = np.ones(500, bool)
    sep[1:] &= gap
    sep[:-1] &= gap
    assert len(set(got.tolist()) ^ set(want.tolist())) <= 4             # fp32 near-ties at the cut only
    assert np.array_equal(got[sep], want[sep])
    np.testing.assert_allclose(_c(metrics), g["eval_metrics"], atol=1.0 / 9 + 1e-6)


@pytest.mark.gpu
def test_cuda_stl_vs_reference_run():
    """STLModel.apply (ID towers), train_step (optax.adam), eval_step and find_top_k against the reference's run."""
    from esrecsys_b200 import optim as O
    from esrecsys_b200.pinterest.make_recommendations import find_top_k
    from esrecsys_b200.pinterest.models import STLModel
    from esrecsys_b200.pinterest.train_shop_the_look import eval_step, loss_and_grads, train_step
    from esrecsys_b200.train_state import TrainState
    g = _load("ref_stl.npz")
    S, P = g["S"], g["P"]
    scene, pos, neg = g["scene"], g["pos"], g["neg"]
    reg, lr, B = float(g["reg"]), float(g["lr"]), scene.shape[1]
    model = STLModel(output_size=S.shape[1], num_scenes=S.shape[0], num_products=P.shape[0])
    variables = {"params": {"scene_cnn": {"embedding": _t(S)}, "product_cnn": {"embedding": _t(P)}}}
    out = model.apply(variables, scene[0], pos[0], neg[0], True)
    for got, k in zip(out, ("pos_score", "neg_score", "scene_embed", "pos_embed", "neg_embed")):
        np.testing.assert_allclose(_c(got), g["fwd0_" + k], rtol=RTOL, atol=ATOL, err_msg=k)
    assert np.array_equal(_c(model.apply(variables, scene[0], method=STLModel.get_scene_embed)), g["scene_embed_method"].astype(np.float32))
    assert np.array_equal(_c(model.apply(variables, pos[0], method=STLModel.get_product_embed)), g["product_embed_method"].astype(np.float32))
    loss, grads = loss_and_grads(model, variables, scene[0], pos[0], neg[0], reg, B)
    np.testing.assert_allclose(float(loss), g["losses"][0], rtol=2e-5, atol=ATOL)
    np.testing.assert_allclose(_c(grads["params"]["scene_cnn"]["embedding"].dense()), g["dS0"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_c(grads["params"]["product_cnn"]["embedding"].dense()), g["dP0"], rtol=RTOL, atol=ATOL)
    state = TrainState.create(apply_fn=model.apply, params=variables, tx=O.adam(lr))
    for s in range(scene.shape[0]):
        ev = eval_step(state, model, scene[s], pos[s], neg[s])
        np.testing.assert_allclose(float(ev), g["eval_losses"][s], rtol=2e-5, atol=ATOL)
        state, loss = train_step(state, model, scene[s], pos[s], neg[s], reg, B)
        np.testing.assert_allclose(float(loss), g["losses"][s], rtol=2e-5, atol=ATOL)
    np.testing.assert_allclose(_c(state.params["params"]["scene_cnn"]["embedding"]), g["S_final"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_c(state.params["params"]["product_cnn"]["embedding"]), g["P_final"], rtol=RTOL, atol=ATOL)
    val, idx = find_top_k(S[7], P, 10)
    assert np.array_equal(_c(idx), g["topk_indices"])
    np.testing.assert_allclose(_c(val), g["topk_scores"], rtol=RTOL, atol=ATOL)


# ------------------------------------------------------------------------------------------------ end to end, from a corpus file
def _e2e_setup():
    """Inputs of the reference's mini training run (make_ref_golden.wikipedia_e2e): corpus part file, dictionary, the
    seeded initial table."""
    from esrecsys_b200.wikipedia.cooccurrence_matrix import CooccurrenceGenerator
    from esrecsys_b200.wikipedia.token_dictionary import TokenDictionary
    g = _load("ref_glove_e2e.npz")
    td = TokenDictionary(os.path.join(G, "token.tstat.pb.b64.bz2"))
    V, D = td.get_embedding_dictionary_size(), int(g["D"])
    assert V == int(g["num_tokens"])
    tokens = np.array([td.get_embedding_index(w) for w in str(g["terms"]).split(",")], np.int32)
    assert np.array_equal(tokens, g["debug_tokens"])
    E = (np.random.default_rng(int(g["seed"])).standard_normal((V, D)) / np.sqrt(D)).astype(np.float32)
    assert E.astype(np.float64).sum() == float(g["E_checksum"])
    it = CooccurrenceGenerator(os.path.join(G, "e2e_cooccur", "part-?????.bz2")).get_batch(int(g["B"]), 0)
    x, _ = next(it)                                                # main() spends the first batch on model.init
    assert np.array_equal(np.stack(x), g["first_batch_x"])
    return g, td, tokens, E, it


def _knn_lines(td, tokens, top, sc):
    name = td.get_token_from_embedding_index
    return ["Nearest neighbors for %s: %s" % (name(int(tok)), " ".join(
        "%s:%f" % (name(int(top[t][k])), sc[t][k]) for k in range(len(top[t])))) for t, tok in enumerate(tokens)]


def test_oracle_pipeline_vs_reference_training_run():
    """Corpus file -> native decoder / CooccurrenceGenerator -> oracle steps -> oracle top-k -> the reference's log lines."""
    g, td, tokens, E, it = _e2e_setup()
    E = E.astype(np.float64)
    b = np.zeros(E.shape[0])
    st = dict(count=0, muE=np.zeros_like(E), nuE=np.zeros_like(E), mub=np.zeros_like(b), nub=np.zeros_like(b))
    for epoch in range(int(g["epochs"])):
        top, sc = og.top_k(E, tokens, 10)
        assert _knn_lines(td, tokens, top, sc) == [str(v) for v in g["knn_lines_%d" % epoch]]
        losses = []
        for _ in range(int(g["steps_per_epoch"])):
            x, y = next(it)
            losses.append(og.step_adam(E, b, st, x[0], x[1], y.astype(np.float64), float(g["lr"])))
        assert abs(np.mean(losses) - g["train_loss"][epoch]) < 1e-12
    top, sc = og.top_k(E, tokens, 10)
    assert _knn_lines(td, tokens, top, sc) == [str(v) for v in g["knn_lines_%d" % int(g["epochs"])]]
    assert np.abs(E[:64] - g["E_final_head"]).max() < 1e-11 and np.abs(b[:64] - g["b_final_head"]).max() < 1e-11


@pytest.mark.gpu
def test_cuda_pipeline_vs_reference_training_run():
    """The same run through the product: native decoder -> Glove / TrainState / train_epoch / dump_knn mirrors on libesr."""
    from esrecsys_b200 import optim as O
    from esrecsys_b200.train_state import TrainState
    from esrecsys_b200.wikipedia.models import Glove
    from esrecsys_b200.wikipedia.train_cooccurence import dump_knn, train_epoch
    g, td, tokens, E, it = _e2e_setup()
    V, D = E.shape
    model = Glove(num_embeddings=V, features=D)
    params = {"_token_embedding": {"embedding": _t(E)}, "_bias": {"embedding": torch.zeros(V, 1, device="cuda")}}
    state = TrainState.create(apply_fn=model.apply, params=params, tx=O.adam(float(g["lr"])))

    def check_lines(epoch):
        knn = dump_knn(model, state.params, tokens, td, k=10)
        for (query, nbrs), line in zip(knn, g["knn_lines_%d" % epoch]):
            head, body = str(line).split(": ", 1)
            assert head == "Nearest neighbors for %s" % query
            got = " ".join("%s:%f" % (w, s) for w, s in nbrs).split(" ")
            want = body.split(" ")
            assert len(got) == len(want)
            for a, b_ in zip(got, want):
                if ":" in a:
                    assert a.rsplit(":", 1)[0] == b_.rsplit(":", 1)[0]
                    assert abs(float(a.rsplit(":", 1)[1]) - float(b_.rsplit(":", 1)[1])) <= 2e-5
                else:
                    assert a == b_

    for epoch in range(int(g["epochs"])):
        check_lines(epoch)
        state, train_loss = train_epoch(state, int(g["steps_per_epoch"]), it)
        np.testing.assert_allclose(train_loss, g["train_loss"][epoch], rtol=2e-5, atol=ATOL)
    check_lines(int(g["epochs"]))
    np.testing.assert_allclose(_c(state.params["_token_embedding"]["embedding"])[:64], g["E_final_head"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_c(state.params["_bias"]["embedding"])[:64, 0], g["b_final_head"], rtol=RTOL, atol=ATOL)
