"""GPU parity of the reference-surface mirrors (Glove / SpotifyModel / STLModel + TrainState) against
the NumPy oracle.  Tolerance 1e-5 (fp32), as the north star states."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from esrecsys_b200 import synth
from oracle import glove as og
from oracle import optim as oopt
from oracle import spotify as osp
from oracle import stl as ostl

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-5


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------
# wikipedia: Glove surface, apply_model / update_model with the reference's optimizer (dense Adam)
# ------------------------------------------------------------------------------------------------
def _glove_setup(V=2000, D=64, B=512, seed=0):
    from esrecsys_b200.wikipedia.models import Glove
    model = Glove(num_embeddings=V, features=D)
    variables = model.init(seed, None)
    params = variables["params"]
    assert set(params) == {"_token_embedding", "_bias"}
    assert params["_token_embedding"]["embedding"].shape == (V, D)
    assert params["_bias"]["embedding"].shape == (V, 1)
    assert float(params["_bias"]["embedding"].abs().sum()) == 0.0
    params["_bias"]["embedding"].copy_(torch.randn(V, 1, generator=torch.Generator().manual_seed(1)) * 0.05)
    ids, counts = synth.glove_batches(V, B, 4, seed + 1)
    return model, params, ids, counts


def test_glove_forward_and_score_all():
    model, params, ids, counts = _glove_setup()
    E, b = _np(params["_token_embedding"]["embedding"]), _np(params["_bias"]["embedding"]).reshape(-1)
    out = model.apply({"params": params}, ids[0])
    assert out.shape == (512, 512)                                   # the reference's (B,B) broadcast
    np.testing.assert_allclose(_np(out), og.forward_literal(E, b, ids[0, 0], ids[0, 1]), rtol=RTOL, atol=ATOL)
    from esrecsys_b200.wikipedia.models import Glove
    tok = np.array([3, 17, 256, 1999, 3], np.int32)
    sc = model.apply({"params": params}, tok, method=Glove.score_all)
    assert sc.shape == (2000, 5)
    np.testing.assert_allclose(_np(sc), og.score_all(E, tok), rtol=RTOL, atol=ATOL)
    from esrecsys_b200.wikipedia.train_cooccurence import find_knn
    scores, idx = find_knn(model, params, tok)
    _, oidx = og.find_knn(_np(scores), np.arange(5)) if False else (None, np.argsort(_np(scores), axis=0, kind="stable"))
    assert np.array_equal(_np(idx), oidx)


@pytest.mark.parametrize("opt", ["adam", "sgd", "adagrad"])
@pytest.mark.parametrize("bias_mode", ["reference_broadcast", "per_pair"])
def test_apply_model_update_model(opt, bias_mode):
    from esrecsys_b200 import optim as O
    from esrecsys_b200.train_state import TrainState
    from esrecsys_b200.wikipedia.train_cooccurence import apply_model, update_model
    model, params, ids, counts = _glove_setup()
    E = _np(params["_token_embedding"]["embedding"]).copy()
    b = _np(params["_bias"]["embedding"]).reshape(-1).copy()
    tx = {"adam": O.adam(1e-3), "sgd": O.sgd(1e-2, momentum=0.9), "adagrad": O.adagrad(0.05)}[opt]
    state = TrainState.create(apply_fn=model.apply, params=params, tx=tx)
    st = dict(count=0, muE=np.zeros_like(E), nuE=np.zeros_like(E), mub=np.zeros_like(b), nub=np.zeros_like(b),
              trE=np.zeros_like(E), trb=np.zeros_like(b))
    aE, ab = np.full_like(E, 0.1), np.full_like(b, 0.1)
    for k in range(3):
        grads, loss = apply_model(state, ids[k], counts[k], bias_mode)
        gr = og.loss_and_grads(E, b, ids[k, 0], ids[k, 1], counts[k], bias_mode)
        dE, db = og.dense_grads(E.shape[0], gr, E.shape[1])
        np.testing.assert_allclose(_np(grads["_token_embedding"]["embedding"].dense()), dE, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(_np(grads["_bias"]["embedding"].dense()).reshape(-1), db, rtol=RTOL, atol=ATOL)
        state = update_model(state, grads)
        if opt == "adam":
            oloss = og.step_adam(E, b, st, ids[k, 0], ids[k, 1], counts[k], 1e-3, bias_mode)
        elif opt == "sgd":
            oloss = og.step_sgdm(E, b, st, ids[k, 0], ids[k, 1], counts[k], 1e-2, 0.9, bias_mode)
        else:
            oloss = og.step_adagrad(E, b, aE, ab, ids[k, 0], ids[k, 1], counts[k], 0.05, bias_mode)
        np.testing.assert_allclose(float(loss), oloss, rtol=2e-5, atol=ATOL)
    assert state.step == 3
    np.testing.assert_allclose(_np(state.params["_token_embedding"]["embedding"]), E, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_np(state.params["_bias"]["embedding"]).reshape(-1), b, rtol=RTOL, atol=ATOL)


# ------------------------------------------------------------------------------------------------
# pinterest: scoring + triplet loss
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,D", [(16, 64), (16, 32), (4096, 256), (1, 8), (33, 100)])
def test_stl_triplet_matches_oracle(B, D):
    from esrecsys_b200.pinterest.models import triplet_loss_and_grads
    rng = np.random.default_rng(B + D)
    s, p, n = (rng.standard_normal((B, D)).astype(np.float32) * sc for sc in (0.2, 0.2, 0.2))
    s[: B // 2] *= 8.0          # some norms above 1 so the regulariser is active
    p[B // 3:] *= 6.0
    loss, ds, dp, dn, ps, ns = triplet_loss_and_grads(*(torch.from_numpy(x).cuda() for x in (s, p, n)), 0.1, 16)
    ol, ods, odp, odn = ostl.triplet_loss_and_grads(s, p, n, 0.1, 16)
    np.testing.assert_allclose(float(loss), ol, rtol=2e-5, atol=ATOL)
    for got, want in ((ds, ods), (dp, odp), (dn, odn)):
        np.testing.assert_allclose(_np(got), want, rtol=RTOL, atol=ATOL)
    ops, ons = ostl.scores(s, p, n)
    np.testing.assert_allclose(_np(ps), ops, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_np(ns), ons, rtol=RTOL, atol=ATOL)


def test_stlmodel_surface():
    from esrecsys_b200.pinterest.models import STLModel
    m = STLModel(output_size=64, num_scenes=500, num_products=700)
    v = m.init(0)
    ids = np.random.default_rng(0).integers(0, 500, size=(3, 16))
    out = m.apply(v, ids[0], ids[1], ids[2], True)
    assert len(out) == 5 and out[0].shape == (16,) and out[2].shape == (16, 64)
    S = _np(v["params"]["scene_cnn"]["embedding"])
    P = _np(v["params"]["product_cnn"]["embedding"])
    np.testing.assert_allclose(_np(out[0]), np.sum(S[ids[0]] * P[ids[1]], -1), rtol=RTOL, atol=ATOL)
    assert np.array_equal(_np(m.apply(v, ids[0], method=STLModel.get_scene_embed)), S[ids[0]])


# ------------------------------------------------------------------------------------------------
# spotify: fused forward/backward of train_step's loss, reference optimizer (dense SGD momentum)
# ------------------------------------------------------------------------------------------------
def _spotify_setup(F=32, VA=1000, VR=3000, seed=0):
    from esrecsys_b200.spotify.models import SpotifyModel
    model = SpotifyModel(feature_size=F, max_albums=VA, num_artists=VR)
    v = model.init(seed)
    assert set(v["params"]) == {"album_embed", "artist_embed"}
    return model, v["params"]


@pytest.mark.parametrize("m", [5, 7, 33, 120])
def test_spotify_forward_and_grads(m):
    model, params = _spotify_setup()
    A, R = _np(params["album_embed"]["embedding"]), _np(params["artist_embed"]["embedding"])
    rng = np.random.default_rng(m)
    x = synth.spotify_example(rng, m, o=64, n_albums=5000, n_artists=3000)
    # scale some rows up so the norm regulariser (reg = 1.0 here) and the self-affinity hinges are active
    out = model.apply({"params": params}, x["track_context"], x["album_context"], x["artist_context"], x["next_track"],
                      x["next_album"], x["next_artist"], x["neg_track"], x["neg_album"], x["neg_artist"])
    want = osp.forward(A, R, x["album_context"], x["artist_context"], x["next_album"], x["next_artist"],
                       x["neg_album"], x["neg_artist"])
    for got, w in zip(out, want):
        np.testing.assert_allclose(_np(got), w, rtol=RTOL, atol=ATOL)
    for reg in (10.0, 0.5):
        loss, grads = model.loss_and_grads(params, [x], regularization=reg)
        gr = osp.loss_and_grads(A, R, x["album_context"], x["artist_context"], x["next_album"], x["next_artist"],
                                x["neg_album"], x["neg_artist"], reg)
        np.testing.assert_allclose(float(loss[0]), gr.loss, rtol=2e-5, atol=ATOL)
        dA, dR = osp.dense_grads(A, R, gr)
        np.testing.assert_allclose(_np(grads["album_embed"]["embedding"].dense()), dA, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(_np(grads["artist_embed"]["embedding"].dense()), dR, rtol=RTOL, atol=ATOL)


def test_spotify_pack_equals_singles_and_train_step():
    from esrecsys_b200 import optim as O
    from esrecsys_b200.train_state import TrainState
    model, params = _spotify_setup()
    rng = np.random.default_rng(9)
    xs = [synth.spotify_example(rng, m, o=64, n_albums=5000, n_artists=3000) for m in (5, 12, 40)]
    loss_pack, _ = model.loss_and_grads(params, xs, 10.0)
    for e, x in enumerate(xs):
        l1, _ = model.loss_and_grads(params, [x], 10.0)
        assert float(l1[0]) == float(loss_pack[e])                 # bit-identical: one CTA per playlist
    # train_step with optax.sgd(lr, momentum) (spotify/train_spotify.py:238-241), dense
    A, R = _np(params["album_embed"]["embedding"]).copy(), _np(params["artist_embed"]["embedding"]).copy()
    trA, trR = np.zeros_like(A), np.zeros_like(R)
    state = TrainState.create(apply_fn=model.apply, params=params, tx=O.sgd(1e-3, momentum=0.98))
    for x in xs:
        loss, grads = model.loss_and_grads(state.params, [x], 10.0)
        state = state.apply_gradients(grads=grads)
        ol = osp.train_step(A, R, trA, trR, x, 10.0, 1e-3, 0.98)
        np.testing.assert_allclose(float(loss[0]), ol, rtol=2e-5, atol=ATOL)
    np.testing.assert_allclose(_np(state.params["album_embed"]["embedding"]), A, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(_np(state.params["artist_embed"]["embedding"]), R, rtol=RTOL, atol=ATOL)


# ------------------------------------------------------------------------------------------------
# retrieval side: find_knn / dump_knn, eval_step, find_top_k
# ------------------------------------------------------------------------------------------------
def test_find_knn_matches_oracle_with_ties():
    from esrecsys_b200.wikipedia.models import Glove
    from esrecsys_b200.wikipedia.train_cooccurence import dump_knn, find_knn
    V, D = 4097, 64
    model = Glove(num_embeddings=V, features=D)
    params = model.init(3)["params"]
    E = params["_token_embedding"]["embedding"]
    E[100] = E[7]                      # duplicate rows -> tied scores: the stable order must match jnp.argsort
    E[2000] = E[7]
    tokens = np.array([7, 19, 4000, 0, 100, 33, 1, 2], np.int32)      # T = 8 as train_cooccurence.py:51-54
    scores, idx = find_knn(model, params, tokens)
    osc, oidx = og.find_knn(_np(E), tokens)
    np.testing.assert_allclose(_np(scores), osc, rtol=RTOL, atol=ATOL)
    got = _np(idx)
    sc = _np(scores)
    # ranks agree with a stable argsort of the DEVICE scores (bit-exact bookkeeping on identical keys)
    assert np.array_equal(got, np.argsort(sc, axis=0, kind="stable").astype(np.int32))
    assert np.array_equal(got[-1], oidx[-1])          # every query's nearest neighbour agrees with the oracle
    knn = dump_knn(model, params, tokens, k=10)
    assert knn[0][0] == 7 and len(knn[0][1]) == 10 and {knn[0][1][j][0] for j in range(3)} == {7, 100, 2000}


@pytest.mark.parametrize("V,D,T,k", [(50001, 128, 8, 10), (300, 64, 3, 10), (20000, 32, 1, 1024), (7000, 256, 64, 5),
                                     (2262, 128, 2, 500), (9000, 16, 5, 7), (4000, 512, 3, 20)])
def test_fused_topk_scan_matches_oracle_both_tie_orders(V, D, T, k):
    """esr_topk_scan_f32 (one table pass, running top-k in shared memory) against the full-sort restatements:
    oracle.glove.top_k = tail of the stable ascending argsort (dump_knn: ties HIGHER index first) and
    jax.lax.top_k order (ties LOWER index first), duplicates planted across CTA slabs."""
    from esrecsys_b200 import engine
    rng = np.random.default_rng(V + k)
    E = (rng.standard_normal((V, D)) / np.sqrt(D)).astype(np.float32)
    for dst, src in ((V - 1, 3), (V // 2, 3), (V // 3, 5), (11, 5)):
        E[dst] = E[src]                                   # exact ties, far apart
    tokens = np.concatenate([[3, 5], rng.integers(0, V, T)])[:T].astype(np.int32)
    table = engine.EmbeddingTable.from_dense(E, np.zeros(V, np.float32), sparse=(V % 2 == 1))
    q = torch.from_numpy(E[tokens]).cuda()
    sc64 = E.astype(np.float64) @ E[tokens].astype(np.float64).T            # (V, T)
    for high_first in (True, False):
        val, idx = engine.table_topk(table, q, k, ties_high_index_first=high_first)
        val, idx = _np(val), _np(idx)
        assert idx.shape == (T, k)
        for t in range(T):
            dev_sc = sc64[:, t]
            # the kernel's own scores order its list exactly: strictly by (score desc, index per the tie rule)
            assert len(set(idx[t].tolist())) == k
            v = val[t]
            assert np.all(v[:-1] >= v[1:])
            ties = v[:-1] == v[1:]
            if ties.any():
                d = np.diff(idx[t].astype(np.int64))[ties]
                assert np.all(d < 0) if high_first else np.all(d > 0)
            np.testing.assert_allclose(v, dev_sc[idx[t]], rtol=RTOL, atol=ATOL)
            # membership: everything clearly above the k-th oracle score is present
            kth = np.sort(dev_sc)[-k]
            must = np.flatnonzero(dev_sc > kth + 1e-5)
            assert set(must.tolist()) <= set(idx[t].tolist())
    if D >= 64:   # (at small D a foreign row can out-score the self dot product; the ordering asserts above still hold)
        # the planted exact ties at the very top of query 0 (token 3): order is the reference's, bit for bit
        # (tail of the stable ascending argsort of EXACTLY tied scores: [V-1, V//2, 3]; the fp32 BLAS scores of the NumPy
        # restatement are not bit-identical for identical rows, the kernel's are: same code path for every row)
        _, idx = engine.table_topk(table, q[:1], min(k, 8), ties_high_index_first=True)
        assert _np(idx)[0][:3].tolist() == [V - 1, V // 2, 3]
        _, idx = engine.table_topk(table, q[:1], min(k, 8), ties_high_index_first=False)       # jax.lax.top_k order
        assert _np(idx)[0][:3].tolist() == [3, V // 2, V - 1]


def test_spotify_eval_step_matches_oracle():
    from esrecsys_b200.spotify.train_spotify import eval_step
    model, params = _spotify_setup(F=32, VA=1000, VR=3000)
    A, R = _np(params["album_embed"]["embedding"]), _np(params["artist_embed"]["embedding"])
    rng = np.random.default_rng(21)
    N = 50000
    all_tracks = np.arange(N, dtype=np.int64)
    all_albums = rng.integers(0, 5000, N)
    all_artists = rng.integers(0, 3000, N)
    y = synth.spotify_example(rng, 9, o=4, n_tracks=N, n_albums=5000, n_artists=3000)
    metrics, top = eval_step(model, params, y, all_tracks, all_albums, all_artists, k=500)
    om, oorder = osp.eval_step(A, R, y, all_tracks, all_albums, all_artists, k=500)
    aff = osp.eval_scores(A, R, y["album_context"], y["artist_context"], all_albums, all_artists)
    got = _np(top).astype(np.int64)
    # same candidate set up to fp32 near-ties at the cut; identical where the oracle scores are separated
    sep = np.abs(aff[oorder][:-1] - aff[oorder][1:]) > 1e-5
    assert len(set(got.tolist()) ^ set(oorder.tolist())) <= 4
    assert np.array_equal(got[:100][sep[:100]], oorder[:100][sep[:100]])
    np.testing.assert_allclose(_np(metrics), om, atol=2.0 / 9)


def test_pinterest_find_top_k_matches_oracle():
    from esrecsys_b200.pinterest.make_recommendations import find_top_k
    rng = np.random.default_rng(5)
    P = rng.standard_normal((20000, 64)).astype(np.float32)
    P[77] = P[5]                        # exact tie -> lower index first
    s = rng.standard_normal(64).astype(np.float32)
    val, idx = find_top_k(s, P, 10)
    oval, oidx = ostl.find_top_k(s, P, 10)
    assert np.array_equal(_np(idx), oidx)
    np.testing.assert_allclose(_np(val), oval, rtol=RTOL, atol=ATOL)


def test_device_negative_sampler_bit_exact_and_uniform():
    from esrecsys_b200 import engine
    from oracle import index as oidx
    got = engine.sample_uniform(7, 3, 1000, 2262291).cpu().numpy()
    assert np.array_equal(got, oidx.sample_uniform(7, 3, 1000, 2262291))
    assert not np.array_equal(got, engine.sample_uniform(7, 4, 1000, 2262291).cpu().numpy())     # new step, new draw
    big = engine.sample_uniform(1, 0, 1 << 20, 64).cpu().numpy()
    assert big.min() == 0 and big.max() == 63                                                    # upper bound exclusive
    cnt = np.bincount(big, minlength=64)
    assert np.abs(cnt - (1 << 14)).max() < 6 * np.sqrt(1 << 14)                                  # flat to 6 sigma
