"""Golden vectors produced by EXECUTING the reference's own source files (tests/golden/ref_*.npz).

    python tests/golden/make_ref_golden.py            # all three trainers, one subprocess each
    python tests/golden/make_ref_golden.py wikipedia  # or wikipedia_e2e / spotify / pinterest

Runs HERE only (needs /root/reference, which does not exist on the GPU box); the vectors are committed.

jax / flax / optax / tensorflow are not installed in this image, so the reference's scripts are imported
UNMODIFIED from /root/reference/<dir> with `tests/golden/refshim` (a minimal torch-float64 stand-in for the part of
those libraries the hot path calls -- see refshim/README.md for exactly what that pins) first on sys.path, and THEIR
functions produce every number below:

  wikipedia: models.Glove.init/apply/score_all, train_cooccurence.apply_model / update_model / train_epoch /
             find_knn / dump_knn, token_dictionary.TokenDictionary
  spotify:   models.SpotifyModel.init/apply, train_spotify.train_step / eval_step
  pinterest: models.STLModel.__call__ (its CNN towers swapped for ID-embedding towers, the north star's
             substitution), train_shop_the_look.train_step / eval_step, make_recommendations.find_top_k

Inputs (tables, id batches, negatives) are seeded numpy draws saved next to the outputs.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def _enter(subdir):
    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    sys.path[:0] = [os.path.join(HERE, "refshim"), os.path.join(REF, subdir)]
    import warnings
    warnings.filterwarnings("ignore")


def _np(t):
    import torch
    return t.detach().numpy() if isinstance(t, torch.Tensor) else t


# ------------------------------------------------------------------------------------------------ wikipedia
def wikipedia():
    _enter("wikipedia")
    import numpy as np
    import jax
    import jax.numpy as jnp
    import optax
    from flax.training import train_state
    import train_cooccurence as tc                      # the reference's trainer module, unmodified
    from models import Glove                            # the reference's model, unmodified
    from token_dictionary import TokenDictionary        # the reference's dictionary, unmodified

    logged = []
    tc.logging.info = lambda fmt, *a: logged.append(fmt % a)         # dump_knn reports through logging.info

    for V, D, B, seed, steps in ((60, 8, 32, 0, 3), (500, 64, 256, 1, 3)):
        rng = np.random.default_rng(seed)
        E = (rng.standard_normal((V, D)) / np.sqrt(D)).astype(np.float32)
        b = (rng.standard_normal(V) * 0.05).astype(np.float32)
        # (i, j, count) batches as CooccurrenceGenerator yields them: ids in [1, V), i > j, fractional counts on both
        # sides of x_max = 100; the Zipf-ish draw repeats head rows many times per batch
        ids = np.zeros((steps, 2, B), np.int32)
        for s in range(steps):
            a = 1 + (rng.zipf(1.3, 4 * B) % (V - 1))
            c = 1 + (rng.zipf(1.3, 4 * B) % (V - 1))
            keep = np.flatnonzero(a != c)[:B]
            ids[s, 0], ids[s, 1] = np.maximum(a[keep], c[keep]), np.minimum(a[keep], c[keep])
        x = np.clip(rng.lognormal(0.0, 1.5, (steps, B)), 1.0 / 9, 1e4).astype(np.float32)
        x[:, :4] = [0.5, 100.0, 250.0, 99.0]
        tokens = np.array([1, 2, 7, V // 2, V - 1], np.int32)

        model = Glove(num_embeddings=V, features=D)
        tree = model.init(jax.random.PRNGKey(seed), ids[0])
        assert {k: {kk: tuple(vv.shape) for kk, vv in v.items()} for k, v in tree["params"].items()} == {
            "_token_embedding": {"embedding": (V, D)}, "_bias": {"embedding": (V, 1)}}
        assert float(tree["params"]["_bias"]["embedding"].abs().max()) == 0.0
        params = {"_token_embedding": {"embedding": jnp.asarray(E)}, "_bias": {"embedding": jnp.asarray(b.reshape(V, 1))}}
        out = dict(E=E, b=b, ids=ids, x=x, tokens=tokens)
        if V <= 64:
            out["forward0"] = _np(model.apply({"params": params}, jnp.asarray(ids[0])))      # (B, B), models.py:37

        def run(tx, tag, lr_note):
            state = train_state.TrainState.create(apply_fn=model.apply, params=params, tx=tx)      # :172
            for s in range(steps):
                grads, loss = tc.apply_model(state, ids[s], x[s])                                   # :71-89
                if s == 0 and tag == "adam":
                    out["dE0"], out["db0"] = _np(grads["_token_embedding"]["embedding"]), _np(grads["_bias"]["embedding"])[:, 0]
                out.setdefault("loss_%s" % tag, []).append(float(loss))
                state = tc.update_model(state, grads)                                               # :99-101
                if s == 0 and V <= 64:
                    out["E_%s_step1" % tag] = _np(state.params["_token_embedding"]["embedding"])
                    out["b_%s_step1" % tag] = _np(state.params["_bias"]["embedding"])[:, 0]
            assert state.step == steps
            out["E_%s" % tag] = _np(state.params["_token_embedding"]["embedding"])
            out["b_%s" % tag] = _np(state.params["_bias"]["embedding"])[:, 0]
            out["loss_%s" % tag] = np.asarray(out["loss_%s" % tag])
            return state

        st = run(optax.adam(1e-3), "adam", "train_cooccurence.py:171 (reference optimizer, default --learning_rate)")
        out["adam_mu_E"], out["adam_nu_E"] = _np(st.opt_state[0].mu["_token_embedding"]["embedding"]), _np(
            st.opt_state[0].nu["_token_embedding"]["embedding"])
        st = run(optax.adagrad(0.05), "adagrad", "north-star optimizer through the reference's update_model")
        out["adagrad_acc_E"] = _np(st.opt_state[0].sum_of_squares["_token_embedding"]["embedding"])
        out["adagrad_acc_b"] = _np(st.opt_state[0].sum_of_squares["_bias"]["embedding"])[:, 0]
        run(optax.sgd(0.05, momentum=0.9), "sgdm", "optax.sgd(momentum) through update_model")

        # train_epoch (:103-112): state after `steps` batches + the epoch's mean loss
        state = train_state.TrainState.create(apply_fn=model.apply, params=params, tx=optax.adam(1e-3))
        it = iter([(ids[s], x[s]) for s in range(steps)])
        state, train_loss = tc.train_epoch(state, steps, it)
        out["epoch_loss"] = np.float64(train_loss)
        assert np.array_equal(_np(state.params["_token_embedding"]["embedding"]), out["E_adam"])

        # find_knn (:91-97) and dump_knn (:114-126) on the adam-trained parameters
        scores, indices = tc.find_knn(model, state.params, jnp.asarray(tokens))
        out["knn_scores"], out["knn_indices"] = _np(scores), _np(indices).astype(np.int32)
        if V <= 64:
            td = TokenDictionary(os.path.join(HERE, "token.tstat.pb.b64.bz2"))
            del logged[:]
            tc.dump_knn(model, state.params, jnp.asarray(tokens), td)
            out["dump_knn_lines"] = np.array(logged)
        name = "ref_glove_V%d_D%d_B%d.npz" % (V, D, B)
        np.savez_compressed(os.path.join(HERE, name), **out)
        print("wrote", name, "losses", out["loss_adam"])


def wikipedia_e2e():
    """The body of train_cooccurence.main (:136-192) on a small corpus FILE, with the reference's own pieces end to end:
    TokenDictionary -> CooccurrenceGenerator.get_batch -> Glove -> TrainState(optax.adam) -> [dump_knn, train_epoch] x
    epochs.  (main itself needs wandb.init and tf.data; get_dataset (:108-115) only wraps get_batch into a tf Dataset.)
    Writes the corpus part (tests/golden/e2e_cooccur/part-00000.bz2) and ref_glove_e2e.npz."""
    _enter("wikipedia")
    import base64
    import bz2
    import numpy as np
    import jax
    import jax.numpy as jnp
    import optax
    from flax.training import train_state
    import nlp_pb2 as nlp_pb
    import train_cooccurence as tc
    from cooccurrence_matrix import CooccurrenceGenerator
    from models import Glove
    from token_dictionary import TokenDictionary

    logged = []
    tc.logging.info = lambda fmt, *a: logged.append(fmt % a)
    D, B, epochs, steps_per_epoch, lr, seed = 8, 64, 2, 3, 0.05, 7
    td = TokenDictionary(os.path.join(HERE, "token.tstat.pb.b64.bz2"))
    num_tokens = td.get_embedding_dictionary_size()                               # :148
    terms = "the,of,zürich,w010,w033"
    debug_tokens = jnp.asarray(np.array([td.get_embedding_index(w) for w in terms.split(",")], np.int32))   # :150-154

    # corpus: rows as make_cooccurrence.py:83-100 emits them (index > every other_index, fractional counts)
    rng = np.random.default_rng(seed)
    part_dir = os.path.join(HERE, "e2e_cooccur")
    os.makedirs(part_dir, exist_ok=True)
    with bz2.open(os.path.join(part_dir, "part-00000.bz2"), "wb") as f:
        for index in range(2, 48):
            row = nlp_pb.CooccurrenceRow()
            row.index = index
            k = int(rng.integers(4, 18))
            others = rng.integers(1, index, k)
            row.other_index.extend(int(v) for v in others)
            row.count.extend(float(np.float32(v)) for v in np.clip(rng.lognormal(0.5, 1.6, k), 1.0 / 9, 400.0))
            f.write(base64.b64encode(row.SerializeToString()) + b"\n")

    model = Glove(num_embeddings=num_tokens, features=D)                             # :156
    train_data = CooccurrenceGenerator(os.path.join(part_dir, "part-?????.bz2"))     # :158
    train_iterator = train_data.get_batch(B, 0)                                      # what get_dataset wraps (:160)
    x, _ = next(train_iterator)                                                      # :164 (consumes the first batch)
    tree = model.init(jax.random.PRNGKey(seed), x)                                   # :165
    assert tuple(tree["params"]["_token_embedding"]["embedding"].shape) == (num_tokens, D)
    E = (np.random.default_rng(seed).standard_normal((num_tokens, D)) / np.sqrt(D)).astype(np.float32)
    params = {"_token_embedding": {"embedding": jnp.asarray(E)}, "_bias": {"embedding": jnp.asarray(np.zeros((num_tokens, 1)))}}
    state = train_state.TrainState.create(apply_fn=model.apply, params=params, tx=optax.adam(lr))     # :166-167
    out = dict(D=np.int64(D), B=np.int64(B), epochs=np.int64(epochs), steps_per_epoch=np.int64(steps_per_epoch),
               lr=np.float64(lr), seed=np.int64(seed), num_tokens=np.int64(num_tokens), terms=np.array(terms),
               debug_tokens=_np(debug_tokens).astype(np.int32), E_checksum=np.float64(E.astype(np.float64).sum()),
               first_batch_x=np.stack(x), train_loss=[])
    for step in range(epochs):                                                       # :174
        del logged[:]
        tc.dump_knn(model, state.params, debug_tokens, td)                          # :178
        out["knn_lines_%d" % step] = np.array(logged)
        state, train_loss = tc.train_epoch(state, steps_per_epoch, train_iterator)  # :179
        out["train_loss"].append(float(train_loss))
    del logged[:]
    tc.dump_knn(model, state.params, debug_tokens, td)
    out["knn_lines_%d" % epochs] = np.array(logged)
    out["train_loss"] = np.asarray(out["train_loss"])
    Ef = _np(state.params["_token_embedding"]["embedding"])
    assert np.array_equal(Ef[64:], E[64:].astype(np.float64))                        # rows beyond the corpus never move
    out["E_final_head"], out["b_final_head"] = Ef[:64], _np(state.params["_bias"]["embedding"])[:64, 0]
    np.savez_compressed(os.path.join(HERE, "ref_glove_e2e.npz"), **out)
    print("wrote ref_glove_e2e.npz train_loss", out["train_loss"], "V", num_tokens)
    print(out["knn_lines_%d" % epochs][0])


# ------------------------------------------------------------------------------------------------ spotify
def spotify():
    _enter("spotify")
    import numpy as np
    import jax
    import jax.numpy as jnp
    import optax
    from flax.training import train_state
    import models                                       # the reference's spotify/models.py, unmodified
    import train_spotify as ts                          # the reference's trainer module, unmodified

    F, o, N, steps, reg = 8, 64, 3000, 3, 3.5           # o = --num_negatives default (:60); reg where about half the rows bite
    rng = np.random.default_rng(11)
    spot = models.SpotifyModel(feature_size=F)
    # the corpus (input_pipeline.make_all_tracks_numpy): track id -> (album, artist); album ids exceed max_albums so
    # the mod in get_embeddings (models.py:42) and the un-modded isin (models.py:75) both matter
    all_tracks = np.arange(N, dtype=np.int64)
    all_albums = rng.integers(0, 300000, N).astype(np.int64)
    all_artists = rng.integers(0, 295861, N).astype(np.int64)

    def playlist(m):
        t = rng.integers(0, N - 1, 5 + m)
        t[4] = t[3]                                      # a duplicated context track (ties of the max over context)
        t[5 + 1] = t[0]                                  # a next track that is also in the context (isin boosts)
        ex = {"track_context": all_tracks[t[:5]], "album_context": all_albums[t[:5]], "artist_context": all_artists[t[:5]],
              "next_track": all_tracks[t[5:]], "next_album": all_albums[t[5:]], "next_artist": all_artists[t[5:]]}
        ex["next_artist"][2] = ex["artist_context"][1]   # artist-only match
        n = rng.integers(0, N - 1, o)                    # sample_negative's draw (:146), given as an input
        n[3] = t[2]                                      # a negative that happens to be a context track
        ex.update(neg_track=all_tracks[n], neg_album=all_albums[n], neg_artist=all_artists[n])
        return ex

    xs = [playlist(m) for m in (7, 5, 12)][:steps]
    y = playlist(9)
    ORDER = ("track_context", "album_context", "artist_context", "next_track", "next_album", "next_artist",
             "neg_track", "neg_album", "neg_artist")
    tree = jax.jit(spot.init)(jax.random.PRNGKey(0), *[xs[0][k] for k in ORDER])                 # :225-229
    assert {k: tuple(v["embedding"].shape) for k, v in tree["params"].items()} == {
        "album_embed": (100000, F), "artist_embed": (295861, F)}
    # tables: zero except the rows the corpus touches (keeps the fixture small; every row any step reads is random)
    arows = np.unique(all_albums % 100000)
    rrows = np.unique(all_artists)
    avals = (rng.standard_normal((arows.size, F)) * 0.9).astype(np.float32)
    rvals = (rng.standard_normal((rrows.size, F)) * 0.9).astype(np.float32)
    A = np.zeros((100000, F), np.float32); A[arows] = avals
    R = np.zeros((295861, F), np.float32); R[rrows] = rvals
    params = {"params": {"album_embed": {"embedding": jnp.asarray(A)}, "artist_embed": {"embedding": jnp.asarray(R)}}}
    out = dict(F=np.int64(F), reg=np.float64(reg), lr=np.float64(0.01), momentum=np.float64(0.98),
               all_tracks=all_tracks, all_albums=all_albums, all_artists=all_artists,
               arows=arows, rrows=rrows, avals=avals, rvals=rvals)
    for s, ex in enumerate(xs + [y]):
        for k in ORDER:
            out["x%d_%s" % (s, k)] = ex[k]

    res = spot.apply(params, *[xs[0][k] for k in ORDER])                                            # :231-235
    for k, v in zip(("pos_affinity", "neg_affinity", "context_self", "next_self", "neg_self", "l2"), res):
        out["fwd0_" + k] = _np(v)

    tx = optax.sgd(learning_rate=0.01, momentum=0.98)                                               # :238-241
    state = train_state.TrainState.create(apply_fn=spot.apply, params=params, tx=tx)                # :242-243
    train_step_fn = jax.jit(ts.train_step)                                                           # :248
    losses = []
    for s in range(steps):
        if s == 0:                                       # the gradient of step 0, for the forward+backward kernel alone
            probe = train_state.TrainState.create(apply_fn=spot.apply, params=params, tx=optax.sgd(learning_rate=1.0))
            p1, _ = train_step_fn(probe, dict(xs[0]), reg)
            out["dA0"] = (A.astype(np.float64) - _np(p1.params["params"]["album_embed"]["embedding"]))[arows]
            out["dR0"] = (R.astype(np.float64) - _np(p1.params["params"]["artist_embed"]["embedding"]))[rrows]
        state, loss = train_step_fn(state, dict(xs[s]), reg)                                        # :256-257
        losses.append(float(loss))
    out["losses"] = np.asarray(losses)
    A3, R3 = _np(state.params["params"]["album_embed"]["embedding"]), _np(state.params["params"]["artist_embed"]["embedding"])
    mask = np.ones(100000, bool); mask[arows] = False
    assert np.all(A3[mask] == 0)                          # rows outside the corpus never move
    out["A_rows_final"], out["R_rows_final"] = A3[arows], R3[rrows]
    tr = state.opt_state[0].trace["params"]
    out["A_trace_final"], out["R_trace_final"] = _np(tr["album_embed"]["embedding"])[arows], _np(tr["artist_embed"]["embedding"])[rrows]

    metrics = jax.jit(ts.eval_step)(state, dict(y), all_tracks, all_albums, all_artists)            # :113-131
    out["eval_metrics"] = _np(metrics)
    res = spot.apply(state.params, *[y[k] for k in ORDER[:6]], all_tracks, all_albums, all_artists)
    out["eval_affinity"] = _np(res[1])
    out["eval_top500"] = _np(jax.lax.top_k(res[1], 500)[1]).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_spotify.npz"), **out)
    print("wrote ref_spotify.npz losses", losses, "eval", out["eval_metrics"])


# ------------------------------------------------------------------------------------------------ pinterest
def pinterest():
    _enter("pinterest")
    import numpy as np
    import jax.numpy as jnp
    import optax
    from flax import linen as nn
    from flax.training import train_state
    import models                                       # the reference's pinterest/models.py, unmodified
    import train_shop_the_look as stl_train             # the reference's trainer module, unmodified
    import make_recommendations as mr                   # the reference's retrieval script, unmodified

    NS, NP, D, B, steps, reg = 120, 150, 32, 16, 3, 0.1   # B = --batch_size default (:59), reg = --regularization (:58)

    class IdTower(nn.Module):
        """ID-embedding tower with the CNN's call signature (north star: the CNN towers of pinterest/models.py:23-46
        are replaced by embedding tables); STLModel.setup constructs it as CNN(filters=, output_size=)."""
        filters: object
        output_size: int

    # STLModel.setup (models.py:52-55) builds scene_cnn then product_cnn; give each its own table height
    class SceneTower(IdTower):
        def __call__(self, x, train: bool = True):
            return self.param("embedding", nn.initializers.default_embed_init, (NS, self.output_size))[x]

    class ProductTower(IdTower):
        def __call__(self, x, train: bool = True):
            return self.param("embedding", nn.initializers.default_embed_init, (NP, self.output_size))[x]

    made = []

    def cnn_factory(filters, output_size):
        made.append(1)
        return (SceneTower if len(made) % 2 == 1 else ProductTower)(filters=filters, output_size=output_size)

    models.CNN = cnn_factory                             # the ONE substitution; STLModel itself runs as written
    stl = models.STLModel(output_size=D)                 # train_shop_the_look.py:169

    rng = np.random.default_rng(21)
    S = (rng.standard_normal((NS, D)) * 0.25).astype(np.float32)
    P = (rng.standard_normal((NP, D)) * 0.25).astype(np.float32)
    S[:6] *= 1.6                                         # some rows with norm > 1 so the regulariser bites
    P[:6] *= 1.6
    scene = rng.integers(0, NS, (steps, B)).astype(np.int64)
    pos = rng.integers(0, NP, (steps, B)).astype(np.int64)
    neg = rng.integers(0, NP, (steps, B)).astype(np.int64)
    scene[:, :3], pos[:, :3], neg[:, :3] = [0, 1, 2], [0, 1, 2], [3, 4, 0]
    neg[:, 5] = pos[:, 5]                                # pos == neg: hinge exactly at the margin
    params = {"params": {"scene_cnn": {"embedding": jnp.asarray(S)}, "product_cnn": {"embedding": jnp.asarray(P)}}}
    out = dict(S=S, P=P, scene=scene, pos=pos, neg=neg, reg=np.float64(reg), lr=np.float64(1e-3))

    r = stl.apply(params, jnp.asarray(scene[0]), jnp.asarray(pos[0]), jnp.asarray(neg[0]), True)       # models.py:63-74
    for k, v in zip(("pos_score", "neg_score", "scene_embed", "pos_embed", "neg_embed"), r):
        out["fwd0_" + k] = _np(v)
    out["scene_embed_method"] = _np(stl.apply(params, jnp.asarray(scene[0]), method=models.STLModel.get_scene_embed))
    out["product_embed_method"] = _np(stl.apply(params, jnp.asarray(pos[0]), method=models.STLModel.get_product_embed))

    probe = train_state.TrainState.create(apply_fn=stl.apply, params=params, tx=optax.sgd(learning_rate=1.0))
    p1, l0 = stl_train.train_step(probe, jnp.asarray(scene[0]), jnp.asarray(pos[0]), jnp.asarray(neg[0]), reg, B)
    out["dS0"] = S.astype(np.float64) - _np(p1.params["params"]["scene_cnn"]["embedding"])
    out["dP0"] = P.astype(np.float64) - _np(p1.params["params"]["product_cnn"]["embedding"])

    state = train_state.TrainState.create(apply_fn=stl.apply, params=params, tx=optax.adam(learning_rate=1e-3))  # :175-177
    losses, evals = [], []
    for s in range(steps):
        args = [jnp.asarray(v[s]) for v in (scene, pos, neg)]
        evals.append(float(stl_train.eval_step(state, *args)))                                          # :111-122
        state, loss = stl_train.train_step(state, *args, reg, B)                                        # :93-109
        losses.append(float(loss))
    out["losses"], out["eval_losses"] = np.asarray(losses), np.asarray(evals)
    out["S_final"] = _np(state.params["params"]["scene_cnn"]["embedding"])
    out["P_final"] = _np(state.params["params"]["product_cnn"]["embedding"])

    sc, idx = mr.find_top_k(jnp.asarray(S[7]), jnp.asarray(P), 10)                                      # make_recommendations.py:49-65
    out["topk_scores"], out["topk_indices"] = _np(sc), _np(idx).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_stl.npz"), **out)
    print("wrote ref_stl.npz losses", losses, "eval", evals)


if __name__ == "__main__":
    which = sys.argv[1:] or ["wikipedia", "wikipedia_e2e", "spotify", "pinterest"]
    if len(which) == 1:
        {"wikipedia": wikipedia, "wikipedia_e2e": wikipedia_e2e, "spotify": spotify, "pinterest": pinterest}[which[0]]()
    else:                                               # module names collide across the three directories
        for w in which:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), w])
