"""Randomised LIVE check of the oracle against the reference's own source files (run by tests/test_ref_live.py, here only:
it needs /root/reference).  Same mechanism as make_ref_golden.py -- the reference's unmodified modules imported on top of
tests/golden/refshim -- but instead of writing fixtures it draws many random shapes / id patterns per trainer and compares
the oracle (float64) with what the reference's functions return, case by case.

    python tests/golden/ref_live_check.py wikipedia|spotify|pinterest|records [n_cases] [seed]

Prints "OK <n>" or raises.  One trainer per process: the three directories reuse the module names `models`,
`input_pipeline`."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_ref_golden as mk  # noqa: E402  (_enter / _np helpers)

TOL = 1e-11


def _close(a, b, what, tol=TOL):
    import numpy as np
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float(np.abs(a - b).max()) if a.size else 0.0
    assert err <= tol * max(1.0, float(np.abs(b).max()) if b.size else 1.0), (what, err)


def wikipedia(n, seed):
    mk._enter("wikipedia")
    import numpy as np
    import jax.numpy as jnp
    import optax
    from flax.training import train_state
    import train_cooccurence as tc
    from models import Glove
    from oracle import glove as og
    rng = np.random.default_rng(seed)
    for case in range(n):
        V = int(rng.integers(3, 200))
        D = int(rng.choice([1, 2, 4, 8, 16, 33]))
        B = int(rng.integers(1, 150))
        E = rng.standard_normal((V, D)) / np.sqrt(D)
        b = rng.standard_normal(V) * 0.1
        steps = 2
        hot = rng.integers(0, V, 3)
        ids = rng.integers(0, V, (steps, 2, B))
        dup = rng.random((steps, 2, B)) < 0.4                      # heavy duplication; i == j allowed
        ids = np.where(dup, hot[rng.integers(0, 3, (steps, 2, B))], ids).astype(np.int32)
        x = np.clip(rng.lognormal(0, 2.0, (steps, B)), 0, 1e4)
        x[rng.random((steps, B)) < 0.1] = 100.0                    # exactly x_max
        x[rng.random((steps, B)) < 0.05] = 0.0                     # zero weight
        model = Glove(num_embeddings=V, features=D)
        params = {"_token_embedding": {"embedding": jnp.asarray(E)}, "_bias": {"embedding": jnp.asarray(b.reshape(V, 1))}}
        _close(og.forward_literal(E, b, ids[0, 0], ids[0, 1]), mk._np(model.apply({"params": params}, jnp.asarray(ids[0]))), "forward")
        state = train_state.TrainState.create(apply_fn=model.apply, params=params, tx=optax.adam(1e-3))
        Eo, bo = E.copy(), b.copy()
        st = dict(count=0, muE=np.zeros_like(Eo), nuE=np.zeros_like(Eo), mub=np.zeros_like(bo), nub=np.zeros_like(bo))
        for s in range(steps):
            grads, loss = tc.apply_model(state, ids[s], x[s])
            gr = og.loss_and_grads(Eo, bo, ids[s, 0], ids[s, 1], x[s])
            dE, db = og.dense_grads(V, gr, D)
            _close(gr.loss, float(loss), "loss")
            _close(dE, mk._np(grads["_token_embedding"]["embedding"]), "dE")
            _close(db, mk._np(grads["_bias"]["embedding"])[:, 0], "db")
            state = tc.update_model(state, grads)
            og.step_adam(Eo, bo, st, ids[s, 0], ids[s, 1], x[s], 1e-3)
            _close(Eo, mk._np(state.params["_token_embedding"]["embedding"]), "E after adam", 1e-9)
            _close(bo, mk._np(state.params["_bias"]["embedding"])[:, 0], "b after adam", 1e-9)
        tok = rng.integers(0, V, int(rng.integers(1, 9))).astype(np.int32)
        sc, idx = tc.find_knn(model, state.params, jnp.asarray(tok))
        osc, oidx = og.find_knn(Eo, tok)
        _close(osc, mk._np(sc), "knn scores", 1e-9)
        ref_idx = mk._np(idx)
        srt = np.take_along_axis(mk._np(sc), ref_idx, axis=0)
        sep = np.ones(ref_idx.shape, bool)
        if V > 1:
            gap = np.abs(np.diff(srt, axis=0)) > 1e-9
            sep[1:] &= gap
            sep[:-1] &= gap
        assert np.array_equal(oidx[sep], ref_idx[sep]), "knn order"
    print("OK", n)


def spotify(n, seed):
    mk._enter("spotify")
    import numpy as np
    import jax.numpy as jnp
    import optax
    from flax.training import train_state
    import models
    import train_spotify as ts
    from oracle import spotify as osp
    rng = np.random.default_rng(seed)
    ORDER = ("track_context", "album_context", "artist_context", "next_track", "next_album", "next_artist",
             "neg_track", "neg_album", "neg_artist")
    NA, NR = osp.MAX_ALBUMS, osp.NUM_ARTISTS
    for case in range(n):
        F = int(rng.choice([1, 2, 4, 8]))
        m, o = int(rng.integers(1, 20)), int(rng.integers(1, 70))
        pool_a = rng.integers(0, 3 * NA, 12)                      # small pools: many shared albums / artists
        pool_r = rng.integers(0, NR, 10)
        x = {}
        for role, cnt in (("context", 5), ("next", m), ("neg", o)):
            x["track_" + role if role == "context" else role + "_track"] = rng.integers(0, 1000, cnt)
            x["album_" + role if role == "context" else role + "_album"] = pool_a[rng.integers(0, 12, cnt)]
            x["artist_" + role if role == "context" else role + "_artist"] = pool_r[rng.integers(0, 10, cnt)]
        A = np.zeros((NA, F))
        R = np.zeros((NR, F))
        A[np.unique(pool_a % NA)] = rng.standard_normal((np.unique(pool_a % NA).size, F)) * rng.choice([0.3, 1.0, 2.0])
        R[np.unique(pool_r)] = rng.standard_normal((np.unique(pool_r).size, F)) * rng.choice([0.3, 1.0, 2.0])
        reg = float(rng.choice([0.5, 1.5, 10.0]))
        spot = models.SpotifyModel(feature_size=F)
        params = {"params": {"album_embed": {"embedding": jnp.asarray(A)}, "artist_embed": {"embedding": jnp.asarray(R)}}}
        res = spot.apply(params, *[x[k] for k in ORDER])
        out = osp.forward(A, R, x["album_context"], x["artist_context"], x["next_album"], x["next_artist"],
                          x["neg_album"], x["neg_artist"])
        for got, want, k in zip(out, res, ("pos", "neg", "ctx_self", "next_self", "neg_self", "l2")):
            _close(got, mk._np(want), k)
        lr, mom = 0.01, 0.98
        state = train_state.TrainState.create(apply_fn=spot.apply, params=params, tx=optax.sgd(learning_rate=lr, momentum=mom))
        Ao, Ro, trA, trR = A.copy(), R.copy(), np.zeros_like(A), np.zeros_like(R)
        for s in range(2):
            state, loss = ts.train_step(state, dict(x), reg)
            ol = osp.train_step(Ao, Ro, trA, trR, x, reg, lr, mom)
            _close(ol, float(loss), "loss")
            _close(Ao, mk._np(state.params["params"]["album_embed"]["embedding"]), "A", 1e-10)
            _close(Ro, mk._np(state.params["params"]["artist_embed"]["embedding"]), "R", 1e-10)
    print("OK", n)


def pinterest(n, seed):
    mk._enter("pinterest")
    import numpy as np
    import jax.numpy as jnp
    import optax
    from flax import linen as nn
    from flax.training import train_state
    import models
    import train_shop_the_look as stl_train
    import make_recommendations as mr
    from oracle import optim as oopt
    from oracle import stl as ostl
    rng = np.random.default_rng(seed)
    shape = {}

    class Tower(nn.Module):
        filters: object
        output_size: int
        which: str = "scene"

        def __call__(self, x, train: bool = True):
            return self.param("embedding", nn.initializers.default_embed_init, (shape[self.which], self.output_size))[x]

    made = []

    def cnn(filters, output_size):
        made.append(1)
        return Tower(filters=filters, output_size=output_size, which="scene" if len(made) % 2 == 1 else "product")

    models.CNN = cnn
    for case in range(n):
        NS, NP = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        D, B = int(rng.choice([1, 3, 8, 32])), int(rng.integers(1, 40))
        shape.update(scene=NS, product=NP)
        S = rng.standard_normal((NS, D)) * rng.choice([0.2, 0.6, 1.5])
        P = rng.standard_normal((NP, D)) * rng.choice([0.2, 0.6, 1.5])
        sc, po, ne = rng.integers(0, NS, B), rng.integers(0, NP, B), rng.integers(0, NP, B)
        reg, bs = float(rng.choice([0.0, 0.1, 1.0])), float(rng.choice([B, 16]))
        stl = models.STLModel(output_size=D)
        params = {"params": {"scene_cnn": {"embedding": jnp.asarray(S)}, "product_cnn": {"embedding": jnp.asarray(P)}}}
        state = train_state.TrainState.create(apply_fn=stl.apply, params=params, tx=optax.adam(learning_rate=1e-3))
        So, Po = S.copy(), P.copy()
        muS, nuS, muP, nuP, count = np.zeros_like(S), np.zeros_like(S), np.zeros_like(P), np.zeros_like(P), 0
        for s in range(2):
            _close(ostl.eval_loss(So[sc], Po[po], Po[ne]), float(stl_train.eval_step(state, jnp.asarray(sc), jnp.asarray(po), jnp.asarray(ne))), "eval")
            state, loss = stl_train.train_step(state, jnp.asarray(sc), jnp.asarray(po), jnp.asarray(ne), reg, bs)
            ol, ds, dp, dn = ostl.triplet_loss_and_grads(So[sc], Po[po], Po[ne], reg, bs)
            _close(ol, float(loss), "loss")
            dS, dP = np.zeros_like(So), np.zeros_like(Po)
            np.add.at(dS, sc, ds)
            np.add.at(dP, po, dp)
            np.add.at(dP, ne, dn)
            So, muS, nuS, _ = oopt.adam_update(So, dS, muS, nuS, count, 1e-3)
            Po, muP, nuP, count = oopt.adam_update(Po, dP, muP, nuP, count, 1e-3)
            _close(So, mk._np(state.params["params"]["scene_cnn"]["embedding"]), "S", 1e-9)
            _close(Po, mk._np(state.params["params"]["product_cnn"]["embedding"]), "P", 1e-9)
        k = int(rng.integers(1, NP + 1))
        v, i = mr.find_top_k(jnp.asarray(So[0]), jnp.asarray(Po), k)
        ov, oi = ostl.find_top_k(So[0], Po, k)
        _close(ov, mk._np(v), "topk values", 1e-9)
    print("OK", n)


def records(n, seed):
    """Wire formats: random CooccurrenceRow / TokenStat messages serialised by the reference's own generated protobuf
    module (wikipedia/nlp_pb2.py) and parsed back by it exactly as cooccurrence_matrix.py:69-83 / token_dictionary.py
    do, against the native decoder (esr_decode_cooccur_b64, host code) and the host TokenStat parser.  Bit-exact."""
    mk._enter("wikipedia")
    import base64
    import numpy as np
    import nlp_pb2 as nlp_pb
    from esrecsys_b200.wikipedia import cooccurrence_matrix as cm
    from esrecsys_b200.wikipedia import token_dictionary as td
    rng = np.random.default_rng(seed)
    for case in range(n):
        lines, ii, jj, cc = [], [], [], []
        for r in range(int(rng.integers(1, 40))):
            row = nlp_pb.CooccurrenceRow()
            row.index = int(rng.choice([0, 1, 127, 128, 16383, 16384, 2 ** 31 - 1, int(rng.integers(0, 2 ** 31))]))
            k = int(rng.choice([0, 1, 2, 63, 64, 200]))
            row.other_index.extend(int(v) for v in rng.integers(0, 2 ** 31, k))
            vals = rng.lognormal(0, 3, k).astype(np.float32)
            if k:
                vals[rng.integers(0, k)] = np.float32(rng.choice([0.0, 1e-42, 3.4e38, 1.0 / 3]))
            row.count.extend(float(v) for v in vals)
            lines.append(base64.b64encode(row.SerializeToString()))
            back = nlp_pb.CooccurrenceRow()
            back.ParseFromString(base64.b64decode(lines[-1]))
            for t in range(len(back.other_index)):
                ii.append(back.index); jj.append(back.other_index[t]); cc.append(back.count[t])
        text = b"".join(ln + b"\n" for ln in lines)
        i, j, c, used = cm.decode_text(text)
        assert used == len(text)
        assert np.array_equal(i, np.asarray(ii, np.int64).astype(np.int32)) and np.array_equal(j, np.asarray(jj, np.int64).astype(np.int32))
        assert np.array_equal(c.view(np.uint32), np.asarray(cc, np.float32).view(np.uint32))
        assert cm.encode_row(5, [int(v) for v in jj[:7]], [float(v) for v in cc[:7]]) == nlp_pb.CooccurrenceRow(
            index=5, other_index=[int(v) for v in jj[:7]], count=[float(v) for v in cc[:7]]).SerializeToString()
        ts = nlp_pb.TokenStat(token="".join(chr(int(v)) for v in rng.choice([97, 233, 0x4e2d, 0x1F600, 45], int(rng.integers(0, 9)))),
                              url="u%d" % case if case % 3 else "", frequency=int(rng.integers(0, 2 ** 40)),
                              doc_frequency=int(rng.integers(0, 2 ** 20)), index=int(rng.integers(0, 2 ** 31)))
        raw = ts.SerializeToString()
        got = td.parse_token_stat(raw)
        assert got == {"token": ts.token, "url": ts.url, "frequency": ts.frequency, "doc_frequency": ts.doc_frequency,
                       "index": ts.index}, (got, ts)
        assert td.encode_token_stat(ts.token, ts.url, ts.frequency, ts.doc_frequency, ts.index) == raw
    # the dictionary class itself: tokenizer, min-hash, embedding indices and their printed names on random strings
    import token_dictionary as ref_td                       # /root/reference/wikipedia/token_dictionary.py
    path = os.path.join(HERE, "token.tstat.pb.b64.bz2")
    want, got = ref_td.TokenDictionary(path), td.TokenDictionary(path)
    alphabet = list("abcxyzABC019 _-.,;:/\\[]{}()'\"!?@#|<>+*&^%$\t\n") + ["é", "ü", "中", "\U0001F600", "the", "w010"]
    for case in range(50 * n):
        text = "".join(rng.choice(alphabet, int(rng.integers(0, 30))))
        toks = want.simple_tokenize(text)
        assert got.simple_tokenize(text) == toks, text
        assert got.get_embedding_indices(toks) == want.get_embedding_indices(toks), toks
        for t in toks:
            assert got.minhash(t) == want.minhash(t) and got.get_token_index(t) == want.get_token_index(t)
    for idx in list(range(1, 60)) + [65583, 70000]:          # 0 is skipped: the reference tests `is 0`
        assert got.get_token_from_embedding_index(idx) == want.get_token_from_embedding_index(idx)
    assert got.get_embedding_dictionary_size() == want.get_embedding_dictionary_size()
    assert got.get_max_doc_frequency() == want.get_max_doc_frequency() and got.get_doc_frequency(3) == want.get_doc_frequency(3)
    print("OK", n)


if __name__ == "__main__":
    which = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    {"wikipedia": wikipedia, "spotify": spotify, "pinterest": pinterest, "records": records}[which](n, seed)
