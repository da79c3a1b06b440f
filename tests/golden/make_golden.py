"""Generates the golden vectors under tests/golden/ (committed as small .npz files).

    python tests/golden/make_golden.py

The reference cannot run in this image (jax / flax / optax / tensorflow are absent and there is no
network -- DESIGN.md section 4), so these vectors do NOT come from the reference itself: they come from
an independent second derivation -- a literal float64 torch-autograd transcription of each reference
forward + loss (the code below follows the cited reference lines statement by statement) and torch.optim
for the optimizer rules.  The oracle (oracle/*.py) and the CUDA path are both checked against them.
"Parity unpinned" stays true: these pin our reading of the reference, not the reference's bytes.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from esrecsys_b200 import synth  # noqa: E402  (seeded generators only; no compute)

F64 = torch.float64


def glove_case(V, D, B, seed):
    """wikipedia/models.py:21-38 (Glove.__call__) + wikipedia/train_cooccurence.py:76-87 (glove_loss, value_and_grad)
    + optax.adam (train_cooccurence.py:171, :101) and optax.adagrad, one step each from the same start."""
    E, _ = synth.init_glove_tables(V, D, seed)
    b = (np.random.default_rng(seed + 5).standard_normal(V) * 0.05).astype(np.float32)
    ids, counts = synth.glove_batches(V, B, 1, seed)
    i, j, x = ids[0, 0], ids[0, 1], counts[0]
    Et = torch.tensor(E, dtype=F64, requires_grad=True)
    bt = torch.tensor(b.reshape(-1, 1), dtype=F64, requires_grad=True)
    it, jt = torch.tensor(i, dtype=torch.long), torch.tensor(j, dtype=torch.long)
    xt = torch.tensor(x, dtype=F64)
    token1, token2 = Et[it], Et[jt]                                    # models.py:31,33
    bias1, bias2 = bt[it], bt[jt]                                      # models.py:32,34   (B,1)
    dot = (token1 * token2).sum(-1)                                    # models.py:35-36   (B,)
    output = dot + bias1 + bias2                                       # models.py:37      (B,B) broadcast
    ones = torch.ones_like(xt)
    weight = torch.minimum(ones, xt / 100.0) ** 0.75                   # train_cooccurence.py:79-81
    log_target = torch.log10(1.0 + xt)                                 # :82
    loss = torch.mean(torch.square(log_target - output) * weight)      # :83
    loss.backward()
    dE, db = Et.grad.numpy(), bt.grad.numpy()[:, 0]
    out = dict(E=E, b=b, i=i, j=j, x=x, loss=np.float64(loss.item()), dE=dE, db=db)
    # one optax.adagrad step (initial_accumulator_value=0.1, eps=1e-7) -- torch.optim.Adagrad differs in eps placement,
    # so the published rule is evaluated directly in float64
    for name, p, g in (("E", E.astype(np.float64), dE), ("b", b.astype(np.float64), db)):
        acc = 0.1 + g * g
        out["adagrad_" + name] = p - 0.05 * g / np.sqrt(acc + 1e-7)
        out["adagrad_acc_" + name] = np.where(g != 0, acc, 0.1)
    # one optax.adam step == torch.optim.Adam defaults (b1=.9, b2=.999, eps=1e-8, bias-corrected, eps outside sqrt)
    P = [torch.tensor(E, dtype=F64, requires_grad=True), torch.tensor(b, dtype=F64, requires_grad=True)]
    opt = torch.optim.Adam(P, lr=1e-3)
    P[0].grad, P[1].grad = torch.tensor(dE), torch.tensor(db)
    opt.step()
    out["adam_E"], out["adam_b"] = P[0].detach().numpy(), P[1].detach().numpy()
    return out


def stl_case(B, D, seed):
    """pinterest/models.py:67-72 (row-wise scores) + pinterest/train_shop_the_look.py:99-107 (loss, grads)."""
    rng = np.random.default_rng(seed)
    s, p, n = (rng.standard_normal((B, D)).astype(np.float32) * 0.6 for _ in range(3))
    st, pt, nt = (torch.tensor(v, dtype=F64, requires_grad=True) for v in (s, p, n))
    pos_score = (st * pt).sum(-1)                                       # models.py:67-68
    neg_score = (st * nt).sum(-1)                                       # models.py:71-72
    triplet = torch.sum(torch.relu(1.0 + neg_score - pos_score))        # train_shop_the_look.py:99-100

    def reg_fn(e):
        return torch.relu(torch.sqrt(torch.sum(torch.square(e), -1)) - 1.0)   # :101

    reg = torch.sum(reg_fn(st) + reg_fn(pt) + reg_fn(nt))               # :102-103
    loss = (triplet + 0.1 * reg) / 16.0                                 # :104 (divisor = batch_size flag, :59)
    loss.backward()
    return dict(scene=s, pos=p, neg=n, loss=np.float64(loss.item()), pos_score=pos_score.detach().numpy(),
                neg_score=neg_score.detach().numpy(), d_scene=st.grad.numpy(), d_pos=pt.grad.numpy(), d_neg=nt.grad.numpy())


def spotify_case(seed, m):
    """spotify/models.py:48-91 (SpotifyModel.__call__) + spotify/train_spotify.py:91-105 (loss) on one playlist."""
    rng = np.random.default_rng(seed)
    F, VA, VR, o, reg = 8, 50, 40, 12, 1.5
    A = (rng.standard_normal((VA, F)) / np.sqrt(F)).astype(np.float32) * 2.0
    R = (rng.standard_normal((VR, F)) / np.sqrt(F)).astype(np.float32) * 2.0
    x = {k: rng.integers(0, hi, n) for k, hi, n in (("album_context", 3 * VA, 5), ("artist_context", VR, 5),
                                                     ("next_album", 3 * VA, m), ("next_artist", VR, m),
                                                     ("neg_album", 3 * VA, o), ("neg_artist", VR, o))}
    x["next_artist"][1] = x["artist_context"][0]          # isin boosts, ties of the max, duplicate rows
    x["next_album"][2] = x["album_context"][1]
    x["artist_context"][4] = x["artist_context"][3]
    x["album_context"][4] = x["album_context"][3]
    At, Rt = torch.tensor(A, dtype=F64, requires_grad=True), torch.tensor(R, dtype=F64, requires_grad=True)

    def emb(album, artist):                               # models.py:33-46
        return torch.cat([At[torch.tensor(album % VA)], Rt[torch.tensor(artist)]], -1)

    ctx, nxt, neg = emb(x["album_context"], x["artist_context"]), emb(x["next_album"], x["next_artist"]), emb(
        x["neg_album"], x["neg_artist"])
    isin = lambda a, b: torch.tensor(np.isin(a, b).astype(np.float64))
    pos_aff = torch.max(nxt @ ctx.T, -1).values + 0.1 * isin(x["next_album"], x["album_context"]) + 0.1 * isin(
        x["next_artist"], x["artist_context"])            # models.py:74-76
    neg_aff = torch.max(neg @ ctx.T, -1).values + 0.1 * isin(x["neg_album"], x["album_context"]) + 0.1 * isin(
        x["neg_artist"], x["artist_context"])             # models.py:78-80
    all_e = torch.cat([ctx, nxt, neg], 0)
    l2 = torch.sqrt(torch.sum(torch.square(all_e), -1))   # models.py:82-83
    ctx_self = torch.flip(ctx, [0]) @ ctx.T               # models.py:85-87
    nxt_self = torch.flip(nxt, [0]) @ nxt.T
    neg_self = torch.flip(neg, [0]) @ neg.T
    relu = torch.relu
    loss = (relu(1.0 + torch.mean(neg_aff) - torch.mean(pos_aff)) + relu(1.0 + torch.max(neg_aff) - torch.min(pos_aff))
            + torch.mean(relu(0.5 - ctx_self)) + torch.mean(relu(0.5 - nxt_self)) + torch.mean(relu(neg_self))
            + torch.sum(relu(l2 - reg)))                  # train_spotify.py:91-105
    loss.backward()
    out = {k: v.astype(np.int64) for k, v in x.items()}
    out.update(A=A, R=R, reg=np.float64(reg), loss=np.float64(loss.item()), dA=At.grad.numpy(), dR=Rt.grad.numpy(),
               pos_aff=pos_aff.detach().numpy(), neg_aff=neg_aff.detach().numpy(), l2=l2.detach().numpy())
    return out


def main():
    torch.manual_seed(0)
    np.savez_compressed(os.path.join(HERE, "glove_V60_D8_B32.npz"), **glove_case(60, 8, 32, 0))
    np.savez_compressed(os.path.join(HERE, "glove_V500_D64_B256.npz"), **glove_case(500, 64, 256, 1))
    np.savez_compressed(os.path.join(HERE, "stl_B16_D32.npz"), **stl_case(16, 32, 2))
    np.savez_compressed(os.path.join(HERE, "spotify_m7.npz"), **spotify_case(3, 7))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
