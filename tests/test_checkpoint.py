"""Checkpoint format (flax msgpack state dict, restated from flax 0.5.2 serialization.py) -- host logic, CPU."""
import os

import msgpack
import numpy as np
import pytest
import torch

from esrecsys_b200 import checkpoint as ck
from esrecsys_b200 import optim as O
from esrecsys_b200.train_state import TrainState


def _state(tx, seed=0):
    g = torch.Generator().manual_seed(seed)
    params = {"_token_embedding": {"embedding": torch.randn(11, 4, generator=g)}, "_bias": {"embedding": torch.randn(11, 1, generator=g)}}
    return TrainState.create(apply_fn=None, params=params, tx=tx)


def test_ndarray_wire_format_known_answer():
    a = np.array([1.0, 2.0], np.float32)
    got = ck.msgpack_serialize({"a": a})
    payload = msgpack.packb(((2,), "float32", a.tobytes()), use_bin_type=True)
    want = b"\x81" + b"\xa1a" + b"\xc7" + bytes([len(payload)]) + b"\x01" + payload     # map1, 'a', ext8(type 1)
    assert got == want
    # payload layout: array(3)[array(1)[2], str 'float32', bin 8 bytes]
    assert payload == b"\x93\x91\x02\xa7float32\xc4\x08" + a.tobytes()
    back = ck.msgpack_restore(got)
    assert back["a"].dtype == np.float32 and back["a"].tolist() == [1.0, 2.0]
    s = ck.msgpack_serialize({"step": np.int32(7)})
    assert s[:6] == b"\x81\xa4step" and s[6] == 0xc7 and s[8] == 3    # ext8, type 3 = npscalar (flax _MsgpackExtType.npscalar)
    assert ck.msgpack_restore(s)["step"] == 7


@pytest.mark.parametrize("tx,first_keys", [(O.adam(1e-3), {"count", "mu", "nu"}), (O.sgd(1e-3, 0.98), {"trace"}),
                                           (O.adagrad(0.05), {"sum_of_squares"})])
def test_trainstate_roundtrip_and_tree(tx, first_keys):
    st = _state(tx)
    st.step = 42
    if tx.kind == "adam":
        st.opt_state["count"] = 42
        st.opt_state["mu"]["_bias/embedding"].fill_(0.25)
    sd = ck.msgpack_restore(ck.to_bytes(st))
    assert set(sd) == {"step", "params", "opt_state"} and int(sd["step"]) == 42
    assert set(sd["opt_state"]) == {"0", "1"} and sd["opt_state"]["1"] == {} and set(sd["opt_state"]["0"]) == first_keys
    assert set(sd["params"]) == {"_token_embedding", "_bias"} and sd["params"]["_bias"]["embedding"].shape == (11, 1)
    other = _state(tx, seed=9)
    ck.from_bytes(other, ck.to_bytes(st))
    assert other.step == 42
    for a, b in zip(torch.utils._pytree.tree_leaves(other.params), torch.utils._pytree.tree_leaves(st.params)):
        assert torch.equal(a, b)
    if tx.kind == "adam":
        assert other.opt_state["count"] == 42 and float(other.opt_state["mu"]["_bias/embedding"][0, 0]) == 0.25
    bad = _state(tx)
    bad.params["_bias"]["embedding"] = torch.zeros(12, 1)
    with pytest.raises(ValueError):
        ck.from_bytes(bad, ck.to_bytes(st))


def test_chunked_leaves(monkeypatch):
    monkeypatch.setattr(ck, "MAX_CHUNK_SIZE", 64)
    a = np.arange(100, dtype=np.float32).reshape(10, 10)
    data = ck.msgpack_serialize({"w": a})
    raw = msgpack.unpackb(data, ext_hook=ck._ext_unpack, raw=False, strict_map_key=False)
    assert raw["w"]["__msgpack_chunked_array__"] is True and len(raw["w"]["chunks"]) == 7
    assert np.array_equal(ck.msgpack_restore(data)["w"], a)


def test_save_restore_keep(tmp_path):
    st = _state(O.sgd(1e-3, 0.98))
    d = str(tmp_path / "ck")
    for step in (10, 20, 30, 40):
        st.step = step
        ck.save_checkpoint(d, st, step, keep=3)
    assert sorted(os.listdir(d)) == ["checkpoint_20", "checkpoint_30", "checkpoint_40"]
    fresh = _state(O.sgd(1e-3, 0.98), seed=5)
    assert ck.restore_checkpoint(d, fresh).step == 40
    assert ck.restore_checkpoint(str(tmp_path / "none"), fresh) is fresh
    with pytest.raises(ValueError):
        ck.save_checkpoint(d, st, 40)


def test_sharded_manifest(tmp_path):
    V, D, world = 10, 4, 3
    E = np.arange(V * D, dtype=np.float32).reshape(V, D)
    for r in range(world):
        ck.save_sharded(str(tmp_path), 5, r, world, {"rows": E[r::world], "bias": E[r::world, 0]}, V, D)
    out = ck.load_sharded_dense(str(tmp_path), 5)
    assert np.array_equal(out["rows"], E) and np.array_equal(out["bias"], E[:, 0])


def test_train_cooccurence_save_and_resume_names(tmp_path):
    from esrecsys_b200.wikipedia.train_cooccurence import resume_state, save_state
    st = _state(O.adam(1e-3))
    st.step = 7
    path = save_state(st, 7, str(tmp_path))
    assert os.path.basename(path) == "checkpoint-00007.flax"
    other = _state(O.adam(1e-3), seed=3)
    assert resume_state(other, path).step == 7
    assert torch.equal(other.params["_bias"]["embedding"], st.params["_bias"]["embedding"])
