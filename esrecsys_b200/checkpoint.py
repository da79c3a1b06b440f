"""Checkpoint / artifact format of the reference: ``flax.serialization.to_bytes / from_bytes`` of a
``TrainState`` (wikipedia/train_cooccurence.py:129-134, :173-177, :188-192; the ``.flax`` files
pinterest/make_embeddings.py:82-85 loads) and ``flax.training.checkpoints.save_checkpoint /
restore_checkpoint`` (spotify/train_spotify.py:244-245, :262-264).  SURVEY.md 8(f) N2.

flax is not importable in this image, so the byte format is RESTATED from flax 0.5.2 ``serialization.py``
(the pinned version, wikipedia/requirements.txt:20) and is unverified against flax-produced bytes:

* the state dict of a TrainState is ``{'step', 'params', 'opt_state'}`` (``apply_fn`` and ``tx`` are not pytree
  nodes); tuples / NamedTuples of the optax state become dicts -- tuple items keyed '0', '1', ... and
  NamedTuple fields by name: ``optax.adam`` -> ``{'0': {'count', 'mu', 'nu'}, '1': {}}``,
  ``optax.sgd(momentum)`` -> ``{'0': {'trace'}, '1': {}}``, ``optax.adagrad`` -> ``{'0': {'sum_of_squares'}, '1': {}}``;
* the dict is msgpack-packed (``strict_types``), ndarrays as ExtType 1 whose payload is
  ``msgpack.packb((shape, dtype.name, bytes), use_bin_type=True)``, NumPy scalars as ExtType 3 (same payload);
* leaves above 2**30 bytes are split into ``{'__msgpack_chunked_array__': True, 'shape': ..., 'chunks': {'0':..}}``.

Tensors live on the GPU; (de)serialisation goes through host NumPy.  Row-sharded tables write one file per
rank plus a manifest (``save_sharded``) -- the reference has no sharded format.
"""
from __future__ import annotations

import json
import os
import re

import msgpack
import numpy as np
import torch

_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
MAX_CHUNK_SIZE = 2 ** 30


def _ndarray_to_bytes(arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)
    return msgpack.packb((arr.shape, arr.dtype.name, arr.tobytes("C")), use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
    return np.frombuffer(buf, dtype=np.dtype(dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name)).reshape(
        tuple(shape)).copy()


def _ext_pack(x):
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(_EXT_NDARRAY, _ndarray_to_bytes(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(_EXT_COMPLEX, msgpack.packb((x.real, x.imag)))
    return x


def _ext_unpack(code, data):
    if code == _EXT_NDARRAY:
        return _ndarray_from_bytes(data)
    if code == _EXT_NPSCALAR:
        return _ndarray_from_bytes(data)[()]
    if code == _EXT_COMPLEX:
        re_, im = msgpack.unpackb(data)
        return complex(re_, im)
    return msgpack.ExtType(code, data)


def _chunk(arr: np.ndarray):
    if arr.size * arr.dtype.itemsize <= MAX_CHUNK_SIZE:
        return arr
    flat = arr.reshape(-1)
    per = max(1, MAX_CHUNK_SIZE // arr.dtype.itemsize)
    return {"__msgpack_chunked_array__": True, "shape": {str(i): int(d) for i, d in enumerate(arr.shape)},   # _tuple_to_dict
            "chunks": {str(i): flat[s:s + per] for i, s in enumerate(range(0, flat.size, per))}}


def _unchunk(d):
    if isinstance(d, dict):
        if d.get("__msgpack_chunked_array__"):
            parts = [d["chunks"][str(i)] for i in range(len(d["chunks"]))]
            return np.concatenate(parts).reshape(tuple(d["shape"][str(i)] for i in range(len(d["shape"]))))
        return {k: _unchunk(v) for k, v in d.items()}
    return d


def _to_host(tree):
    if isinstance(tree, dict):
        return {str(k): _to_host(v) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return {str(i): _to_host(v) for i, v in enumerate(tree)}
    if torch.is_tensor(tree):
        return _chunk(tree.detach().cpu().numpy())
    if isinstance(tree, np.ndarray):
        return _chunk(tree)
    return tree


def msgpack_serialize(pytree) -> bytes:
    return msgpack.packb(_to_host(pytree), default=_ext_pack, strict_types=True)


def msgpack_restore(data: bytes):
    return _unchunk(msgpack.unpackb(data, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


# ------------------------------------------------------------------------------------------------
# TrainState <-> state dict
# ------------------------------------------------------------------------------------------------
def _nest(flat):
    out = {}
    for path, v in flat.items():
        d = out
        keys = path.split("/")
        for k in keys[:-1]:
            d = d.setdefault(k, {})
        d[keys[-1]] = v
    return out


def to_state_dict(state):
    """``flax.serialization.to_state_dict(TrainState)`` for the TrainState shim (esrecsys_b200/train_state.py)."""
    tx, opt = state.tx, state.opt_state
    if tx.kind == "adam":
        first = {"count": np.asarray(opt["count"], np.int32), "mu": _nest(opt["mu"]), "nu": _nest(opt["nu"])}
    elif tx.kind == "sgd":
        first = {"trace": _nest(opt["trace"])}
    elif tx.kind == "adagrad":
        first = {"sum_of_squares": _nest(opt["acc"])}
    else:
        raise ValueError(tx.kind)
    return {"step": np.asarray(state.step, np.int32), "params": state.params, "opt_state": {"0": first, "1": {}}}


def to_bytes(state) -> bytes:
    """``flax.serialization.to_bytes(state)`` (train_cooccurence.py:133, :190)."""
    return msgpack_serialize(to_state_dict(state))


def _assign(dst, src, path=""):
    """Copy the arrays of ``src`` into the tensors of ``dst`` in place; structure and shapes must match (flax raises too)."""
    if isinstance(dst, dict):
        if set(dst) != set(src):
            raise ValueError("checkpoint keys %s do not match the target's %s at '%s'" % (sorted(src), sorted(dst), path))
        for k in dst:
            _assign(dst[k], src[k], path + "/" + k)
    else:
        a = np.asarray(src)
        if tuple(a.shape) != tuple(dst.shape):
            raise ValueError("shape mismatch at '%s': %s vs %s" % (path, a.shape, tuple(dst.shape)))
        dst.copy_(torch.from_numpy(a).to(dst.dtype))


def from_bytes(state, data: bytes):
    """``flax.serialization.from_bytes(target, encoded)`` (train_cooccurence.py:177): restores ``data`` into ``state``
    (tensors are overwritten in place) and returns it."""
    sd = msgpack_restore(data)
    if set(sd) != {"step", "params", "opt_state"}:
        raise ValueError("not a TrainState checkpoint: keys %s" % sorted(sd))
    state.step = int(sd["step"])
    _assign(state.params, sd["params"], "params")
    first, tx, opt = sd["opt_state"]["0"], state.tx, state.opt_state
    if tx.kind == "adam":
        opt["count"] = int(first["count"])
        _assign(_nest(opt["mu"]), first["mu"], "opt_state/0/mu")
        _assign(_nest(opt["nu"]), first["nu"], "opt_state/0/nu")
    elif tx.kind == "sgd":
        _assign(_nest(opt["trace"]), first["trace"], "opt_state/0/trace")
    else:
        _assign(_nest(opt["acc"]), first["sum_of_squares"], "opt_state/0/sum_of_squares")
    return state


# ------------------------------------------------------------------------------------------------
# flax.training.checkpoints naming: <dir>/<prefix><step>, keep the newest `keep`
# ------------------------------------------------------------------------------------------------
def _steps(ckpt_dir, prefix):
    out = []
    for f in os.listdir(ckpt_dir) if os.path.isdir(ckpt_dir) else []:
        m = re.fullmatch(re.escape(prefix) + r"(\d+)", f)
        if m:
            out.append((int(m.group(1)), os.path.join(ckpt_dir, f)))
    return sorted(out)


def save_checkpoint(ckpt_dir, target, step, prefix="checkpoint_", keep=1, overwrite=False):
    """``checkpoints.save_checkpoint(ckpt_dir, target=state, step=i, keep=3)`` (train_spotify.py:262-264)."""
    os.makedirs(ckpt_dir, exist_ok=True)
    have = _steps(ckpt_dir, prefix)
    if have and have[-1][0] >= step and not overwrite:
        raise ValueError("a checkpoint at step %d >= %d exists; pass overwrite=True" % (have[-1][0], step))
    path = os.path.join(ckpt_dir, "%s%d" % (prefix, step))
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(to_bytes(target) if hasattr(target, "opt_state") else msgpack_serialize(target))
    os.replace(tmp, path)
    for _, old in _steps(ckpt_dir, prefix)[:-keep]:
        os.remove(old)
    return path


def restore_checkpoint(ckpt_dir, target, step=None, prefix="checkpoint_"):
    """``checkpoints.restore_checkpoint(ckpt_dir, state)`` (train_spotify.py:244-245): newest (or ``step``) checkpoint
    into ``target``; returns ``target`` unchanged when the directory holds none (as flax does)."""
    have = _steps(ckpt_dir, prefix)
    if step is not None:
        have = [h for h in have if h[0] == step]
    if not have:
        return target
    data = open(have[-1][1], "rb").read()
    return from_bytes(target, data) if hasattr(target, "opt_state") else msgpack_restore(data)


def save_sharded(ckpt_dir, step, rank, world, shard_state: dict, V, D):
    """Row-sharded table: one msgpack file per rank + a JSON manifest written by rank 0 (cyclic ownership:
    global row = local * world + rank).  No reference counterpart."""
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, "shard_%d-of-%d_%d" % (rank, world, step))
    with open(path + ".tmp", "wb") as f:
        f.write(msgpack_serialize(shard_state))
    os.replace(path + ".tmp", path)
    if rank == 0:
        json.dump({"step": int(step), "world": int(world), "V": int(V), "D": int(D), "ownership": "cyclic",
                   "files": ["shard_%d-of-%d_%d" % (r, world, step) for r in range(world)]},
                  open(os.path.join(ckpt_dir, "manifest_%d.json" % step), "w"))
    return path


def load_sharded_dense(ckpt_dir, step):
    """Reassembles the dense ``{'rows', 'bias', ...}`` arrays from a sharded checkpoint (host NumPy)."""
    man = json.load(open(os.path.join(ckpt_dir, "manifest_%d.json" % step)))
    world, V = man["world"], man["V"]
    out = {}
    for r, fn in enumerate(man["files"]):
        sd = msgpack_restore(open(os.path.join(ckpt_dir, fn), "rb").read())
        for k, a in sd.items():
            a = np.asarray(a)
            if k not in out:
                out[k] = np.zeros((V,) + a.shape[1:], a.dtype)
            out[k][r::world] = a
    return out
