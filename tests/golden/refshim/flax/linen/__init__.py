"""flax.linen stand-in: dataclass-style Module with setup(), init(), apply(method=), nn.Embed, nn.relu.

Mechanics restated from flax 0.5/0.6's documented behaviour: class annotations are constructor fields;
`setup()` runs lazily on a bound copy; a submodule assigned to `self.<name>` in setup owns `params[<name>]`;
`nn.Embed(num_embeddings, features)` holds one parameter `embedding` of shape (num_embeddings, features),
default init variance_scaling(1.0, 'fan_in', 'normal', out_axis=0) = N(0, 1/features), and `__call__(ids)` is
`jnp.take(embedding, ids, axis=0)`.
"""
import copy as _copy

import numpy as _np
import torch as _t

from jax.numpy import FLOAT as _FLOAT
from jax.numpy import asarray as _a

from . import initializers  # noqa: F401


def relu(x):
    return _t.relu(_a(x))            # d/dx at 0 is 0 in both torch and jax.nn.relu


def swish(x):
    x = _a(x)
    return x * _t.sigmoid(x)


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = []
        for klass in reversed(cls.__mro__):
            for name in getattr(klass, "__annotations__", {}):
                if name not in fields and not name.startswith("_flax"):
                    fields.append(name)
        cls._fields_ = tuple(fields)

    def __init__(self, *args, **kwargs):
        names = list(self._fields_)
        for n, v in zip(names, args):
            kwargs[n] = v
        for n in names:
            if n in kwargs:
                object.__setattr__(self, n, kwargs[n])
            elif not hasattr(type(self), n):
                raise TypeError(f"{type(self).__name__}: missing field {n!r}")
        object.__setattr__(self, "_scope", None)

    # -- binding ----------------------------------------------------------------------------
    def _bind(self, params, mode, rng):
        m = _copy.copy(self)
        object.__setattr__(m, "_scope", {"params": params, "mode": mode, "rng": rng})
        object.__setattr__(m, "_in_setup", True)
        m.setup()
        object.__setattr__(m, "_in_setup", False)
        return m

    def setup(self):
        pass

    def __setattr__(self, name, value):
        sc = self.__dict__.get("_scope")
        if isinstance(value, Module) and sc is not None and self.__dict__.get("_in_setup"):
            if sc["mode"] == "init":
                sub = sc["params"].setdefault(name, {})
            else:
                sub = sc["params"][name]
            value = value._bind(sub, sc["mode"], sc["rng"])
        object.__setattr__(self, name, value)

    def param(self, name, init_fn, shape):
        sc = self._scope
        if sc["mode"] == "init":
            if name not in sc["params"]:
                sc["params"][name] = init_fn(sc["rng"], tuple(shape))
        return _a(sc["params"][name])

    # -- public surface ---------------------------------------------------------------------
    def init(self, rngs, *args, method=None, **kwargs):
        rng = _np.random.default_rng([int(v) for v in _np.asarray(rngs).reshape(-1)])
        params = {}
        m = self._bind(params, "init", rng)
        fn = getattr(m, method.__name__) if method is not None else m
        fn(*args, **kwargs)
        return {"params": params}

    def apply(self, variables, *args, method=None, mutable=False, rngs=None, **kwargs):
        m = self._bind(variables["params"], "apply", None)
        fn = getattr(m, method.__name__) if method is not None else m
        out = fn(*args, **kwargs)
        return (out, {}) if mutable else out


class Embed(Module):
    num_embeddings: int
    features: int
    embedding_init: object = None

    def __init__(self, num_embeddings, features, embedding_init=None, **kw):
        super().__init__(num_embeddings=num_embeddings, features=features,
                         embedding_init=embedding_init or initializers.default_embed_init, **kw)

    def __call__(self, inputs):
        table = self.param("embedding", self.embedding_init, (self.num_embeddings, self.features))
        return table[_a(inputs)]                                   # jnp.take(embedding, inputs, axis=0)


def compact(fn):
    """Only so that class bodies decorated with @nn.compact import; compact modules (the CNN towers, out of scope) are
    never called through the shim."""
    return fn
