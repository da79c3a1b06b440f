"""jax stand-in (see ../README.md): jit = convert-inputs-and-call, vmap via torch.vmap, value_and_grad via
torch.autograd, lax.top_k with jax's tie order, random = seeded numpy stand-in (NOT threefry)."""
import functools as _ft

import numpy as _np
import torch as _t

from . import numpy  # noqa: F401  (jax.numpy)
from .numpy import asarray as _a

__version__ = "0.0-refshim"


# ---- pytrees (dicts / lists / tuples / objects exposing tree_flatten-ish replace) ----
def _tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: _tree_map(f, tree[k], *(r[k] for r in rest)) for k in tree}
    if isinstance(tree, (list, tuple)) and not hasattr(tree, "_fields"):
        return type(tree)(_tree_map(f, v, *(r[i] for r in rest)) for i, v in enumerate(tree))
    if hasattr(tree, "_fields"):                                   # namedtuple
        return type(tree)(*(_tree_map(f, v, *(getattr(r, n) for r in rest)) for n, v in zip(tree._fields, tree)))
    if tree is None:
        return None
    return f(tree, *rest)


class tree_util:                                                   # noqa: N801
    tree_map = staticmethod(_tree_map)


tree_map = _tree_map


def _to_arrays(x):
    """What jit tracing does to host inputs: numpy arrays / python numbers in containers become device arrays."""
    if isinstance(x, _np.ndarray):
        return _a(x)
    if isinstance(x, dict):
        return {k: _to_arrays(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)) and not hasattr(x, "_fields"):
        return type(x)(_to_arrays(v) for v in x)
    return x


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: jit(f)

    @_ft.wraps(fun)
    def wrapped(*args, **kwargs):
        return fun(*[_to_arrays(a) for a in args], **{k: _to_arrays(v) for k, v in kwargs.items()})
    return wrapped


def vmap(fun, in_axes=0, out_axes=0):
    in_dims = tuple(in_axes) if isinstance(in_axes, (list, tuple)) else in_axes
    return _t.vmap(fun, in_dims=in_dims, out_dims=out_axes)


def value_and_grad(fun, has_aux=False):
    def wrapped(params, *args, **kwargs):
        leaves = []

        def leaf(p):
            q = _a(p).detach().clone().requires_grad_(True)
            leaves.append(q)
            return q
        p2 = _tree_map(leaf, params)
        out = fun(p2, *args, **kwargs)
        val, aux = (out if has_aux else (out, None))
        grads = _t.autograd.grad(val, leaves, allow_unused=True)
        it = iter(g if g is not None else _t.zeros_like(q) for g, q in zip(grads, leaves))
        gtree = _tree_map(lambda _p: next(it), p2)
        val = val.detach()
        return ((val, aux), gtree) if has_aux else (val, gtree)
    return wrapped


def grad(fun):
    return lambda *a, **k: value_and_grad(fun)(*a, **k)[1]


class lax:                                                         # noqa: N801
    @staticmethod
    def top_k(x, k):
        """values descending; equal values keep ascending index order (XLA TopK / jax.lax.top_k)."""
        x = _a(x)
        order = _t.argsort(-x, dim=-1, stable=True)[..., :k]
        return _t.gather(x, -1, order), order


class random:                                                      # noqa: N801
    """Seeded numpy stand-in.  NOT jax's threefry: streams differ from the reference's; the parity harness passes
    initial tables and sampled negatives in as given inputs."""

    @staticmethod
    def PRNGKey(seed):                                             # noqa: N802
        return _np.array([0, seed], dtype=_np.uint32)

    @staticmethod
    def split(key, num=2):
        ss = _np.random.SeedSequence([int(key[0]), int(key[1])])
        return [_np.array(s.generate_state(2), dtype=_np.uint32) for s in ss.spawn(num)]

    @staticmethod
    def _rng(key):
        return _np.random.default_rng([int(key[0]), int(key[1])])

    @staticmethod
    def randint(key, shape, minval, maxval):
        return _a(random._rng(key).integers(minval, maxval, size=tuple(shape)))   # maxval exclusive, as in jax

    @staticmethod
    def normal(key, shape):
        return _a(random._rng(key).standard_normal(tuple(shape)))
