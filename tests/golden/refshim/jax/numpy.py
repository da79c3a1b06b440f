"""jax.numpy stand-in on torch float64 tensors (see ../README.md).  Only what the reference's hot path calls."""
import numpy as _np
import torch as _t

FLOAT = _t.float64
_t.set_default_dtype(FLOAT)      # python-scalar * bool/int array promotes to the array float type, as in jax
float32 = _t.float32
float64 = _t.float64
int32 = _t.int32
int64 = _t.int64
ndarray = _t.Tensor

if not hasattr(_t.Tensor, "astype"):
    _t.Tensor.astype = lambda self, dt: self.to(FLOAT if dt in (_t.float32, _t.float64, float) else dt)


def asarray(x, dtype=None):
    """numpy / python values -> tensors; floats become float64, integers int64 (index arithmetic is exact)."""
    if isinstance(x, _t.Tensor):
        t = x
    else:
        a = _np.asarray(x)
        if a.dtype.kind == "f":
            t = _t.tensor(a, dtype=FLOAT)
        elif a.dtype.kind in "iu":
            t = _t.tensor(a.astype(_np.int64))
        elif a.dtype.kind == "b":
            t = _t.tensor(a)
        else:
            raise TypeError(f"refshim: unsupported dtype {a.dtype}")
    if dtype is not None:
        t = t.astype(dtype)
    return t


array = asarray
_a = asarray


def ones_like(x):
    return _t.ones_like(_a(x))


def zeros_like(x):
    return _t.zeros_like(_a(x))


def minimum(a, b):
    return _t.minimum(_a(a), _a(b))


def maximum(a, b):
    return _t.maximum(_a(a), _a(b))


def power(a, p):
    return _t.pow(_a(a), p)


def log10(a):
    return _t.log10(_a(a))


def square(a):
    return _t.square(_a(a))


def sqrt(a):
    return _t.sqrt(_a(a))


def mean(a, axis=None):
    a = _a(a)
    return a.mean() if axis is None else a.mean(dim=axis)


def sum(a, axis=None):          # noqa: A001  (mirrors jnp.sum)
    a = _a(a)
    return a.sum() if axis is None else a.sum(dim=axis)


def max(a, axis=None):          # noqa: A001
    # jax's reduce_max VJP splits the cotangent equally among tied maxima; so does torch.amax
    a = _a(a)
    return _t.amax(a) if axis is None else _t.amax(a, dim=axis)


def min(a, axis=None):          # noqa: A001
    a = _a(a)
    return _t.amin(a) if axis is None else _t.amin(a, dim=axis)


def dot(a, b):
    """numpy.dot semantics for the ranks the reference uses (1-D.1-D inner, N-D.1-D, 2-D.2-D)."""
    a, b = _a(a), _a(b)
    if a.dim() == 1 and b.dim() == 1:
        return (a * b).sum()
    return _t.matmul(a, b)


def isin(element, test_elements):
    return _t.isin(_a(element), _a(test_elements))


def mod(a, n):
    return _t.remainder(_a(a), n)


def concatenate(arrays, axis=0):
    return _t.cat([_a(x) for x in arrays], dim=axis)


def stack(arrays, axis=0):
    return _t.stack([_a(x) for x in arrays], dim=axis)


def flip(a, axis=None):
    a = _a(a)
    if axis is None:
        return _t.flip(a, dims=list(range(a.dim())))
    return _t.flip(a, dims=[axis % a.dim()])


def arange(start, stop=None, step=1, dtype=None):
    if stop is None:
        start, stop = 0, start
    return _t.arange(start, stop, step, dtype=_t.int64)


def argsort(a, axis=-1):
    # jax sorts are stable (ties keep index order)
    return _t.argsort(_a(a), dim=axis, stable=True)


def take(a, idx, axis=0):
    return _t.index_select(_a(a), axis, _a(idx).reshape(-1)).reshape(tuple(_a(idx).shape) + tuple(_a(a).shape[1:]))
