"""Inert stand-in so `import tensorflow as tf` and the module-level schema literals of the reference's
input pipelines evaluate; nothing on the hot path calls tf (it is tf.data iterators and image decoding only).
Any attribute resolves to an inert recorder; calling into it from the hot path would return another recorder and
fail loudly at the first arithmetic use."""
import sys as _sys
import types as _types


class _Inert:
    def __init__(self, name):
        self._name = name

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Inert(self._name + "." + k)

    def __call__(self, *a, **k):
        return _Inert(self._name + "()")

    def __repr__(self):
        return f"<refshim inert {self._name}>"


class _Mod(_types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Inert("tf." + k)


_sys.modules[__name__].__class__ = _Mod
