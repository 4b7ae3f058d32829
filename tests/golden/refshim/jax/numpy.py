"""jax.numpy stand-in: NumPy with arrays that carry JAX's functional-update `.at[...]`."""
import sys as _sys
import types as _types

import numpy as _np


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = _np.array(self.arr, copy=True).view(Arr)
        out[self.idx] = v
        return out

    def add(self, v):
        out = _np.array(self.arr, copy=True).view(Arr)
        out[self.idx] += v
        return out


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class Arr(_np.ndarray):
    @property
    def at(self):
        return _At(self)


def _wrap(x):
    if isinstance(x, _np.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, (list, tuple)) and x and all(isinstance(e, _np.ndarray) for e in x):
        return type(x)(_wrap(e) for e in x)
    return x


def _wrapping(f):
    def g(*a, **k):
        return _wrap(f(*a, **k))
    g.__name__ = getattr(f, "__name__", "f")
    return g


class _Module(_types.ModuleType):
    def __init__(self, name, target):
        super().__init__(name)
        self.__dict__["_target"] = target

    def __getattr__(self, name):
        v = getattr(self.__dict__["_target"], name)
        if isinstance(v, type) or not callable(v):
            return v
        return _wrapping(v)


linalg = _Module(__name__ + ".linalg", _np.linalg)
_self = _sys.modules[__name__]


def __getattr__(name):
    v = getattr(_np, name)
    if isinstance(v, type) or not callable(v):
        return v
    return _wrapping(v)
