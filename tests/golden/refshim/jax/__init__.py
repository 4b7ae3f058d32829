"""NumPy-backed stand-in for the parts of `jax` the reference imports (see ../README.md)."""
import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import random  # noqa: F401
from . import scipy  # noqa: F401
from .numpy import _wrap


class _Config:
    def update(self, *a, **k):   # jax.config.update("jax_enable_x64", True): NumPy is FP64 already
        return None


config = _Config()


def jit(fun, *a, **k):
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for j in range(n):
            call = [a if ax is None else _wrap(_np.asarray(_np.take(a, j, axis=ax))) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if isinstance(outs[0], tuple):
            return tuple(_wrap(_np.stack([_np.asarray(o[k]) for o in outs])) for k in range(len(outs[0])))
        return _wrap(_np.stack([_np.asarray(o) for o in outs]))
    return mapped


def grad(fun, argnums=0):
    """d fun / d args[argnums] by 4th-order central differences (Richardson extrapolation of two steps)."""
    def g(*args):
        x0 = _np.array(args[argnums], dtype=float)
        out = _np.zeros_like(x0)
        flat = out.reshape(-1)
        for k in range(x0.size):
            h = 1e-3 * max(1.0, abs(float(x0.reshape(-1)[k])))

            def f(step):
                x = x0.copy()
                x.reshape(-1)[k] += step
                a = list(args)
                a[argnums] = _wrap(x)
                return float(fun(*a))
            d1 = (f(h) - f(-h)) / (2 * h)
            d2 = (f(h / 2) - f(-h / 2)) / h
            flat[k] = (4.0 * d2 - d1) / 3.0
        return _wrap(out)
    return g
