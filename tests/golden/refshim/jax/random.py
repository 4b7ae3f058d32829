"""jax.random stand-in (NumPy generator keyed by an integer; NOT threefry-compatible)."""
import numpy as _np

from .numpy import _wrap


def PRNGKey(seed):
    return _np.array([0, int(seed)], dtype=_np.uint32)


def split(key, num=2):
    ss = _np.random.SeedSequence([int(k) for k in _np.asarray(key).ravel()])
    return [_np.array(c.generate_state(2), dtype=_np.uint32) for c in ss.spawn(num)]


def _gen(key):
    return _np.random.default_rng([int(k) for k in _np.asarray(key).ravel()])


def normal(key, shape=()):
    return _wrap(_np.asarray(_gen(key).standard_normal(shape)))


def uniform(key, shape=()):
    return _wrap(_np.asarray(_gen(key).random(shape)))
