"""safe-bayesian-optimization_b200 -- B200-native grid hot path of dleeim/Safe-Bayesian-Optimization.

Layout: ``csrc/`` (sm_100a kernels + the C ABI of include/sbo_b200.h), ``_capi`` (ctypes binding),
``engine`` (host driver) and the host-side mirror of the reference interface:
``models.GP_Safe``, ``models.SafeOpt``, ``models.GoOSE``, ``models.GP_TR``, ``utils.utils_SafeOpt``, ``utils.utils_GoOSE``,
``drivers`` (the reference's BO loops),
``problems`` (NumPy restatements of the black-box plants used as fixtures).

The directory name is not a Python identifier; import it as ``sbo_b200`` (repo-root shim) or put this
directory on ``sys.path`` and use ``from models import SafeOpt`` exactly like the reference.
"""
from . import _capi  # noqa: F401
from .engine import GridEngine, SboError  # noqa: F401

__all__ = ["GridEngine", "SboError"]
