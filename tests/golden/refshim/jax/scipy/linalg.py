import scipy.linalg as _sl

from ..numpy import _wrap


def solve_triangular(a, b, lower=False, **k):
    return _wrap(_sl.solve_triangular(a, b, lower=lower, **k))
