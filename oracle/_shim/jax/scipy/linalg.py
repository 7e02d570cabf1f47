"""``jax.scipy.linalg`` stand-in on SciPy (dtype preserving). TEST INFRASTRUCTURE ONLY."""
import numpy as _np
import scipy.linalg as _sl

from ..numpy import _wrap


def cholesky(a, lower=False):
    return _wrap(_sl.cholesky(_np.asarray(a), lower=lower, check_finite=False))


def solve_triangular(a, b, lower=False):
    return _wrap(_sl.solve_triangular(_np.asarray(a), _np.asarray(b), lower=lower, check_finite=False))
