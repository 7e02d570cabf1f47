"""``jax.random`` stand-in on the oracle's Threefry restatement (partitionable mode). TEST INFRASTRUCTURE ONLY."""
import numpy as _np

from oracle import prng as _p
from .numpy import _cfg, _wrap

key = _p.key
PRNGKey = _p.key
fold_in = _p.fold_in


def split(k, num=2):
    return _p.split(k, num)


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    n = int(_np.prod(shape))
    return _wrap(_p.uniform(k, n, minval, maxval).reshape(shape))


def randint(k, shape, minval, maxval, dtype=_np.int32):
    n = int(_np.prod(shape))
    return _wrap(_p.randint(k, n, minval, maxval).reshape(shape))
