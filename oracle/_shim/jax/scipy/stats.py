"""``jax.scipy.stats.norm`` stand-in (ndtr with jax's erf/erfc branches; pdf = exp(logpdf)). TEST INFRASTRUCTURE ONLY."""
import numpy as _np
from scipy.special import erf as _erf, erfc as _erfc

from ..numpy import _wrap


class norm:
    @staticmethod
    def cdf(x):
        x = _np.asarray(x)
        dt = x.dtype.type
        h = dt(_np.sqrt(0.5))
        w = x * h
        z = _np.abs(w)
        with _np.errstate(invalid="ignore"):
            y = _np.where(z < h, dt(1) + _erf(w), _np.where(w > 0, dt(2) - _erfc(z), _erfc(z)))
        return _wrap((dt(0.5) * y).astype(x.dtype))

    @staticmethod
    def pdf(x):
        x = _np.asarray(x)
        dt = x.dtype.type
        return _wrap(_np.exp((dt(_np.log(2 * _np.pi)) + x * x) / dt(-2)).astype(x.dtype))
