"""``jax.numpy`` stand-in: NumPy with a configurable default float and ``.at[].set``. TEST INFRASTRUCTURE ONLY."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import float32, float64, int32, uint32, bool_, pi, inf, newaxis  # noqa: F401

_cfg = {"float": _np.float32}


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, value):
                out = arr.copy()
                out[idx] = value
                return out
        return _Setter()


class ShimArray(_np.ndarray):
    @property
    def at(self):
        return _At(self)


ndarray = _np.ndarray


def _wrap(a):
    return _np.asarray(a).view(ShimArray)


def _fdt(dtype):
    return _cfg["float"] if dtype is None else dtype


def _canon(a):
    a = _np.asarray(a)
    if a.dtype == _np.float64 and _cfg["float"] is _np.float32:
        a = a.astype(_np.float32)
    if a.dtype == _np.int64:
        a = a.astype(_np.int32)
    return a


def _canon_out(x):
    if isinstance(x, tuple):
        return tuple(_canon_out(v) for v in x)
    if isinstance(x, (_np.ndarray, _np.generic)) and _np.asarray(x).dtype in (_np.float64, _np.int64):
        return _wrap(_canon(x))
    if isinstance(x, _np.ndarray):
        return _wrap(x)
    return x


def _canonicalising(fn):
    def wrapped(*a, **k):
        return _canon_out(fn(*a, **k))
    wrapped.__name__ = getattr(fn, "__name__", "fn")
    return wrapped


for _name in ("sum", "exp", "outer", "logaddexp", "mean", "dot", "log", "diag", "sqrt", "maximum", "minimum",
              "max", "min", "squeeze", "argmin", "abs", "sin", "cos", "square", "negative", "all", "any",
              "isclose", "allclose", "issubdtype", "equal", "matmul", "transpose", "reshape", "take"):
    globals()[_name] = _canonicalising(getattr(_np, _name))


def asarray(a, dtype=None):
    return _wrap(_canon(a) if dtype is None else _np.asarray(a, dtype=dtype))


array = asarray


def zeros(shape, dtype=None):
    return _wrap(_np.zeros(shape, dtype=_fdt(dtype)))


def ones(shape, dtype=None):
    return _wrap(_np.ones(shape, dtype=_fdt(dtype)))


def full(shape, fill_value, dtype=None):
    return _wrap(_np.full(shape, fill_value, dtype=_fdt(dtype) if dtype is None and isinstance(fill_value, float) else dtype))


def eye(n, dtype=None):
    return _wrap(_np.eye(n, dtype=_fdt(dtype)))


def linspace(a, b, num=50, dtype=None):
    return _wrap(_np.linspace(a, b, num).astype(_fdt(dtype)))


def arange(*a, dtype=None):
    return _wrap(_canon(_np.arange(*a)) if dtype is None else _np.arange(*a, dtype=dtype))


def stack(arrs, axis=0):
    return _wrap(_np.stack([_np.asarray(x) for x in arrs], axis=axis))


def concatenate(arrs, axis=0):
    return _wrap(_np.concatenate([_np.asarray(x) for x in arrs], axis=axis))


def where(c, a, b):
    a, b = _np.asarray(a), _np.asarray(b)
    out = _np.where(c, a, b)
    # weak-type promotion: a python scalar must not widen the array operand
    for ref in (a, b):
        if ref.ndim > 0 and out.dtype != ref.dtype and (a.ndim == 0 or b.ndim == 0):
            other = b if ref is a else a
            if other.ndim == 0 and _np.can_cast(other.dtype, ref.dtype, "same_kind"):
                out = out.astype(ref.dtype)
    return _wrap(_canon(out))


def pad(a, widths):
    return _wrap(_np.pad(_np.asarray(a), widths))


def argsort(a):
    return _wrap(_np.argsort(_np.asarray(a) + 0, kind="stable"))


def argmax(a):
    a = _np.asarray(a)
    if a.dtype.kind == "f" and _np.isnan(a).any():
        return _np.argmax(_np.isnan(a))
    return _np.argmax(a)


def round(a):  # noqa: A001
    return _wrap(_np.round(_np.asarray(a)))


def clip(a, lo, hi):
    a = _np.asarray(a)
    return _wrap(_np.clip(a, _np.asarray(lo, dtype=a.dtype) if a.dtype.kind == "f" else lo,
                          _np.asarray(hi, dtype=a.dtype) if a.dtype.kind == "f" else hi))
