"""``bayex.domain`` with the same types, asserts and dtypes (reference: bayex/domain.py), GPU-backed sampling.

``transform`` is host-side NumPy glue (clip / round-half-even); ``sample`` draws on the GPU with the Threefry
kernel (``bx_threefry_fill``), bit-compatible with ``jax.random.uniform`` / ``randint`` in partitionable mode.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import random as _random


class Domain:
    """domain.py:5-19."""
    is_integer = False

    def __init__(self, dtype):
        self.dtype = dtype

    def __hash__(self):
        return hash(self.dtype)

    def __eq__(self, other):
        return self.dtype == other.dtype

    def transform(self, x):
        raise NotImplementedError

    def sample(self, key, shape: Tuple):
        raise NotImplementedError

    def _sample(self, key, shape):
        from . import engine
        count = int(np.prod(shape))
        dims = engine.dims_array({"_": self})
        out = engine.domain_sample(_random.as_key(key), dims, count)  # the key itself, no split (domain.py:74,130)
        return out.cpu().numpy().reshape(shape)


class Real(Domain):
    """Continuous domain with clipping (domain.py:22-75)."""

    def __init__(self, lower, upper):
        assert isinstance(lower, float) or isinstance(lower, int), "Lower bound must be a float"
        # the reference's second assert tests `lower` again (domain.py:38); same acceptance set kept
        assert isinstance(upper, float) or isinstance(lower, int), "Upper bound must be a float"
        assert lower < upper, "Lower bound must be less than upper bound"
        self.lower = float(lower)
        self.upper = float(upper)
        super().__init__(dtype=np.float32)

    def __hash__(self):
        return hash((self.lower, self.upper))

    def __eq__(self, other):
        return self.lower == other.lower and self.upper == other.upper

    def transform(self, x):
        """Clip to [lower, upper] (domain.py:61)."""
        x = np.asarray(x)
        if not np.issubdtype(x.dtype, np.floating):
            x = x.astype(np.float32)
        return np.clip(x, x.dtype.type(self.lower), x.dtype.type(self.upper))

    def sample(self, key, shape: Tuple):
        """uniform(lower, upper) then clip (domain.py:74-75)."""
        return self._sample(key, shape)


class Integer(Domain):
    """Integer domain with round-half-even + clip (domain.py:78-131)."""
    is_integer = True

    def __init__(self, lower, upper):
        assert isinstance(lower, int), "Lower bound must be an integer"
        assert isinstance(upper, int), "Upper bound must be an integer"
        assert lower < upper, "Lower bound must be less than upper bound"
        self.lower = int(lower)
        self.upper = int(upper)
        super().__init__(dtype=np.int32)

    def __hash__(self):
        return hash((self.lower, self.upper))

    def __eq__(self, other):
        return self.lower == other.lower and self.upper == other.upper

    def transform(self, x):
        """clip(round(x)) as float32 (domain.py:117)."""
        return np.clip(np.round(np.asarray(x)), self.lower, self.upper).astype(np.float32)

    def sample(self, key, shape: Tuple):
        """randint(lower, upper + 1) then transform (domain.py:130-131)."""
        return self._sample(key, shape)


class ParamSpace:
    """Named collection of domains (domain.py:134-191)."""

    def __init__(self, space: dict):
        self.space = space

    def sample_params(self, key, shape: Tuple) -> dict:
        """split(key, d); j-th key -> j-th domain in dict order (domain.py:159-161).  Drawn on the GPU."""
        from . import engine
        count = int(np.prod(shape))
        dims = engine.dims_array(self.space)
        arr = engine.threefry_fill(_random.as_key(key), dims, len(self.space), 0, count).cpu().numpy()
        return {name: arr[:, j].reshape(shape).copy() for j, name in enumerate(self.space)}

    def to_array(self, tree: dict) -> np.ndarray:
        """Stack transformed columns in domain order (domain.py:175)."""
        return np.stack([np.asarray(self.space[k].transform(tree[k]), dtype=np.float32) for k in self.space], axis=1)

    def to_dict(self, xs) -> dict:
        """Column i -> transform of domain i (domain.py:191)."""
        xs = np.asarray(xs)
        return {k: self.space[k].transform(xs[:, i]) for i, k in enumerate(self.space)}
