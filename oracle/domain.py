"""Restatement of ``/root/reference/bayex/domain.py`` on NumPy (TEST INFRASTRUCTURE ONLY)."""

from __future__ import annotations

import numpy as np

from . import prng


class Domain:
    """domain.py:5-19."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __hash__(self):
        return hash(self.dtype)

    def __eq__(self, other):
        return self.dtype == other.dtype


class Real(Domain):
    """domain.py:22-75."""

    def __init__(self, lower, upper):
        assert isinstance(lower, float) or isinstance(lower, int), "Lower bound must be a float"
        # domain.py:38 tests `lower` again in the second clause (typo kept: quirk Q14)
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
        """domain.py:61: clip."""
        x = np.asarray(x)
        if not np.issubdtype(x.dtype, np.floating):
            x = x.astype(np.float32)
        return np.clip(x, x.dtype.type(self.lower), x.dtype.type(self.upper))

    def sample(self, key, shape, offset=0):
        """domain.py:74-75: uniform(lo, hi) then clip."""
        (count,) = shape
        return self.transform(prng.uniform(key, count, self.lower, self.upper, offset=offset))


class Integer(Domain):
    """domain.py:78-131."""

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
        """domain.py:117: clip(round_half_even(x), lo, hi) -> float32."""
        x = np.asarray(x)
        return np.clip(np.round(x), self.lower, self.upper).astype(np.float32)

    def sample(self, key, shape, offset=0):
        """domain.py:130-131: randint(lo, hi+1) then transform."""
        (count,) = shape
        return self.transform(prng.randint(key, count, self.lower, self.upper + 1, offset=offset))


class ParamSpace:
    """domain.py:134-191."""

    def __init__(self, space: dict):
        self.space = space

    def sample_params(self, key, shape, offset=0) -> dict:
        """domain.py:159-161: j-th split key feeds the j-th domain (dict order)."""
        keys = prng.split(key, len(self.space))
        return {name: self.space[name].sample(k, shape, offset=offset) for name, k in zip(self.space, keys)}

    def to_array(self, tree: dict) -> np.ndarray:
        """domain.py:175."""
        return np.stack([np.asarray(self.space[k].transform(tree[k]), dtype=np.float32) for k in self.space], axis=1)

    def to_dict(self, xs) -> dict:
        """domain.py:191."""
        xs = np.asarray(xs)
        return {k: self.space[k].transform(xs[:, i]) for i, k in enumerate(self.space)}
