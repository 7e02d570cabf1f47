"""The subset of ``jax.random`` bayex users touch (``key``/``PRNGKey``/``fold_in``/``split``), without JAX.

Keys are Threefry-2x32 key data ``(hi, lo)`` as a ``numpy.uint32[2]`` array, bit-compatible with
``jax.random.key_data(jax.random.key(seed))`` in JAX's partitionable mode (README.md:49-51,
tests/test_optimizer.py:43-47).  The block function below is the scalar host twin of the device code in
``csrc/common.cuh``; candidate draws themselves happen on the GPU.
"""
from __future__ import annotations

import numpy as np

_M = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _threefry2x32(k0, k1, c0, c1):
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0, x1 = (c0 + ks[0]) & _M, (c1 + ks[1]) & _M
    for b in range(5):
        for r in _ROT[b % 2]:
            x0 = (x0 + x1) & _M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & _M
            x1 ^= x0
        x0 = (x0 + ks[(b + 1) % 3]) & _M
        x1 = (x1 + ks[(b + 2) % 3] + b + 1) & _M
    return x0, x1


def key(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & _M], dtype=np.uint32)


PRNGKey = key


def as_key(k) -> np.ndarray:
    """int seed, uint32[2] array-like, or an object exposing JAX key data -> uint32[2]."""
    if isinstance(k, (int, np.integer)):
        return key(int(k))
    if hasattr(k, "dtype") and "key" in str(getattr(k, "dtype", "")):  # a live jax typed key, if JAX ever exists
        import jax  # pragma: no cover
        k = jax.random.key_data(k)  # pragma: no cover
    a = np.asarray(k)
    if a.shape != (2,):
        raise ValueError(f"a PRNG key is two uint32 words, got shape {a.shape}")
    return a.astype(np.uint32)


def fold_in(k, data: int) -> np.ndarray:
    k = as_key(k)
    return np.array(_threefry2x32(int(k[0]), int(k[1]), 0, int(data) & _M), dtype=np.uint32)


def split(k, num: int = 2) -> np.ndarray:
    k = as_key(k)
    return np.array([_threefry2x32(int(k[0]), int(k[1]), 0, i) for i in range(num)], dtype=np.uint32)
