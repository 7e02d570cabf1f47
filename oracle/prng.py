"""Restatement of the JAX PRNG surface bayex uses (TEST INFRASTRUCTURE ONLY).

jax / jaxlib are unpinned third-party dependencies of the reference
(``/root/reference/pyproject.toml:20``) and are absent from this image, so the
published algorithm is restated here:

* Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as
  1, 2, 3", SC'11; Random123 known-answer vectors in ``tests/test_prng.py``).
* ``jax.random.key`` / ``PRNGKey``, ``fold_in``, ``split``, 32-bit
  ``random_bits``, ``uniform`` and ``randint`` in JAX's *partitionable* mode
  (the default since JAX 0.5): element ``i`` of a draw depends only on
  ``(key, i)``, which is what lets the candidate axis shard across GPUs.

Reference call sites: ``bayex/domain.py:74`` (uniform), ``:130`` (randint),
``:160`` (split); ``README.md:49-51`` / ``tests/test_optimizer.py:43-47`` (key,
fold_in).
"""

from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = np.uint32(0x1BD11BDA)
_U32 = np.uint32


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32-20 block function on uint32 arrays (broadcasting)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=np.uint32)
        k1 = np.asarray(k1, dtype=np.uint32)
        x0 = np.asarray(c0, dtype=np.uint32).copy()
        x1 = np.asarray(c1, dtype=np.uint32).copy()
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for block in range(5):
            for r in _ROT[block % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(block + 1) % 3]
            x1 = x1 + ks[(block + 2) % 3] + _U32(block + 1)
        return x0, x1


def key(seed: int) -> np.ndarray:
    """``jax.random.key(seed)`` / ``PRNGKey(seed)`` key data: (hi32, lo32)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


PRNGKey = key


def as_key(k) -> np.ndarray:
    if isinstance(k, (int, np.integer)):
        return key(int(k))
    a = np.asarray(k)
    assert a.shape == (2,), "a Threefry key is two uint32 words"
    return a.astype(np.uint32)


def fold_in(k, data: int) -> np.ndarray:
    """``jax.random.fold_in``: threefry(key, counter=(0, data))."""
    k = as_key(k)
    y0, y1 = threefry2x32(k[0], k[1], _U32(0), _U32(int(data) & 0xFFFFFFFF))
    return np.array([y0, y1], dtype=np.uint32)


def split(k, num: int = 2, *, legacy: bool = False) -> np.ndarray:
    """``jax.random.split`` -> (num, 2) uint32.

    partitionable (default): key_i = threefry(key, (0, i)).
    legacy (JAX < 0.5 default; oracle-only option): counts = iota(2*num), first
    half feeds word 0, second half word 1, outputs concatenated and re-paired.
    """
    k = as_key(k)
    if not legacy:
        i = np.arange(num, dtype=np.uint32)
        y0, y1 = threefry2x32(k[0], k[1], np.zeros_like(i), i)
        return np.stack([y0, y1], axis=1)
    cnt = np.arange(2 * num, dtype=np.uint32)
    y0, y1 = threefry2x32(k[0], k[1], cnt[:num], cnt[num:])
    return np.concatenate([y0, y1]).reshape(num, 2)


def random_bits(k, count: int, *, offset: int = 0) -> np.ndarray:
    """32-bit ``random_bits(key, 32, (count,))``, partitionable mode.

    ``offset`` selects the index window [offset, offset+count) of a longer draw
    (identical to slicing the full draw) - the property sharding relies on.
    """
    k = as_key(k)
    i = np.arange(offset, offset + count, dtype=np.uint64)
    hi = (i >> np.uint64(32)).astype(np.uint32)
    lo = (i & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    y0, y1 = threefry2x32(k[0], k[1], hi, lo)
    return y0 ^ y1


def uniform(k, count: int, minval: float, maxval: float, *, offset: int = 0) -> np.ndarray:
    """``jax.random.uniform(key, (count,), float32, minval, maxval)``.

    bits>>9 | 0x3F800000 -> [1,2) -> -1 -> *(max-min) + min, then max(min, .).
    The multiply and add are rounded separately (no FMA contraction).
    """
    bits = random_bits(k, count, offset=offset)
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.float32(minval)
    scale = np.float32(np.float32(maxval) - lo)
    return np.maximum(lo, (f * scale).astype(np.float32) + lo).astype(np.float32)


def randint(k, count: int, minval: int, maxval: int, *, offset: int = 0) -> np.ndarray:
    """``jax.random.randint(key, (count,), minval, maxval, int32)`` (maxval exclusive).

    Two 32-bit draws from split(key) are combined as a 64-bit value modulo span,
    all in wrapping uint32 arithmetic.
    """
    k1, k2 = split(k, 2)
    hi = random_bits(k1, count, offset=offset)
    lo = random_bits(k2, count, offset=offset)
    span = (int(maxval) - int(minval)) & 0xFFFFFFFF
    if maxval <= minval:
        span = 1
    span = _U32(span)
    with np.errstate(over="ignore"):
        mult = _U32(65536) % span
        mult = (mult * mult) % span
        off = ((hi % span) * mult + (lo % span)) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)
