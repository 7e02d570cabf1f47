"""Threefry / jax.random restatement: Random123 KATs + JAX-documented vectors."""
import numpy as np

from oracle import prng


def test_threefry_random123_kats():
    kats = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
            ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
            ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for k, c, want in kats:
        y0, y1 = prng.threefry2x32(k[0], k[1], c[0], c[1])
        assert (int(y0), int(y1)) == want


def test_key_split_fold_in_vectors():
    assert prng.key(42).tolist() == [0, 42]
    assert prng.key((7 << 32) | 9).tolist() == [7, 9]
    # jax documentation example, partitionable threefry (default since JAX 0.5)
    assert prng.split(prng.key(0)).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]
    # legacy (pre-0.5) mode
    assert prng.split(prng.key(0), legacy=True).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert prng.fold_in(prng.key(42), 0).tolist() == [1832780943, 270669613]


def test_draws_are_index_addressable():
    """Element i depends only on (key, i): windows equal slices (what GPU sharding relies on)."""
    k = prng.key(3)
    full_u = prng.uniform(k, 1000, -2.0, 5.0)
    full_i = prng.randint(k, 1000, -10, 11)
    assert np.array_equal(prng.uniform(k, 100, -2.0, 5.0, offset=300), full_u[300:400])
    assert np.array_equal(prng.randint(k, 100, -10, 11, offset=300), full_i[300:400])
    assert full_u.dtype == np.float32 and full_u.min() >= -2.0 and full_u.max() < 5.0
    assert full_i.dtype == np.int32 and full_i.min() >= -10 and full_i.max() <= 10
    assert set(np.unique(full_i)) == set(range(-10, 11))


def test_uniform_bit_recipe():
    k = prng.key(0)
    bits = prng.random_bits(k, 8)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)
    assert np.array_equal(prng.uniform(k, 8, 0.0, 1.0), f)


def test_randint_matches_64bit_modulo():
    """The wrapping-u32 recipe equals ((hi << 32 | lo) mod span) whenever 2^32 mod span is used as multiplier."""
    k = prng.key(11)
    k1, k2 = prng.split(k, 2)
    hi = prng.random_bits(k1, 500).astype(object)
    lo = prng.random_bits(k2, 500).astype(object)
    for span in (3, 21, 1000):
        want = np.array([((int(h) % span) * ((1 << 32) % span) + int(l) % span) % span for h, l in zip(hi, lo)])
        got = prng.randint(k, 500, 5, 5 + span) - 5
        assert np.array_equal(got, want)
