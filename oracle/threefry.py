"""Oracle: `jax.random.normal(PRNGKey(seed), shape)` in FP64 (TEST INFRASTRUCTURE ONLY).

Needed only to reproduce the reference's end-to-end ETKF golden test, whose initial
ensemble noise is `jrand.normal(jrand.PRNGKey(42), shape=(8, 5))`
(tests/dacycler_etkf_test.py:10,60).  JAX is absent from this image; its published
generator is restated: Threefry-2x32 (20 rounds, Random123), the pre-0.5
*non-partitionable* counter layout, 52-bit mantissa fill, and sqrt(2)*erfinv(u).
"""
import numpy as np
from scipy.special import erfinv

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def threefry2x32(key, x0, x1):
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
    x0 = (x0.astype(np.uint32) + ks[0]).astype(np.uint32)
    x1 = (x1.astype(np.uint32) + ks[1]).astype(np.uint32)
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1).astype(np.uint32)
            x1 = _rotl(x1, r)
            x1 = x1 ^ x0
        x0 = (x0 + ks[(g + 1) % 3]).astype(np.uint32)
        x1 = (x1 + ks[(g + 2) % 3] + np.uint32(g + 1)).astype(np.uint32)
    return x0, x1


def prng_key(seed):
    return (np.uint32((seed >> 32) & 0xFFFFFFFF), np.uint32(seed & 0xFFFFFFFF))


def random_bits64(key, size):
    counts = np.arange(2 * size, dtype=np.uint32)
    o0, o1 = threefry2x32(key, counts[:size], counts[size:])
    out = np.concatenate([o0, o1])
    hi, lo = out[:size].astype(np.uint64), out[size:].astype(np.uint64)
    return (hi << np.uint64(32)) | lo


def normal(seed, shape):
    size = int(np.prod(shape))
    with np.errstate(over='ignore'):
        bits = random_bits64(prng_key(seed), size)
    one_bits = np.array(1.0, dtype=np.float64).view(np.uint64)
    floats = ((bits >> np.uint64(12)) | one_bits).view(np.float64) - 1.0
    lo = np.nextafter(np.float64(-1.0), np.float64(0.0))
    u = np.maximum(lo, floats * (1.0 - lo) + lo)
    return (np.sqrt(2.0) * erfinv(u)).reshape(shape)
