"""Seeded synthetic inputs for tests, golden vectors and bench.

Everything is derived from ``numpy.random.Generator(PCG64(seed))`` so that the
development container (which has the reference) and the GPU box (which does
not) build identical inputs.  Pure host-side byte shuffling; no field math.
"""
import numpy as np

GROUP_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001

SEED_PAIRING = 0xB2000002
SEED_AGGREGATE = 0xB2000003
SEED_AGG_VERIFY = 0xB2000004
SEED_BATCH_VERIFY = 0xB2000005


def scalars(seed, count):
    """``count`` scalars in [1, n) as a (count, 32) uint8 array, big-endian.

    Draws 32 random bytes and clears the top two bits (so the value is < 2^254
    < n); zero is mapped to one.  Cheap enough for 8 M scalars.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.integers(0, 256, size=(count, 32), dtype=np.uint8)
    raw[:, 0] &= 0x3f
    zero = ~raw.any(axis=1)
    raw[zero, 31] = 1
    return raw


def scalar_ints(seed, count):
    return [int.from_bytes(bytes(row), "big") for row in scalars(seed, count)]


def message_hashes(seed, count):
    """``count`` 32-byte message digests as a (count, 32) uint8 array."""
    rng = np.random.Generator(np.random.PCG64(seed ^ 0x5a5a5a5a))
    return rng.integers(0, 256, size=(count, 32), dtype=np.uint8)


def corrupted_indices(seed, count):
    """1% of the indices, sorted (config 5)."""
    rng = np.random.Generator(np.random.PCG64(seed ^ 0xc0440000))
    k = max(1, count // 100) if count >= 2 else 0
    return np.sort(rng.choice(count, size=k, replace=False))
