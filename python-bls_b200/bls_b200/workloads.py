"""BASELINE.json's configs as reusable workloads: seeded synthetic inputs built ON THE DEVICE,
the launches that process them, and the construction ground truth each result must match.  Used by bench.py
(timing + parity booleans), the full-size `-m gpu` tests and tools/; nothing here computes field arithmetic on
the host -- the only host-side numbers are Python-int scalar sums that predict a result (sum k_i G = (sum k_i) G).

  config 2  independent ate pairings            P_i = a_i G1, Q_i = b_i G2
  config 3  aggregate_sigs_simple / aggregate_pub_keys(secure=False) over points k_i G2 / k_i G1
  config 4  aggregate verification of n distinct messages: sk_i, m_i, sigma = sum sk_i H(m_i)
  config 5  independent verifications, 1 % of the signatures replaced by another signer's (valid point, wrong pair)
"""
import numpy as np

from . import engine, synth
from ._lib import check, lib
from .programs.curve import G1_GEN
from .programs.hashg2 import G2_GEN

N = synth.GROUP_ORDER
G1_BYTES = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in G1_GEN), dtype=np.uint8)
G2_BYTES = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (G2_GEN[0] + G2_GEN[1])), dtype=np.uint8)


def ints(scalars):
    return [int.from_bytes(bytes(r), "big") for r in np.asarray(scalars).reshape(-1, 32)]


def dev_scalar_mul(scalars, g2, base=None):
    """[k_i * base] (base = the generator by default) for an (n, 32) uint8 array of big-endian scalars ->
    DeviceBuffer of n affine points, produced by the scalar-multiplication kernel"""
    scalars = np.ascontiguousarray(scalars).reshape(-1, 32)
    n = scalars.shape[0]
    w = 192 if g2 else 96
    gen = G2_BYTES if g2 else G1_BYTES
    d_base = base if base is not None else engine.DeviceBuffer(w * n).upload(np.tile(gen, n))
    d_sc = engine.DeviceBuffer(32 * n).upload(scalars)
    d_out = engine.DeviceBuffer(w * n)
    fn = lib.b200bls_g2_scalar_mul_batch_dev if g2 else lib.b200bls_g1_scalar_mul_batch_dev
    check(fn(d_base.ptr, d_sc.ptr, d_out.ptr, n))
    check(lib.b200bls_sync())
    if base is None:
        d_base.free()
    d_sc.free()
    return d_out


def timed(fn, reps=3):
    """best of `reps` device times (CUDA events over all library streams) in ms"""
    best = 1e30
    for _ in range(reps):
        engine.timer_start()
        fn()
        best = min(best, engine.timer_stop())
    return best


# ---- config 2 -----------------------------------------------------------------------------------------------------
def config2_inputs(n, rank=0):
    """-> (dP, dQ, a, b): device buffers of n G1 / G2 points and the seeded scalars that made them"""
    a = synth.scalars(synth.SEED_PAIRING + 2 * rank, n)
    b = synth.scalars(synth.SEED_PAIRING + 2 * rank + 1, n)
    return dev_scalar_mul(a, False), dev_scalar_mul(b, True), a, b


# ---- config 3 -----------------------------------------------------------------------------------------------------
def config3_slice(n_total, g2, rank=0, world=1, seed=synth.SEED_AGGREGATE):
    """this rank's contiguous slice of the n_total points k_i G -> (DeviceBuffer, count, sum of ALL n_total scalars
    mod n: the scalar of the expected total)"""
    from .distributed import shard_range
    sc = synth.scalars(seed, n_total)
    total = sum(ints(sc)) % N
    lo, hi = shard_range(n_total, rank, world)
    if hi == lo:
        return engine.DeviceBuffer(1), 0, total
    return dev_scalar_mul(sc[lo:hi], g2), hi - lo, total


def expected_multiple(k, g2):
    """k * generator as affine bytes, computed by the (separately parity-tested) scalar-multiplication kernel"""
    gen = G2_BYTES if g2 else G1_BYTES
    return engine.scalar_mul(gen, (k % N).to_bytes(32, "big"), g2).tobytes()


# ---- config 4 -----------------------------------------------------------------------------------------------------
def config4_inputs(n, seed=synth.SEED_AGG_VERIFY):
    """n signers on n distinct messages -> (aggregate signature 192 B, public keys n x 96 B, message hashes
    (n, 32), secret keys (n, 32)) as host arrays; sigma = sum_i sk_i H(m_i) (bls.py:13-26 on distinct messages)"""
    sks = synth.scalars(seed, n)
    hs = synth.message_hashes(seed, n)
    sigs = engine.scalar_mul(engine.hash_to_g2(hs), sks, True)
    agg = engine.point_sum(sigs, True)
    pks = engine.scalar_mul(np.tile(G1_BYTES, n), sks, False)
    return agg, pks, hs, sks


# ---- config 5 -----------------------------------------------------------------------------------------------------
def config5_inputs(n, rank=0, seed=synth.SEED_BATCH_VERIFY):
    """n independent (pk, message hash, signature) triples on the device, 1 % of the signatures replaced by the
    signature of a different triple -> (d_pk, d_hs, d_sig, want): want[i] = 0 exactly for the corrupted ones.
    Also returns the host copies needed for oracle spot checks: (pk_host, hs, sig_host)."""
    sks = synth.scalars(seed + 16 * rank, n)
    hs = synth.message_hashes(seed + 16 * rank, n)
    d_hs = engine.DeviceBuffer(32 * n).upload(hs)
    d_H = engine.DeviceBuffer(192 * n)
    check(lib.b200bls_hash_to_g2_batch_dev(d_hs.ptr, d_H.ptr, n))
    d_sig = dev_scalar_mul(sks, True, base=d_H)
    d_H.free()
    d_pk = dev_scalar_mul(sks, False)
    sig_host = d_sig.download().reshape(n, 192)
    bad = synth.corrupted_indices(seed + 16 * rank, n)
    bad_set = set(int(i) for i in bad)
    for i in bad:
        j = (int(i) + 1) % n
        while j in bad_set:
            j = (j + 1) % n
        sig_host[i] = sig_host[j]
    d_sig.upload(sig_host)
    want = np.ones(n, dtype=np.uint8)
    want[bad] = 0
    return d_pk, d_hs, d_sig, want, (d_pk.download().reshape(n, 96), hs, sig_host)
