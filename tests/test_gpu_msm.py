"""GPU: SURVEY 8(f2)/(f3) -- aggregation exponents on the device and secure aggregation as one
multi-scalar multiplication (bls_py/util.py:36-50, bls.py:29-56, 132-144, 217-221)."""
import hashlib

import numpy as np
import pytest

import bls_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu
N = O.N


def ser1(p):
    return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")


def ser2(p):
    return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))


def test_hash_pks_device_matches_host():
    from bls_b200 import engine
    from bls_b200.util import hash_pks, hash_pks_bytes, DEVICE_HASH_PKS_MIN
    pk_hash = hashlib.sha256(b"keys").digest()
    n = 5000
    out = engine.hash_pks(pk_hash, n).tobytes()
    for i in (0, 1, 2, 77, 4095, 4999):
        want = int.from_bytes(hashlib.sha256(i.to_bytes(4, "big") + pk_hash).digest(), "big") % N
        assert int.from_bytes(out[32 * i:32 * (i + 1)], "big") == want
    assert hashlib.sha256(out).digest() == hashlib.sha256(b"".join(
        (int.from_bytes(hashlib.sha256(i.to_bytes(4, "big") + pk_hash).digest(), "big") % N).to_bytes(32, "big")
        for i in range(n))).digest()
    assert engine.hash_pks(pk_hash, 3, first=4997).tobytes() == out[32 * 4997:]
    assert engine.hash_pks(pk_hash, 0).size == 0
    # the scheme-layer helper switches to the device for many keys and must agree with hash_pks
    pks = [bytes([i & 255, i >> 8]) + bytes(46) for i in range(DEVICE_HASH_PKS_MIN + 5)]
    ts = hash_pks(len(pks), pks)
    raw = hash_pks_bytes(len(pks), pks)
    assert [int.from_bytes(raw[32 * i:32 * (i + 1)], "big") for i in range(len(pks))] == ts
    assert hash_pks_bytes(7, pks[:7]) == b"".join(t.to_bytes(32, "big") for t in hash_pks(7, pks[:7]))


@pytest.mark.parametrize("g2", [False, True])
def test_msm_small_vs_oracle(g2):
    """below the bucket-method threshold: per-point ladders + reduction, against the oracle"""
    from bls_b200 import engine
    G = O.G2 if g2 else O.G1
    ser = ser2 if g2 else ser1
    w = 192 if g2 else 96
    rng = np.random.default_rng(5)
    ks = [int.from_bytes(rng.bytes(32), "big") % N for _ in range(6)]
    es = [int.from_bytes(rng.bytes(32), "big") % N for _ in range(6)]
    es[2] = 0
    pts = [O.aff_mul(k, G) for k in ks]
    raw = b"".join(ser(p) for p in pts)
    sc = b"".join(e.to_bytes(32, "big") for e in es)
    want = ser(O.aff_mul(sum(k * e for k, e in zip(ks, es)) % N, G))
    assert engine.msm(raw, sc, g2).tobytes() == want
    assert engine.msm(b"", b"", g2).tobytes() == bytes(w)
    assert engine.msm(raw[:w], (1).to_bytes(32, "big"), g2).tobytes() == raw[:w]


@pytest.mark.parametrize("g2", [False, True])
def test_msm_buckets_identity(g2):
    """bucket method (n >= 65,536): sum_i e_i (k_i G) == (sum_i e_i k_i) G with full-width scalars,
    duplicated points (P + P inside a bucket), inverse pairs, infinities and zero scalars"""
    from bls_b200 import engine, synth
    G = O.G2 if g2 else O.G1
    ser = ser2 if g2 else ser1
    w = 192 if g2 else 96
    n = 70000
    ksc = synth.scalars(901, n)
    base = np.frombuffer(ser(G), dtype=np.uint8)
    pts = engine.scalar_mul(np.tile(base, n), ksc, g2).reshape(n, w).copy()
    ks = [int.from_bytes(bytes(r), "big") for r in ksc]
    rng = np.random.default_rng(902)
    esc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    es = [int.from_bytes(bytes(r), "big") % N for r in esc]
    # structure: the first 3,000 points are one and the same point with one and the same scalar
    # digit pattern in the low window (forces P + P in a bucket); inverse pair; infinities; zeros
    pts[1:3000] = pts[0]
    for i in range(1, 3000):
        ks[i] = ks[0]
        es[i] = (es[i] & ~0x7ff) | 5
    neg = np.frombuffer(ser(O.aff_neg(O.aff_mul(ks[0], G))), dtype=np.uint8)
    pts[3000] = neg
    ks[3000] = N - ks[0]
    es[3000] = (es[3000] & ~0x7ff) | 5
    pts[3001:3011] = 0
    for i in range(3001, 3011):
        ks[i] = 0
    for i in range(3011, 3021):
        es[i] = 0
    es[3021] = N - 1
    es[3022] = 1
    sc = b"".join(e.to_bytes(32, "big") for e in es)
    want = ser(O.aff_mul(sum(k * e for k, e in zip(ks, es)) % N, G))
    got = engine.msm(pts, sc, g2).tobytes()
    assert got == want
    # skewed digits: every scalar the same, so 24 buckets hold all the points (they are cut into
    # bounded segments folded by different threads)
    e0 = es[7] | 1
    same = e0.to_bytes(32, "big") * n
    assert engine.msm(pts, same, g2).tobytes() == ser(O.aff_mul(sum(ks) * e0 % N, G))
    assert engine.msm(pts, bytes(32 * n), g2).tobytes() == bytes(w)          # all scalars zero
    # the ladder path on a slice agrees with the same identity (cross-check of both paths)
    m = 300
    assert engine.msm(pts[:m], sc[:32 * m], g2).tobytes() == \
        ser(O.aff_mul(sum(k * e for k, e in zip(ks[:m], es[:m])) % N, G))


def test_secure_aggregation_reference_vectors():
    """the scheme-layer secure paths now run through the multi-scalar multiplication"""
    from bls_b200 import BLS, PublicKey
    g = load_golden("agg_kat.json")["secure_pk_agg"]
    pks = [PublicKey.from_bytes(bytes.fromhex(p)) for p in g["pks"]]
    assert BLS.aggregate_pub_keys(pks, True).serialize().hex() == g["out"]


def test_many_keys_through_the_scheme_layer_in_few_launches():
    """row f2: 3,000 public keys decoded in one call, then aggregated securely (sort by serialised
    bytes, exponents, multi-scalar multiplication) -- a handful of kernel launches, not one per key,
    and the same point as the identity sum_i T_i k_i G"""
    from bls_b200 import BLS, PublicKey, engine, synth
    from bls_b200._lib import lib
    from bls_b200.util import hash_pks
    n = 3000
    ksc = synth.scalars(4242, n)
    base = np.frombuffer(ser1(O.G1), dtype=np.uint8)
    aff = engine.scalar_mul(np.tile(base, n), ksc, False)
    comp = engine.compress(aff, False).tobytes()
    bufs = [comp[48 * i:48 * (i + 1)] for i in range(n)]
    l0 = lib.b200bls_launch_count()
    pks = PublicKey.from_bytes_batch(bufs)
    assert [pk.serialize() for pk in pks[:50]] == bufs[:50]
    agg = BLS.aggregate_pub_keys(list(pks), True)
    launches = lib.b200bls_launch_count() - l0
    assert launches < 40, launches
    order = sorted(range(n), key=lambda i: bufs[i])
    ts = hash_pks(n, [bufs[i] for i in order])
    ks = [int.from_bytes(bytes(ksc[i]), "big") for i in order]
    want = O.aff_mul(sum(t * k for t, k in zip(ts, ks)) % N, O.G1)
    assert agg.value.raw == ser1(want)
    # a buffer that does not decode makes the batch raise, like the reference's from_bytes
    bad = None
    for i in range(1, 60):
        cand = bytearray(bufs[0])
        cand[17] ^= i
        try:
            O.g1_deserialize(bytes(cand))
        except ValueError:
            bad = bytes(cand)
            break
    with pytest.raises(ValueError):
        PublicKey.from_bytes_batch(bufs[:3] + [bad])
    # points created on the device get their serialisations in one batched call as well
    from bls_b200 import ec
    pts = [ec.Point(aff[96 * i:96 * (i + 1)].tobytes(), False) for i in range(200)]
    l0 = lib.b200bls_launch_count()
    ec.serialize_many(pts)
    assert lib.b200bls_launch_count() - l0 <= 3
    assert [p.serialize() for p in pts] == bufs[:200]
