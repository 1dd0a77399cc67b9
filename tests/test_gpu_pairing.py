"""GPU parity: pairings through the C ABI vs golden vectors from the live reference and vs
the oracle on seeded random inputs."""
import numpy as np
import pytest

import bls_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _pq_bytes(c):
    return (bytes.fromhex(c["p"]["x"]) + bytes.fromhex(c["p"]["y"]),
            bytes.fromhex(c["q"]["x"]) + bytes.fromhex(c["q"]["y"]))


def test_pairing_golden():
    from bls_b200 import engine
    g = load_golden("pairing_kat.json")
    cases = g["pairs"] + g["degenerate"]
    P = b"".join(_pq_bytes(c)[0] for c in cases)
    Qb = b"".join(_pq_bytes(c)[1] for c in cases)
    out = engine.pairing_batch(P, Qb).tobytes()
    for i, c in enumerate(cases):
        assert out[576 * i:576 * (i + 1)].hex() == c["out"], i
    # SURVEY.md 8c digest of e(G1, G2)
    import hashlib
    assert hashlib.sha256(out[:576]).hexdigest() == \
        "70f0561453673ff155a40ba3618727f8a411c492748d845280dd71dce099905a"


def test_miller_then_final_exp_is_canonical():
    from bls_b200 import engine
    g = load_golden("pairing_kat.json")
    cases = g["pairs"][:4]
    P = b"".join(_pq_bytes(c)[0] for c in cases)
    Qb = b"".join(_pq_bytes(c)[1] for c in cases)
    ml = engine.miller_loop_batch(P, Qb)
    out = engine.final_exp_batch(ml).tobytes()
    for i, c in enumerate(cases):
        assert out[576 * i:576 * (i + 1)].hex() == c["out"], i
    # final exponentiation of the REFERENCE's Miller value gives the same bytes
    fe = engine.final_exp_batch(bytes.fromhex(g["miller_loop_g1_g2"])).tobytes()
    assert fe.hex() == g["final_exp_of_miller"] == g["pairs"][0]["out"]


def test_pairing_random_vs_oracle():
    """a ragged batch (> 1 CTA) of small multiples of the generators against the oracle"""
    from bls_b200 import engine
    n = 150
    base = []
    for k in range(6):
        p = O.aff_mul(3 + 5 * k, O.G1)
        q = O.aff_mul(7 + 11 * k, O.G2)
        base.append((p, q, O.f12_serialize(O.ate_pairing(p, q))))
    def pb(p):
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
    def qb(q):
        return b"".join(c.to_bytes(48, "big") for c in (q[0][0], q[0][1], q[1][0], q[1][1]))
    P = b"".join(pb(base[i % 6][0]) for i in range(n))
    Qb = b"".join(qb(base[i % 6][1]) for i in range(n))
    out = engine.pairing_batch(P, Qb).tobytes()
    for i in range(n):
        assert out[576 * i:576 * (i + 1)] == base[i % 6][2], i


def test_bilinearity_property():
    """e(aP, bQ) == e(P, Q)^(ab) checked as e(2P, 3Q) == e(6P, Q) == e(P, 6Q) on the GPU"""
    from bls_b200 import engine
    def pb(p):
        return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
    def qb(q):
        return b"".join(c.to_bytes(48, "big") for c in (q[0][0], q[0][1], q[1][0], q[1][1]))
    ps = [O.aff_mul(2, O.G1), O.aff_mul(6, O.G1), O.G1]
    qs = [O.aff_mul(3, O.G2), O.G2, O.aff_mul(6, O.G2)]
    out = engine.pairing_batch(b"".join(map(pb, ps)), b"".join(map(qb, qs))).tobytes()
    assert out[:576] == out[576:1152] == out[1152:]
    assert out[:576] != O.f12_serialize(O.F12_ONE)


def test_config2_full_size_checksum_of_checksums():
    """BASELINE config 2 at full size (65,536 pairs): the product of all per-pair outputs of
    pairing_batch, the single-final-exponentiation pairing_multi of the same pairs, and the
    oracle's e((sum a_i b_i) G1, G2) must be the same 576 bytes."""
    from bls_b200 import engine, synth
    n = 65536
    a = synth.scalars(synth.SEED_PAIRING, n)
    b = synth.scalars(synth.SEED_PAIRING + 1, n)
    g1 = np.frombuffer(O.G1[0].to_bytes(48, "big") + O.G1[1].to_bytes(48, "big"), dtype=np.uint8)
    g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (O.G2[0] + O.G2[1])), dtype=np.uint8)
    P = engine.scalar_mul(np.tile(g1, n), a, False)
    Q = engine.scalar_mul(np.tile(g2, n), b, True)
    e = engine.pairing_batch(P, Q)
    # 16-level product tree over the GT elements (batched Fq12 multiplications on the GPU)
    level = e.reshape(n, 576)
    while level.shape[0] > 1:
        half = level.shape[0] // 2
        level = engine.field_op(12, "mul", level[:half].copy(), level[half:].copy()).reshape(half, 576)
    prod = level.tobytes()
    assert engine.pairing_multi(P, Q).tobytes() == prod
    s = sum(int.from_bytes(bytes(x), "big") * int.from_bytes(bytes(y), "big") for x, y in zip(a, b)) % O.N
    assert prod == O.f12_serialize(O.ate_pairing(O.aff_mul(s, O.G1), O.G2))
    # and two individual outputs against the oracle
    for i in (0, n - 1):
        p = O.aff_mul(int.from_bytes(bytes(a[i]), "big"), O.G1)
        q = O.aff_mul(int.from_bytes(bytes(b[i]), "big"), O.G2)
        assert e[576 * i:576 * (i + 1)].tobytes() == O.f12_serialize(O.ate_pairing(p, q))
