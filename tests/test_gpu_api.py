"""GPU: the bls_py-compatible Python API reproduces the reference's own test vectors
(bls_py/tests.py:113-198 test_vectors / test_vectors2, without divide_by)."""
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_reference_test_vectors():
    from bls_b200 import BLS, AggregationInfo, PrivateKey, PublicKey, Signature
    g = load_golden("sig_kat.json")
    v = g["test_vectors"]
    sk1 = PrivateKey.from_seed(bytes([1, 2, 3, 4, 5]))
    sk2 = PrivateKey.from_seed(bytes([1, 2, 3, 4, 5, 6]))
    pk1, pk2 = sk1.get_public_key(), sk2.get_public_key()
    sig1, sig2 = sk1.sign(bytes([7, 8, 9])), sk2.sign(bytes([7, 8, 9]))
    assert sk1.serialize().hex() == g["keys"][0]["sk"]
    assert pk1.get_fingerprint() == 0x26d53247 and pk2.get_fingerprint() == 0x289bb56e
    signed = {(c["key"], c["msg"]): c["sig"] for c in g["sign"]}
    assert sig1.serialize().hex() == signed[(0, "070809")]
    assert sig2.serialize().hex() == signed[(1, "070809")]
    assert sig1.size() == 96 and pk1.size() == 48 and sk1.size() == 32

    agg_sig = BLS.aggregate_sigs([sig1, sig2])
    agg_pk = BLS.aggregate_pub_keys([pk1, pk2], True)
    agg_sk = BLS.aggregate_priv_keys([sk1, sk2], [pk1, pk2], True)
    assert agg_sig.serialize().hex() == v["secure_agg_sig"]
    assert agg_pk.serialize().hex() == v["secure_agg_pk"]
    assert BLS.aggregate_pub_keys([pk1, pk2], False).serialize().hex() == v["simple_agg_pk"]
    assert agg_sk.serialize().hex() == v["secure_agg_sk"]
    assert agg_sk.sign(bytes([7, 8, 9])).serialize() == agg_sig.serialize()

    assert BLS.verify(sig1) is True
    assert BLS.verify(agg_sig) is True
    agg_sig.set_aggregation_info(AggregationInfo.from_msg(agg_pk, bytes([7, 8, 9])))
    assert BLS.verify(agg_sig) is True
    sig1.set_aggregation_info(sig2.aggregation_info)
    assert BLS.verify(sig1) is False

    sig3, sig4, sig5 = sk1.sign(bytes([1, 2, 3])), sk1.sign(bytes([1, 2, 3, 4])), sk2.sign(bytes([1, 2]))
    agg_sig2 = BLS.aggregate_sigs([sig3, sig4, sig5])
    assert BLS.verify(agg_sig2) is True
    assert agg_sig2.serialize().hex() == v["distinct_agg_sig"]

    # serialisation round trips (tests.py:223-260)
    pk_rt = PublicKey.from_bytes(pk1.serialize())
    sig_rt = Signature.from_bytes(sig3.serialize())
    assert pk_rt == pk1 and sig_rt == sig3
    sig_rt.set_aggregation_info(AggregationInfo.from_msg(pk_rt, bytes([1, 2, 3])))
    assert BLS.verify(sig_rt) is True
    with pytest.raises(ValueError):
        bad = bytearray(sig3.serialize())
        for flip in range(1, 40):
            bad[5] ^= flip
            Signature.from_bytes(bytes(bad))


def test_reference_test_vectors2_nested_aggregation():
    from bls_b200 import BLS, PrivateKey
    g = load_golden("sig_kat.json")["test_vectors2"]
    m1, m2, m3, m4 = bytes([1, 2, 3, 40]), bytes([5, 6, 70, 201]), bytes([9, 10, 11, 12, 13]), bytes([15, 63, 244, 92, 0, 1])
    sk1 = PrivateKey.from_seed(bytes([1, 2, 3, 4, 5]))
    sk2 = PrivateKey.from_seed(bytes([1, 2, 3, 4, 5, 6]))
    sig1, sig2, sig3 = sk1.sign(m1), sk2.sign(m2), sk2.sign(m1)
    sig4, sig5, sig6 = sk1.sign(m3), sk1.sign(m1), sk1.sign(m4)
    sig_l = BLS.aggregate_sigs([sig1, sig2])
    sig_r = BLS.aggregate_sigs([sig3, sig4, sig5])
    assert sig_l.serialize().hex() == g["sig_L"] and sig_r.serialize().hex() == g["sig_R"]
    assert BLS.verify(sig_l) and BLS.verify(sig_r)
    sig_final = BLS.aggregate_sigs([sig_l, sig_r, sig6])
    assert sig_final.serialize().hex() == g["sig_final"]
    assert BLS.verify(sig_final) is True
    got = sorted([k[0].hex(), k[1].serialize().hex(), hex(e)] for k, e in sig_final.aggregation_info.tree.items())
    assert got == g["final_tree"]


def test_ate_pairing_multi_api():
    from bls_b200 import ate_pairing_multi, ec
    from bls_b200.fields import Fq12
    from bls_b200.pairing import ate_pairing
    g1, g2 = ec.generator_Fq(), ec.generator_Fq2()
    g = load_golden("pairing_kat.json")
    e = ate_pairing(g1, g2)
    assert e.serialize().hex() == g["pairs"][0]["out"]
    # e(3 G1, 5 G2) * e(-15 G1, G2) == 1
    res = ate_pairing_multi([g1 * 3, (g1 * 15).negate()], [g2 * 5, g2])
    assert res == Fq12.one()
    assert e ** 15 == ate_pairing(g1 * 3, g2 * 5)


def test_batch_verify_api():
    from bls_b200 import BLS, PrivateKey
    sks = [PrivateKey.from_seed(bytes([i])) for i in range(4)]
    msgs = [bytes([i, i + 1]) for i in range(4)]
    sigs = [sk.sign(m) for sk, m in zip(sks, msgs)]
    pks = [sk.get_public_key() for sk in sks]
    from bls_b200.util import hash256
    hs = [hash256(m) for m in msgs]
    assert BLS.verify_batch(pks, hs, sigs) == [True] * 4
    sigs[2] = sigs[1]
    assert BLS.verify_batch(pks, hs, sigs) == [True, True, False, True]


def test_ec_module_mirrors_untwist_twist_psi_sw_encode_and_point_classes():
    """bls_b200.ec offers the reference's names (ec.py:18-188 AffinePoint / JacobianPoint, 402-444 untwist / twist /
    psi, 449-507 sw_encode) with the reference's values"""
    import bls_oracle as O
    from bls_b200 import ec
    from bls_b200.fields import Fq2, Fq12, Q
    from conftest import load_golden
    g2 = ec.generator_Fq2()
    assert isinstance(g2, ec.AffinePoint) and isinstance(g2.to_jacobian(), ec.JacobianPoint)
    p = g2.to_jacobian() * 5
    assert isinstance(p, ec.JacobianPoint) and isinstance(p.to_affine(), ec.AffinePoint)
    assert p == p.to_affine() and hash(p) == hash(p.to_affine()) and p.to_jacobian().z == (1, 0)
    op = O.aff_mul(5, O.G2)
    u = ec.untwist(p)
    ux, uy, _ = O.untwist((op[0], op[1], False))
    assert u.x.ZT == ux and u.y.ZT == uy
    back = ec.twist(u)
    assert back.x == Fq12.from_fq(Q, Fq2(Q, *op[0])) or back.x.ZT == tuple(op[0]) + (0,) * 10
    assert back.y.ZT == tuple(op[1]) + (0,) * 10
    assert ec.untwist(ec.twist(u)) == u                       # fq12_untwist . fq12_twist
    for c in load_golden("hash_kat.json")["psi"]:
        pt = ec.AffinePoint(bytes.fromhex(c["p"]["x"] + c["p"]["y"]), True)
        assert ec.psi(pt).raw.hex() == c["out"]["x"] + c["out"]["y"]
    for c in load_golden("hash_kat.json")["sw_encode"][:6]:
        t = bytes.fromhex(c["t"])
        got = ec.sw_encode(Fq2(t))
        assert got.infinity == c["out"]["inf"]
        if not got.infinity:
            assert got.raw.hex() == c["out"]["x"] + c["out"]["y"]


def test_point_scalars_outside_256_bits_follow_the_reference():
    """AffinePoint.__mul__ / JacobianPoint.__mul__ (ec.py:61-66, 151-156) accept any Python int: scalars of 2^256 and
    more, multiples of the group order and negative ones reach the 32-byte device ladder reduced mod n"""
    from bls_b200 import ec
    import bls_oracle as O
    ks = [O.N + 5, (1 << 256) + 5, (1 << 300) + 12345, 3 * O.N, -7, O.N - 1]
    for g2, gen, G in ((False, ec.generator_Fq(), O.G1), (True, ec.generator_Fq2(), O.G2)):
        many = ec.scalar_mul_many([gen] * len(ks), ks, g2)
        for k, m in zip(ks, many):
            want = O.aff_mul(k % O.N, G)
            if g2:
                wb = bytes(192) if want[2] else b"".join(c.to_bytes(48, "big") for c in (want[0][0], want[0][1], want[1][0], want[1][1]))
            else:
                wb = bytes(96) if want[2] else want[0].to_bytes(48, "big") + want[1].to_bytes(48, "big")
            assert m.raw == wb, (g2, k)
            assert (gen * k).raw == wb, (g2, k)
