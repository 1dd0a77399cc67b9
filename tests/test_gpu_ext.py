"""GPU: SURVEY 8(f4) -- Signature.divide_by, HD keys and threshold signatures reproduce the
reference (bls_py/tests.py:150-220 test_vectors2 / test_vectors3, 350-427 test_threshold);
expected values were recorded from the live reference by tools/gen_golden.py (ext_kat.json)."""
import random
from itertools import combinations

import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _tree(sig):
    return sorted([k[0].hex(), k[1].serialize().hex(), hex(e)] for k, e in sig.aggregation_info.tree.items())


def test_divide_by_reference_vectors():
    from bls_b200 import BLS, PrivateKey
    g = load_golden("ext_kat.json")["divide_by"]
    m1, m2, m3, m4 = bytes([1, 2, 3, 40]), bytes([5, 6, 70, 201]), bytes([9, 10, 11, 12, 13]), bytes([15, 63, 244, 92, 0, 1])
    sk1 = PrivateKey.from_seed(bytes([1, 2, 3, 4, 5]))
    sk2 = PrivateKey.from_seed(bytes([1, 2, 3, 4, 5, 6]))
    sig1, sig2, sig3 = sk1.sign(m1), sk2.sign(m2), sk2.sign(m1)
    sig4, sig5, sig6 = sk1.sign(m3), sk1.sign(m1), sk1.sign(m4)
    sig_l = BLS.aggregate_sigs([sig1, sig2])
    sig_r = BLS.aggregate_sigs([sig3, sig4, sig5])
    sig_final = BLS.aggregate_sigs([sig_l, sig_r, sig6])
    before = sig_final.serialize()
    quotient = sig_final.divide_by([sig2, sig5, sig6])
    assert quotient.serialize().hex() == g["quotient"]            # tests.py:177
    assert _tree(quotient) == g["quotient_tree"]
    assert BLS.verify(quotient) is True and BLS.verify(sig_final) is True
    assert sig_final.serialize() == before                         # the dividend is not modified
    assert quotient.divide_by([]) == quotient
    with pytest.raises(Exception, match="not a subset"):
        quotient.divide_by([sig6])
    assert sig_final.divide_by([sig1]).serialize().hex() == g["by_sig1"]
    with pytest.raises(Exception, match="not unique"):
        sig_final.divide_by([sig_l])
    # divide by an aggregate (tests.py:191-198)
    sig_r2 = BLS.aggregate_sigs([sk2.sign(m3), sk2.sign(m4)])
    sig_final2 = BLS.aggregate_sigs([sig_final, sig_r2])
    assert sig_final2.serialize().hex() == g["sig_final2"]
    quotient2 = sig_final2.divide_by([sig_r2])
    assert quotient2.serialize().hex() == g["quotient2"]
    assert _tree(quotient2) == g["quotient2_tree"]
    assert BLS.verify(quotient2) is True


def test_hd_keys_reference_vectors():
    from bls_b200 import ExtendedPrivateKey, ExtendedPublicKey
    g = load_golden("ext_kat.json")["hd"]
    esk = ExtendedPrivateKey.from_seed(bytes.fromhex(g["seed"]))
    assert esk.private_key.get_public_key().get_fingerprint() == 0xa4700b27          # tests.py:206
    assert esk.chain_code.hex() == "d8b12555b4cc5578951e4a7c80031e22019cc0dce168b3ed88115311b8feb1e3"
    for node in g["nodes"]:
        cur = esk
        for i in node["path"]:
            cur = cur.private_child(i)
        assert cur.serialize().hex() == node["xprv"] and cur.size() == 77
        epk = cur.get_extended_public_key()
        assert epk.serialize().hex() == node["xpub"] and epk.size() == 93
        assert cur.get_public_key().get_fingerprint() == node["fingerprint"]
        assert cur.chain_code.hex() == node["chain_code"]
        assert ExtendedPublicKey.from_bytes(epk.serialize()) == epk
        if "xpub_public_derivation" in node:
            pub = esk.get_extended_public_key()
            for i in node["path"]:
                pub = pub.public_child(i)
            assert pub.serialize().hex() == node["xpub_public_derivation"] == node["xpub"]
            if node["path"]:
                assert esk.public_child(node["path"][0]).serialize() == \
                    esk.get_extended_public_key().public_child(node["path"][0]).serialize()
    with pytest.raises(Exception, match="hardened"):
        esk.get_extended_public_key().public_child(2 ** 31)


def test_threshold_reference_vectors():
    from bls_b200 import BLS, AggregationInfo, PrivateKey, PublicKey, Signature, Threshold
    from bls_b200 import ec
    for t in load_golden("ext_kat.json")["threshold"]:
        T, N = t["T"], t["N"]
        commitments = [[ec.point_from_bytes(bytes.fromhex(c), False) for c in cs] for cs in t["commitments"]]
        fragments = [[int(f, 16) for f in row] for row in t["fragments"]]
        # the dealing itself: commitments are g1 * coefficient
        for poly, cs in zip(t["polys"], commitments):
            assert [c.serialize().hex() for c in ec.scalar_mul_many([ec.generator_Fq()] * T, [int(c, 16) for c in poly], False)] \
                == [c.serialize().hex() for c in cs]
        for src in range(1, N + 1):
            for tgt in range(1, N + 1):
                assert Threshold.verify_secret_fragment(T, fragments[tgt - 1][src - 1], tgt, commitments[src - 1])
        if N > 1:
            assert not Threshold.verify_secret_fragment(T, fragments[0][0] + 1, 1, commitments[0])
        master_pk = BLS.aggregate_pub_keys([PublicKey.from_g1(cs[0]) for cs in commitments], False)
        assert master_pk.serialize().hex() == t["master_pk"]
        shares = [BLS.aggregate_priv_keys(map(PrivateKey, row), None, False) for row in fragments]
        assert [s.serialize().hex() for s in shares] == t["shares"]
        X = t["players"]
        assert [hex(l) for l in Threshold.lagrange_coeffs_at_zero(X)] == t["lagrange"]
        assert Threshold.interpolate_at_zero(X, [shares[x - 1].value for x in X]) == int(t["master_sk"], 16)
        sig_shares = [shares[x - 1].sign_threshold("Test", x, X) for x in X]
        assert [s.serialize().hex() for s in sig_shares] == t["sig_shares"]
        sig = BLS.aggregate_sigs_simple(sig_shares)
        assert sig.serialize().hex() == t["signature"]
        unit = [shares[x - 1].sign("Test") for x in X]
        assert [s.serialize().hex() for s in unit] == t["unit_sigs"]
        assert Threshold.aggregate_unit_sigs(unit, X, T).serialize().hex() == t["signature"]
        sig.set_aggregation_info(AggregationInfo.from_msg(master_pk, "Test"))
        assert BLS.verify(sig) is True
        assert isinstance(sig, Signature)


def test_threshold_fresh_dealing_every_subset():
    """the reference's own flow (tests.py:350-419) on a fresh 2-of-4 dealing, every subset of players"""
    from bls_b200 import BLS, AggregationInfo, PrivateKey, PublicKey, Threshold
    T, N = 2, 4
    rng = random.Random(20251018)
    dealt = [PrivateKey.new_threshold(T, N, rng) for _ in range(N)]
    fragments = [[dealt[src][2][tgt] for src in range(N)] for tgt in range(N)]
    for src in range(N):
        for tgt in range(N):
            assert Threshold.verify_secret_fragment(T, fragments[tgt][src], tgt + 1, dealt[src][1])
    master_pk = BLS.aggregate_pub_keys([PublicKey.from_g1(d[1][0]) for d in dealt], False)
    shares = [BLS.aggregate_priv_keys(map(PrivateKey, row), None, False) for row in fragments]
    master_sk = BLS.aggregate_priv_keys([d[0] for d in dealt], None, False)
    assert master_sk.get_public_key() == master_pk
    actual = master_sk.sign("Test")
    for X in combinations(range(1, N + 1), T):
        X = list(X)
        assert PrivateKey(Threshold.interpolate_at_zero(X, [shares[x - 1].value for x in X])) == master_sk
        assert BLS.aggregate_sigs_simple([shares[x - 1].sign_threshold("Test", x, X) for x in X]) == actual
        assert Threshold.aggregate_unit_sigs([shares[x - 1].sign("Test") for x in X], X, T) == actual
    actual.set_aggregation_info(AggregationInfo.from_msg(master_pk, "Test"))
    assert BLS.verify(actual) is True
    sk, commitments, frags = PrivateKey.new_threshold(3, 5)       # system randomness path
    assert len(commitments) == 3 and len(frags) == 5 and 0 < sk.value
