"""CPU tests of the host-side scheme logic (no GPU needed): hashing helpers, aggregation
exponents and AggregationInfo merging against vectors generated from the live reference."""
import bls_oracle as O
import pytest
from conftest import load_golden

from bls_b200 import ec, synth
from bls_b200.aggregation_info import AggregationInfo
from bls_b200.keys import PublicKey
from bls_b200.util import hash256, hash512, hash_pks, hmac256


def stub_pk(ser_hex):
    """a PublicKey whose serialisation is preset, so no GPU call is needed"""
    p = ec.Point(bytes(96), False)
    p._ser = bytes.fromhex(ser_hex)
    return PublicKey(p)


def test_hash_helpers_match_oracle():
    for m in (b"", b"abc", bytes(range(200))):
        assert hash256(m) == O.hash256(m)
        assert hash512(m) == O.hash512(m)
        assert hmac256(m, b"BLS private key seed") == O.hmac256(m, b"BLS private key seed")
    assert hmac256(b"x", bytes(100)) == O.hmac256(b"x", bytes(100))      # long key is hashed first


def test_private_key_from_seed_vector():
    from bls_b200.keys import PrivateKey
    g = load_golden("sig_kat.json")
    for k in g["keys"]:
        assert PrivateKey.from_seed(bytes.fromhex(k["seed"])).serialize().hex() == k["sk"]


def test_hash_pks_matches_oracle():
    g = load_golden("sig_kat.json")
    pks = [stub_pk(k["pk"]) for k in g["keys"]]
    assert hash_pks(3, pks) == O.hash_pks(3, [bytes.fromhex(k["pk"]) for k in g["keys"]])


def test_merge_infos_reproduces_reference_tree():
    """tests.py:150-172: sig_final = agg([agg([s1, s2]), agg([s3, s4, s5]), s6])"""
    g = load_golden("sig_kat.json")
    pk1, pk2 = [stub_pk(k["pk"]) for k in g["keys"]]
    m1, m2, m3, m4 = [hash256(bytes(m)) for m in ([1, 2, 3, 40], [5, 6, 70, 201], [9, 10, 11, 12, 13],
                                                  [15, 63, 244, 92, 0, 1])]
    i1, i2, i3 = (AggregationInfo.from_msg_hash(pk1, m1), AggregationInfo.from_msg_hash(pk2, m2),
                  AggregationInfo.from_msg_hash(pk2, m1))
    i4, i5, i6 = (AggregationInfo.from_msg_hash(pk1, m3), AggregationInfo.from_msg_hash(pk1, m1),
                  AggregationInfo.from_msg_hash(pk1, m4))
    left = AggregationInfo.merge_infos([i1, i2])
    right = AggregationInfo.merge_infos([i3, i4, i5])
    final = AggregationInfo.merge_infos([left, right, i6])
    got = sorted([k[0].hex(), k[1].serialize().hex(), hex(e)] for k, e in final.tree.items())
    assert got == g["test_vectors2"]["final_tree"]
    assert final.message_hashes == sorted(final.message_hashes) or len(set(final.message_hashes)) < len(final.message_hashes)
    assert [(mh, pk) for mh, pk in zip(final.message_hashes, final.public_keys)] == sorted(final.tree.keys())


def test_aggregation_info_ordering():
    pk1, pk2 = stub_pk("00" * 47 + "01"), stub_pk("00" * 47 + "02")
    a = AggregationInfo.from_msg_hash(pk1, b"\x01" * 32)
    b = AggregationInfo.from_msg_hash(pk2, b"\x01" * 32)
    c = AggregationInfo.merge_infos([a, AggregationInfo.from_msg_hash(pk2, b"\x02" * 32)])
    assert a < b and not b < a and a == a.copy()
    assert a < c                                  # a proper prefix sorts first
    assert sorted([b, c, a])[0] is a


def test_synthetic_inputs_are_reproducible():
    s = synth.scalars(synth.SEED_PAIRING, 5)
    assert s.shape == (5, 32) and (s[:, 0] < 0x40).all()
    assert (synth.scalars(synth.SEED_PAIRING, 5) == s).all()
    bad = synth.corrupted_indices(7, 1000)
    assert len(bad) == 10 and len(set(bad.tolist())) == 10


def test_lagrange_and_interpolation_match_reference():
    """threshold.py:57-102 on the host (ints mod n), against values recorded from the reference"""
    from bls_b200.threshold import Threshold
    from bls_b200.util import GROUP_ORDER
    for t in load_golden("ext_kat.json")["threshold"]:
        X = t["players"]
        lambs = Threshold.lagrange_coeffs_at_zero(X)
        assert [hex(l) for l in lambs] == t["lagrange"]
        shares = [int(s, 16) for s in t["shares"]]
        assert Threshold.interpolate_at_zero(X, [shares[x - 1] for x in X]) == int(t["master_sk"], 16)
    # P(x) = 7 + 3x + 5x^2 through three points
    P = lambda x: (7 + 3 * x + 5 * x * x) % GROUP_ORDER
    assert Threshold.interpolate_at_zero([2, 9, 11], [P(2), P(9), P(11)]) == 7
    with pytest.raises(AssertionError):
        Threshold.lagrange_coeffs_at_zero([1, 1])
    with pytest.raises(AssertionError):
        Threshold.lagrange_coeffs_at_zero([0, 1])


def test_extended_key_containers_without_gpu():
    """header layout and value semantics of the HD-key containers (keys.py:167-316); derivation itself
    needs the GPU and is covered by tests/test_gpu_ext.py"""
    from bls_b200.keys import ExtendedPrivateKey, ExtendedPublicKey, PrivateKey, HARDENED
    e = ExtendedPrivateKey(1, 3, 0xa4700b27, HARDENED + 77, bytes(range(32)), PrivateKey(0x1234))
    raw = e.serialize()
    assert len(raw) == e.size() == ExtendedPrivateKey.EXTENDED_PRIVATE_KEY_SIZE == 77
    assert raw[:4] == (1).to_bytes(4, "big") and raw[4] == 3 and raw[5:9] == bytes.fromhex("a4700b27")
    assert raw[9:13] == (HARDENED + 77).to_bytes(4, "big") and raw[13:45] == bytes(range(32))
    assert raw[45:] == (0x1234).to_bytes(32, "big")
    assert e == ExtendedPrivateKey(1, 3, 0xa4700b27, HARDENED + 77, bytes(range(32)), PrivateKey(0x1234))
    assert hash(e) == hash(int.from_bytes(raw, "big"))
    assert ExtendedPublicKey.EXTENDED_PUBLIC_KEY_SIZE == 93
    deep = ExtendedPrivateKey(1, 255, 0, 0, bytes(32), PrivateKey(1))
    with pytest.raises(Exception, match="255 levels"):
        deep.private_child(0)
