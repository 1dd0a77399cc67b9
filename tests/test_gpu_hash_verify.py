"""GPU parity: hash_to_point_prehashed_Fq2, single-message verification (accept / reject),
multi-pairing and aggregate verification through the C ABI."""
import numpy as np
import pytest

import bls_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


def ser1(p):
    return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")


def ser2(p):
    return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))


def test_hash_to_g2_golden():
    from bls_b200 import engine
    g = load_golden("hash_kat.json")
    cases = g["hash_to_g2_prehashed"]
    hs = b"".join(bytes.fromhex(c["h"]) for c in cases)
    out = engine.hash_to_g2(hs * 13).tobytes()             # 312 items: ragged, several CTAs
    for rep in (0, 12):
        for i, c in enumerate(cases):
            k = rep * len(cases) + i
            assert out[192 * k:192 * (k + 1)].hex() == c["out"]["x"] + c["out"]["y"], (rep, i)
    comp = engine.compress(out[:192 * len(cases)], True).tobytes()
    for i, c in enumerate(cases):
        assert comp[96 * i:96 * (i + 1)].hex() == c["ser"]


def test_sign_vectors_of_the_reference():
    """tests.py:113-122: sk * H(m) serialises to the reference's signature bytes"""
    from bls_b200 import engine
    g = load_golden("sig_kat.json")
    sks = [int(k["sk"], 16) for k in g["keys"]]
    hs = b"".join(O.hash256(bytes.fromhex(c["msg"])) for c in g["sign"])
    H = engine.hash_to_g2(hs)
    sc = b"".join(sks[c["key"]].to_bytes(32, "big") for c in g["sign"])
    sig = engine.compress(engine.scalar_mul(H, sc, True), True).tobytes()
    for i, c in enumerate(g["sign"]):
        assert sig[96 * i:96 * (i + 1)].hex() == c["sig"], i
    # public keys: sk * G1 (keys.py:119-121)
    pk = engine.compress(engine.scalar_mul(ser1(O.G1) * 2, b"".join(s.to_bytes(32, "big") for s in sks), False), False).tobytes()
    for i, k in enumerate(g["keys"]):
        assert pk[48 * i:48 * (i + 1)].hex() == k["pk"]


def test_verify_truth_table():
    from bls_b200 import engine
    g = load_golden("sig_kat.json")
    tab = g["verify_table"]
    pk, _ = engine.decompress(b"".join(bytes.fromhex(t["pk"]) for t in tab), False)
    sg, ok = engine.decompress(b"".join(bytes.fromhex(t["sig"]) for t in tab), True)
    assert ok.all()
    res = engine.verify_batch(pk, b"".join(bytes.fromhex(t["h"]) for t in tab), sg)
    assert [bool(r) for r in res] == [t["ok"] for t in tab]
    assert any(t["ok"] for t in tab) and not all(t["ok"] for t in tab)


def test_verify_bitflipped_signatures():
    """corrupted signature bytes: decode failure or pairing failure, never acceptance"""
    from bls_b200 import engine
    g = load_golden("sig_kat.json")["bitflips"]
    cases = g["cases"]
    sg, ok = engine.decompress(b"".join(bytes.fromhex(c["sig"]) for c in cases), True)
    pk, _ = engine.decompress(bytes.fromhex(g["pk"]) * len(cases), False)
    res = engine.verify_batch(pk, bytes.fromhex(g["h"]) * len(cases), sg)
    for i, c in enumerate(cases):
        assert bool(ok[i]) == c["decodes"], i
        accepted = bool(ok[i]) and bool(res[i])
        assert accepted == c["ok"], i
    assert not any(c["ok"] for c in cases)


def test_pairing_multi_golden():
    from bls_b200 import engine
    g = load_golden("pairing_kat.json")
    for m in g["multi"]:
        cs = [g["pairs"][i] for i in m["idx"]]
        P = b"".join(bytes.fromhex(c["p"]["x"]) + bytes.fromhex(c["p"]["y"]) for c in cs)
        Q = b"".join(bytes.fromhex(c["q"]["x"]) + bytes.fromhex(c["q"]["y"]) for c in cs)
        assert engine.pairing_multi(P, Q).tobytes().hex() == m["out"]
        # the two-step form ranks use: Miller product, then one final exponentiation
        assert engine.final_exp_batch(engine.miller_product(P, Q)).tobytes().hex() == m["out"]


def test_pairing_multi_many_pairs_property():
    """prod_i e(a_i G1, b_i G2) == e((sum a_i b_i) G1, G2) over 700 pairs (several CTAs, ragged)"""
    from bls_b200 import engine, synth
    n = 700
    a = synth.scalars(11, n)
    b = synth.scalars(12, n)
    P = engine.scalar_mul(np.tile(np.frombuffer(ser1(O.G1), dtype=np.uint8), n), a, False)
    Q = engine.scalar_mul(np.tile(np.frombuffer(ser2(O.G2), dtype=np.uint8), n), b, True)
    got = engine.pairing_multi(P, Q).tobytes()
    s = sum(int.from_bytes(bytes(x), "big") * int.from_bytes(bytes(y), "big") for x, y in zip(a, b)) % O.N
    want = O.f12_serialize(O.ate_pairing(O.aff_mul(s, O.G1), O.G2))
    assert got == want


def test_aggregate_verify():
    """distinct-message aggregate (bls.py:194-201): accept, then reject after swapping a message"""
    from bls_b200 import engine, synth
    n = 300
    sks = synth.scalars(21, n)
    hs = synth.message_hashes(22, n)
    H = engine.hash_to_g2(hs)
    sigs = engine.scalar_mul(H, sks, True)
    agg = engine.point_sum(sigs, True)
    pks = engine.scalar_mul(np.tile(np.frombuffer(ser1(O.G1), dtype=np.uint8), n), sks, False)
    assert engine.aggregate_verify(agg, pks, hs) is True
    hs2 = hs.copy()
    hs2[5] = hs[6]
    assert engine.aggregate_verify(agg, pks, hs2) is False
    # one pair against the oracle
    assert O.aggregate_verify([O.pk_of(int.from_bytes(bytes(sks[0]), "big"))], [bytes(hs[0])],
                              O.sign_prehashed(int.from_bytes(bytes(sks[0]), "big"), bytes(hs[0])))
    assert engine.aggregate_verify(sigs[:192], pks[:96], hs[:1]) is True


def test_batch_verify_with_corruption():
    """config-5 semantics at small scale: 1% of the signatures replaced by another signer's"""
    from bls_b200 import engine, synth
    n = 1000
    sks = synth.scalars(31, n)
    hs = synth.message_hashes(32, n)
    H = engine.hash_to_g2(hs)
    sigs = engine.scalar_mul(H, sks, True).reshape(n, 192).copy()
    pks = engine.scalar_mul(np.tile(np.frombuffer(ser1(O.G1), dtype=np.uint8), n), sks, False)
    bad = synth.corrupted_indices(33, n)
    for i in bad:
        sigs[i] = sigs[(i + 1) % n]
    res = engine.verify_batch(pks, hs, sigs)
    want = np.ones(n, dtype=np.uint8)
    want[bad] = 0
    assert np.array_equal(res, want)


def test_verify_from_wire_bytes():
    """row f1: verification straight from the serialised formats; a signature or key that does not
    decode (the reference raises in from_bytes) is a rejection, never an error"""
    from bls_b200 import BLS, engine
    g = load_golden("sig_kat.json")
    tab = g["verify_table"]
    pks = [bytes.fromhex(t["pk"]) for t in tab]
    hs = [bytes.fromhex(t["h"]) for t in tab]
    sigs = [bytes.fromhex(t["sig"]) for t in tab]
    want = [t["ok"] for t in tab]
    # byte-level corruption of a valid signature (about half do not decode), same key and message
    flips = g["bitflips"]
    for c in flips["cases"]:
        pks.append(bytes.fromhex(flips["pk"]))
        hs.append(bytes.fromhex(flips["h"]))
        sigs.append(bytes.fromhex(c["sig"]))
        want.append(c["ok"])
    # corrupted public keys: x values with no point on the curve, and a decodable wrong key
    bad_pk = None
    for i in range(1, 60):
        cand = bytearray(pks[0])
        cand[20] ^= i
        try:
            O.g1_deserialize(bytes(cand))
        except ValueError:
            bad_pk = bytes(cand)
            break
    assert bad_pk is not None
    pks += [bad_pk, pks[1]]
    hs += [hs[0], hs[0]]
    sigs += [sigs[0], sigs[0]]
    want += [False, False]
    assert BLS.verify_batch_bytes(pks, hs, sigs) == want
    assert any(want) and not all(want)
    assert sum(1 for c in flips["cases"] if not c["decodes"]) > 0
    assert engine.verify_batch_wire(b"", b"", b"").size == 0


def test_aggregate_miller_partials_combine_to_aggregate_verify():
    """the sharded form: per-slice Miller partials (signature pair on the first slice only), their
    Fq12 product and ONE final exponentiation give the same verdict as b200bls_aggregate_verify"""
    from bls_b200 import engine, synth
    from bls_b200.distributed import aggregate_verify as sharded
    n = 37
    sks = synth.scalars(71, n)
    hs = synth.message_hashes(71, n)
    g1 = np.frombuffer(ser1(O.G1), dtype=np.uint8)
    sigs = engine.scalar_mul(engine.hash_to_g2(hs), sks, True)
    agg = engine.point_sum(sigs, True)
    pks = engine.scalar_mul(np.tile(g1, n), sks, False)
    assert engine.aggregate_verify(agg, pks, hs)
    cut = 15
    parts = [engine.aggregate_miller(agg, pks[:96 * cut], hs[:cut]).tobytes(),
             engine.aggregate_miller(None, pks[96 * cut:], hs[cut:]).tobytes()]
    prod = engine.field_op(12, "mul", parts[0], parts[1])
    one = (1).to_bytes(48, "big") + bytes(528)
    assert engine.final_exp_batch(prod).tobytes() == one
    assert sharded(agg, pks, hs) is True                      # world size 1
    bad = hs.copy()
    bad[3] = hs[4]
    assert sharded(agg, pks, bad) is False
    # only the signature pair: e(-G1, sig) alone is not one
    lone = engine.aggregate_miller(agg, b"", b"")
    assert engine.final_exp_batch(lone).tobytes() != one


def test_aggregate_verify_many_overlapping_jobs():
    """asynchronous aggregate verifications on all library streams: verdicts per job, good and bad"""
    from bls_b200 import engine, synth
    g1 = np.frombuffer(ser1(O.G1), dtype=np.uint8)
    jobs, want = [], []
    for j in range(7):
        n = 20 + 13 * j
        sks = synth.scalars(500 + j, n)
        hs = synth.message_hashes(500 + j, n)
        agg = engine.point_sum(engine.scalar_mul(engine.hash_to_g2(hs), sks, True), True)
        pks = engine.scalar_mul(np.tile(g1, n), sks, False)
        if j % 3 == 1:
            hs = hs.copy()
            hs[2] = hs[3]
        jobs.append((agg, pks, hs))
        want.append(j % 3 != 1)
    assert engine.aggregate_verify_many(jobs) == want
    assert [engine.aggregate_verify(*job) for job in jobs] == want
    assert engine.aggregate_verify_many([]) == []
