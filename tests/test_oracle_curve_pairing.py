"""Oracle vs golden vectors generated from the live reference (tools/gen_golden.py)."""
import hashlib

import bls_oracle as O
from conftest import load_golden, unhex_elems

Q, N = O.Q, O.N


def g1pt(d):
    return (unhex_elems(d["x"])[0], unhex_elems(d["y"])[0], d["inf"])


def g2pt(d):
    return (unhex_elems(d["x"]), unhex_elems(d["y"]), d["inf"])


def test_scalar_mult_and_add():
    g = load_golden("curve_kat.json")
    for c in g["g1_mul"]:
        out = O.aff_mul(int(c["k"], 16), g1pt(c["p"]))
        assert out == g1pt(c["out"])
        if "ser" in c:
            assert O.g1_serialize(out).hex() == c["ser"]
    for c in g["g2_mul"][::2]:
        out = O.aff_mul(int(c["k"], 16), g2pt(c["p"]))
        assert out == g2pt(c["out"])
        if "ser" in c:
            assert O.g2_serialize(out).hex() == c["ser"]
    for c in g["g1_add"]:
        out = O.to_aff(O.jac_add(O.to_jac(g1pt(c["a"])), O.to_jac(g1pt(c["b"]))))
        assert out == g1pt(c["out"]) and O.g1_serialize(out).hex() == c["ser"]
    for c in g["g2_add"]:
        out = O.to_aff(O.jac_add(O.to_jac(g2pt(c["a"])), O.to_jac(g2pt(c["b"]))))
        assert out == g2pt(c["out"]) and O.g2_serialize(out).hex() == c["ser"]


def test_pairing_known_answers():
    g = load_golden("pairing_kat.json")
    assert O.f12_serialize(O.miller_loop(O.G1, O.G2)).hex() == g["miller_loop_g1_g2"]
    pts = []
    for c in g["pairs"]:
        p, q = g1pt(c["p"]), g2pt(c["q"])
        pts.append((p, q))
    for c, (p, q) in list(zip(g["pairs"], pts))[:4]:
        assert O.f12_serialize(O.ate_pairing(p, q)).hex() == c["out"]
    e = O.ate_pairing(O.G1, O.G2)
    # digests recorded in SURVEY.md 8c
    assert hashlib.sha256(O.f12_serialize(e)).hexdigest() == \
        "70f0561453673ff155a40ba3618727f8a411c492748d845280dd71dce099905a"
    for m in g["multi"][:2]:
        ps, qs = [pts[i][0] for i in m["idx"]], [pts[i][1] for i in m["idx"]]
        assert O.f12_serialize(O.ate_pairing_multi(ps, qs)).hex() == m["out"]
    for d in g["degenerate"]:
        assert O.f12_serialize(O.ate_pairing(g1pt(d["p"]), g2pt(d["q"]))).hex() == d["out"]
        assert unhex_elems(d["out"]) == O.F12_ONE


def test_bilinearity():
    e1 = O.ate_pairing(O.aff_mul(3, O.G1), O.aff_mul(5, O.G2))
    e2 = O.f12_pow(O.ate_pairing(O.G1, O.G2), 15)
    assert e1 == e2


def test_hash_to_g2():
    g = load_golden("hash_kat.json")
    for c in g["hash_to_g2_prehashed"][:8]:
        p = O.hash_to_g2_prehashed(bytes.fromhex(c["h"]))
        assert p == g2pt(c["out"]) and O.g2_serialize(p).hex() == c["ser"]
    for c in g["sw_encode"]:
        assert O.sw_encode_g2(unhex_elems(c["t"])) == g2pt(c["out"])
    for c in g["psi"]:
        assert O.psi(g2pt(c["p"])) == g2pt(c["out"])


def test_signatures_and_verify():
    g = load_golden("sig_kat.json")
    sks = []
    for k in g["keys"]:
        sk = O.sk_from_seed(bytes.fromhex(k["seed"]))
        sks.append(sk)
        assert sk.to_bytes(32, "big").hex() == k["sk"]
        pk = O.g1_serialize(O.pk_of(sk))
        assert pk.hex() == k["pk"]
        assert int.from_bytes(O.hash256(pk)[:4], "big") == k["fingerprint"]
    for c in g["sign"][:6]:
        sig = O.sign(sks[c["key"]], bytes.fromhex(c["msg"]))
        assert O.g2_serialize(sig).hex() == c["sig"]
    for t in g["verify_table"][:6]:
        pk = O.g1_deserialize(bytes.fromhex(t["pk"]))
        sig = O.g2_deserialize(bytes.fromhex(t["sig"]))
        assert O.verify(pk, bytes.fromhex(t["h"]), sig) == t["ok"]
    # secure aggregate of test_vectors (tests.py:124-127): sum T_i * sig_i
    sig_bytes = {(c["key"], c["msg"]): c["sig"] for c in g["sign"]}
    m = bytes([7, 8, 9]).hex()
    pks = [bytes.fromhex(k["pk"]) for k in g["keys"]]
    order = sorted(range(2), key=lambda i: pks[i])
    ts = O.hash_pks(2, [pks[i] for i in order])
    acc = O.jac_inf(O._F2)
    for t, i in zip(ts, order):
        s = O.g2_deserialize(bytes.fromhex(sig_bytes[(i, m)]))
        acc = O.jac_add(acc, O.jac_mul(t, O.to_jac(s)))
    assert O.g2_serialize(acc).hex() == g["test_vectors"]["secure_agg_sig"]


def test_bitflip_decoding():
    g = load_golden("sig_kat.json")["bitflips"]
    n_raise = 0
    for c in g["cases"]:
        try:
            p = O.g2_deserialize(bytes.fromhex(c["sig"]))
            assert c["decodes"] and p == g2pt(c["point"])
        except ValueError:
            assert not c["decodes"]
            n_raise += 1
    assert 0 < n_raise < len(g["cases"])
    pk = O.g1_deserialize(bytes.fromhex(g["pk"]))
    done = 0
    for c in g["cases"]:
        if c["decodes"] and done < 2:
            sig = O.g2_deserialize(bytes.fromhex(c["sig"]))
            assert O.verify(pk, bytes.fromhex(g["h"]), sig) == c["ok"]
            done += 1


def test_aggregation_sums():
    from bls_b200 import synth
    g = load_golden("agg_kat.json")
    ks = synth.scalar_ints(synth.SEED_AGGREGATE, 1000)
    for s in g["sums"]:
        n = s["n"]
        if n <= 3:
            p1 = [O.aff_mul(k, O.G1) for k in ks[:n]]
            p2 = [O.aff_mul(k, O.G2) for k in ks[:n]]
            assert O.g1_serialize(O.g1_sum(p1)).hex() == s["g1_sum"]
            assert O.g2_serialize(O.g2_sum(p2)).hex() == s["g2_sum"]
        # size-independent identity: sum(k_i G) == (sum k_i) G
        tot = sum(ks[:n]) % N
        assert O.g1_serialize(O.aff_mul(tot, O.G1)).hex() == s["g1_sum"]
        assert O.g2_serialize(O.aff_mul(tot, O.G2)).hex() == s["g2_sum"]


def test_cyclotomic_subgroup_identities_behind_compressed_squaring():
    """The relations bls_b200/programs/tower.py: decompress_many solves, checked with the ORACLE's
    arithmetic on elements of the cyclotomic subgroup (f^((q^6-1)(q^2+1)) of random f).  With Fq12 =
    A + B w + C w^2 over Fq2[s], s = w^3, z0..z5 as in F12.cyclotomic_sqr:
        4 z2 z1 = 3 z4^2 + xi z5^2 - 2 z3        (general branch)
        z0 z2 + xi z1 z3 = 2 xi z4 z5 + z2       (gives z1 = 2 z4 z5 / z3 on the z2 = 0 branch)
        z0 = xi (2 z1^2 + z2 z5 - 3 z3 z4) + 1
    and the compressed squaring itself (B, C of the square from B, C alone)."""
    import random
    rnd = random.Random(5)
    Q = O.Q
    for _ in range(3):
        f = tuple(rnd.randrange(Q) for _ in range(12))
        t = O.f12_mul(O.f12_frob(f, 6), O.f12_inv(f))
        m = O.f12_mul(O.f12_frob(t, 2), t)
        c = [(m[2 * k], m[2 * k + 1]) for k in range(6)]        # flat ZT order: c0.a0 c0.a1 c0.a2 c1.a0 c1.a1 c1.a2
        z0, z4, z3, z2, z1, z5 = c
        mul, add, sub, xi = O.f2_mul, O.f2_add, O.f2_sub, O.f2_mul_xi
        k = lambda a, n: O.f2_scale(a, n)
        assert k(mul(z2, z1), 4) == sub(add(k(mul(z4, z4), 3), xi(mul(z5, z5))), k(z3, 2))
        assert add(mul(z0, z2), xi(mul(z1, z3))) == add(k(xi(mul(z4, z5)), 2), z2)
        inner = sub(add(k(mul(z1, z1), 2), mul(z2, z5)), k(mul(z3, z4), 3))
        assert z0 == add(xi(inner), (1, 0))
        m2 = O.f12_mul(m, m)
        d = [(m2[2 * j], m2[2 * j + 1]) for j in range(6)]
        sq4 = lambda a, b: (add(mul(a, a), xi(mul(b, b))), k(mul(a, b), 2))
        t0, t1 = sq4(z2, z3)
        t2, t3 = sq4(z4, z5)
        assert d[3] == add(k(xi(t3), 3), k(z2, 2)) and d[2] == sub(k(t2, 3), k(z3, 2))      # new z2, z3
        assert d[1] == sub(k(t0, 3), k(z4, 2)) and d[5] == add(k(t1, 3), k(z5, 2))          # new z4, z5
