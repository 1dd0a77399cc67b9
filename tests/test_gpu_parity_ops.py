"""GPU parity of the single-function entry points: every golden vector the live
reference produced for qi_pow, pow, modsqrt, sw_encode, psi (tests/golden/field_kat.json, hash_kat.json) through the
C ABI, plus seeded random cases against the oracle, incl. the branches no honest input reaches."""
import random

import numpy as np
import pytest

import bls_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu
Q = O.Q


def ser(elems):
    return b"".join(int(c).to_bytes(48, "big") for c in elems)


def g2b(p):
    return ser((p[0][0], p[0][1], p[1][0], p[1][1]))


def test_frobenius_pow_sqrt_golden():
    """fields_t.py:344-364 and fields.py:199-205, 463-482: the 20 + 3 + 16 reference cases"""
    from bls_b200 import engine
    g = load_golden("field_kat.json")
    seen = {"frob": 0, "pow": 0, "sqrt": 0}
    for c in g["cases"]:
        level, op = c["level"], c["op"]
        if op not in seen:
            continue
        a = bytes.fromhex(c["a"])
        w = 48 * level
        assert len(a) == w
        if op == "frob":
            assert engine.field_frob(level, c["i"], a).tobytes().hex() == c["out"], (level, c["i"])
        elif op == "pow":
            assert engine.field_pow(level, a, [int(c["e"], 16)]).tobytes().hex() == c["out"], level
        else:
            out, ok = engine.field_sqrt(level, a)
            if c["out"] is None:
                assert ok[0] == 0 and not out.any()
            else:
                assert ok[0] == 1 and out.tobytes().hex() == c["out"], level
        seen[op] += 1
    assert seen == {"frob": 20, "pow": 3, "sqrt": 16}


def test_frobenius_and_pow_random_vs_oracle():
    from bls_b200 import engine
    rnd = random.Random(0xF0B)
    n = 300                                   # ragged, several CTAs
    xs = [tuple(rnd.randrange(Q) for _ in range(12)) for _ in range(n)]
    a = b"".join(ser(x) for x in xs)
    for i in (1, 2, 3, 6, 11):
        out = engine.field_frob(12, i, a).tobytes()
        for k in (0, 1, 137, n - 1):
            assert out[576 * k:576 * (k + 1)] == ser(O.f12_frob(xs[k], i)), (i, k)
    # x^(q^i) really is the power (tests.py:60-68)
    assert engine.field_pow(12, ser(xs[0]), [Q]).tobytes() == engine.field_frob(12, 1, ser(xs[0])).tobytes()
    # powers: per-item exponents, edge exponents 0, 1, 2^384 - 1
    es = [0, 1, (1 << 384) - 1] + [rnd.getrandbits(rnd.choice((8, 64, 255, 384))) for _ in range(29)]
    for level in (1, 2, 6, 12):
        m = len(es)
        vals = [xs[k][:level] for k in range(m)]
        out = engine.field_pow(level, b"".join(ser(v) for v in vals), es).tobytes()
        w = 48 * level
        for k in range(m):
            want = O.f12_pow(tuple(vals[k]) + (0,) * (12 - level), es[k])[:level] if level > 1 else (pow(vals[k][0], es[k], Q),)
            assert out[w * k:w * (k + 1)] == ser(want), (level, k)


def test_sqrt_random_and_edges_vs_oracle():
    """Fq / Fq2 modsqrt: the reference's own root (not just a root), 'no sqrt' as a flag, zero, real inputs"""
    from bls_b200 import engine
    rnd = random.Random(0x50)
    vals = [0, 1, 4, Q - 1] + [rnd.randrange(Q) for _ in range(200)]
    out, ok = engine.field_sqrt(1, ser(vals))
    n_no = 0
    for k, v in enumerate(vals):
        try:
            want = O.fq_sqrt(v)
        except ValueError:
            want = None
        if want is None:
            n_no += 1
            assert ok[k] == 0 and not out[48 * k:48 * (k + 1)].any(), k
        else:
            assert ok[k] == 1 and out.tobytes()[48 * k:48 * (k + 1)] == want.to_bytes(48, "big"), k
    assert 60 < n_no < 140
    v2 = [(0, 0), (4, 0), (Q - 1, 0), (0, 1), (0, Q - 1), (3, 4)] + [(rnd.randrange(Q), rnd.randrange(Q)) for _ in range(200)]
    v2 += [O.f2_mul(v, v) for v in v2[6:60]]             # guaranteed squares
    out, ok = engine.field_sqrt(2, b"".join(ser(v) for v in v2))
    raw = out.tobytes()
    n_no = 0
    for k, v in enumerate(v2):
        try:
            want = O.f2_sqrt(v)
        except ValueError:
            want = None
        if isinstance(want, int):                        # the reference hands a real element to the Fq root
            want = (want, 0)
        if want is None:
            n_no += 1
            assert ok[k] == 0 and not out[96 * k:96 * (k + 1)].any(), k
        else:
            assert ok[k] == 1 and raw[96 * k:96 * (k + 1)] == ser(want), k
            assert O.f2_mul(want, want) == (v[0] % Q, v[1] % Q)
    assert n_no > 50


def test_sw_encode_all_golden_vectors_and_unreachable_branches():
    """ec.py:449-507: the 29 golden vectors (t = 0 -> infinity among them), seeded random t vs the oracle"""
    from bls_b200 import engine
    g = load_golden("hash_kat.json")
    cases = g["sw_encode"]
    assert len(cases) == 29 and any(c["out"]["inf"] for c in cases)
    out = engine.sw_encode_g2(b"".join(bytes.fromhex(c["t"]) for c in cases)).tobytes()
    for i, c in enumerate(cases):
        want = bytes(192) if c["out"]["inf"] else bytes.fromhex(c["out"]["x"] + c["out"]["y"])
        assert out[192 * i:192 * (i + 1)] == want, i
    rnd = random.Random(0x5E2)
    ts = [(rnd.randrange(Q), rnd.randrange(Q)) for _ in range(400)] + [(rnd.randrange(Q), 0) for _ in range(8)] + \
         [(0, rnd.randrange(Q)) for _ in range(8)]
    out = engine.sw_encode_g2(b"".join(ser(t) for t in ts)).tobytes()
    idx = list(range(0, 400, 9)) + list(range(400, len(ts)))
    for k in idx:
        x, y, inf = O.sw_encode_g2(ts[k])
        assert not inf and out[192 * k:192 * (k + 1)] == ser((x[0], x[1], y[0], y[1])), k
    # every output is on the twist y^2 = x^3 + 4(1 + u)
    for k in range(0, len(ts), 5):
        p = out[192 * k:192 * (k + 1)]
        c = [int.from_bytes(p[48 * j:48 * (j + 1)], "big") for j in range(4)]
        x, y = (c[0], c[1]), (c[2], c[3])
        assert O.f2_mul(y, y) == O.f2_add(O.f2_mul(O.f2_mul(x, x), x), (4, 4)), k
    # w0 = t^2 + 5 + 4u = 0 (-> generator) has no solution over Fq2: the branch is dead for every input
    assert pow(41, (Q - 1) // 2, Q) == Q - 1


def test_twist_maps_and_psi():
    """fields_t.py:936-943 / 1018-1031, ec.py:402-444: golden psi vectors, the oracle's Fq12 arithmetic, and the
    identities twist(untwist(P)) = P, psi = twist . frob . untwist computed from the separate entry points"""
    from bls_b200 import engine
    g = load_golden("hash_kat.json")
    for c in g["psi"]:
        got = engine.g2_psi(bytes.fromhex(c["p"]["x"] + c["p"]["y"])).tobytes()
        assert got.hex() == c["out"]["x"] + c["out"]["y"]
    rnd = random.Random(0x7157)
    pts = [O.aff_mul(rnd.randrange(1, O.N), O.G2) for _ in range(6)]
    P = b"".join(g2b(p) for p in pts)
    ut = engine.g2_untwist(P).tobytes()
    for i, p in enumerate(pts):
        ux, uy, _ = O.untwist((p[0], p[1], False))
        assert ut[1152 * i:1152 * (i + 1)] == ser(ux) + ser(uy), i
    back = engine.fq12_twist(ut).tobytes()
    for i, p in enumerate(pts):
        assert back[1152 * i:1152 * (i + 1)] == ser(tuple(p[0]) + (0,) * 10) + ser(tuple(p[1]) + (0,) * 10), i
    # psi from its parts: Frobenius of both untwisted coordinates, then twist
    n = len(pts)
    coords = b"".join(ut[576 * k:576 * (k + 1)] for k in range(2 * n))
    fr = engine.field_frob(12, 1, coords).tobytes()
    tw = engine.fq12_twist(fr).tobytes()
    psi = engine.g2_psi(P).tobytes()
    for i, p in enumerate(pts):
        q = O.psi((p[0], p[1], False))
        assert psi[192 * i:192 * (i + 1)] == g2b(q), i
        assert tw[1152 * i:1152 * i + 96] == psi[192 * i:192 * i + 96]
        assert tw[1152 * i + 576:1152 * i + 672] == psi[192 * i + 96:192 * (i + 1)]
        assert not any(tw[1152 * i + 96:1152 * i + 576]) and not any(tw[1152 * i + 672:1152 * (i + 1)])
    # twist of arbitrary Fq12 coordinates
    xs = [tuple(rnd.randrange(Q) for _ in range(12)) for _ in range(4)]
    got = engine.fq12_twist(ser(xs[0]) + ser(xs[1]) + ser(xs[2]) + ser(xs[3])).tobytes()
    for i in range(2):
        tx, ty, _ = O.twist((xs[2 * i], xs[2 * i + 1], False))
        assert got[1152 * i:1152 * (i + 1)] == ser(tx) + ser(ty)
