"""GPU: the reference's plugin seam (bls_py/fields_t.py:1218-1265) served by bls_b200.fields_t_c with
the reference's own signatures and tuple conventions, against golden values from the live reference
and the oracle."""
import pytest

import bls_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _p(c):
    return (int(c["p"]["x"], 16), int(c["p"]["y"], 16), False)


def _q(c):
    x, y = bytes.fromhex(c["q"]["x"]), bytes.fromhex(c["q"]["y"])
    t = lambda b: (int.from_bytes(b[:48], "big"), int.from_bytes(b[48:], "big"))
    return (t(x), t(y), False)


def _f12(hexstr):
    raw = bytes.fromhex(hexstr)
    return tuple(int.from_bytes(raw[i:i + 48], "big") for i in range(0, 576, 48))


def test_pairing_functions_with_reference_signatures():
    from bls_b200 import fields_t_c as C
    g = load_golden("pairing_kat.json")
    pairs = g["pairs"][:4]
    for c in pairs:
        px, py, pinf = _p(c)
        qx, qy, qinf = _q(c)
        want = _f12(c["out"])
        assert C.fq_ate_pairing_multi([(px, py, pinf)], [(qx, qy, qinf)]) == want
        # the Miller value is only defined up to what the final exponentiation removes
        assert C.fq12_final_exp(C.fq_miller_loop(px, py, pinf, qx, qy, qinf)) == want
    assert C.fq_ate_pairing_batch([_p(c) for c in pairs], [_q(c) for c in pairs]) == [_f12(c["out"]) for c in pairs]
    for m in g["multi"]:                      # multi-pairings over pairs[idx]
        cs = [g["pairs"][i] for i in m["idx"]]
        assert C.fq_ate_pairing_multi([_p(c) for c in cs], [_q(c) for c in cs]) == _f12(m["out"])
    # infinity on either side pairs to one, as the reference's computation on (0, 0) does
    one = (1,) + (0,) * 11
    px, py, _ = _p(pairs[0])
    qx, qy, _ = _q(pairs[0])
    assert C.fq_ate_pairing_multi([(0, 0, True)], [(qx, qy, False)]) == one
    assert C.fq_ate_pairing_multi([(px, py, False)], [((0, 0), (0, 0), True)]) == one


def test_scalar_mult_jacobian_with_reference_signatures():
    from bls_b200 import fields_t_c as C
    k = 0x1234567890abcdef1234567890abcdef1234567890abcdef1234567890abcdef % O.N
    # Jacobian input with z != 1: (x z^2, y z^3, z)
    z = 0x1d0c5
    gx, gy = O.G1[0], O.G1[1]
    xr, yr, zr, inf = C.fq_scalar_mult_jacobian(k, gx * z * z % O.Q, gy * z * z * z % O.Q, z, False)
    want = O.aff_mul(k, O.G1)
    assert (xr, yr, zr, inf) == (want[0], want[1], 1, False)
    assert C.fq_scalar_mult_jacobian(0, gx, gy, 1, False) == (1, 1, 0, True)
    assert C.fq_scalar_mult_jacobian(5, gx, gy, 1, True) == (1, 1, 0, True)
    assert C.fq_scalar_mult_jacobian(O.N, gx, gy, 1, False) == (1, 1, 0, True)
    # wider than 256 bits (the reference takes any non-negative int)
    big = (1 << 300) + 12345
    xr, yr, zr, inf = C.fq_scalar_mult_jacobian(big, gx, gy, 1, False)
    want = O.aff_mul(big % O.N, O.G1)
    assert (xr, yr) == (want[0], want[1]) and not inf
    # twist
    z2 = (3, 7)
    z2s = O.f2_mul(z2, z2)
    X = O.f2_mul(O.G2[0], z2s)
    Y = O.f2_mul(O.G2[1], O.f2_mul(z2s, z2))
    xr, yr, zr, inf = C.fq2_scalar_mult_jacobian(k, X, Y, z2, False)
    want = O.aff_mul(k, O.G2)
    assert (xr, yr, zr, inf) == (want[0], want[1], (1, 0), False)
    assert C.fq2_scalar_mult_jacobian(0, X, Y, z2, False) == ((1, 0), (1, 0), (0, 0), True)
    res = C.fq2_scalar_mult_jacobian_batch([k, 0, 3], [(O.G2[0], O.G2[1], False)] * 3)
    assert res[0] == (want[0], want[1], False) and res[1][2] is True
    w3 = O.aff_mul(3, O.G2)
    assert res[2] == (w3[0], w3[1], False)


def _unhex(v):
    if isinstance(v, bool):
        return v
    if isinstance(v, str):
        return int(v, 16)
    return tuple(_unhex(x) for x in v)


def _normalise(name, res):
    """Jacobian triples -> the affine point they represent, through the ORACLE (the seam returns z = 1, the reference
    whatever its formulas leave: the same point, not the same triple)"""
    level = 12 if name.startswith("fq12") else (2 if name.startswith("fq2") else 1)
    x, y, z = res[0], res[1], res[2]
    if len(res) > 3 and res[3]:
        return "infinity"
    if level == 1:
        zi = O.fq_inv(z)
        return (x * zi * zi % O.Q, y * zi * zi * zi % O.Q)
    mul, inv = (O.f2_mul, O.f2_inv) if level == 2 else (O.f12_mul, O.f12_inv)
    zi = inv(z)
    zi2 = mul(zi, zi)
    return (mul(x, zi2), mul(y, mul(zi2, zi)))


def test_all_35_seam_functions_against_the_live_reference_vectors():
    """tests/golden/seam_kat.json (tools/gen_seam_golden.py): every function name of fields_t.py:1218-1265, inputs and
    outputs recorded from the reference's pure-Python definitions"""
    from bls_b200 import fields_t_c as C
    g = load_golden("seam_kat.json")
    seen = set()
    for c in g["cases"]:
        fn = getattr(C, c["fn"])
        args = [_unhex(a) for a in c["args"]]
        got = fn(*args)
        want = _unhex(c["out"])
        if c.get("norm"):
            assert _normalise(c["fn"], got) == _normalise(c["fn"], want), c["fn"]
        else:
            assert got == want, c["fn"]
        seen.add(c["fn"])
    names = """fq_invert fq_floordiv fq_pow fq2_invert fq2_floordiv fq2_qi_pow fq2_pow fq6_invert fq6_floordiv fq6_qi_pow
               fq6_add fq6_mul fq12_invert fq12_floordiv fq12_qi_pow fq12_pow fq12_mul_fq fq12_add fq12_mul fq2_to_affine
               fq2_double_point fq2_add_points fq_double_point_jacobian fq2_double_point_jacobian fq12_double_point_jacobian
               fq_add_points_jacobian fq2_add_points_jacobian fq12_add_points_jacobian fq2_scalar_mult_jacobian
               fq2_double_line_eval fq2_add_line_eval fq2_untwist fq_miller_loop fq12_final_exp fq_ate_pairing_multi""".split()
    assert len(names) == 35
    for nm in names:
        assert callable(getattr(C, nm)), nm
    # the three pairing functions are covered by test_pairing_functions_with_reference_signatures
    assert set(names) - seen == {"fq_miller_loop", "fq12_final_exp", "fq_ate_pairing_multi"}


def test_field_classes_mirror_the_reference_surface():
    """bls_b200.fields.Fq / Fq2 / Fq6 / Fq12 (fields.py:35-764): operators, ~, pow, qi_power, modsqrt, mixed levels"""
    import random
    from bls_b200.fields import Fq, Fq2, Fq6, Fq12, Q
    rnd = random.Random(0xF1E1D)
    a12 = tuple(rnd.randrange(Q) for _ in range(12))
    b12 = tuple(rnd.randrange(Q) for _ in range(12))
    A, B = Fq12(Q, a12), Fq12(Q, b12)
    assert (A * B).ZT == O.f12_mul(a12, b12) and (A + B).ZT == O.f12_add(a12, b12) and (A - B).ZT == O.f12_sub(a12, b12)
    assert (~A).ZT == O.f12_inv(a12) and (A / B).ZT == O.f12_mul(a12, O.f12_inv(b12))
    assert (A ** 0) == Fq12.one() and (A ** 5).ZT == O.f12_pow(a12, 5) and (A ** -3).ZT == O.f12_pow(O.f12_inv(a12), 3)
    e = rnd.getrandbits(900)                                 # wider than one 384-bit device exponent
    assert (A ** e).ZT == O.f12_pow(a12, e)
    assert A.qi_power(3).ZT == O.f12_frob(a12, 3)
    assert (A * 7).ZT == O.f12_mul(a12, (7,) + (0,) * 11) and (7 * A) == (A * 7) and (A * Fq(Q, 7)) == (A * 7)
    x, y = Fq2(Q, 3, 4), Fq2(Q, 5, 6)
    assert (x * y).ZT == O.f2_mul((3, 4), (5, 6)) and (-x + x) == Fq2.zero() and not (x - x)
    assert [c.Z for c in x] == [3, 4] and x[1] == Fq(Q, 4)
    s = (x * x).modsqrt()
    assert s * s == x * x and s.ZT == O.f2_sqrt(O.f2_mul((3, 4), (3, 4)))
    assert Fq(Q, 16).modsqrt().Z == O.fq_sqrt(16) and isinstance(Fq2(Q, 16, 0).modsqrt(), Fq)
    with pytest.raises(ValueError):
        nonres = next(v for v in range(2, 50) if pow(v, (Q - 1) // 2, Q) != 1)
        Fq(Q, nonres).modsqrt()
    a6 = tuple(rnd.randrange(Q) for _ in range(6))
    F = Fq6(Q, a6)
    assert (F * F).ZT == O.f12_mul(a6 + (0,) * 6, a6 + (0,) * 6)[:6]
    assert [c.ZT for c in F] == [a6[0:2], a6[2:4], a6[4:6]]
    assert A.serialize() == b"".join(c.to_bytes(48, "big") for c in a12) and hash(A) == hash(Fq12(Q, a12))
