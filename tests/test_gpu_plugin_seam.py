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
