"""The reference's plugin seam, served by the GPU library.

bls_py/fields_t.py:1218-1265 re-imports its heavy functions from an optional accelerator module
`bls_py.fields_t_c`; consumers bind late through `fields_t` (fields.py:3-18, ec.py:558-574,
pairing.py:95-100).  This module exports the functions of that list that are worth a GPU call,
with the reference's own signatures and value conventions (Fq = int, Fq2 / Fq12 = tuples of 2 / 12
ints, infinity = bool, results fully reduced), so a maintainer can write

    from bls_b200.fields_t_c import (fq_ate_pairing_multi, fq12_final_exp, fq_miller_loop,
                                     fq2_scalar_mult_jacobian, fq_scalar_mult_jacobian)

next to the existing `from .fields_t_c import (...)` block -- all 35 names of that block are
served here.  One element per call
is the reference's granularity, not the GPU's: the *_batch variants below take lists and are what
the throughput numbers are measured on.

Conventions that differ from the pure-Python functions, none of them observable after the
operations that follow in the reference:
  * Jacobian results are returned normalised (z = 1): the same point, not the same triple;
  * fq_miller_loop's value differs from the reference's by a factor that every final
    exponentiation removes (Miller loop), so only fq12_final_exp(fq_miller_loop(...))
    and fq_ate_pairing_multi are byte-comparable.
"""
from . import engine

Q = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
        "1eabfffeb153ffffb9feffffffffaaab", 16)
FQ2_ONE_TUPLE = (1, 0)
FQ2_ZERO_TUPLE = (0, 0)


def _b(z):
    return (int(z) % Q).to_bytes(48, "big")


def _g1_bytes(x, y, inf):
    return bytes(96) if inf else _b(x) + _b(y)


def _g2_bytes(x, y, inf):
    return bytes(192) if inf else _b(x[0]) + _b(x[1]) + _b(y[0]) + _b(y[1])


def _ints(raw):
    return tuple(int.from_bytes(raw[i:i + 48], "big") for i in range(0, len(raw), 48))


def _jac1(x, y, z, inf):
    """Jacobian point over Fq as 144 bytes; infinity is Z = 0 on the device"""
    return bytes(144) if inf else _b(x) + _b(y) + _b(z)


def _jac2(x, y, z, inf):
    return bytes(288) if inf else b"".join(_b(c) for c in (x[0], x[1], y[0], y[1], z[0], z[1]))


def _ser(t):
    return b"".join(_b(c) for c in ((t,) if isinstance(t, int) else t))


# ---- pairing ------------------------------------------------------------------------------------
def fq_ate_pairing_multi(Ps, Qs):
    """fields_t.py:1114-1121: Ps = [(px, py, pinf)], Qs = [((x0, x1), (y0, y1), qinf)] -> 12-tuple;
    all Miller loops in one batch, ONE final exponentiation"""
    if len(Ps) != len(Qs):
        raise ValueError("Ps and Qs differ in length")
    P = b"".join(_g1_bytes(*p) for p in Ps)
    Qb = b"".join(_g2_bytes(*q) for q in Qs)
    return _ints(engine.pairing_multi(P, Qb).tobytes())


def fq_ate_pairing_batch(Ps, Qs):
    """n independent pairings -> list of 12-tuples (the data-parallel form of ate_pairing)"""
    P = b"".join(_g1_bytes(*p) for p in Ps)
    Qb = b"".join(_g2_bytes(*q) for q in Qs)
    out = engine.pairing_batch(P, Qb).tobytes()
    return [_ints(out[576 * i:576 * (i + 1)]) for i in range(len(Ps))]


def fq_miller_loop(px, py, pinf, qx_t, qy_t, qinf):
    """fields_t.py:1091-1111 (value defined up to the factor a final exponentiation removes)"""
    return _ints(engine.miller_loop_batch(_g1_bytes(px, py, pinf), _g2_bytes(qx_t, qy_t, qinf)).tobytes())


def fq12_final_exp(t_x):
    """fields_t.py:1124-1128"""
    return _ints(engine.final_exp_batch(b"".join(_b(c) for c in t_x)).tobytes())


# ---- scalar multiplication ------------------------------------------------------------------------
def _scalar_chunks(c):
    """c as 256-bit chunks, most significant first (the device ladder takes 32-byte scalars)"""
    out = []
    while True:
        out.append(c & ((1 << 256) - 1))
        c >>= 256
        if not c:
            return out[::-1]


def _mul_bytes(point, c, g2):
    w = 192 if g2 else 96
    chunks = _scalar_chunks(int(c))
    acc = engine.scalar_mul(point, chunks[0].to_bytes(32, "big"), g2).tobytes()
    for ch in chunks[1:]:                   # acc = 2^256 acc + ch P  (scalars wider than 256 bits)
        acc = engine.scalar_mul(acc, (1 << 255).to_bytes(32, "big"), g2).tobytes()
        acc = engine.point_add(acc, acc, g2).tobytes()
        acc = engine.point_add(acc, engine.scalar_mul(point, ch.to_bytes(32, "big"), g2).tobytes(), g2).tobytes()
    assert len(acc) == w
    return acc


def _jmul(c, raw, g2):
    """c * P for a Jacobian point given as bytes: the normalisation and the ladder run in ONE device program; scalars
    wider than 256 bits are folded 256 bits at a time (acc = 2^256 acc + chunk P)"""
    chunks = _scalar_chunks(int(c))
    acc = engine.jacobian_op("mul", raw, chunks[0].to_bytes(32, "big"), g2).tobytes()
    if len(chunks) > 1:
        base = engine.jacobian_op("affine", raw, None, g2).tobytes()
        for ch in chunks[1:]:
            acc = engine.scalar_mul(acc, (1 << 255).to_bytes(32, "big"), g2).tobytes()
            acc = engine.point_add(acc, acc, g2).tobytes()
            acc = engine.point_add(acc, engine.scalar_mul(base, ch.to_bytes(32, "big"), g2).tobytes(), g2).tobytes()
    return acc


def _out1(r):
    if not any(r):
        return 1, 1, 0, True
    x, y = _ints(r)
    return x, y, 1, False


def _out2(r):
    if not any(r):
        return FQ2_ONE_TUPLE, FQ2_ONE_TUPLE, FQ2_ZERO_TUPLE, True
    v = _ints(r)
    return (v[0], v[1]), (v[2], v[3]), FQ2_ONE_TUPLE, False


def fq_scalar_mult_jacobian(c, x1, y1, z1, inf1):
    """fields_t.py:705-721: c * (x1, y1, z1) on E(Fq) -> (x, y, 1, inf)"""
    if inf1 or c % Q == 0:
        return 1, 1, 0, True
    if c < 0:
        raise ValueError("negative scalar")
    return _out1(_jmul(c, _jac1(x1, y1, z1, False), False))


def fq2_scalar_mult_jacobian(c, x1, y1, z1, inf1):
    """fields_t.py:724-740: c * (x1, y1, z1) on the twist -> (x, y, (1, 0), inf)"""
    if inf1 or c % Q == 0:
        return FQ2_ONE_TUPLE, FQ2_ONE_TUPLE, FQ2_ZERO_TUPLE, True
    if c < 0:
        raise ValueError("negative scalar")
    return _out2(_jmul(c, _jac2(x1, y1, z1, False), True))


# ---- Jacobian / affine point functions (inputs normalised on the device) ------------------------------------------
def fq_add_points_jacobian(x1, y1, z1, inf1, x2, y2, z2, inf2):
    """fields_t.py:762-797 (equal points double: the reference's 4-argument call at line 781 is a defect)"""
    return _out1(engine.jacobian_op("add", _jac1(x1, y1, z1, inf1), _jac1(x2, y2, z2, inf2), False).tobytes())


def fq2_add_points_jacobian(x1, y1, z1, inf1, x2, y2, z2, inf2):
    """fields_t.py:800-819"""
    return _out2(engine.jacobian_op("add", _jac2(x1, y1, z1, inf1), _jac2(x2, y2, z2, inf2), True).tobytes())


def fq_double_point_jacobian(X, Y, Z):
    """fields_t.py:878-897 -> (x, y, z) with z = 1"""
    return _out1(engine.jacobian_op("dbl", _jac1(X, Y, Z, False), None, False).tobytes())[:3]


def fq2_double_point_jacobian(X, Y, Z):
    """fields_t.py:900-903"""
    return _out2(engine.jacobian_op("dbl", _jac2(X, Y, Z, False), None, True).tobytes())[:3]


def fq2_to_affine(px, py, pz, pinf):
    """fields_t.py:617-622"""
    if pinf:
        return FQ2_ZERO_TUPLE, FQ2_ZERO_TUPLE, pinf
    v = _ints(engine.jacobian_op("affine", _jac2(px, py, pz, False), None, True).tobytes())
    return (v[0], v[1]), (v[2], v[3]), pinf


def fq2_add_points(x1, y1, inf1, x2, y2, inf2):
    """fields_t.py:673-686 (affine in, affine out; equal points double)"""
    r = engine.point_add(_g2_bytes(x1, y1, inf1), _g2_bytes(x2, y2, inf2), True).tobytes()
    v = _ints(r)
    return (v[0], v[1]), (v[2], v[3]), not any(r)


def fq2_double_point(px, py, pinf):
    """fields_t.py:641-646"""
    x, y, _ = fq2_add_points(px, py, pinf, px, py, pinf)
    return x, y, False


def fq2_scalar_mult_jacobian_batch(cs, points):
    """[c_i * P_i] for affine twist points [((x0, x1), (y0, y1), inf)] and scalars < 2^256, one call"""
    raw = b"".join(_g2_bytes(*p) for p in points)
    sc = b"".join(int(c).to_bytes(32, "big") for c in cs)
    out = engine.scalar_mul(raw, sc, True).tobytes()
    res = []
    for i in range(len(points)):
        r = out[192 * i:192 * (i + 1)]
        v = _ints(r)
        res.append(((v[0], v[1]), (v[2], v[3]), not any(r)))
    return res


# ---- field functions of the seam: one device call each ---------------------------------------------------------------
def _field(level, op, a, b=None):
    return _ints(engine.field_op(level, op, _ser(a), _ser(b) if b is not None else None).tobytes())


def fq_invert(P, a):
    """fields_t.py:47-55 (0 -> 0)"""
    return _field(1, "inv", a % Q)[0]


def fq_floordiv(P, a, X):
    """fields_t.py:71-72: a / X"""
    return _field(1, "mul", a % Q, fq_invert(P, X))[0]


def fq_pow(P, a, X):
    """fields_t.py:58-68"""
    if X < 0:
        return fq_pow(P, fq_invert(P, a), -X)
    return _pow(1, (a % Q,), X)[0]


def _pow(level, t, e):
    from . import fields
    cls = {1: fields.Fq, 2: fields.Fq2, 6: fields.Fq6, 12: fields.Fq12}[level]
    return (cls(_ser(t)) ** e).ZT


def fq2_invert(t_a):
    return _field(2, "inv", t_a)


def fq2_floordiv(t_a, t_x):
    return _field(2, "mul", t_a, fq2_invert(t_x))


def fq2_pow(t_a, e):
    return _pow(2, t_a, e)


def fq2_qi_pow(t_x, i):
    """fields_t.py:104-110"""
    return _ints(engine.field_frob(2, i % 2, _ser(t_x)).tobytes())


def fq6_invert(t_x):
    return _field(6, "inv", t_x)


def fq6_floordiv(t_a, t_x):
    return _field(6, "mul", t_a, fq6_invert(t_x))


def fq6_qi_pow(t_x, i):
    """fields_t.py:203-212"""
    return _ints(engine.field_frob(6, i % 6, _ser(t_x)).tobytes())


def fq6_add(t_a, t_m):
    return _field(6, "add", t_a, t_m)


def fq6_mul(t_a, t_m):
    """fields_t.py:293-318"""
    return _field(6, "mul", t_a, t_m)


def fq12_invert(t_x):
    """fields_t.py:328-337"""
    return _field(12, "inv", t_x)


def fq12_floordiv(t_a, t_x):
    return _field(12, "mul", t_a, fq12_invert(t_x))


def fq12_qi_pow(t_x, i):
    """fields_t.py:355-364"""
    return _ints(engine.field_frob(12, i % 12, _ser(t_x)).tobytes())


def fq12_pow(t_a, e):
    """fields_t.py:344-352"""
    return _pow(12, t_a, e)


def fq12_mul_fq(t_a, m):
    """fields_t.py:448-452"""
    return _field(12, "mul", t_a, (m % Q,) + (0,) * 11)


def fq12_add(t_a, t_m):
    return _field(12, "add", t_a, t_m)


def fq12_mul(t_a, t_m):
    """fields_t.py:503-554"""
    return _field(12, "mul", t_a, t_m)


# ---- twist and the reference's dense line functions ----------------------------------------------------------------
def fq2_untwist(x_t, y_t):
    """fields_t.py:936-943: (x, y) on the twist -> (x / w^2, y / w^3) as two 12-tuples"""
    v = _ints(engine.g2_untwist(_g2_bytes(x_t, y_t, False)).tobytes())
    return v[:12], v[12:]


def _f12(t):
    from . import fields
    return fields.Fq12(_ser(t))


def fq12_double_point_jacobian(X, Y, Z):
    """fields_t.py:906-933 on Fq12 coordinates (a = 0): composed from device field operations"""
    X, Y, Z = _f12(X), _f12(Y), _f12(Z)
    yy = Y * Y
    s = (X * yy) * 4
    m = (X * X) * 3
    x3 = m * m - s * 2
    y3 = m * (s - x3) - (yy * yy) * 8
    z3 = (Y * Z) * 2
    return x3.ZT, y3.ZT, z3.ZT


def fq12_add_points_jacobian(x1, y1, z1, inf1, x2, y2, z2, inf2):
    """fields_t.py:822-841 on Fq12 coordinates: composed from device field operations"""
    from . import fields
    one, zero = fields.Fq12.one().ZT, fields.Fq12.zero().ZT
    if inf1:
        return x2, y2, z2, inf2
    if inf2:
        return x1, y1, z1, inf1
    X1, Y1, Z1, X2, Y2, Z2 = (_f12(t) for t in (x1, y1, z1, x2, y2, z2))
    z1z1, z2z2 = Z1 * Z1, Z2 * Z2
    u1, u2 = X1 * z2z2, X2 * z1z1
    s1, s2 = Y1 * (z2z2 * Z2), Y2 * (z1z1 * Z1)
    if u1 == u2:
        if s1 != s2:
            return one, one, zero, True
        return fq12_double_point_jacobian(x1, y1, z1) + (False,)
    h, r = u2 - u1, s2 - s1
    hh = h * h
    hhh, v = hh * h, u1 * hh
    x3 = r * r - hhh - v * 2
    y3 = r * (v - x3) - s1 * hhh
    z3 = (Z1 * Z2) * h
    return x3.ZT, y3.ZT, z3.ZT, False


def fq2_double_line_eval(rx_t, ry_t, px, py):
    """fields_t.py:1035-1049: the reference's dense tangent-line value (slope through an Fq12 inversion).  The Miller
    loop of this library never forms it (sparse projective lines); offered for parity of the function."""
    ux, uy = (_f12(t) for t in fq2_untwist(rx_t, ry_t))
    slope = (ux * ux * 3) * ~(uy * 2)
    v = uy - slope * ux
    return ((-(slope * px)) + py - v).ZT


def fq2_add_line_eval(rx_t, ry_t, qx_t, qy_t, px, py):
    """fields_t.py:1052-1078.  The special case is tested exactly as the reference tests it -- BOTH coordinates of
    untwist(R) equal to the negated coordinates of untwist(Q) (lines 1062-1065), not the geometric R = -Q -- so the
    two functions agree on every input."""
    rx, ry = (_f12(t) for t in fq2_untwist(rx_t, ry_t))
    qx, qy = (_f12(t) for t in fq2_untwist(qx_t, qy_t))
    if rx == -qx and ry == -qy:
        return ((-rx) + px).ZT
    slope = (qy - ry) * ~(qx - rx)
    v = (qy * rx - ry * qx) * ~(rx - qx)
    return ((-(slope * px)) + py - v).ZT
