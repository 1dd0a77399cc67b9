"""The reference's plugin seam, served by the GPU library.

bls_py/fields_t.py:1218-1265 re-imports its heavy functions from an optional accelerator module
`bls_py.fields_t_c`; consumers bind late through `fields_t` (fields.py:3-18, ec.py:558-574,
pairing.py:95-100).  This module exports the functions of that list that are worth a GPU call,
with the reference's own signatures and value conventions (Fq = int, Fq2 / Fq12 = tuples of 2 / 12
ints, infinity = bool, results fully reduced), so a maintainer can write

    from bls_b200.fields_t_c import (fq_ate_pairing_multi, fq12_final_exp, fq_miller_loop,
                                     fq2_scalar_mult_jacobian, fq_scalar_mult_jacobian)

next to the existing `from .fields_t_c import (...)` block.  One element per call
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


# ---- Fq2 helpers on the host (Jacobian -> affine of ONE input point; a handful of int products) ----
def _f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def _f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % Q, -1, Q)
    return (a[0] * n % Q, -a[1] * n % Q)


def _affine1(x, y, z):
    zi = pow(int(z) % Q, -1, Q)
    return x * zi * zi % Q, y * zi * zi * zi % Q


def _affine2(x, y, z):
    zi = _f2_inv(z)
    zi2 = _f2_mul(zi, zi)
    return _f2_mul(x, zi2), _f2_mul(y, _f2_mul(zi2, zi))


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


def fq_scalar_mult_jacobian(c, x1, y1, z1, inf1):
    """fields_t.py:705-721: c * (x1, y1, z1) on E(Fq) -> (x, y, 1, inf)"""
    if inf1 or c % Q == 0:
        return 1, 1, 0, True
    if c < 0:
        raise ValueError("negative scalar")
    x, y = _affine1(x1, y1, z1)
    r = _mul_bytes(_g1_bytes(x, y, False), c, False)
    if not any(r):
        return 1, 1, 0, True
    xr, yr = _ints(r)
    return xr, yr, 1, False


def fq2_scalar_mult_jacobian(c, x1, y1, z1, inf1):
    """fields_t.py:724-740: c * (x1, y1, z1) on the twist -> (x, y, (1, 0), inf)"""
    if inf1 or c % Q == 0:
        return FQ2_ONE_TUPLE, FQ2_ONE_TUPLE, FQ2_ZERO_TUPLE, True
    if c < 0:
        raise ValueError("negative scalar")
    x, y = _affine2(tuple(x1), tuple(y1), tuple(z1))
    r = _mul_bytes(_g2_bytes(x, y, False), c, True)
    if not any(r):
        return FQ2_ONE_TUPLE, FQ2_ONE_TUPLE, FQ2_ZERO_TUPLE, True
    v = _ints(r)
    return (v[0], v[1]), (v[2], v[3]), FQ2_ONE_TUPLE, False


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
