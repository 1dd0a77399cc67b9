#!/usr/bin/env python3
"""tests/golden/seam_kat.json: inputs and outputs of the functions behind the reference's plugin seam
(bls_py/fields_t.py:1218-1265), recorded from the LIVE reference's pure-Python definitions in the development
container.  Tuples as lists of hex ints.  Usage: python tools/gen_seam_golden.py"""
import json
import logging
import os
import random
import sys

logging.disable(logging.CRITICAL)
sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from bls_py import fields_t as T                       # noqa: E402
from bls_py import ec as rec                           # noqa: E402

Q = rec.default_ec.q
rnd = random.Random(0x5EA)


def hx(v):
    if isinstance(v, bool):
        return v
    if isinstance(v, int):
        return hex(v)
    return [hx(x) for x in v]


def el(level):
    return rnd.randrange(Q) if level == 1 else tuple(rnd.randrange(Q) for _ in range(level))


cases = []


def rec_call(name, *args):
    out = getattr(T, name)(*args)
    cases.append({"fn": name, "args": hx(list(args)), "out": hx(out)})
    return out


for _ in range(2):
    a, x = el(1), el(1)
    rec_call("fq_invert", Q, a)
    rec_call("fq_floordiv", Q, a, x)
    rec_call("fq_pow", Q, a, rnd.getrandbits(300))
    a2, x2 = el(2), el(2)
    rec_call("fq2_invert", a2)
    rec_call("fq2_floordiv", a2, x2)
    rec_call("fq2_pow", a2, rnd.getrandbits(260))
    rec_call("fq2_qi_pow", a2, 1)
    a6, x6 = el(6), el(6)
    rec_call("fq6_invert", a6)
    rec_call("fq6_floordiv", a6, x6)
    rec_call("fq6_qi_pow", a6, rnd.randrange(1, 6))
    rec_call("fq6_add", a6, x6)
    rec_call("fq6_mul", a6, x6)
    a12, x12 = el(12), el(12)
    rec_call("fq12_invert", a12)
    rec_call("fq12_floordiv", a12, x12)
    rec_call("fq12_qi_pow", a12, rnd.randrange(1, 12))
    rec_call("fq12_pow", a12, rnd.getrandbits(200))
    rec_call("fq12_mul_fq", a12, el(1))
    rec_call("fq12_add", a12, x12)
    rec_call("fq12_mul", a12, x12)

g1, g2 = rec.generator_Fq(), rec.generator_Fq2()


def jac2(k, z):
    p = (g2.to_jacobian() * k).to_affine()
    x, y = p.x.ZT, p.y.ZT
    zz = T.fq2_mul(z, z)
    return T.fq2_mul(x, zz), T.fq2_mul(y, T.fq2_mul(zz, z)), z


def jac1(k, z):
    p = (g1.to_jacobian() * k).to_affine()
    return p.x.Z * z * z % Q, p.y.Z * z * z * z % Q, z


def norm(name, out):
    """Jacobian results are compared after normalisation (the same point, not the same triple)"""
    cases[-1]["norm"] = True


for k1, k2 in ((5, 9), (1234567, 7654321)):
    z1, z2 = el(2), el(2)
    X1, Y1, Z1 = jac2(k1, z1)
    X2, Y2, Z2 = jac2(k2, z2)
    rec_call("fq2_to_affine", X1, Y1, Z1, False)
    rec_call("fq2_add_points_jacobian", X1, Y1, Z1, False, X2, Y2, Z2, False); norm(*[None] * 2)
    rec_call("fq2_add_points_jacobian", X1, Y1, Z1, False, X1, Y1, Z1, False); norm(*[None] * 2)
    rec_call("fq2_double_point_jacobian", X1, Y1, Z1); norm(*[None] * 2)
    rec_call("fq2_scalar_mult_jacobian", rnd.getrandbits(254), X1, Y1, Z1, False); norm(*[None] * 2)
    p1 = (g2.to_jacobian() * k1).to_affine()
    p2 = (g2.to_jacobian() * k2).to_affine()
    rec_call("fq2_double_point", p1.x.ZT, p1.y.ZT, False)
    rec_call("fq2_add_points", p1.x.ZT, p1.y.ZT, False, p2.x.ZT, p2.y.ZT, False)
    rec_call("fq2_untwist", p1.x.ZT, p1.y.ZT)
    q1 = (g1.to_jacobian() * k2).to_affine()
    rec_call("fq2_double_line_eval", p1.x.ZT, p1.y.ZT, q1.x.Z, q1.y.Z)
    rec_call("fq2_add_line_eval", p1.x.ZT, p1.y.ZT, p2.x.ZT, p2.y.ZT, q1.x.Z, q1.y.Z)
    a1, b1 = jac1(k1, el(1)), jac1(k2, el(1))
    rec_call("fq_add_points_jacobian", *a1, False, *b1, False); norm(*[None] * 2)
    rec_call("fq_double_point_jacobian", *a1); norm(*[None] * 2)
    # Fq12 coordinates: an untwisted point
    ux, uy = T.fq2_untwist(p1.x.ZT, p1.y.ZT)
    vx, vy = T.fq2_untwist(p2.x.ZT, p2.y.ZT)
    one12 = (1,) + (0,) * 11
    rec_call("fq12_double_point_jacobian", ux, uy, one12); norm(*[None] * 2)
    rec_call("fq12_add_points_jacobian", ux, uy, one12, False, vx, vy, one12, False); norm(*[None] * 2)

path = os.path.join(ROOT, "tests", "golden", "seam_kat.json")
with open(path, "w") as fh:
    json.dump({"source": "bls_py/fields_t.py pure-Python definitions, live reference", "cases": cases}, fh, indent=0,
              separators=(",", ":"))
    fh.write("\n")
print("wrote", path, os.path.getsize(path), "bytes,", len(cases), "cases")
