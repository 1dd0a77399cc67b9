"""Single-function parity programs: every reference function of SURVEY.md rows a4 / a10 / a15 that the
hot programs only use *inside* a larger computation gets an entry point of its own here, so that the
reference's golden vectors for it can be replayed on the device (and on the host simulation).

  sw_encode_g2      ec.py:449-507 sw_encode for Fq2 (t = 0 -> infinity, w0 = 0 -> generator)
  f{2,6,12}_frob{i} fields_t.py:104-110, 203-212, 355-364 qi_pow, i = 0 .. level - 1
  f{1,2,6,12}_pow   fields_t.py:58-68, 92-101, 344-352 pow with a 384-bit exponent read from the item
  f{1,2}_sqrt       fields.py:199-205 Fq.modsqrt, fields.py:463-482 Fq2.modsqrt -- the reference's own root, bit
                    for bit (which of the two roots its exponentiations pick), and "no sqrt" as a flag
  g2_untwist, f12_twist, g2_psi   fields_t.py:936-943, 1018-1031 and ec.py:402-444
"""
from ..vm.builder import Program, Q
from .curve import Curve
from .hashg2 import HALF, PSI_CX, PSI_CY, _root_pow, sw_encode
from .tower import F6, F12, f2_inv, fp_inverter, f12_one

POW_BYTES = 48              # exponents of the pow programs: 48 big-endian bytes


def build_sw_encode():
    """buffers: 0 = t (Fq2, 96 B), 1 = out (affine point on the twist, 192 B; infinity = zero bytes)"""
    prog = Program("sw_encode_g2")
    prog.begin_body()
    t = prog.load2_be48(0, 0)
    one = prog.const2((1, 0))
    tt = t.sqr()
    w0 = tt + prog.const2((5, 4))                   # t^2 + b + 1
    t3 = tt.dbl() + tt
    s0 = prog.sel2(w0.is_zero(), one, w0)           # zero factors must not poison the shared inversion
    s1 = prog.sel2(t3.is_zero(), one, t3)
    inv = f2_inv(s0 * s1, fp_inverter(prog))
    x, y, inf = sw_encode(prog, t, inv * s1, inv * s0, w0)
    zero = prog.const2((0, 0))
    prog.store2_be48(1, 0, prog.sel2(inf, zero, x))
    prog.store2_be48(1, 96, prog.sel2(inf, zero, y))
    return prog


# ---- tower elements as flat lists of V2 (level >= 2) or one V1 (level 1) ---------------------------------
def _load(prog, buf, level):
    if level == 1:
        return prog.load1_be48(buf, 0)
    c = [prog.load2_be48(buf, 96 * k) for k in range(level // 2)]
    return c[0] if level == 2 else (F6(*c) if level == 6 else F12.from_coeffs(c))


def _coeffs(level, x):
    return [x] if level <= 2 else ([x.a0, x.a1, x.a2] if level == 6 else x.coeffs())


def _from_coeffs(level, c):
    return c[0] if level <= 2 else (F6(*c) if level == 6 else F12.from_coeffs(c))


def _store(prog, buf, level, x):
    if level == 1:
        prog.store1_be48(buf, 0, x)
    else:
        for k, c in enumerate(_coeffs(level, x)):
            prog.store2_be48(buf, 96 * k, c)


def build_frob(level, i):
    """x -> x^(q^i).  buffers: 0 = a, 1 = out"""
    def build():
        prog = Program("f%d_frob%d" % (level, i))
        prog.begin_body()
        a = _load(prog, 0, level)
        if level == 2:
            r = a.conj() if (i & 1) else a
        else:
            zero = prog.const2((0, 0))
            emb = F12(a, F6(zero, zero, zero)) if level == 6 else a
            r = emb.frob(prog, i) if i else emb
            if level == 6:
                r = r.c0
        _store(prog, 1, level, r)
        return prog
    return build


def build_pow(level):
    """x -> x^e, e = 48 big-endian bytes per item (left to right: square, multiply, select -- every lane runs
    the same code; the reference walks the bits from the other end, fields_t.py:344-352, same value).
    buffers: 0 = a, 1 = e, 2 = out"""
    def build():
        prog = Program("f%d_pow" % level)
        prog.begin_body()
        a = _load(prog, 0, level)
        if level == 1:
            acc = prog.const1(1)
        elif level == 2:
            acc = prog.const2((1, 0))
        else:
            one12 = f12_one(prog)
            acc = one12 if level == 12 else one12.c0
        for bit in range(8 * POW_BYTES - 1, -1, -1):
            if bit != 8 * POW_BYTES - 1:
                acc = acc.sqr()
            t = acc * a
            f = prog.flag_bit(1, bit, nbytes=POW_BYTES)
            if level == 1:
                acc = prog.sel1(f, t, acc)
            else:
                acc = _from_coeffs(level, [prog.sel2(f, x, y) for x, y in zip(_coeffs(level, t), _coeffs(level, acc))])
        _store(prog, 2, level, acc)
        return prog
    return build


def build_sqrt(level):
    """The reference's modsqrt.  buffers: 0 = a, 1 = root (48 / 96 B, zero where there is none), 2 = ok byte
    (0 where the reference raises ValueError('No sqrt exists')).
    Fq (fields.py:199-205): 0 -> 0; the root is a^((q+1)/4), i.e. the one that is itself a square.
    Fq2 (fields.py:463-482, 'complex method'): alpha = norm^((q+1)/4); delta = (a0 + alpha)/2, or (a0 - alpha)/2 when
    that is no square; x0 = delta^((q+1)/4), x1 = a1 / (2 x0).  An element with a1 = 0 is handed to the Fq root
    (the reference then returns an Fq object; here the root sits in c0 and c1 = 0)."""
    def build():
        prog = Program("f%d_sqrt" % level)
        prog.begin_body()
        one = prog.const1(1)

        def fq_root(z):
            c = _root_pow(prog, z)                   # z^((q-3)/4)
            r = c * z                                # z^((q+1)/4)
            ok = (r * c).eq(one) | z.is_zero()       # z^((q-1)/2) == 1, or z == 0 (then r = 0)
            return r, ok

        if level == 1:
            a = prog.load1_be48(0, 0)
            r, ok = fq_root(a)
            prog.store1_be48(1, 0, prog.sel1(ok, r, prog.const1(0)))
            prog.store_flag(2, 0, ok)
            return prog
        u = prog.load2_be48(0, 0)
        real = u.c1.is_zero()
        r_real, ok_real = fq_root(u.c0)
        n = u.c0.sqr() + u.c1.sqr()
        alpha, ok_n = fq_root(n)                     # norm 0 only for u = 0, which is the real case
        half = prog.const1(HALF)
        delta = (u.c0 + alpha) * half
        c1 = _root_pow(prog, delta)
        s = c1 * delta                               # s^2 = +-delta
        is_sq = (s * c1).eq(one)
        hlf = (u.c1 * c1) * half
        # delta a square: x0 = s (= delta^((q+1)/4) exactly), x1 = a1 / (2 s) = hlf (c1 = 1 / s).
        # delta no square: c1 = -1 / s, hlf^2 = (a0 - alpha) / 2, and the reference's x0 is whichever of +-hlf is
        # itself a square (that is what an exponentiation by (q+1)/4 returns); x1 = a1 / (2 x0) = -+s.
        hsq = hlf.is_square()
        x0 = prog.sel1(is_sq, s, prog.sel1(hsq, hlf, -hlf))
        x1 = prog.sel1(is_sq, hlf, prog.sel1(hsq, -s, s))
        zero = prog.const1(0)
        ok = (real & ok_real) | (~real & ok_n)
        y0 = prog.sel1(real, r_real, x0)
        y1 = prog.sel1(real, zero, x1)
        prog.store1_be48(1, 0, prog.sel1(ok, y0, zero))
        prog.store1_be48(1, 48, prog.sel1(ok, y1, zero))
        prog.store_flag(2, 0, ok)
        return prog
    return build


# ---- twist maps (fields_t.py:936-1031): Fq12 = Fq2[w] / (w^6 - xi); flat coefficient p multiplies w^W_OF[p] ---
W_OF = [0, 2, 4, 1, 3, 5]
XI_INV = ((Q + 1) // 2, (Q - 1) // 2)            # 1 / (1 + u) = (1 - u) / 2: the reference's tw1 = 1/2, tw2 = -1/2
assert ((XI_INV[0] - XI_INV[1]) % Q, (XI_INV[0] + XI_INV[1]) % Q) == (1, 0)


def _mul_w_power(prog, coeffs, k):
    """(sum_p c_p w^W_OF[p]) * w^k, k in 0..5, with w^6 = xi"""
    out = [None] * 6
    for p in range(6):
        e = W_OF[p] + k
        c = coeffs[p]
        if e >= 6:
            e -= 6
            c = c.mul_xi()
        out[W_OF.index(e)] = c
    return out


def build_untwist():
    """fq2_untwist (fields_t.py:936-943): (x, y) on E'(Fq2) -> (x / w^2, y / w^3) on E(Fq12).
    1 / w^2 = w^4 / xi and 1 / w^3 = w^3 / xi, so the images are sparse: x / xi at w^4, y / xi at w^3.
    buffers: 0 = affine twist point (192 B), 1 = x' || y' (2 x 576 B)"""
    prog = Program("g2_untwist")
    prog.begin_body()
    x = prog.load2_be48(0, 0)
    y = prog.load2_be48(0, 96)
    xi_inv = prog.const2(XI_INV)
    zero = prog.const2((0, 0))
    xs = [zero] * 6
    ys = [zero] * 6
    xs[W_OF.index(4)] = x * xi_inv
    ys[W_OF.index(3)] = y * xi_inv
    for k in range(6):
        prog.store2_be48(1, 96 * k, xs[k])
        prog.store2_be48(1, 576 + 96 * k, ys[k])
    return prog


def build_twist12():
    """fq12_twist (fields_t.py:1018-1031): (x, y) with Fq12 coordinates -> (x w^2, y w^3).
    buffers: 0 = x || y (2 x 576 B), 1 = x w^2 || y w^3"""
    prog = Program("f12_twist")
    prog.begin_body()
    x = [prog.load2_be48(0, 96 * k) for k in range(6)]
    y = [prog.load2_be48(0, 576 + 96 * k) for k in range(6)]
    xo = _mul_w_power(prog, x, 2)
    yo = _mul_w_power(prog, y, 3)
    for k in range(6):
        prog.store2_be48(1, 96 * k, xo[k])
        prog.store2_be48(1, 576 + 96 * k, yo[k])
    return prog


def build_psi():
    """ec.psi (ec.py:440-444: untwist, Frobenius, twist) on the twist's own coordinates:
    (conj(x) / xi^((q-1)/3), conj(y) / xi^((q-1)/2)).  buffers: 0 = affine point (192 B), 1 = out (192 B)"""
    prog = Program("g2_psi")
    prog.begin_body()
    x = prog.load2_be48(0, 0)
    y = prog.load2_be48(0, 96)
    prog.store2_be48(1, 0, x.conj() * prog.const2(PSI_CX))
    prog.store2_be48(1, 96, y.conj() * prog.const2(PSI_CY))
    return prog
