"""Ate pairing on BLS12-381 as VM programs.

What the reference computes (bls_py/fields_t.py:1091-1128, bls_py/pairing.py:51-92):
    e(P, Q) = f_{|x|,Q}(P) ^ ((q^12 - 1) / n)
with |x| = 0xd201000000010000 and NO conjugation for the negative curve parameter.  The
reference walks an affine R with dense Fq12 line values (one Fq12 inversion per line) and
raises to the 1268-bit exponent bit by bit.  Only the value AFTER the final exponentiation
is canonical, so this implementation is free to use

  * homogeneous projective R, denominator-free sparse lines scaled by w^3 and by Fq2
    factors (all killed by the final exponentiation),
  * the exact hard-part decomposition
        (q^4 - q^2 + 1)/n = ((a+1)^2 / 3) (q - a) (a^2 + q^2 - 1) + 1,   a = |x|,
    i.e. five 64-bit exponentiations instead of one 1268-bit one.
"""
from ..vm.builder import Program, Q
from .tower import F6, F12, f12_one, fp_inv_fermat

X_ABS = 0xd201000000010000
X_BITS = bin(X_ABS)[3:]                     # below the leading one, MSB first
Y_EXP = (X_ABS + 1) // 3                    # 0x460055555555aaab
assert 3 * Y_EXP == X_ABS + 1
N_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
assert (Q ** 4 - Q ** 2 + 1) // N_ORDER == Y_EXP * (X_ABS + 1) * (Q - X_ABS) * (X_ABS ** 2 + Q ** 2 - 1) + 1

BUF_P, BUF_Q, BUF_OUT = 0, 1, 2


def _mul12_xi(prog, c):
    """12 * xi * c  (3 b' with b' = 4 xi)"""
    t = c.mul_xi()
    t4 = t.dbl().dbl()
    return t4.dbl() + t4


def double_step(prog, r, xp, yp):
    """R <- 2R on E'(Fq2) in homogeneous projective coordinates; returns the tangent
    line at (the old) R evaluated at P as sparse coefficients (l0, l1, l4)."""
    x, y, z = r
    b = y.sqr()
    c = z.sqr()
    e = _mul12_xi(prog, c)              # 3 b' Z^2
    f = e.dbl() + e                     # 9 b' Z^2
    h = (y * z).dbl()                   # 2 Y Z
    xx = x.sqr()
    l0 = b - e                          # Y^2 - 3 b' Z^2
    l1 = -((xx.dbl() + xx) * xp)        # -3 X^2 xP
    l4 = h * yp                         # 2 Y Z yP
    x3 = ((x * y) * (b - f)).dbl()      # 2 X Y (Y^2 - 9 b' Z^2)
    e2 = e.sqr()
    e2_4 = e2.dbl().dbl()
    y3 = (b + f).sqr() - (e2_4.dbl() + e2_4)   # (B + F)^2 - 12 E^2
    z3 = (b * h).dbl().dbl()            # 8 Y^3 Z
    return (x3, y3, z3), (l0, l1, l4)


def add_step(prog, r, q, xp, yp):
    """R <- R + Q (Q affine); returns the chord line through R and Q at P."""
    x, y, z = r
    xq, yq = q
    theta = y - yq * z
    lam = x - xq * z
    l0 = theta * xq - lam * yq
    l1 = -(theta * xp)
    l4 = lam * yp
    c = theta.sqr()
    d = lam.sqr()
    e = lam * d
    f = z * c
    g = x * d
    h = e + f - g.dbl()
    x3 = lam * h
    y3 = theta * (g - h) - e * y
    z3 = z * e
    return (x3, y3, z3), (l0, l1, l4)


def miller_loop(prog, xp, yp, xq, yq):
    """f_{|x|,Q}(P) up to factors that the final exponentiation removes."""
    one = prog.const2((1, 0))
    r = (xq, yq, one)
    f = None
    for bit in X_BITS:
        r, (l0, l1, l4) = double_step(prog, r, xp, yp)
        if f is None:
            zero = prog.const2((0, 0))
            f = F12(F6(l0, l1, zero), F6(zero, l4, zero))
        else:
            f = f.sqr().mul_by_014(l0, l1, l4)
        if bit == "1":
            r, (l0, l1, l4) = add_step(prog, r, (xq, yq), xp, yp)
            f = f.mul_by_014(l0, l1, l4)
    return f


def _pow_bits(f, e, sqr):
    """f^e, MSB-first square and multiply"""
    acc = f
    for bit in bin(e)[3:]:
        acc = sqr(acc)
        if bit == "1":
            acc = acc * f
    return acc


def final_exponentiation(prog, f, cyclotomic=True):
    """f^((q^12-1)/n) with the exact exponent (reference: fields_t.py:1124-1128)."""
    fp_inv = fp_inv_fermat(prog)
    # easy part: f^((q^6 - 1)(q^2 + 1))
    t = f.conj() * f.inv(fp_inv)
    m = t.frob(prog, 2) * t
    # after the easy part m lies in the cyclotomic subgroup: cheap squarings, inverse = conj
    sqr = (lambda x: x.cyclotomic_sqr()) if cyclotomic else (lambda x: x.sqr())
    # hard part: m^(y (a+1) (q-a) (a^2+q^2-1)) * m
    t1 = _pow_bits(m, Y_EXP, sqr)
    t2 = _pow_bits(t1, X_ABS, sqr) * t1                    # ^(a+1)
    t3 = t2.frob(prog, 1) * _pow_bits(t2, X_ABS, sqr).conj()      # ^(q-a)
    t3a = _pow_bits(t3, X_ABS, sqr)
    t4 = _pow_bits(t3a, X_ABS, sqr) * t3.frob(prog, 2) * t3.conj()   # ^(a^2+q^2-1)
    return t4 * m


def load_g1(prog, buf):
    return prog.load1_be48(buf, 0), prog.load1_be48(buf, 48)


def load_g2(prog, buf):
    return prog.load2_be48(buf, 0), prog.load2_be48(buf, 96)


def store_f12(prog, buf, f):
    for k, c in enumerate(f.coeffs()):
        prog.store2_be48(buf, 96 * k, c)


def select_f12(prog, flag, a, b):
    return F12.from_coeffs([prog.sel2(flag, x, y) for x, y in zip(a.coeffs(), b.coeffs())])


def build_pairing():
    """one full pairing per item: P (96 B) x Q (192 B) -> Fq12 (576 B).
    A point given as all-zero coordinates is the point at infinity and pairs to 1, which
    is what the reference's blind computation on (0, 0) produces."""
    prog = Program("pairing")
    prog.begin_body()
    xp, yp = load_g1(prog, BUF_P)
    xq, yq = load_g2(prog, BUF_Q)
    inf = (xp.is_zero() & yp.is_zero()) | (xq.is_zero() & yq.is_zero())
    f = miller_loop(prog, xp, yp, xq, yq)
    e = final_exponentiation(prog, f)
    e = select_f12(prog, inf, f12_one(prog), e)
    store_f12(prog, BUF_OUT, e)
    return prog


def build_miller_only():
    """Miller loop only (debug / per-stage benchmarking); output is NOT canonical."""
    prog = Program("miller")
    prog.begin_body()
    xp, yp = load_g1(prog, BUF_P)
    xq, yq = load_g2(prog, BUF_Q)
    store_f12(prog, BUF_OUT, miller_loop(prog, xp, yp, xq, yq))
    return prog


def build_final_exp():
    """Fq12 (576 B) -> Fq12 (576 B)"""
    prog = Program("final_exp")
    prog.begin_body()
    f = F12.from_coeffs([prog.load2_be48(0, 96 * k) for k in range(6)])
    store_f12(prog, 1, final_exponentiation(prog, f))
    return prog
