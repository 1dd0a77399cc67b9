"""Fq6 / Fq12 tower arithmetic expressed as VM programs over Fq2 values.

Tower (same as the reference, bls_py/fields.py:322/486/625):
  Fq6  = Fq2[v]/(v^3 - xi),  xi = 1 + u
  Fq12 = Fq6[w]/(w^2 - v)
The reference multiplies with hand-expanded schoolbook formulas (36 / 144 integer products,
bls_py/fields_t.py:293-318, 503-554); here Karatsuba at every level (18 Fq2 products per
Fq12 product) -- results are canonical field elements either way.
"""
from ..vm.builder import Q

XI = (1, 1)


def f2_pow_int(a, e):
    """plain-int Fq2 power for constant generation"""
    r = (1, 0)
    while e:
        if e & 1:
            r = ((r[0] * a[0] - r[1] * a[1]) % Q, (r[0] * a[1] + r[1] * a[0]) % Q)
        a = ((a[0] * a[0] - a[1] * a[1]) % Q, 2 * a[0] * a[1] % Q)
        e >>= 1
    return r


def frob_gamma(i):
    """gamma[k] = xi^(k (q^i - 1) / 6): (c w^k)^(q^i) = conj^i(c) gamma[k] w^k
    (the values the reference tabulates at fields_t.py:1133-1216)"""
    g = f2_pow_int(XI, (Q ** i - 1) // 6)
    tab = [(1, 0)]
    for _ in range(5):
        t = tab[-1]
        tab.append(((t[0] * g[0] - t[1] * g[1]) % Q, (t[0] * g[1] + t[1] * g[0]) % Q))
    return tab


class F6:
    """a0 + a1 v + a2 v^2"""
    __slots__ = ("a0", "a1", "a2")

    def __init__(self, a0, a1, a2):
        self.a0, self.a1, self.a2 = a0, a1, a2

    def __add__(self, o):
        return F6(self.a0 + o.a0, self.a1 + o.a1, self.a2 + o.a2)

    def __sub__(self, o):
        return F6(self.a0 - o.a0, self.a1 - o.a1, self.a2 - o.a2)

    def __neg__(self):
        return F6(-self.a0, -self.a1, -self.a2)

    def dbl(self):
        return F6(self.a0.dbl(), self.a1.dbl(), self.a2.dbl())

    def mul_v(self):
        return F6(self.a2.mul_xi(), self.a0, self.a1)

    def __mul__(self, o):
        a0, a1, a2, b0, b1, b2 = self.a0, self.a1, self.a2, o.a0, o.a1, o.a2
        v0, v1, v2 = a0 * b0, a1 * b1, a2 * b2
        c0 = v0 + ((a1 + a2) * (b1 + b2) - v1 - v2).mul_xi()
        c1 = (a0 + a1) * (b0 + b1) - v0 - v1 + v2.mul_xi()
        c2 = (a0 + a2) * (b0 + b2) - v0 - v2 + v1
        return F6(c0, c1, c2)

    def sqr(self):
        a0, a1, a2 = self.a0, self.a1, self.a2
        s0 = a0.sqr()
        s1 = (a0 * a1).dbl()
        s2 = (a0 - a1 + a2).sqr()
        s3 = (a1 * a2).dbl()
        s4 = a2.sqr()
        return F6(s0 + s3.mul_xi(), s1 + s4.mul_xi(), s1 + s2 + s3 - s0 - s4)

    def mul_by_01(self, c0, c1):
        """times (c0 + c1 v)"""
        a0, a1, a2 = self.a0, self.a1, self.a2
        t0, t1 = a0 * c0, a1 * c1
        r0 = t0 + (a2 * c1).mul_xi()
        r1 = (a0 + a1) * (c0 + c1) - t0 - t1
        r2 = t1 + a2 * c0
        return F6(r0, r1, r2)

    def mul_by_12(self, c1, c2):
        """times (c1 v + c2 v^2): 5 Fq2 products"""
        a0, a1, a2 = self.a0, self.a1, self.a2
        t1, t2 = a1 * c1, a2 * c2
        r0 = ((a1 + a2) * (c1 + c2) - t1 - t2).mul_xi()      # xi (a1 c2 + a2 c1)
        r1 = a0 * c1 + t2.mul_xi()
        r2 = a0 * c2 + t1
        return F6(r0, r1, r2)

    def mul_by_1(self, c1):
        """times c1 v"""
        return F6((self.a2 * c1).mul_xi(), self.a0 * c1, self.a1 * c1)

    def mul_fp2(self, c):
        return F6(self.a0 * c, self.a1 * c, self.a2 * c)

    def inv(self, fp_inv):
        """norm down to Fq2 (as bls_py/fields_t.py:170-184), one Fq inversion via fp_inv"""
        a0, a1, a2 = self.a0, self.a1, self.a2
        g0 = a0.sqr() - (a1 * a2).mul_xi()
        g1 = a2.sqr().mul_xi() - a0 * a1
        g2 = a1.sqr() - a0 * a2
        n = a0 * g0 + (a2 * g1 + a1 * g2).mul_xi()
        t = f2_inv(n, fp_inv)
        return F6(g0 * t, g1 * t, g2 * t)


def f2_inv(a, fp_inv):
    """conj(a) / (c0^2 + c1^2)  (bls_py/fields_t.py:81-85)"""
    n = a.c0.sqr() + a.c1.sqr()
    t = fp_inv(n)
    return a.conj() * t


class F12:
    """c0 + c1 w, c0, c1 in Fq6"""
    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1):
        self.c0, self.c1 = c0, c1

    def coeffs(self):
        """the six Fq2 coefficients in the reference's flat (ZT) order"""
        return [self.c0.a0, self.c0.a1, self.c0.a2, self.c1.a0, self.c1.a1, self.c1.a2]

    @staticmethod
    def from_coeffs(c):
        return F12(F6(c[0], c[1], c[2]), F6(c[3], c[4], c[5]))

    def __add__(self, o):
        return F12(self.c0 + o.c0, self.c1 + o.c1)

    def __sub__(self, o):
        return F12(self.c0 - o.c0, self.c1 - o.c1)

    def __neg__(self):
        return F12(-self.c0, -self.c1)

    def __mul__(self, o):
        t0 = self.c0 * o.c0
        t1 = self.c1 * o.c1
        c1 = (self.c0 + self.c1) * (o.c0 + o.c1) - t0 - t1
        return F12(t0 + t1.mul_v(), c1)

    def sqr(self):
        a0, a1 = self.c0, self.c1
        t = a0 * a1
        c0 = (a0 + a1) * (a0 + a1.mul_v()) - t - t.mul_v()
        return F12(c0, t.dbl())

    def sqr_x2(self):
        """2 x^2 -- for callers that may scale by an Fq constant (the Miller loop: the final
        exponentiation removes it).  Chung-Hasan SQR3 over Fq4: with x = A + B w + C w^2, A, B, C in
        Fq4 = Fq2[s], s = w^3 (the z_i pairs of cyclotomic_sqr),
            S0 = A^2, S1 = (A + B + C)^2, S2 = (A - B + C)^2, P = B C, S4 = C^2,
            2 c0 = 2 S0 + 4 s P,   2 c1 = S1 - S2 - 4 P + 2 s S4,   2 c2 = S1 + S2 - 2 S0 - 2 S4:
        four Fq4 squarings (3 Fq2 squarings each) and one Fq4 product = 33 M instead of the 36 M of
        the complex method; the halving the exact formula needs is what the factor 2 avoids."""
        def sqr4(x):
            t0, t1 = x[0].sqr(), x[1].sqr()
            return (t1.mul_xi() + t0, (x[0] + x[1]).sqr() - t0 - t1)

        def mul4(x, y):
            p0, p1 = x[0] * y[0], x[1] * y[1]
            return (p1.mul_xi() + p0, (x[0] + x[1]) * (y[0] + y[1]) - p0 - p1)

        def add4(x, y):
            return (x[0] + y[0], x[1] + y[1])

        def sub4(x, y):
            return (x[0] - y[0], x[1] - y[1])

        def dbl4(x):
            return (x[0].dbl(), x[1].dbl())

        def times_s(x):
            return (x[1].mul_xi(), x[0])

        a = (self.c0.a0, self.c1.a1)
        b = (self.c1.a0, self.c0.a2)
        c = (self.c0.a1, self.c1.a2)
        s0, s4 = sqr4(a), sqr4(c)
        ac = add4(a, c)
        s1, s2 = sqr4(add4(ac, b)), sqr4(sub4(ac, b))
        p2 = dbl4(mul4(b, c))                               # 2 P
        r0 = dbl4(add4(s0, times_s(p2)))                    # 2 S0 + 4 s P
        r1 = add4(sub4(sub4(s1, s2), dbl4(p2)), dbl4(times_s(s4)))
        r2 = sub4(add4(s1, s2), dbl4(add4(s0, s4)))
        return F12(F6(r0[0], r2[0], r1[1]), F6(r1[0], r0[1], r2[1]))

    def conj(self):
        """x^(q^6)"""
        return F12(self.c0, -self.c1)

    def cyclotomic_sqr(self):
        """x^2 for x in the cyclotomic subgroup (x^(q^6+1) = 1, true after the easy part of
        the final exponentiation): Granger-Scott squaring, 9 Fq2 squarings instead of 12
        Fq2 products.  Fq12 is seen as three Fq4 = Fq2[s]/(s^2 - xi) components."""
        def fp4_sqr(a, b):
            t0, t1 = a.sqr(), b.sqr()
            return t1.mul_xi() + t0, (a + b).sqr() - t0 - t1

        z0, z4, z3 = self.c0.a0, self.c0.a1, self.c0.a2
        z2, z1, z5 = self.c1.a0, self.c1.a1, self.c1.a2
        prog = z0.prog
        t0, t1 = fp4_sqr(z0, z1)
        r0 = prog.tri2(t0, z0, False)          # 3 t0 - 2 z0
        r1 = prog.tri2(t1, z1, True)           # 3 t1 + 2 z1
        t0, t1 = fp4_sqr(z2, z3)
        t2, t3 = fp4_sqr(z4, z5)
        r4 = prog.tri2(t0, z4, False)
        r5 = prog.tri2(t1, z5, True)
        r2 = prog.tri2(t3.mul_xi(), z2, True)
        r3 = prog.tri2(t2, z3, False)
        return F12(F6(r0, r4, r3), F6(r2, r1, r5))

    def mul_by_014(self, l0, l1, l4):
        """times the sparse line value (l0 + l1 v) + (l4 v) w"""
        t0 = self.c0.mul_by_01(l0, l1)
        t1 = self.c1.mul_by_1(l4)
        c1 = (self.c0 + self.c1).mul_by_01(l0, l1 + l4) - t0 - t1
        return F12(t0 + t1.mul_v(), c1)

    def mul_by_014_pair(self, la, lb):
        """times the product of TWO sparse line values (a0 + a1 v) + (a4 v) w and (b0 + b1 v) + (b4 v) w
        (two Miller loops sharing their squarings): the lines are multiplied first, 6 Fq2 products,
        using v^2 w^2 = v^3 = xi --
            (a0 b0 + xi a4 b4) + (a0 b1 + a1 b0) v + a1 b1 v^2  +  ((a0 b4 + a4 b0) v + (a1 b4 + a4 b1) v^2) w
        -- and the result, whose w-part has no constant term, costs 17 instead of 18 products against
        self: 23 Fq2 products in all instead of 2 x 13."""
        a0, a1, a4 = la
        b0, b1, b4 = lb
        p00, p11, p44 = a0 * b0, a1 * b1, a4 * b4
        m01 = (a0 + a1) * (b0 + b1) - p00 - p11
        m04 = (a0 + a4) * (b0 + b4) - p00 - p44
        m14 = (a1 + a4) * (b1 + b4) - p11 - p44
        c0 = F6(p00 + p44.mul_xi(), m01, p11)
        d1, d2 = m04, m14                                   # c1 = d1 v + d2 v^2
        t0 = self.c0 * c0
        t1 = self.c1.mul_by_12(d1, d2)
        s = self.c0 + self.c1
        mid = s * F6(c0.a0, c0.a1 + d1, c0.a2 + d2) - t0 - t1
        return F12(t0 + t1.mul_v(), mid)

    def inv(self, fp_inv):
        """(c0 - c1 w) / (c0^2 - v c1^2)  (bls_py/fields_t.py:328-337)"""
        t = (self.c0.sqr() - self.c1.sqr().mul_v()).inv(fp_inv)
        return F12(self.c0 * t, -(self.c1 * t))

    def frob(self, prog, i):
        """x^(q^i) for i in (1, 2, 3)  (bls_py/fields_t.py:355-364)"""
        gam = frob_gamma(i)
        c = self.coeffs()
        w_of = [0, 2, 4, 1, 3, 5]           # flat coefficient p multiplies w^(w_of[p])
        out = []
        for p in range(6):
            x = c[p].conj() if (i & 1) else c[p]
            g = gam[w_of[p]]
            if g == (1, 0):
                out.append(x)
            elif g[1] == 0:
                out.append(x * prog.const1(g[0]))
            else:
                out.append(x * prog.const2(g))
        return F12.from_coeffs(out)


class CompressedCyc:
    """Karabina's compressed form of an element of the cyclotomic subgroup: with Fq12 seen as
    A + B w + C w^2 over Fq4 = Fq2[s], s = w^3 (A = z0 + z1 s, B = z2 + z3 s, C = z4 + z5 s, the
    z_i of F12.cyclotomic_sqr), Granger-Scott squaring sends B -> 3 s C^2 + 2 conj(B) and
    C -> 3 B^2 - 2 conj(C): B and C never look at A.  A run of squarings therefore needs 6 instead of
    9 Fq2 squarings (12 M instead of 18 M) and four live values instead of six; A is recovered from
    the subgroup relations when the run is over (decompress_many)."""
    __slots__ = ("z2", "z3", "z4", "z5")

    def __init__(self, z2, z3, z4, z5):
        self.z2, self.z3, self.z4, self.z5 = z2, z3, z4, z5

    @staticmethod
    def of(f):
        return CompressedCyc(f.c1.a0, f.c0.a2, f.c0.a1, f.c1.a2)

    def sqr(self):
        def fp4_sqr(a, b):
            t0, t1 = a.sqr(), b.sqr()
            return t1.mul_xi() + t0, (a + b).sqr() - t0 - t1

        prog = self.z2.prog
        t0, t1 = fp4_sqr(self.z2, self.z3)
        t2, t3 = fp4_sqr(self.z4, self.z5)
        return CompressedCyc(prog.tri2(t3.mul_xi(), self.z2, True), prog.tri2(t2, self.z3, False),
                             prog.tri2(t0, self.z4, False), prog.tri2(t1, self.z5, True))


def decompress_many(prog, items, fp_inv):
    """[CompressedCyc] -> [F12] with ONE shared inversion.  From conj6(x) x = 1 and the
    Granger-Scott identities (A B = s C^2 + conj(B), ...):
        z1 = (3 z4^2 + xi z5^2 - 2 z3) / (4 z2)          (z2 != 0)
        z1 = 2 z4 z5 / z3                                (z2 == 0)
        z0 = xi (2 z1^2 + z2 z5 - 3 z3 z4) + 1
    z2 = z3 = 0 only happens for x = 1 (no other element of the subgroup has B = 0), where the
    numerator is 0 as well and the formulas give z1 = 0, z0 = 1 whatever stands in for 1 / 0;
    zero denominators are replaced by 1 so that they cannot poison the shared inversion."""
    one = prog.const2((1, 0))
    nums, dens = [], []
    for c in items:
        s4, s5 = c.z4.sqr(), c.z5.sqr()
        num_a = prog.tri2(s4, c.z3, False) + s5.mul_xi()
        num_b = (c.z4 + c.z5).sqr() - s4 - s5
        z2_zero = c.z2.is_zero()
        nums.append(prog.sel2(z2_zero, num_b, num_a))
        den = prog.sel2(z2_zero, c.z3, c.z2.dbl().dbl())
        dens.append(prog.sel2(den.is_zero(), one, den))
    prefix = [dens[0]]
    for d in dens[1:]:
        prefix.append(prefix[-1] * d)
    inv = f2_inv(prefix[-1], fp_inv)
    invs = [None] * len(dens)
    for i in range(len(dens) - 1, 0, -1):
        invs[i] = inv * prefix[i - 1]
        inv = inv * dens[i]
    invs[0] = inv
    out = []
    for c, num, di in zip(items, nums, invs):
        z1 = num * di
        u = c.z3 * c.z4
        t = z1.sqr().dbl() + c.z2 * c.z5 - (u.dbl() + u)
        z0 = t.mul_xi() + one
        out.append(F12(F6(z0, c.z4, c.z3), F6(c.z2, z1, c.z5)))
    return out


def f12_one(prog):
    one = prog.const2((1, 0))
    zero = prog.const2((0, 0))
    return F12(F6(one, zero, zero), F6(zero, zero, zero))


def fp_pow_chain(prog, a, e, window=4):
    """a^e for a fixed exponent e > 0: left-to-right sliding window over Fq cells.
    The odd-power table is packed two entries per Fq2 cell pair."""
    if e == 1:
        return a
    nt = 1 << (window - 1)
    bits = bin(e)[2:]
    # which odd powers are actually needed
    i, need, plan = 0, set(), []
    while i < len(bits):
        if bits[i] == "0":
            plan.append(("s", 1))
            i += 1
            continue
        j = min(i + window, len(bits))
        while bits[j - 1] == "0":
            j -= 1
        val = int(bits[i:j], 2)
        plan.append(("m", j - i, val))
        need.add(val)
        i = j
    top = max(need)
    table = {1: a}
    if top > 1:
        a2 = a.sqr()
        prev = a
        packs = []
        pending = None
        for k in range(3, top + 1, 2):
            prev = prev * a2
            if pending is None:
                pending = (k, prev)
            else:
                pk = prog.pack(pending[1], prev)
                table[pending[0]] = pk.c0
                table[k] = pk.c1
                pending = None
        if pending is not None:
            table[pending[0]] = pending[1]
    acc = None
    for step in plan:
        if step[0] == "s":
            if acc is not None:
                acc = acc.sqr()
        else:
            _, n, val = step
            if acc is None:
                acc = table[val]
            else:
                for _ in range(n):
                    acc = acc.sqr()
                acc = acc * table[val]
    return acc


def fp_inverter(prog):
    """returns a function a -> 1 / a in Fq (0 -> 0, like bls_py/fields_t.py:47-55): the INV1
    instruction -- a binary almost-inverse (shifts and subtractions on the ALU pipe, beside the other
    warps' multiplications) and two Montgomery products, instead of the 465 products of a^(q-2)"""
    return lambda a: a.inv()
