"""CPU oracle for the b200-bls hot path -- TEST INFRASTRUCTURE ONLY.

This module is a pure-Python restatement of the algorithm that the reference
(zebra-lucky/python-bls, ``bls_py``) runs on its CPU path.  It exists so that

  * ``tests/``                      can check the CUDA library bit-for-bit,
  * ``__graft_entry__.smoke()``     can check one small GPU call,
  * ``bench.py``'s CPU-baseline leg can time "what the reference does" on the
    GPU box's host cores (the reference itself lives under ``/root/reference``
    in the development container and does not travel).

Nothing in the product path (``python-bls_b200/``) imports it; the product has
no CPU fallback.

Parity status: PINNED.  ``tools/gen_golden.py`` imports the real reference in
the development container, replays the reference's own known answers
(``bls_py/tdata.py`` through ``bls_py/tests.py:434-1036``, the signature and
aggregate vectors of ``tests.py:113-198``), adds live outputs of the reference
for pairings / hashing / curve ops, and writes them to ``tests/golden/*.json``.
``tests/test_oracle_*.py`` check this file against every one of those vectors.

Every function cites the reference lines (relative to ``/root/reference``) whose
behaviour it restates.  The *algorithms* are the reference's (affine Miller loop
with Fq12 slopes, square-and-multiply final exponentiation, Fouque-Tibouchi
hashing with exception-driven branch choice); the code is written from scratch
on one generic "polynomial over Fq2" representation instead of the reference's
per-level hand-expanded formulas.

Representation (same as the reference's ``ZT`` tuples, fields.py:322/486/625):
  Fq   : int in [0, Q)
  Fq2  : (c0, c1)            c0 + c1*u,            u^2 = -1
  Fq6  : 6 ints  = 3 x Fq2   a0 + a1*v + a2*v^2,   v^3 = u + 1
  Fq12 : 12 ints = 2 x Fq6   b0 + b1*w,            w^2 = v
"""
import hashlib

# ---------------------------------------------------------------------------
# constants (bls_py/bls12381.py:7-44, fields_t.py:20-25)
# ---------------------------------------------------------------------------
Q = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
        "1eabfffeb153ffffb9feffffffffaaab", 16)
N = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X_ABS = 0xd201000000010000          # |x|, the (negative) BLS parameter
G1 = (int("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
          "6c55e83ff97a1aeffb3af00adb22c6bb", 16),
      int("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3ed"
          "d03cc744a2888ae40caa232946c5e7e1", 16), False)
G2 = ((int("024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d177"
           "0bac0326a805bbefd48056c8c121bdb8", 16),
       int("13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
           "334cf11213945d57e5ac7d055d042b7e", 16)),
      (int("0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c"
           "923ac9cc3baca289e193548608b82801", 16),
       int("0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab"
           "3f370d275cec1da1aaa9075ff05f79be", 16)), False)
B1 = 4                               # E : y^2 = x^3 + 4
B2 = (4, 4)                          # E': y^2 = x^3 + 4(u+1)
SQRT_M3 = 1586958781458431025242759403266842894121773480562120986020912974854563298150952611241517463240701
SQRT_M3_M1_O2 = 793479390729215512621379701633421447060886740281060493010456487427281649075476305620758731620350
FINAL_EXP_HARD = (Q ** 4 - Q ** 2 + 1) // N   # fields_t.py:44


# ---------------------------------------------------------------------------
# Fq   (fields.py:35-243, fields_t.py:47-72)
# ---------------------------------------------------------------------------
def fq_inv(a):
    """Extended Euclid, returns 0 for a == 0 (fields_t.py:47-55)."""
    a %= Q
    r0, r1, s0, s1 = Q, a, 0, 1
    while r1:
        k = r0 // r1
        r0, r1 = r1, r0 - k * r1
        s0, s1 = s1, s0 - k * s1
    # for a == 0 the loop never runs and s0 == 0
    return s0 % Q if a else 0


def fq_is_square(a):
    return a % Q == 0 or pow(a, (Q - 1) // 2, Q) == 1


def fq_sqrt(a):
    """fields.py:199-205: Q % 4 == 3 -> a^((Q+1)/4); raises when no root."""
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        raise ValueError("No sqrt exists")
    return pow(a, (Q + 1) // 4, Q)


# ---------------------------------------------------------------------------
# Fq2  (fields.py:321-482, fields_t.py:75-161)
# ---------------------------------------------------------------------------
F2_ZERO, F2_ONE = (0, 0), (1, 0)


def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a):
    return (-a[0] % Q, -a[1] % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_scale(a, k):
    return (a[0] * k % Q, a[1] * k % Q)


def f2_mul_xi(a):
    """times the Fq6 non-residue xi = 1 + u (fields_t.py:113-116)."""
    return ((a[0] - a[1]) % Q, (a[0] + a[1]) % Q)


def f2_inv(a):
    """conj(a) / norm(a) (fields_t.py:81-85)."""
    t = fq_inv(a[0] * a[0] + a[1] * a[1])
    return (a[0] * t % Q, -a[1] * t % Q)


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_mul(a, a)
        e >>= 1
    return r


def f2_sqrt(a):
    """'Complex method' root (fields.py:463-482).

    Mirrors the reference's quirks: a purely real input is delegated to the Fq
    root (and therefore comes back as a plain int, which the caller treats as a
    failure), and 'no root' is a ValueError.
    """
    a0, a1 = a
    if a1 == 0:
        return fq_sqrt(a0)                       # an int, not a pair!
    alpha = (a0 * a0 + a1 * a1) % Q
    if pow(alpha, (Q - 1) // 2, Q) == Q - 1:
        raise ValueError("No sqrt exists")
    alpha = fq_sqrt(alpha)
    half = fq_inv(2)
    delta = (a0 + alpha) * half % Q
    if pow(delta, (Q - 1) // 2, Q) == Q - 1:
        delta = (a0 - alpha) * half % Q
    x0 = fq_sqrt(delta)
    x1 = a1 * fq_inv(2 * x0) % Q
    return (x0, x1)


# ---------------------------------------------------------------------------
# polynomial extensions over Fq2:  Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq2[w]/(w^6 - xi)
# A flat ZT tuple of level 6 is (a0,a1,a2) in v; of level 12 it is
# ((b00,b01,b02),(b10,b11,b12)) with element sum b_ij v^j w^i = sum b_ij w^(2j+i).
# ---------------------------------------------------------------------------
def _pairs(flat):
    return [(flat[i], flat[i + 1]) for i in range(0, len(flat), 2)]


def _flat(pairs):
    out = []
    for p in pairs:
        out.extend(p)
    return tuple(out)


def _poly_mul(a, b, deg):
    """product of two degree<deg polynomials over Fq2 modulo X^deg - xi, with the
    reductions mod Q deferred to the end like fields_t.py:293-318 / 503-554."""
    lo = [[0, 0] for _ in range(2 * deg - 1)]
    for i, (x0, x1) in enumerate(a):
        if x0 == 0 and x1 == 0:
            continue
        for j, (y0, y1) in enumerate(b):
            s = lo[i + j]
            s[0] += x0 * y0 - x1 * y1
            s[1] += x0 * y1 + x1 * y0
    out = []
    for k in range(deg):
        c0, c1 = lo[k]
        if k + deg < 2 * deg - 1:
            h0, h1 = lo[k + deg]
            c0 += h0 - h1
            c1 += h0 + h1
        out.append((c0 % Q, c1 % Q))
    return out


def f6_add(a, b):
    return tuple((x + y) % Q for x, y in zip(a, b))


def f6_sub(a, b):
    return tuple((x - y) % Q for x, y in zip(a, b))


def f6_neg(a):
    return tuple(-x % Q for x in a)


def f6_mul(a, b):
    return _flat(_poly_mul(_pairs(a), _pairs(b), 3))


def f6_mul_v(a):
    """times v (fields_t.py:215-220)."""
    a0, a1, a2 = _pairs(a)
    return _flat([f2_mul_xi(a2), a0, a1])


def f6_inv(a):
    """fields_t.py:170-184 (norm to Fq2, one Fq2 inversion)."""
    a0, a1, a2 = _pairs(a)
    g0 = f2_sub(f2_mul(a0, a0), f2_mul_xi(f2_mul(a1, a2)))
    g1 = f2_sub(f2_mul_xi(f2_mul(a2, a2)), f2_mul(a0, a1))
    g2 = f2_sub(f2_mul(a1, a1), f2_mul(a0, a2))
    nrm = f2_add(f2_mul(a0, g0),
                 f2_mul_xi(f2_add(f2_mul(a2, g1), f2_mul(a1, g2))))
    t = f2_inv(nrm)
    return _flat([f2_mul(g0, t), f2_mul(g1, t), f2_mul(g2, t)])


F12_ONE = (1,) + (0,) * 11
F12_ZERO = (0,) * 12
# flat index (pair units) <-> power of w:  pair p = 3*i + j  <->  w^(2j+i)
_W_OF_PAIR = [0, 2, 4, 1, 3, 5]
_PAIR_OF_W = [0, 3, 1, 4, 2, 5]


def _to_w(a):
    p = _pairs(a)
    return [p[_PAIR_OF_W[k]] for k in range(6)]


def _from_w(c):
    return _flat([c[_W_OF_PAIR[p]] for p in range(6)])


def f12_add(a, b):
    return tuple((x + y) % Q for x, y in zip(a, b))


def f12_sub(a, b):
    return tuple((x - y) % Q for x, y in zip(a, b))


def f12_neg(a):
    return tuple(-x % Q for x in a)


def f12_scale(a, k):
    return tuple(x * k % Q for x in a)


def f12_mul(a, b):
    """fields_t.py:503-554."""
    return _from_w(_poly_mul(_to_w(a), _to_w(b), 6))


def f12_inv(a):
    """fields_t.py:328-337: (b0 - b1 w) / (b0^2 - v b1^2)."""
    b0, b1 = a[:6], a[6:]
    d = f6_inv(f6_sub(f6_mul(b0, b0), f6_mul_v(f6_mul(b1, b1))))
    return f6_mul(b0, d) + f6_mul(f6_neg(b1), d)


def f12_pow(a, e):
    """LSB-first square and multiply (fields_t.py:344-352)."""
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


_FROB_CACHE = {}


def _frob_consts(i):
    """gamma_k = xi^(k (Q^i - 1) / 6), k = 0..5: w^k -> gamma_k w^k under x -> x^(Q^i)
    (what the table at fields_t.py:1133-1216 holds)."""
    if i not in _FROB_CACHE:
        g = f2_pow((1, 1), (Q ** i - 1) // 6)
        tab = [F2_ONE]
        for _ in range(5):
            tab.append(f2_mul(tab[-1], g))
        _FROB_CACHE[i] = tab
    return _FROB_CACHE[i]


def f12_frob(a, i):
    """x -> x^(Q^i) (fields_t.py:355-364)."""
    i %= 12
    if i == 0:
        return tuple(a)
    tab = _frob_consts(i)
    out = []
    for k, c in enumerate(_to_w(a)):
        if i & 1:
            c = (c[0], -c[1] % Q)
        out.append(f2_mul(c, tab[k]))
    return _from_w(out)


def f12_serialize(a):
    """48-byte big-endian per coefficient, ZT order (fields.py:273-278)."""
    return b"".join(int(c).to_bytes(48, "big") for c in a)


# ---------------------------------------------------------------------------
# curve arithmetic, generic over a small field descriptor
# ---------------------------------------------------------------------------
class _F1:
    zero, one = 0, 1
    add = staticmethod(lambda a, b: (a + b) % Q)
    sub = staticmethod(lambda a, b: (a - b) % Q)
    mul = staticmethod(lambda a, b: a * b % Q)
    neg = staticmethod(lambda a: -a % Q)
    inv = staticmethod(fq_inv)
    scale = staticmethod(lambda a, k: a * k % Q)
    b = B1


class _F2:
    zero, one = F2_ZERO, F2_ONE
    add, sub, mul, neg = map(staticmethod, (f2_add, f2_sub, f2_mul, f2_neg))
    inv = staticmethod(f2_inv)
    scale = staticmethod(f2_scale)
    b = B2


class _F12:
    zero, one = F12_ZERO, F12_ONE
    add, sub, mul, neg = map(staticmethod, (f12_add, f12_sub, f12_mul, f12_neg))
    inv = staticmethod(f12_inv)
    scale = staticmethod(f12_scale)
    b = (4,) + (0,) * 11


def _field_of(x):
    if isinstance(x, int):
        return _F1
    return _F2 if len(x) == 2 else _F12


def aff_double(p):
    """fields_t.py:625-646 (affine tangent; no special cases, like the reference)."""
    x, y, _ = p
    F = _field_of(x)
    s = F.mul(F.scale(F.mul(x, x), 3), F.inv(F.scale(y, 2)))
    xr = F.sub(F.mul(s, s), F.scale(x, 2))
    yr = F.sub(F.mul(s, F.sub(x, xr)), y)
    return (xr, yr, False)


def aff_add(p1, p2):
    """fields_t.py:657-686."""
    if p1[2]:
        return p2
    if p2[2]:
        return p1
    x1, y1, _ = p1
    x2, y2, _ = p2
    F = _field_of(x1)
    if x1 == x2:
        if y1 == y2:
            return aff_double(p1)
        return (F.zero, F.zero, True)
    s = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    xr = F.sub(F.sub(F.mul(s, s), x1), x2)
    yr = F.sub(F.mul(s, F.sub(x1, xr)), y1)
    return (xr, yr, False)


def aff_neg(p):
    F = _field_of(p[0])
    return (p[0], F.neg(p[1]), p[2])


def jac_double(p):
    """fields_t.py:878-933 (a = 0)."""
    x, y, z, inf = p
    F = _field_of(x)
    if inf:
        return p
    yy = F.mul(y, y)
    s = F.scale(F.mul(x, yy), 4)
    m = F.scale(F.mul(x, x), 3)
    xr = F.sub(F.mul(m, m), F.scale(s, 2))
    yr = F.sub(F.mul(m, F.sub(s, xr)), F.scale(F.mul(yy, yy), 8))
    zr = F.scale(F.mul(y, z), 2)
    return (xr, yr, zr, False)


def jac_add(p1, p2):
    """fields_t.py:762-875.  P + P doubles (the reference's G1 TypeError at
    fields_t.py:781 is a defect that is not reproduced); P + (-P) = (1,1,0,inf)."""
    if p1[3]:
        return p2
    if p2[3]:
        return p1
    x1, y1, z1, _ = p1
    x2, y2, z2, _ = p2
    F = _field_of(x1)
    z1z1, z2z2 = F.mul(z1, z1), F.mul(z2, z2)
    u1, u2 = F.mul(x1, z2z2), F.mul(x2, z1z1)
    s1, s2 = F.mul(y1, F.mul(z2z2, z2)), F.mul(y2, F.mul(z1z1, z1))
    if u1 == u2:
        if s1 != s2:
            return (F.one, F.one, F.zero, True)
        return jac_double(p1)
    h, r = F.sub(u2, u1), F.sub(s2, s1)
    hh = F.mul(h, h)
    hhh = F.mul(hh, h)
    v = F.mul(u1, hh)
    xr = F.sub(F.sub(F.mul(r, r), hhh), F.scale(v, 2))
    yr = F.sub(F.mul(r, F.sub(v, xr)), F.mul(s1, hhh))
    zr = F.mul(F.mul(z1, z2), h)
    return (xr, yr, zr, False)


def to_jac(p):
    F = _field_of(p[0])
    return (p[0], p[1], F.one, p[2])


def jac_inf(F):
    return (F.one, F.one, F.zero, True)


def to_aff(p):
    """fields_t.py:609-632: infinity -> (0, 0, True)."""
    x, y, z, inf = p
    F = _field_of(x)
    if inf:
        return (F.zero, F.zero, True)
    zi = F.inv(z)
    zi2 = F.mul(zi, zi)
    return (F.mul(x, zi2), F.mul(y, F.mul(zi2, zi)), False)


def jac_mul(k, p):
    """LSB-first double and add (fields_t.py:705-759), including the early
    out on k % Q == 0 (field prime, not group order -- a reference quirk)."""
    F = _field_of(p[0])
    acc = jac_inf(F)
    if p[3] or k % Q == 0:
        return acc
    while k > 0:
        if k & 1:
            acc = jac_add(acc, p)
        p = jac_double(p)
        k >>= 1
    return acc


def aff_mul(k, p):
    """AffinePoint.__mul__ (ec.py:79-82)."""
    return to_aff(jac_mul(k, to_jac(p)))


def on_curve(p):
    x, y, inf = p
    if inf:
        return True
    F = _field_of(x)
    return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), F.b)


# ---------------------------------------------------------------------------
# serialization (ec.py:94-111, keys.py:29-40, signature.py:22-38)
# ---------------------------------------------------------------------------
def g1_serialize(p):
    """48 bytes: x big-endian, top bit set iff y > Q//2; infinity -> zero bytes."""
    x, y, _ = p if len(p) == 3 else to_aff(p)
    out = bytearray(x.to_bytes(48, "big"))
    if y > Q // 2:
        out[0] |= 0x80
    return bytes(out)


def g2_serialize(p):
    """96 bytes: x.c0 || x.c1, top bit set iff y.c1 > Q//2 (only c1 examined)."""
    x, y, _ = p if len(p) == 3 else to_aff(p)
    out = bytearray(x[0].to_bytes(48, "big") + x[1].to_bytes(48, "big"))
    if y[1] > Q // 2:
        out[0] |= 0x80
    return bytes(out)


def g1_y_for_x(x):
    """ec.py:255-269 for Fq: both roots or ValueError."""
    u = (x * x * x + B1) % Q
    y = fq_sqrt(u)
    if y == 0:
        raise ValueError("No y for point x")
    return [y, Q - y]


def g2_y_for_x(x):
    """ec.py:255-269 for Fq2.  A root that is not an Fq2 pair (the real-input
    quirk of f2_sqrt) makes the reference raise; so does y == 0."""
    u = f2_add(f2_mul(f2_mul(x, x), x), B2)
    y = f2_sqrt(u)
    if isinstance(y, int) or y == F2_ZERO or f2_mul(y, y) != u:
        raise ValueError("No y for point x")
    return [y, f2_neg(y)]


def g1_deserialize(buf):
    """PublicKey.from_bytes (keys.py:29-40)."""
    big = buf[0] & 0x80
    x = int.from_bytes(bytes([buf[0] & 0x1f]) + buf[1:], "big") % Q
    ys = sorted(g1_y_for_x(x))
    return (x, ys[1] if big else ys[0], False)


def g2_deserialize(buf):
    """Signature.from_bytes (signature.py:22-38)."""
    big = buf[0] & 0x80
    buf = bytes([buf[0] & 0x1f]) + buf[1:]
    x = (int.from_bytes(buf[:48], "big") % Q, int.from_bytes(buf[48:], "big") % Q)
    ys = g2_y_for_x(x)
    y = ys[0]
    if (big and ys[1][1] > Q // 2) or (not big and ys[1][1] < Q // 2):
        y = ys[1]
    return (x, y, False)


# ---------------------------------------------------------------------------
# twist maps and psi (fields_t.py:936-1031, ec.py:402-444)
# ---------------------------------------------------------------------------
_INV_W2 = None
_INV_W3 = None


def _untwist_consts():
    global _INV_W2, _INV_W3
    if _INV_W2 is None:
        w = (0,) * 6 + (1,) + (0,) * 5
        w2 = f12_mul(w, w)
        _INV_W2 = f12_inv(w2)
        _INV_W3 = f12_inv(f12_mul(w2, w))
    return _INV_W2, _INV_W3


def _embed2(a):
    return tuple(a) + (0,) * 10


def untwist(p):
    """E'(Fq2) -> E(Fq12): (x / w^2, y / w^3) (fields_t.py:936-943)."""
    i2, i3 = _untwist_consts()
    return (f12_mul(_embed2(p[0]), i2), f12_mul(_embed2(p[1]), i3), False)


def twist(p):
    """E(Fq12) -> E'(Fq12 coordinates): (x w^2, y w^3) (fields_t.py:1018-1031)."""
    w = (0,) * 6 + (1,) + (0,) * 5
    w2 = f12_mul(w, w)
    return (f12_mul(p[0], w2), f12_mul(p[1], f12_mul(w2, w)), False)


def psi(p):
    """untwist -> Frobenius -> twist, back in Fq2 coordinates (ec.py:440-444)."""
    ut = untwist(p)
    t = twist((f12_frob(ut[0], 1), f12_frob(ut[1], 1), False))
    return (t[0][:2], t[1][:2], False)


# ---------------------------------------------------------------------------
# pairing (fields_t.py:1035-1128, pairing.py:51-92)
# ---------------------------------------------------------------------------
def _line_tangent(r, px, py):
    """fields_t.py:1035-1049: tangent at untwist(R) evaluated at P."""
    x, y, _ = untwist(r)
    slope = f12_mul(f12_scale(f12_mul(x, x), 3), f12_inv(f12_scale(y, 2)))
    v = f12_sub(y, f12_mul(slope, x))
    ell = f12_neg(f12_scale(slope, px))
    ell = ((ell[0] + py) % Q,) + ell[1:]
    return f12_sub(ell, v)


def _line_chord(r, q, px, py):
    """fields_t.py:1052-1078 incl. the vertical-line case R == -Q."""
    rx, ry, _ = untwist(r)
    qx, qy, _ = untwist(q)
    if rx == qx and ry == f12_neg(qy):
        out = f12_neg(rx)
        return ((out[0] + px) % Q,) + out[1:]
    slope = f12_mul(f12_sub(qy, ry), f12_inv(f12_sub(qx, rx)))
    v = f12_mul(f12_sub(f12_mul(qy, rx), f12_mul(ry, qx)),
                f12_inv(f12_sub(rx, qx)))
    ell = f12_neg(f12_scale(slope, px))
    ell = ((ell[0] + py) % Q,) + ell[1:]
    return f12_sub(ell, v)


def miller_loop(p, q):
    """f_{|x|,Q}(P) with affine R and dense Fq12 lines (fields_t.py:1091-1111).
    Infinity flags are ignored exactly as in the reference."""
    px, py, _ = p
    r = (q[0], q[1], q[2])
    f = F12_ONE
    for bit in bin(X_ABS)[3:]:
        f = f12_mul(f12_mul(f, f), _line_tangent(r, px, py))
        r = aff_double(r)
        if bit == "1":
            f = f12_mul(f, _line_chord(r, q, px, py))
            r = aff_add(r, q)
    return f


def final_exp(f):
    """fields_t.py:1124-1128: f^E, then ^(Q^2+1), then ^(Q^6-1)."""
    a = f12_pow(f, FINAL_EXP_HARD)
    a = f12_mul(f12_frob(a, 2), a)
    return f12_mul(f12_frob(a, 6), f12_inv(a))


def ate_pairing(p, q):
    return final_exp(miller_loop(p, q))


def ate_pairing_multi(ps, qs):
    """fields_t.py:1114-1121: product of Miller loops, one final exponentiation."""
    f = F12_ONE
    for p, q in zip(ps, qs):
        f = f12_mul(f, miller_loop(p, q))
    return final_exp(f)


# ---------------------------------------------------------------------------
# hashing to G2 (util.py:7-16, ec.py:449-555)
# ---------------------------------------------------------------------------
def hash256(m):
    return hashlib.sha256(m).digest()


def hash512(m):
    return hash256(m + b"\x00") + hash256(m + b"\x01")


def sw_encode_g2(t):
    """Shallue-van de Woestijne / Fouque-Tibouchi map to E'(Fq2) (ec.py:449-507).
    Returns an affine point (infinity for t == 0)."""
    if t == F2_ZERO:
        return (F2_ZERO, F2_ZERO, True)
    parity = t[1] > (-t[1] % Q)
    w0 = f2_add(f2_add(f2_mul(t, t), B2), F2_ONE)
    if w0 == F2_ZERO:
        return G2                       # no parity negation for Fq2 (ec.py:466-470)
    w = f2_mul(f2_scale(f2_inv(w0), SQRT_M3), t)
    x1 = f2_neg(f2_mul(w, t))
    x1 = ((x1[0] + SQRT_M3_M1_O2) % Q, x1[1])
    x2 = f2_sub((Q - 1, 0), x1)
    x3 = f2_inv(f2_mul(w, w))
    x3 = ((x3[0] + 1) % Q, x3[1])
    chi = []
    for x in (x1, x2):
        try:
            g2_y_for_x(x)
            chi.append(1)
        except ValueError:
            chi.append(-1)
    idx = ((chi[0] - 1) * chi[1]) % 3
    x = (x1, x2, x3)[idx]
    y = g2_y_for_x(x)[0]
    if (y[1] > Q // 2) is not parity:
        y = f2_neg(y)
    return (x, y, False)


def hash_to_g2_prehashed(h):
    """ec.py:528-550: two SW encodings, affine add, Budroni-Pintore clearing."""
    ts = []
    for j in (b"0", b"1"):
        c0 = int.from_bytes(hash512(h + b"G2_" + j + b"_c0"), "big") % Q
        c1 = int.from_bytes(hash512(h + b"G2_" + j + b"_c1"), "big") % Q
        ts.append((c0, c1))
    p = aff_add(sw_encode_g2(ts[0]), sw_encode_g2(ts[1]))
    a = X_ABS
    psi2p = psi(psi(aff_mul(2, p)))
    t0 = aff_mul(a, p)
    t1 = aff_mul(a, t0)
    t2 = aff_add(aff_add(t1, t0), aff_neg(p))
    t3 = psi(aff_mul(a + 1, p))
    return aff_add(aff_add(t2, aff_neg(t3)), psi2p)


def hash_to_g2(msg):
    return hash_to_g2_prehashed(hash256(msg))


# ---------------------------------------------------------------------------
# scheme level (keys.py:90-132, bls.py:154-223, util.py:19-50)
# ---------------------------------------------------------------------------
def hmac256(m, k):
    """util.py:19-33."""
    if len(k) > 64:
        k = hash256(k)
    k = k + bytes(64 - len(k))
    return hash256(bytes(b ^ 0x5c for b in k) + hash256(bytes(b ^ 0x36 for b in k) + m))


def sk_from_seed(seed):
    """keys.py:90-92."""
    return int.from_bytes(hmac256(seed, b"BLS private key seed"), "big") % N


def pk_of(sk):
    """keys.py:119-121 -> affine G1."""
    return aff_mul(sk, G1)


def sign_prehashed(sk, h):
    """keys.py:128-132 -> affine G2."""
    return aff_mul(sk, hash_to_g2_prehashed(h))


def sign(sk, msg):
    return sign_prehashed(sk, hash256(msg))


def hash_pks(num_outputs, pk_bytes_list):
    """util.py:36-50 on already-serialized public keys."""
    pk_hash = hash256(b"".join(pk_bytes_list))
    return [int.from_bytes(hash256(i.to_bytes(4, "big") + pk_hash), "big") % N
            for i in range(num_outputs)]


NEG_G1 = None


def _neg_g1():
    global NEG_G1
    if NEG_G1 is None:
        NEG_G1 = aff_mul(N - 1, G1)        # bls.py:197
    return NEG_G1


def aggregate_verify(pks, hashes, sig):
    """bls.py:194-201 for distinct message hashes and unit exponents:
    e(-G1, sig) * prod e(pk_i, H(h_i)) == 1."""
    ps = [_neg_g1()] + list(pks)
    qs = [sig] + [hash_to_g2_prehashed(h) for h in hashes]
    return ate_pairing_multi(ps, qs) == F12_ONE


def verify(pk, h, sig):
    return aggregate_verify([pk], [h], sig)


def g1_sum(points):
    """left fold of bls.py:217-221 (secure=False) on affine inputs -> affine."""
    acc = jac_inf(_F1)
    for p in points:
        acc = jac_add(acc, to_jac(p))
    return to_aff(acc)


def g2_sum(points):
    """left fold of bls.py:13-26 on affine inputs -> affine."""
    acc = jac_inf(_F2)
    for p in points:
        acc = jac_add(acc, to_jac(p))
    return to_aff(acc)
