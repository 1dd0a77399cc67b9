"""Field elements of the scheme layer: Fq, Fq2, Fq6, Fq12 with the reference's class surface (bls_py/fields.py:35-764:
operators, ~x, pow, qi_power, modsqrt, serialize, one / zero / from_fq, ZT / Z) -- every operation one call into the
batched C ABI.  A value is its canonical serialisation (48-byte big-endian coefficients in the reference's flat ZT
order, fields.py:273-278); nothing here computes field arithmetic on the CPU.  One element per call is the
reference's granularity, not the GPU's: bulk work goes through bls_b200.engine directly."""
from . import engine

Q = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
        "1eabfffeb153ffffb9feffffffffaaab", 16)


class _Field:
    """an element of the tower level `extension` in {1, 2, 6, 12}"""
    extension = 1
    __slots__ = ("raw",)

    def __init__(self, Q_or_raw, *coeffs):
        """Fq(Q, z) / Fq2(Q, c0, c1) / Fq12(Q, zt_tuple) like the reference, or Cls(raw bytes)"""
        if not coeffs:
            raw = bytes(Q_or_raw)
        else:
            if len(coeffs) == 1 and isinstance(coeffs[0], (tuple, list)):
                coeffs = tuple(coeffs[0])
            flat = []
            for c in coeffs:                       # Fq6(Q, a0, a1, a2) with Fq2 arguments, as the reference allows
                flat.extend(c.ZT if isinstance(c, _Field) else (int(c),))
            raw = b"".join((int(c) % Q).to_bytes(48, "big") for c in flat)
        if len(raw) != 48 * self.extension:
            raise ValueError("%s needs %d coefficients" % (type(self).__name__, self.extension))
        self.raw = raw

    # -- constructors / accessors of the reference ---------------------------------------------
    @classmethod
    def zero(cls, q=Q):
        return cls(bytes(48 * cls.extension))

    @classmethod
    def one(cls, q=Q):
        return cls((1).to_bytes(48, "big") + bytes(48 * (cls.extension - 1)))

    @classmethod
    def from_fq(cls, q, fq):
        return cls(fq.raw[:48] + bytes(48 * (cls.extension - 1)))

    @property
    def ZT(self):
        return tuple(int.from_bytes(self.raw[i:i + 48], "big") for i in range(0, len(self.raw), 48))

    @property
    def Z(self):
        return self.ZT[0]

    @property
    def Q(self):
        return Q

    def serialize(self):
        return self.raw

    def __iter__(self):
        """coefficients one tower level down (Fq12 -> two Fq6, Fq6 -> three Fq2, Fq2 -> two Fq), fields.py:280-284"""
        sub = {2: (Fq, 1), 6: (Fq2, 2), 12: (Fq6, 6)}.get(self.extension)
        if sub is None:
            raise TypeError("Fq is not iterable")
        cls, w = sub
        return iter([cls(self.raw[48 * w * k:48 * w * (k + 1)]) for k in range(self.extension // w)])

    def __getitem__(self, k):
        return list(self)[k]

    # -- arithmetic: the other operand is lifted to this level at coefficient 0 (fields.py mixed-type operators) ---
    def _lift(self, other):
        if isinstance(other, _Field):
            if other.extension > self.extension:
                return None
            return other.raw + bytes(len(self.raw) - len(other.raw))
        if isinstance(other, int):
            return (other % Q).to_bytes(48, "big") + bytes(len(self.raw) - 48)
        return None

    def _op(self, name, other=None):
        b = None
        if other is not None:
            b = self._lift(other)
            if b is None:
                return NotImplemented
        return type(self)(engine.field_op(self.extension, name, self.raw, b).tobytes())

    def __add__(self, o):
        return self._op("add", o)

    def __radd__(self, o):
        return self._op("add", o)

    def __sub__(self, o):
        return self._op("sub", o)

    def __rsub__(self, o):
        b = self._lift(o)
        if b is None:
            return NotImplemented
        return type(self)(engine.field_op(self.extension, "sub", b, self.raw).tobytes())

    def __mul__(self, o):
        if isinstance(o, _Field) and o.extension > self.extension:
            return o * self
        return self._op("mul", o)

    def __rmul__(self, o):
        return self._op("mul", o)

    def __neg__(self):
        return self._op("neg")

    def __invert__(self):
        return self._op("inv")

    def __truediv__(self, o):
        b = self._lift(o)
        if b is None:
            return NotImplemented
        inv = engine.field_op(self.extension, "inv", b).tobytes()
        return type(self)(engine.field_op(self.extension, "mul", self.raw, inv).tobytes())

    __floordiv__ = __truediv__

    def __pow__(self, e):
        """fields_t.py:58-68, 92-101, 344-352; 384 bits of the exponent per device call"""
        e = int(e)
        if e < 0:
            return (~self) ** (-e)
        cls, base, acc = type(self), self, None
        while True:
            part = cls(engine.field_pow(self.extension, base.raw, [e & ((1 << 384) - 1)]).tobytes())
            acc = part if acc is None else acc * part
            e >>= 384
            if not e:
                return acc
            half = cls(engine.field_pow(self.extension, base.raw, [1 << 383]).tobytes())
            base = half * half                     # base^(2^384)

    def qi_power(self, i):
        """x^(q^i) (fields.py:286-293 -> fields_t.py:104-110, 203-212, 355-364)"""
        if self.extension == 1:
            return self
        return type(self)(engine.field_frob(self.extension, i % self.extension, self.raw).tobytes())

    def modsqrt(self):
        """Fq: fields.py:199-205, Fq2: fields.py:463-482 -- the reference's own root; ValueError('No sqrt exists')"""
        if self.extension > 2:
            raise NotImplementedError("modsqrt is defined for Fq and Fq2")
        out, ok = engine.field_sqrt(self.extension, self.raw)
        if not ok[0]:
            raise ValueError("No sqrt exists")
        if self.extension == 2 and not any(self.raw[48:]):
            return Fq(out.tobytes()[:48])         # the reference hands a real Fq2 element to Fq.modsqrt
        return type(self)(out.tobytes())

    def __eq__(self, other):
        if isinstance(other, _Field):
            if other.extension > self.extension:
                return other == self
            return self.raw == other.raw + bytes(len(self.raw) - len(other.raw))
        if isinstance(other, int):
            return self.raw == (other % Q).to_bytes(48, "big") + bytes(len(self.raw) - 48)
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __lt__(self, other):
        """the reference compares coefficient tuples from the highest coefficient down (fields.py:295-303)"""
        return self.ZT[::-1] < other.ZT[::-1]

    def __gt__(self, other):
        return self.ZT[::-1] > other.ZT[::-1]

    def __hash__(self):
        return hash(self.raw)

    def __bool__(self):
        return any(self.raw)

    def __int__(self):
        if self.extension != 1:
            raise TypeError("only Fq converts to int")
        return self.Z

    def __repr__(self):
        return "%s(%s...)" % (type(self).__name__, self.raw[:8].hex())


class Fq(_Field):
    extension = 1
    __slots__ = ()


class Fq2(_Field):
    extension = 2
    __slots__ = ()

    def mul_by_nonresidue(self):
        """times (1 + u) (fields_t.py:113-116)"""
        return self * Fq2(Q, 1, 1)


class Fq6(_Field):
    extension = 6
    __slots__ = ()

    def mul_by_nonresidue(self):
        """times v (fields_t.py:215-220)"""
        return self * Fq6(Q, 0, 0, 1, 0, 0, 0)


class Fq12(_Field):
    extension = 12
    __slots__ = ()
