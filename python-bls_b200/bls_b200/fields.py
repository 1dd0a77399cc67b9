"""Fq12 value object returned by the pairing entry points (mirrors the parts of
bls_py/fields.py:624-764 that callers of ate_pairing_multi touch: equality, one(),
serialize(), multiplication, inversion, pow).  Arithmetic runs on the GPU."""
from . import engine

Q = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
        "1eabfffeb153ffffb9feffffffffaaab", 16)


class Fq12:
    extension = 12
    __slots__ = ("raw",)

    def __init__(self, Q_or_raw, coeffs=None):
        if coeffs is None:
            raw = bytes(Q_or_raw)
        else:
            raw = b"".join((int(c) % Q).to_bytes(48, "big") for c in coeffs)
        if len(raw) != 576:
            raise ValueError("Fq12 needs 12 coefficients")
        self.raw = raw

    @staticmethod
    def one(q=Q):
        return Fq12(q, (1,) + (0,) * 11)

    @staticmethod
    def zero(q=Q):
        return Fq12(q, (0,) * 12)

    @property
    def ZT(self):
        return tuple(int.from_bytes(self.raw[i:i + 48], "big") for i in range(0, 576, 48))

    def serialize(self):
        """48-byte big-endian per coefficient, ZT order (fields.py:273-278)"""
        return self.raw

    def __eq__(self, other):
        return isinstance(other, Fq12) and self.raw == other.raw

    def __hash__(self):
        return hash(self.raw)

    def __mul__(self, other):
        return Fq12(engine.field_op(12, "mul", self.raw, other.raw).tobytes())

    def __invert__(self):
        return Fq12(engine.field_op(12, "inv", self.raw).tobytes())

    def __truediv__(self, other):
        return self * ~other

    def __pow__(self, e):
        e = int(e)
        if e < 0:
            return (~self) ** (-e)
        acc, base = Fq12.one(), self
        while e:
            if e & 1:
                acc = acc * base
            base = base * base
            e >>= 1
        return acc

    def __repr__(self):
        return "Fq12(%s...)" % self.raw[:8].hex()
