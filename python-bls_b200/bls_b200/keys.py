"""PrivateKey / PublicKey with the reference's interface (bls_py/keys.py:17-165); curve work
runs on the GPU.  HD keys (keys.py:167-316) and threshold helpers are out of scope
(SURVEY.md 8)."""
from . import ec
from .aggregation_info import AggregationInfo
from .util import GROUP_ORDER, hash256, hmac256


class PublicKey:
    PUBLIC_KEY_SIZE = 48

    def __init__(self, value):
        self.value = value                      # ec.Point on G1

    @staticmethod
    def from_bytes(buffer):
        return PublicKey(ec.point_from_bytes(bytes(buffer), False))

    @staticmethod
    def from_g1(g1_el):
        assert isinstance(g1_el, ec.Point) and not g1_el.g2
        return PublicKey(g1_el)

    def serialize(self):
        return self.value.serialize()

    def get_fingerprint(self):
        return int.from_bytes(hash256(self.serialize())[:4], "big")

    def size(self):
        return self.PUBLIC_KEY_SIZE

    def __eq__(self, other):
        return self.serialize() == other.serialize()

    def __hash__(self):
        return int.from_bytes(self.serialize(), "big")

    def __lt__(self, other):
        return self.serialize() < other.serialize()

    def __repr__(self):
        return "PublicKey(%s)" % self.serialize().hex()

    __str__ = __repr__


class PrivateKey:
    PRIVATE_KEY_SIZE = 32

    def __init__(self, value):
        self.value = int(value)

    @staticmethod
    def from_bytes(buffer):
        return PrivateKey(int.from_bytes(buffer, "big"))

    @staticmethod
    def from_seed(seed):
        return PrivateKey(int.from_bytes(hmac256(seed, b"BLS private key seed"), "big") % GROUP_ORDER)

    def get_public_key(self):
        return PublicKey(ec.generator_Fq() * self.value)

    def sign(self, m):
        return self.sign_prehashed(hash256(m))

    def sign_prehashed(self, h):
        from .signature import Signature
        r = ec.hash_to_point_prehashed_Fq2(h)
        return Signature.from_g2(r * self.value, AggregationInfo.from_msg_hash(self.get_public_key(), h))

    def serialize(self):
        return self.value.to_bytes(self.PRIVATE_KEY_SIZE, "big")

    def size(self):
        return self.PRIVATE_KEY_SIZE

    def __lt__(self, other):
        return self.value < other.value

    def __eq__(self, other):
        return self.value == other.value

    def __hash__(self):
        return self.value

    def __repr__(self):
        return "PrivateKey(%s)" % hex(self.value)

    __str__ = __repr__
