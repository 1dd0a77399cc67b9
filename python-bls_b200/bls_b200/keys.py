"""PrivateKey / PublicKey / ExtendedPrivateKey / ExtendedPublicKey with the reference's
interface (bls_py/keys.py:17-316); curve work runs on the GPU."""
import secrets

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
    def from_bytes_batch(buffers):
        """many keys decoded in one GPU call; their serialised forms are cached"""
        return [PublicKey(p) for p in ec.points_from_bytes(buffers, False)]

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

    @staticmethod
    def new_threshold(T, N, rng=None):
        """one player's dealing of a T-of-N Joint-Feldman scheme (keys.py:93-117): a random
        degree T-1 polynomial, commitments g1 * coefficient (one batched scalar multiplication)
        and the fragments P(1..N).  `rng` (random.Random-like) makes the dealing reproducible."""
        assert 1 <= T <= N
        draw = (lambda: rng.randrange(1, GROUP_ORDER)) if rng is not None else \
            (lambda: 1 + secrets.randbelow(GROUP_ORDER - 1))
        poly = [draw() for _ in range(T)]
        commitments = ec.scalar_mul_many([ec.generator_Fq()] * T, poly, False)
        fragments = [sum(c * pow(x, i, GROUP_ORDER) for i, c in enumerate(poly)) % GROUP_ORDER
                     for x in range(1, N + 1)]
        return PrivateKey(poly[0]), commitments, fragments

    def get_public_key(self):
        return PublicKey(ec.generator_Fq() * self.value)

    def sign_threshold(self, m, player, players):
        """this player's signature share, already weighted by its Lagrange coefficient
        (keys.py:137-147)"""
        from .signature import Signature
        from .threshold import Threshold
        assert player in players
        lamb = Threshold.lagrange_coeffs_at_zero(players)[players.index(player)]
        return Signature.from_g2(ec.hash_to_point_Fq2(m) * (self.value * lamb % GROUP_ORDER))

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


def _hd_halves(data, chain_code):
    return hmac256(data + bytes([0]), chain_code), hmac256(data + bytes([1]), chain_code)


class ExtendedPrivateKey:
    """BIP32-style hierarchical key (keys.py:167-253): child secret = parent secret + HMAC half"""
    version = 1
    EXTENDED_PRIVATE_KEY_SIZE = 77

    def __init__(self, version, depth, parent_fingerprint, child_number, chain_code, private_key):
        self.version = version
        self.depth = depth
        self.parent_fingerprint = parent_fingerprint
        self.child_number = child_number
        self.chain_code = chain_code
        self.private_key = private_key

    @staticmethod
    def from_seed(seed):
        left, right = _hd_halves(bytes(seed), b"BLS HD seed")
        return ExtendedPrivateKey(ExtendedPrivateKey.version, 0, 0, 0, right,
                                  PrivateKey(int.from_bytes(left, "big") % GROUP_ORDER))

    def private_child(self, i):
        if self.depth >= 255:
            raise Exception("Cannot go further than 255 levels")
        parent_pk = self.private_key.get_public_key()
        # hardened children (i >= 2^31) commit to the secret key, the others to the public key
        data = (self.private_key.serialize() if i >= 2 ** 31 else parent_pk.serialize()) + i.to_bytes(4, "big")
        left, right = _hd_halves(data, self.chain_code)
        sk = PrivateKey((int.from_bytes(left, "big") + self.private_key.value) % GROUP_ORDER)
        return ExtendedPrivateKey(ExtendedPrivateKey.version, self.depth + 1, parent_pk.get_fingerprint(), i, right, sk)

    def public_child(self, i):
        return self.private_child(i).get_extended_public_key()

    def _header(self):
        return (self.version.to_bytes(4, "big") + bytes([self.depth]) + self.parent_fingerprint.to_bytes(4, "big") +
                self.child_number.to_bytes(4, "big") + self.chain_code)

    def get_extended_public_key(self):
        return ExtendedPublicKey.from_bytes(self._header() + self.private_key.get_public_key().serialize())

    def get_private_key(self):
        return self.private_key

    def get_public_key(self):
        return self.private_key.get_public_key()

    def size(self):
        return self.EXTENDED_PRIVATE_KEY_SIZE

    def serialize(self):
        return self._header() + self.private_key.serialize()

    def __eq__(self, other):
        return self.serialize() == other.serialize()

    def __hash__(self):
        return int.from_bytes(self.serialize(), "big")


class ExtendedPublicKey:
    """public half of the hierarchy (keys.py:256-316): non-hardened children only,
    child key = parent key + g1 * HMAC half (one scalar multiplication and one addition)"""
    EXTENDED_PUBLIC_KEY_SIZE = 93

    def __init__(self, version, depth, parent_fingerprint, child_number, chain_code, public_key):
        self.version = version
        self.depth = depth
        self.parent_fingerprint = parent_fingerprint
        self.child_number = child_number
        self.chain_code = chain_code
        self.public_key = public_key

    @staticmethod
    def from_bytes(serialized):
        serialized = bytes(serialized)
        return ExtendedPublicKey(int.from_bytes(serialized[:4], "big"), serialized[4],
                                 int.from_bytes(serialized[5:9], "big"), int.from_bytes(serialized[9:13], "big"),
                                 serialized[13:45], PublicKey.from_bytes(serialized[45:]))

    def public_child(self, i):
        if self.depth >= 255:
            raise Exception("Cannot go further than 255 levels")
        if i >= 2 ** 31:
            raise Exception("Cannot derive hardened children from public key")
        left, right = _hd_halves(self.public_key.serialize() + i.to_bytes(4, "big"), self.chain_code)
        tweak = ec.generator_Fq() * (int.from_bytes(left, "big") % GROUP_ORDER)
        return ExtendedPublicKey(self.version, self.depth + 1, self.public_key.get_fingerprint(), i, right,
                                 PublicKey.from_g1(tweak + self.public_key.value))

    def get_public_key(self):
        return self.public_key

    def size(self):
        return self.EXTENDED_PUBLIC_KEY_SIZE

    def serialize(self):
        return (self.version.to_bytes(4, "big") + bytes([self.depth]) + self.parent_fingerprint.to_bytes(4, "big") +
                self.child_number.to_bytes(4, "big") + self.chain_code + self.public_key.serialize())

    def __eq__(self, other):
        return self.serialize() == other.serialize()

    def __hash__(self):
        return int.from_bytes(self.serialize(), "big")
