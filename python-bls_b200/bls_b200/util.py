"""SHA-256 helpers of the scheme layer (host side; mirrors bls_py/util.py:7-50).

The field-element derivation of hash-to-G2 runs on the GPU (csrc/sha256.cuh); these
functions serve key derivation and the aggregation exponents, which hash a handful of
bytes per call."""
import hashlib

GROUP_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _b(m):
    return m if isinstance(m, (bytes, bytearray)) else m.encode("utf-8")


def hash256(m):
    return hashlib.sha256(_b(m)).digest()


def hash512(m):
    m = bytes(_b(m))
    return hash256(m + b"\x00") + hash256(m + b"\x01")


def hmac256(m, k):
    """HMAC-SHA256 with key k over message m (bls_py/util.py:19-33)"""
    m, k = bytes(_b(m)), bytes(_b(k))
    if len(k) > 64:
        k = hash256(k)
    k = k.ljust(64, b"\x00")
    inner = hash256(bytes(x ^ 0x36 for x in k) + m)
    return hash256(bytes(x ^ 0x5c for x in k) + inner)


DEVICE_HASH_PKS_MIN = 2048     # from this many exponents on, the per-key hashes run on the GPU


def hash_pks_bytes(num_outputs, public_keys):
    """hash_pks as num_outputs x 32 big-endian bytes (the scalar format of the GPU entry points).
    The hash over all keys is one sequential SHA-256 (host); the num_outputs per-key hashes and
    reductions mod n are independent and run on the GPU when there are many (row f2)."""
    blob = b"".join(pk if isinstance(pk, (bytes, bytearray)) else pk.serialize() for pk in public_keys)
    pk_hash = hash256(blob)
    if num_outputs >= DEVICE_HASH_PKS_MIN:
        from . import engine
        return engine.hash_pks(pk_hash, num_outputs).tobytes()
    return b"".join((int.from_bytes(hash256(i.to_bytes(4, "big") + pk_hash), "big") % GROUP_ORDER).to_bytes(32, "big")
                    for i in range(num_outputs))


def hash_pks(num_outputs, public_keys):
    """aggregation exponents T_i = H(i || H(pk_1 || ... || pk_n)) mod n (bls_py/util.py:36-50).
    `public_keys` are PublicKey objects or already-serialised 48-byte strings."""
    blob = b"".join(pk if isinstance(pk, (bytes, bytearray)) else pk.serialize() for pk in public_keys)
    pk_hash = hash256(blob)
    return [int.from_bytes(hash256(i.to_bytes(4, "big") + pk_hash), "big") % GROUP_ORDER
            for i in range(num_outputs)]
