"""b200-bls: the python-bls (bls_py) pairing / aggregation hot path on NVIDIA B200.

Same entry points as the reference package (BLS.verify / aggregate_sigs / aggregate_pub_keys,
PrivateKey.sign, PublicKey / Signature serialisation, AggregationInfo, ate_pairing_multi),
served by libb200bls.so through ctypes.  Nothing here computes field arithmetic on the CPU:
without the CUDA library and a GPU every operation raises."""
__all__ = ["BLS", "PrivateKey", "PublicKey", "ExtendedPrivateKey", "ExtendedPublicKey", "Signature",
           "AggregationInfo", "Threshold", "ate_pairing_multi"]


def __getattr__(name):
    if name == "BLS":
        from .bls import BLS
        return BLS
    if name in ("PrivateKey", "PublicKey", "ExtendedPrivateKey", "ExtendedPublicKey"):
        from . import keys
        return getattr(keys, name)
    if name == "Signature":
        from .signature import Signature
        return Signature
    if name == "Threshold":
        from .threshold import Threshold
        return Threshold
    if name == "AggregationInfo":
        from .aggregation_info import AggregationInfo
        return AggregationInfo
    if name == "ate_pairing_multi":
        from .pairing import ate_pairing_multi
        return ate_pairing_multi
    raise AttributeError(name)
