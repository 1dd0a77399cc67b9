"""Signature with the reference's interface (bls_py/signature.py:8-130); divide_by is out of
scope."""
from . import ec


class Signature:
    SIGNATURE_SIZE = 96

    def __init__(self, value, aggregation_info=None):
        self.value = value                      # ec.Point on G2
        self.aggregation_info = aggregation_info

    @staticmethod
    def from_bytes(buffer, aggregation_info=None):
        return Signature(ec.point_from_bytes(bytes(buffer), True), aggregation_info)

    @staticmethod
    def from_g2(g2_el, aggregation_info=None):
        return Signature(g2_el, aggregation_info)

    def set_aggregation_info(self, aggregation_info):
        self.aggregation_info = aggregation_info

    def get_aggregation_info(self):
        return self.aggregation_info

    def serialize(self):
        return self.value.serialize()

    def size(self):
        return self.SIGNATURE_SIZE

    def __eq__(self, other):
        return self.serialize() == other.serialize()

    def __hash__(self):
        return int.from_bytes(self.serialize(), "big")

    def __lt__(self, other):
        return self.serialize() < other.serialize()

    def __repr__(self):
        return "Signature(%s)" % self.serialize().hex()

    __str__ = __repr__
