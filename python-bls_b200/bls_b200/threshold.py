"""Threshold (T-of-N Joint-Feldman) helpers with the reference's interface
(bls_py/threshold.py:8-136).  Scalar work (Lagrange coefficients mod n) stays on the host as
plain ints; the curve work -- the commitment check and the weighted signature sum -- is one
batched scalar multiplication and one point sum on the GPU."""
from . import ec
from .signature import Signature
from .util import GROUP_ORDER


class Threshold:
    @staticmethod
    def lagrange_coeffs_at_zero(X, n=GROUP_ORDER):
        """L_i with P(0) = sum L_i P(X[i]) for a degree len(X)-1 polynomial (threshold.py:57-89,
        second barycentric form).  Returns ints mod n."""
        X = [int(x) for x in X]
        k = len(X)
        assert len(set(X)) == k and all(0 != x < n for x in X)
        shifts = []
        for j in range(k):
            w = 1
            for i in range(k):
                if i != j:
                    w = w * (X[j] - X[i]) % n
            shifts.append(pow(w * (-X[j]) % n, -1, n))
        inv_den = pow(sum(shifts) % n, -1, n)
        return [s * inv_den % n for s in shifts]

    @staticmethod
    def interpolate_at_zero(X, Y, n=GROUP_ORDER):
        """P(0) from the points (X[i], Y[i]) (threshold.py:92-102)"""
        return sum(l * int(y) for l, y in zip(Threshold.lagrange_coeffs_at_zero(X, n), Y)) % n

    @staticmethod
    def verify_secret_fragment(T, secret_fragment, player, commitment):
        """g1 * fragment == sum_k commitment[k] * player^k (threshold.py:105-125)"""
        assert len(commitment) == T
        assert int(secret_fragment) != 0
        assert player != 0
        lhs = ec.generator_Fq() * (int(secret_fragment) % GROUP_ORDER)
        powers = [pow(player, k, GROUP_ORDER) for k in range(len(commitment))]
        rhs = ec.sum_points(ec.scalar_mul_many(list(commitment), powers, False), False)
        return lhs == rhs

    @staticmethod
    def aggregate_unit_sigs(signatures, players, T):
        """sum_i lambda_i * sig_i (threshold.py:128-136)"""
        lambs = Threshold.lagrange_coeffs_at_zero(players)
        pts = ec.scalar_mul_many([s.value for s in signatures], lambs, True)
        return Signature.from_g2(ec.sum_points(pts, True))
