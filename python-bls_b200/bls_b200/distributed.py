"""Multi-GPU layer: one process per GPU, contiguous batch slices, no data-path collective.

Independent units (pairings, verifications) need no exchange at all.  Reductions exchange one
tiny partial per rank -- a 576-byte Miller product (ate_pairing_multi / aggregate verification)
or one affine point (signature / public-key aggregation) -- with a single all_gather, then
every rank finishes locally (one Fq12 product chain + ONE final exponentiation, or one point
sum).  all_reduce cannot be used: the group operations are not built-in reductions
(SURVEY.md 8e).  `torch.distributed` is plumbing only: with the gloo backend the gather goes
through host memory, with nccl through NVLink; payloads are <= 4.6 KB per job, so latency,
not bandwidth, decides (bench.py reports both when run under torchrun).
"""
import numpy as np


def shard_range(n, rank, world):
    """contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_bytes(payload, dist=None):
    """all_gather of equally sized byte strings -> list in rank order (on every rank)"""
    payload = bytes(payload)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [payload]
    import torch
    world = dist.get_world_size()
    device = "cpu"
    if dist.get_backend() == "nccl":
        device = "cuda:%d" % torch.cuda.current_device()
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def combine_miller_partials(partials, f12_mul, final_exp):
    """product of the per-rank Miller products, then one final exponentiation"""
    acc = partials[0]
    for p in partials[1:]:
        acc = f12_mul(acc, p)
    return final_exp(acc)


def pairing_multi(P, Q, dist=None):
    """ate_pairing_multi over pairs sharded across ranks: P, Q are THIS rank's slice.
    Returns the same 576 bytes on every rank."""
    from . import engine
    part = engine.miller_product(P, Q).tobytes()
    parts = gather_bytes(part, dist)
    return combine_miller_partials(
        parts,
        lambda a, b: engine.field_op(12, "mul", a, b).tobytes(),
        lambda f: engine.final_exp_batch(f).tobytes())


def aggregate_verify(sig, pks, hashes, rank=0, dist=None):
    """aggregate verification with (pk_i, message hash_i) sharded across ranks: `pks`, `hashes` are
    THIS rank's slice, `sig` the aggregate signature (used by rank 0 only).  Every rank hashes and
    pairs its slice in one fused launch, the 576-byte Miller products are gathered, multiplied and
    final-exponentiated once.  Returns the same bool on every rank."""
    from . import engine
    part = engine.aggregate_miller(sig if rank == 0 else None, pks, hashes).tobytes()
    parts = gather_bytes(part, dist)
    res = combine_miller_partials(
        parts,
        lambda a, b: engine.field_op(12, "mul", a, b).tobytes(),
        lambda f: engine.final_exp_batch(f).tobytes())
    return res == (1).to_bytes(48, "big") + bytes(528)


def point_sum(points, g2, dist=None):
    """sum of points sharded across ranks (aggregate_sigs_simple / aggregate_pub_keys)"""
    from . import engine
    part = engine.point_sum(points, g2).tobytes()
    parts = gather_bytes(part, dist)
    if len(parts) == 1:
        return parts[0]
    return engine.point_sum(b"".join(parts), g2).tobytes()


def secure_sum(points, pk_hash, first_index, g2, dist=None):
    """secure aggregation sum_i T_i * P_i sharded across ranks (bls.py:29-56, 217-221): `points` is
    THIS rank's contiguous slice starting at global index `first_index`; the exponents
    T_i = H(i || pk_hash) mod n are computed on the device for exactly that index range, the
    slice goes through one multi-scalar multiplication, and the per-rank points are gathered
    and summed like plain partial sums."""
    from . import engine
    w = 192 if g2 else 96
    n = np.asarray(points).size // w if not isinstance(points, (bytes, bytearray)) else len(points) // w
    ts = engine.hash_pks(pk_hash, n, first=first_index)
    part = engine.msm(points, ts, g2).tobytes()
    parts = gather_bytes(part, dist)
    if len(parts) == 1:
        return parts[0]
    return engine.point_sum(b"".join(parts), g2).tobytes()


def verify_batch(pks, hashes, sigs, dist=None):
    """independent verifications: every rank checks its own slice, no exchange"""
    from . import engine
    return engine.verify_batch(pks, hashes, sigs)
