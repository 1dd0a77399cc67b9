"""Multi-GPU layer: one process per GPU, contiguous batch slices.

Independent units (pairings, verifications) need no exchange at all.  Reductions exchange one tiny partial per
rank -- a 576-byte Miller product (ate_pairing_multi / aggregate verification) or one affine point (signature /
public-key aggregation) -- with a single all-gather, then every rank finishes locally (Fq12 products + ONE final
exponentiation, or one point sum).  all_reduce cannot be used: the group operations are not built-in reductions.

The exchange is native (csrc/comm.cuh, no torch): `ncclAllGather` over NVLink on the library stream, NCCL loaded
with dlopen, or a host gather through a POSIX shared-memory segment of the node; payloads are <= 4.6 KB per job,
so latency decides, and bench.py reports both.  The launcher's environment (RANK, WORLD_SIZE, LOCAL_RANK,
MASTER_PORT -- what `python -m torch.distributed.run` sets) is all that is needed.
"""
import os

import numpy as np

from . import _lib
from ._lib import as_u8, check, lib, ptr

_state = {"rank": 0, "world": 1, "ready": False}


def shard_range(n, rank, world):
    """contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def nccl_library():
    """the NCCL the image ships next to PyTorch (nvidia-nccl wheel), else None: the loader then tries the system's"""
    try:
        import nvidia.nccl as pkg                    # a namespace package: no torch import
        for base in pkg.__path__:
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except ImportError:
        pass
    return None


def init(rank=None, world=None, key=None, gpu=True):
    """join the ranks of this node; rank / world default to the launcher's RANK / WORLD_SIZE.  gpu=False joins the
    host gather only (CPU-side tests of the multi-rank logic)."""
    if _state["ready"]:
        return _state["rank"], _state["world"]
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    key = key or os.environ.get("MASTER_PORT", "default")
    if gpu:
        _lib.init()
    path = nccl_library() if gpu else None
    check(lib.b200bls_comm_init(rank, world, str(key).encode(), path.encode() if path else None))
    _state.update(rank=rank, world=world, ready=True)
    return rank, world


def shutdown():
    if _state["ready"]:
        lib.b200bls_comm_shutdown()
        _state.update(rank=0, world=1, ready=False)


def has_nccl():
    return bool(lib.b200bls_comm_has_nccl())


def gather_bytes(payload):
    """host all-gather of equally sized byte strings (<= 4096 bytes) -> list in rank order, on every rank"""
    payload = as_u8(payload)
    world = _state["world"]
    out = np.empty(world * payload.size, dtype=np.uint8)
    check(lib.b200bls_allgather_host(ptr(payload), ptr(out), payload.size))
    return [out[r * payload.size:(r + 1) * payload.size].tobytes() for r in range(world)]


def combine_miller_partials(partials, f12_mul, final_exp):
    """product of the per-rank Miller products, then one final exponentiation"""
    acc = partials[0]
    for p in partials[1:]:
        acc = f12_mul(acc, p)
    return final_exp(acc)


def pairing_multi(P, Q):
    """ate_pairing_multi over pairs sharded across ranks: P, Q are THIS rank's slice (possibly empty).
    Returns the same 576 bytes on every rank."""
    from . import engine
    part = engine.miller_product(P, Q).tobytes()
    return combine_miller_partials(
        gather_bytes(part),
        lambda a, b: engine.field_op(12, "mul", a, b).tobytes(),
        lambda f: engine.final_exp_batch(f).tobytes())


def aggregate_verify(sig, pks, hashes, use_nccl=False):
    """aggregate verification with (pk_i, message hash_i) sharded across ranks: `pks`, `hashes` are THIS rank's
    slice (possibly empty), `sig` the aggregate signature (read on rank 0).  One fused hash-and-pair launch per
    rank, one all-gather of 576 bytes, the product and ONE final exponentiation on every rank -> the same bool
    everywhere (b200bls_aggregate_verify_sharded)."""
    _lib.init()
    pks, hashes = as_u8(pks), as_u8(hashes)
    n = hashes.size // 32
    if pks.size != 96 * n or hashes.size != 32 * n:
        raise ValueError("bad buffer sizes")
    s = as_u8(sig, 192) if (sig is not None and _state["rank"] == 0) else None
    if _state["rank"] == 0 and s is None:
        raise ValueError("rank 0 needs the aggregate signature")
    ok = np.zeros(1, dtype=np.uint8)
    check(lib.b200bls_aggregate_verify_sharded(ptr(s) if s is not None else None, ptr(pks) if n else None,
                                               ptr(hashes) if n else None, n, int(bool(use_nccl)), ptr(ok)))
    return bool(ok[0])


def point_sum(points, g2, use_nccl=False):
    """sum of points sharded across ranks (aggregate_sigs_simple / aggregate_pub_keys): `points` is THIS rank's
    slice as host bytes or an engine.DeviceBuffer holding affine points -> the total as bytes on every rank"""
    from . import engine
    w = 192 if g2 else 96
    if isinstance(points, engine.DeviceBuffer):
        dev, n, own = points, points.nbytes // w, False
    else:
        host = as_u8(points)
        n = host.size // w
        dev, own = engine.DeviceBuffer(max(host.size, 1)), True
        if n:
            dev.upload(host)
    out = np.empty(w, dtype=np.uint8)
    check(lib.b200bls_point_sum_sharded_dev(int(bool(g2)), dev.ptr, n, int(bool(use_nccl)), ptr(out)))
    if own:
        dev.free()
    return out.tobytes()


def secure_sum(points, pk_hash, first_index, g2):
    """secure aggregation sum_i T_i * P_i sharded across ranks (bls.py:29-56, 217-221): `points` is THIS rank's
    contiguous slice starting at global index `first_index`; the exponents T_i = H(i || pk_hash) mod n are
    computed on the device for exactly that index range, the slice goes through one multi-scalar multiplication,
    and the per-rank points are gathered and summed like plain partial sums."""
    from . import engine
    w = 192 if g2 else 96
    pts = as_u8(points)
    n = pts.size // w
    if n:
        ts = engine.hash_pks(pk_hash, n, first=first_index)
        part = engine.msm(pts, ts, g2).tobytes()
    else:
        part = bytes(w)
    parts = gather_bytes(part)
    if len(parts) == 1:
        return parts[0]
    return engine.point_sum(b"".join(parts), g2).tobytes()


def verify_batch(pks, hashes, sigs):
    """independent verifications: every rank checks its own slice, no exchange"""
    from . import engine
    return engine.verify_batch(pks, hashes, sigs)
