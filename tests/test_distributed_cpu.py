"""world_size-2 gloo test (CPU) of the multi-GPU host logic: slice ownership, the
all_gather of per-rank partials and the local combine.  The group arithmetic of the
combine is injected (oracle on the CPU here, the GPU engine in production)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for sub in ("python-bls_b200", "oracle"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    import bls_oracle as O
    from bls_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 5 pairs sharded over 2 ranks: rank 0 owns [0, 3), rank 1 owns [3, 5)
        pairs = [(O.aff_mul(k + 2, O.G1), O.aff_mul(2 * k + 3, O.G2)) for k in range(5)]
        lo, hi = D.shard_range(len(pairs), rank, world)
        part = O.F12_ONE
        for p, q in pairs[lo:hi]:
            part = O.f12_mul(part, O.miller_loop(p, q))
        parts = D.gather_bytes(O.f12_serialize(part), dist)
        assert len(parts) == world

        def de(b):
            return tuple(int.from_bytes(b[i:i + 48], "big") for i in range(0, 576, 48))
        res = D.combine_miller_partials(parts, lambda a, b: O.f12_serialize(O.f12_mul(de(a), de(b))),
                                        lambda f: O.f12_serialize(O.final_exp(de(f))))
        want = O.f12_serialize(O.ate_pairing_multi([p for p, _ in pairs], [q for _, q in pairs]))
        ok = res == want
        # point partials: rank-order gather of equally sized payloads
        pts = D.gather_bytes(bytes([rank]) * 96, dist)
        ok = ok and pts == [bytes([r]) * 96 for r in range(world)]
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(bool(t.item() == 1.0))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
    from bls_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 65536, 1000001):
        for world in (1, 2, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_multi_pairing_combine():
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True
