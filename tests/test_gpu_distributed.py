"""GPU: two ranks (one process each, gloo for the plumbing) compute a sharded ate_pairing_multi and
a sharded signature aggregation; every rank must end with the single-process result.  On a
one-GPU box both ranks share cuda:0 (independent launches, nothing waits on the other rank)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for sub in ("python-bls_b200", "oracle"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    import numpy as np
    import bls_oracle as O
    from bls_b200 import _lib, distributed as D, engine, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _lib.init(rank % max(1, torch.cuda.device_count()))
        n = 301
        a, b = synth.scalars(101, n), synth.scalars(102, n)
        g1 = np.frombuffer(O.G1[0].to_bytes(48, "big") + O.G1[1].to_bytes(48, "big"), dtype=np.uint8)
        g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (O.G2[0] + O.G2[1])), dtype=np.uint8)
        P = engine.scalar_mul(np.tile(g1, n), a, False)
        Q = engine.scalar_mul(np.tile(g2, n), b, True)
        lo, hi = D.shard_range(n, rank, world)
        got = D.pairing_multi(P[96 * lo:96 * hi], Q[192 * lo:192 * hi], dist)
        want = engine.pairing_multi(P, Q).tobytes()
        ok = got == want
        s = D.point_sum(Q[192 * lo:192 * hi], True, dist)
        ok = ok and s == engine.point_sum(Q, True).tobytes()
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(bool(t.item() == 1.0))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_reductions():
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get() is True
