"""GPU: two ranks (one process each, the library's native communicator) compute a sharded ate_pairing_multi, a
sharded aggregate verification and a sharded signature aggregation over both transports (host gather, and NCCL
when two GPUs are present); every rank must end with the single-process result.  On a one-GPU box both ranks share
cuda:0 (independent launches, nothing waits on the other rank inside a kernel) and only the host gather runs --
NCCL refuses two ranks on one device."""
import multiprocessing as mp
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, key, n_gpus, ret):
    for sub in ("python-bls_b200", "oracle"):
        sys.path.insert(0, os.path.join(ROOT, sub))
    import numpy as np
    import bls_oracle as O
    from bls_b200 import _lib, distributed as D, engine, synth
    _lib.init(rank % max(1, n_gpus))
    if n_gpus < world:
        os.environ["B200BLS_NO_NCCL"] = "1"
    D.init(rank, world, key)
    try:
        n = 301
        a, b = synth.scalars(101, n), synth.scalars(102, n)
        g1 = np.frombuffer(O.G1[0].to_bytes(48, "big") + O.G1[1].to_bytes(48, "big"), dtype=np.uint8)
        g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (O.G2[0] + O.G2[1])), dtype=np.uint8)
        P = engine.scalar_mul(np.tile(g1, n), a, False)
        Q = engine.scalar_mul(np.tile(g2, n), b, True)
        lo, hi = D.shard_range(n, rank, world)
        ok = D.pairing_multi(P[96 * lo:96 * hi], Q[192 * lo:192 * hi]) == engine.pairing_multi(P, Q).tobytes()
        transports = [False] + ([True] if D.has_nccl() else [])
        want_sum = engine.point_sum(Q, True).tobytes()
        # aggregate verification over distinct messages: sk_i = a_i, sig = sum a_i H(m_i)
        hs = synth.message_hashes(103, n)
        sigs = engine.scalar_mul(engine.hash_to_g2(hs), a, True)
        agg = engine.point_sum(sigs, True)
        hs_bad = hs.copy()
        hs_bad[5] = hs[6]
        for nccl in transports:
            ok = ok and D.point_sum(Q[192 * lo:192 * hi], True, use_nccl=nccl) == want_sum
            ok = ok and D.aggregate_verify(agg, P[96 * lo:96 * hi], hs[lo:hi], use_nccl=nccl) is True
            ok = ok and D.aggregate_verify(agg, P[96 * lo:96 * hi], hs_bad[lo:hi], use_nccl=nccl) is False
        # a rank with an empty slice still takes part (1 message over 2 ranks)
        one = D.shard_range(1, rank, world)
        ok = ok and D.aggregate_verify(sigs[:192], P[96 * one[0]:96 * one[1]], hs[one[0]:one[1]]) is True
        ret.put((rank, bool(ok), D.has_nccl()))
    finally:
        D.shutdown()


def test_two_rank_sharded_reductions():
    from bls_b200 import _lib
    import ctypes
    n_gpus = ctypes.c_int(0)
    try:
        cudart = ctypes.CDLL("libcudart.so.12")
        cudart.cudaGetDeviceCount(ctypes.byref(n_gpus))
    except OSError:
        n_gpus.value = 1
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    key = "gputest%d" % os.getpid()
    procs = [ctx.Process(target=_worker, args=(r, 2, key, max(1, n_gpus.value), ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = sorted(ret.get() for _ in range(2))
    assert [g[:2] for g in got] == [(0, True), (1, True)]
