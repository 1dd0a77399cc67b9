"""world_size-2 tests (CPU) of the multi-rank host logic: slice ownership, the all-gather of per-rank partials
and the local combine.  Two transports are exercised: the library's own host gather (POSIX shared memory,
csrc/comm.cuh -- needs no GPU) and, as an independent witness of the same protocol, a gloo all_gather.  The group
arithmetic of the combine is injected (oracle on the CPU here, the GPU engine in production)."""
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup():
    for sub in ("python-bls_b200", "oracle"):
        p = os.path.join(ROOT, sub)
        if p not in sys.path:
            sys.path.insert(0, p)


def _check_combine(D, gather, rank, world):
    import bls_oracle as O
    # 5 pairs sharded over 2 ranks: rank 0 owns [0, 3), rank 1 owns [3, 5)
    pairs = [(O.aff_mul(k + 2, O.G1), O.aff_mul(2 * k + 3, O.G2)) for k in range(5)]
    lo, hi = D.shard_range(len(pairs), rank, world)
    part = O.F12_ONE
    for p, q in pairs[lo:hi]:
        part = O.f12_mul(part, O.miller_loop(p, q))
    parts = gather(O.f12_serialize(part))
    assert len(parts) == world

    def de(b):
        return tuple(int.from_bytes(b[i:i + 48], "big") for i in range(0, 576, 48))
    res = D.combine_miller_partials(parts, lambda a, b: O.f12_serialize(O.f12_mul(de(a), de(b))),
                                    lambda f: O.f12_serialize(O.final_exp(de(f))))
    want = O.f12_serialize(O.ate_pairing_multi([p for p, _ in pairs], [q for _, q in pairs]))
    ok = res == want
    # point partials: rank-order gather of equally sized payloads, several rounds (slot banks alternate)
    for rnd in range(5):
        pts = gather(bytes([rank + 16 * rnd]) * 96)
        ok = ok and pts == [bytes([r + 16 * rnd]) * 96 for r in range(world)]
    return ok


def _native_worker(rank, world, key, ret):
    _setup()
    from bls_b200 import distributed as D
    D.init(rank, world, key, gpu=False)
    try:
        ok = _check_combine(D, D.gather_bytes, rank, world)
        # an empty slice takes part in the exchange like any other (3 items over 4 ranks leave one empty)
        assert [D.shard_range(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
        ret.put((rank, bool(ok)))
    finally:
        D.shutdown()


def _gloo_worker(rank, world, port, ret):
    _setup()
    import torch
    import torch.distributed as dist
    from bls_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def gather(payload):
        mine = torch.tensor(list(payload), dtype=torch.uint8)
        out = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(out, mine)
        return [bytes(t.numpy().tobytes()) for t in out]
    try:
        ret.put((rank, bool(_check_combine(D, gather, rank, world))))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    _setup()
    from bls_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 65536, 1000001):
        for world in (1, 2, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _run(target, arg):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    procs = [ctx.Process(target=target, args=(r, 2, arg, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(ret.get() for _ in range(2))
    assert got == [(0, True), (1, True)]


def test_two_rank_combine_over_the_native_host_gather():
    _run(_native_worker, "cputest%d" % os.getpid())


def test_two_rank_combine_over_gloo():
    _run(_gloo_worker, 29600 + os.getpid() % 300)
