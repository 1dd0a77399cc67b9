#!/usr/bin/env python3
"""Multi-GPU reductions under torchrun (one process per GPU): BASELINE configs 3 and 4 sharded
over the ranks, with the per-rank partial (one affine point / one 576-byte Miller product)
combined by an all_gather -- timed with both the gloo (host) and the nccl (NVLink) backend.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_multi.py
Rank 0 prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from bls_b200 import _lib, distributed as D, engine, synth      # noqa: E402
from bls_b200.programs.curve import G1_GEN                      # noqa: E402
from bls_b200.programs.hashg2 import G2_GEN                     # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gloo = dist.new_group(backend="gloo")
_lib.init(local)
g1 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in G1_GEN), dtype=np.uint8)
g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (G2_GEN[0] + G2_GEN[1])), dtype=np.uint8)


class Group:
    """adapter: the tiny subset of torch.distributed that bls_b200.distributed uses, bound to a group"""

    def __init__(self, group, backend):
        self.group, self.backend = group, backend

    def is_initialized(self):
        return True

    def get_world_size(self):
        return world

    def get_backend(self):
        return self.backend

    def all_gather(self, out, t):
        return dist.all_gather(out, t, group=self.group)


def timed(fn, reps=3):
    best = 1e30
    res = None
    for _ in range(reps):
        dist.barrier()
        t0 = time.perf_counter()
        res = fn()
        _lib.check(_lib.lib.b200bls_sync())
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        dist.all_reduce(dt.cuda(), op=dist.ReduceOp.MAX)
        best = min(best, float(dt.item()))
    t = torch.tensor([best], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), res


out = {"n_gpus": world}
# ---- config 3: 1 M G2 signatures in total, contiguous slices --------------------------------
n3 = 1_000_000
lo, hi = D.shard_range(n3, rank, world)
sc = synth.scalars(synth.SEED_AGGREGATE, n3)[lo:hi]
pts = engine.scalar_mul(np.tile(g2, hi - lo), sc, True)
for name, grp in (("gloo", Group(gloo, "gloo")), ("nccl", Group(None, "nccl"))):
    sec, res = timed(lambda: D.point_sum(pts, True, grp))
    out["config3_g2_sum_%s" % name] = {"seconds": sec, "adds_per_s": (n3 - 1) / sec}
    ref = res
import bls_oracle as O                                         # noqa: E402  (checker)
tot = sum(int.from_bytes(bytes(r), "big") for r in synth.scalars(synth.SEED_AGGREGATE, n3)) % O.N
want = O.aff_mul(tot, O.G2)
out["config3_parity"] = ref == b"".join(c.to_bytes(48, "big") for c in (want[0][0], want[0][1], want[1][0], want[1][1]))
# ---- config 3, secure variant: sum T_i * sig_i over the same 1 M points (row f3) -------------------
import hashlib                                                 # noqa: E402
pk_hash = hashlib.sha256(b"config 3 secure").digest()
sec, res = timed(lambda: D.secure_sum(pts, pk_hash, lo, True, Group(None, "nccl")), reps=2)
ks = [int.from_bytes(bytes(r), "big") for r in synth.scalars(synth.SEED_AGGREGATE, n3)]
dot = sum(k * (int.from_bytes(hashlib.sha256(i.to_bytes(4, "big") + pk_hash).digest(), "big") % O.N)
          for i, k in enumerate(ks)) % O.N
want = O.aff_mul(dot, O.G2)
out["config3_secure_g2_msm_nccl"] = {"seconds": sec, "points_per_s": n3 / sec,
                                     "parity": res == b"".join(c.to_bytes(48, "big") for c in (want[0][0], want[0][1], want[1][0], want[1][1]))}
# ---- config 4 (large): 400,000 messages in total ------------------------------------------------
n4 = 400_000
lo, hi = D.shard_range(n4, rank, world)
sks = synth.scalars(synth.SEED_AGG_VERIFY + 1, n4)
hs = synth.message_hashes(synth.SEED_AGG_VERIFY + 1, n4)
H = engine.hash_to_g2(hs[lo:hi])
sig_part = engine.point_sum(engine.scalar_mul(H, sks[lo:hi], True), True)
agg = D.point_sum(sig_part, True, Group(gloo, "gloo"))          # the aggregate signature
pks = engine.scalar_mul(np.tile(g1, hi - lo), sks[lo:hi], False)
neg_g1 = O.aff_neg((O.G1[0], O.G1[1], False))
neg_g1_b = neg_g1[0].to_bytes(48, "big") + neg_g1[1].to_bytes(48, "big")
one = O.f12_serialize(O.F12_ONE)


def verify(grp):
    return D.aggregate_verify(agg, pks, hs[lo:hi], rank, grp)


for name, grp in (("gloo", Group(gloo, "gloo")), ("nccl", Group(None, "nccl"))):
    sec, ok = timed(lambda: verify(grp), reps=2)
    out["config4_aggregate_verify_%s" % name] = {"seconds": sec, "miller_loops_per_s": (n4 + 1) / sec, "accepts": bool(ok)}
# ---- config 5: 4 M independent signatures sharded over the ranks, 1 % corrupted -------------------
n5 = int(os.environ.get("B200BLS_CONFIG5_N", 4_000_000))
lo, hi = D.shard_range(n5, rank, world)
m = hi - lo
sks5 = synth.scalars(synth.SEED_BATCH_VERIFY, n5)[lo:hi]
hs5 = synth.message_hashes(synth.SEED_BATCH_VERIFY, n5)[lo:hi]
lib = _lib.lib
d_hs = engine.DeviceBuffer(32 * m).upload(hs5)
d_H = engine.DeviceBuffer(192 * m)
_lib.check(lib.b200bls_hash_to_g2_batch_dev(d_hs.ptr, d_H.ptr, m))
d_sk = engine.DeviceBuffer(32 * m).upload(sks5)
d_sig = engine.DeviceBuffer(192 * m)
_lib.check(lib.b200bls_g2_scalar_mul_batch_dev(d_H.ptr, d_sk.ptr, d_sig.ptr, m))
d_g = engine.DeviceBuffer(96 * m).upload(np.tile(g1, m))
d_pk = engine.DeviceBuffer(96 * m)
_lib.check(lib.b200bls_g1_scalar_mul_batch_dev(d_g.ptr, d_sk.ptr, d_pk.ptr, m))
_lib.check(lib.b200bls_sync())
sig_host = d_sig.download().reshape(m, 192)
bad = synth.corrupted_indices(synth.SEED_BATCH_VERIFY, n5)
bad = bad[(bad >= lo) & (bad < hi)] - lo
bad_set = set(int(i) for i in bad)
for i in bad:                       # another signer's valid signature: decodes, must fail the pairing check
    j = (int(i) + 1) % m
    while j in bad_set:
        j = (j + 1) % m
    sig_host[i] = sig_host[j]
d_sig.upload(sig_host)
d_ok = engine.DeviceBuffer(m)
_lib.check(lib.b200bls_verify_batch_dev(d_pk.ptr, d_hs.ptr, d_sig.ptr, d_ok.ptr, m))   # warm-up
_lib.check(lib.b200bls_sync())
dist.barrier()
engine.timer_start()
_lib.check(lib.b200bls_verify_batch_dev(d_pk.ptr, d_hs.ptr, d_sig.ptr, d_ok.ptr, m))
ms5 = engine.timer_stop()
res = d_ok.download()
want = np.ones(m, dtype=np.uint8)
want[bad] = 0
t5 = torch.tensor([ms5, 0.0 if np.array_equal(res, want) else 1.0, float(len(bad))], dtype=torch.float64, device="cuda")
tmax = t5.clone()
dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tsum = t5.clone()
dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
out["config5_batch_verify"] = {"n": n5, "corrupted": int(tsum[2].item()), "ms_max_over_ranks": float(tmax[0].item()),
                               "signatures_per_s": n5 / (float(tmax[0].item()) * 1e-3),
                               "all_booleans_match_ground_truth": bool(tmax[1].item() == 0.0)}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
