#!/usr/bin/env python3
"""BASELINE.json configs 3, 4 and 5 on one GPU: full-size parity through size-independent
properties plus oracle spot checks, and device-timed throughput.  Writes gpurun_out/configs.json.

  config 3  aggregate 1 M G2 signatures + 1 M G1 keys (tree reduction)
  config 4  aggregate verify of 10,000 distinct messages (10,001 Miller loops, ONE final exp)
  config 5  batch verification of 500,000 independent signatures (this GPU's 1/8 of 4 M), 1% corrupted
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bls_oracle as O                                      # noqa: E402  (checker only)
from bls_b200 import _lib, engine, synth                    # noqa: E402
from bls_b200._lib import check, lib                        # noqa: E402
from bls_b200.programs.curve import G1_GEN                  # noqa: E402
from bls_b200.programs.hashg2 import G2_GEN                 # noqa: E402

_lib.init(0)
out = {}
scale = float(os.environ.get("B200BLS_CONFIG_SCALE", "1"))
g1 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in G1_GEN), dtype=np.uint8)
g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (G2_GEN[0] + G2_GEN[1])), dtype=np.uint8)


def dev_scalar_mul(base, scalars, g2flag):
    n = scalars.shape[0]
    w = 192 if g2flag else 96
    d_base = engine.DeviceBuffer(w * n).upload(np.tile(base, n))
    d_sc = engine.DeviceBuffer(32 * n).upload(scalars)
    d_out = engine.DeviceBuffer(w * n)
    fn = lib.b200bls_g2_scalar_mul_batch_dev if g2flag else lib.b200bls_g1_scalar_mul_batch_dev
    check(fn(d_base.ptr, d_sc.ptr, d_out.ptr, n))
    check(lib.b200bls_sync())
    d_base.free()
    d_sc.free()
    return d_out


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        engine.timer_start()
        fn()
        best = min(best, engine.timer_stop())
    return best


def ser1(p):
    return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")


def ser2(p):
    return b"".join(c.to_bytes(48, "big") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))


# ---- config 3 -------------------------------------------------------------------------------
n3 = int(1_000_000 * scale)
sc = synth.scalars(synth.SEED_AGGREGATE, n3)
tot = sum(int.from_bytes(bytes(r), "big") for r in sc) % O.N
res3 = {"n": n3}
for name, base, flag, G, ser in (("g2_signatures", g2, True, O.G2, ser2), ("g1_public_keys", g1, False, O.G1, ser1)):
    w = 192 if flag else 96
    d_pts = dev_scalar_mul(base, sc, flag)
    d_sum = engine.DeviceBuffer(w)
    fn = lib.b200bls_g2_sum_dev if flag else lib.b200bls_g1_sum_dev
    ms = timed(lambda: check(fn(d_pts.ptr, d_sum.ptr, n3)))
    got = d_sum.download().tobytes()
    want = ser(O.aff_mul(tot, G))
    res3[name] = {"ms": ms, "adds_per_s": (n3 - 1) / (ms * 1e-3), "parity_sum_identity": got == want}
    d_pts.free()
out["config3_aggregate"] = res3
print(json.dumps(res3), flush=True)

# ---- config 3, secure variant (row f3): sum T_i * P_i with the aggregation exponents T_i ---------
import hashlib                                              # noqa: E402
pk_hash = hashlib.sha256(b"config 3 secure").digest()
d_pkh = engine.DeviceBuffer(32).upload(np.frombuffer(pk_hash, dtype=np.uint8))
d_T = engine.DeviceBuffer(32 * n3)
ms_T = timed(lambda: check(lib.b200bls_hash_pks_dev(d_pkh.ptr, 0, d_T.ptr, n3)))
T_host = d_T.download().reshape(n3, 32)
t0 = time.perf_counter()
T_ref = [int.from_bytes(hashlib.sha256(i.to_bytes(4, "big") + pk_hash).digest(), "big") % O.N for i in range(n3)]
host_T_seconds = time.perf_counter() - t0
ks = [int.from_bytes(bytes(r), "big") for r in sc]
exp_ok = all(int.from_bytes(bytes(T_host[i]), "big") == T_ref[i] for i in range(0, n3, max(1, n3 // 5000)))
dot = sum(k * t for k, t in zip(ks, T_ref)) % O.N
res3s = {"n": n3, "exponents": {"ms_device": ms_T, "seconds_host_hashlib": host_T_seconds, "parity_sampled": bool(exp_ok)}}
for name, base, flag, G, ser in (("g2_signatures", g2, True, O.G2, ser2), ("g1_public_keys", g1, False, O.G1, ser1)):
    w = 192 if flag else 96
    d_pts = dev_scalar_mul(base, sc, flag)
    d_sum = engine.DeviceBuffer(w)
    fn = lib.b200bls_g2_msm_dev if flag else lib.b200bls_g1_msm_dev
    ms = timed(lambda: check(fn(d_pts.ptr, d_T.ptr, d_sum.ptr, n3)), reps=2)
    got = d_sum.download().tobytes()
    # the previous path: one ladder per point, then the reduction
    d_mul = engine.DeviceBuffer(w * n3)
    mul = lib.b200bls_g2_scalar_mul_batch_dev if flag else lib.b200bls_g1_scalar_mul_batch_dev
    red = lib.b200bls_g2_sum_dev if flag else lib.b200bls_g1_sum_dev
    d_sum2 = engine.DeviceBuffer(w)

    def ladder():
        check(mul(d_pts.ptr, d_T.ptr, d_mul.ptr, n3))
        check(red(d_mul.ptr, d_sum2.ptr, n3))
    ms_ladder = timed(ladder, reps=1)
    res3s[name] = {"ms_msm": ms, "points_per_s": n3 / (ms * 1e-3), "ms_ladders_plus_sum": ms_ladder,
                   "parity_identity": got == ser(O.aff_mul(dot, G)),
                   "ladder_path_agrees": d_sum2.download().tobytes() == got}
    d_pts.free()
    d_mul.free()
out["config3_secure_aggregate"] = res3s
print(json.dumps(res3s), flush=True)

# ---- config 4 -------------------------------------------------------------------------------
n4 = int(10_000 * scale)
sks = synth.scalars(synth.SEED_AGG_VERIFY, n4)
hs = synth.message_hashes(synth.SEED_AGG_VERIFY, n4)
H = engine.hash_to_g2(hs)
sigs = engine.scalar_mul(H, sks, True)
agg = engine.point_sum(sigs, True)
pks = engine.scalar_mul(np.tile(g1, n4), sks, False)
t0 = time.perf_counter()
ok = engine.aggregate_verify(agg, pks, hs)
wall = time.perf_counter() - t0
best = 1e30
for _ in range(3):
    t0 = time.perf_counter()
    ok = ok and engine.aggregate_verify(agg, pks, hs)
    best = min(best, time.perf_counter() - t0)
hs_bad = hs.copy()
hs_bad[7] = hs[8]
rejected = not engine.aggregate_verify(agg, pks, hs_bad)
# oracle spot check: a 3-message aggregate built from the same keys
sub = [0, 1, 2]
osig = O.g2_sum([O.sign_prehashed(int.from_bytes(bytes(sks[i]), "big"), bytes(hs[i])) for i in sub])
oracle_ok = O.aggregate_verify([O.pk_of(int.from_bytes(bytes(sks[i]), "big")) for i in sub], [bytes(hs[i]) for i in sub], osig)
gpu_small = engine.aggregate_verify(engine.point_sum(sigs[:3 * 192], True), pks[:3 * 96], hs[:3])
out["config4_aggregate_verify"] = {"n_messages": n4, "accepts": bool(ok), "rejects_swapped_message": bool(rejected),
                                   "oracle_3msg_agrees": bool(oracle_ok and gpu_small),
                                   "seconds_host_to_bool": best, "miller_loops_per_s": (n4 + 1) / best,
                                   "first_call_seconds": wall}
print(json.dumps(out["config4_aggregate_verify"]), flush=True)

# ---- config 4, many jobs in flight: 16 aggregate verifications of n4 messages each on the library's streams
# (the throughput shape: a 10,000-message job is then 27 wide CTAs, so four jobs share the GPU; the
# default picks the lowest-latency shape, whose CTAs take a whole SM each)
jobs = [(agg, pks, hs)] * 32
out["config4_concurrent_jobs"] = {"jobs": len(jobs), "n_messages_each": n4}
for shape in (0, 4):
    check(lib.b200bls_set_ctas_per_sm(shape))
    engine.aggregate_verify_many(jobs[:8])
    dt = 1e30
    for _ in range(4):          # host-side enqueueing jitters (0.30 - 0.7 s for the same 32 jobs): best of 4
        t0 = time.perf_counter()
        res = engine.aggregate_verify_many(jobs)
        dt = min(dt, time.perf_counter() - t0)
    out["config4_concurrent_jobs"]["shape_%d" % shape] = {"all_accept": all(res), "seconds": dt, "jobs_per_s": len(jobs) / dt,
                                                          "miller_loops_per_s": len(jobs) * (n4 + 1) / dt}
check(lib.b200bls_set_ctas_per_sm(0))
print(json.dumps(out["config4_concurrent_jobs"]), flush=True)

# ---- config 4 at a batch that fills the GPU (the 10,000-message case is latency bound: one
# under-filled pass per stage plus ONE single-thread final exponentiation) -----------------------
n4b = int(400_000 * scale)
sks = synth.scalars(synth.SEED_AGG_VERIFY + 1, n4b)
hs = synth.message_hashes(synth.SEED_AGG_VERIFY + 1, n4b)
sigs = engine.scalar_mul(engine.hash_to_g2(hs), sks, True)
agg = engine.point_sum(sigs, True)
pks = engine.scalar_mul(np.tile(g1, n4b), sks, False)
check(lib.b200bls_set_ctas_per_sm(3))
ok = engine.aggregate_verify(agg, pks, hs)
best = 1e30
for _ in range(2):
    t0 = time.perf_counter()
    ok = ok and engine.aggregate_verify(agg, pks, hs)
    best = min(best, time.perf_counter() - t0)
hs_bad = hs.copy()
hs_bad[7] = hs[8]
out["config4_large_aggregate_verify"] = {"n_messages": n4b, "accepts": bool(ok),
                                         "rejects_swapped_message": not engine.aggregate_verify(agg, pks, hs_bad),
                                         "seconds_host_to_bool": best, "miller_loops_per_s": (n4b + 1) / best}
check(lib.b200bls_set_ctas_per_sm(0))
print(json.dumps(out["config4_large_aggregate_verify"]), flush=True)

# ---- config 5 -------------------------------------------------------------------------------
n5 = int(500_000 * scale)
sks = synth.scalars(synth.SEED_BATCH_VERIFY, n5)
hs = synth.message_hashes(synth.SEED_BATCH_VERIFY, n5)
d_hs = engine.DeviceBuffer(32 * n5).upload(hs)
d_H = engine.DeviceBuffer(192 * n5)
check(lib.b200bls_hash_to_g2_batch_dev(d_hs.ptr, d_H.ptr, n5))
d_sk = engine.DeviceBuffer(32 * n5).upload(sks)
d_sig = engine.DeviceBuffer(192 * n5)
check(lib.b200bls_g2_scalar_mul_batch_dev(d_H.ptr, d_sk.ptr, d_sig.ptr, n5))
check(lib.b200bls_sync())
d_pk = dev_scalar_mul(g1, sks, False)
sig_host = d_sig.download().reshape(n5, 192)
bad = synth.corrupted_indices(synth.SEED_BATCH_VERIFY, n5)
for i in bad:
    sig_host[i] = sig_host[(i + 1) % n5] if (i + 1) % n5 not in bad else sig_host[(i + 2) % n5]
d_sig.upload(sig_host)
d_ok = engine.DeviceBuffer(n5)
ms = timed(lambda: check(lib.b200bls_verify_batch_dev(d_pk.ptr, d_hs.ptr, d_sig.ptr, d_ok.ptr, n5)))
res = d_ok.download()
want = np.ones(n5, dtype=np.uint8)
want[bad] = 0
# oracle on 4 sampled indices (2 good, 2 corrupted)
pk_host = d_pk.download().reshape(n5, 96)
oracle_agree = True
good = [i for i in range(n5) if want[i]][:2]
for i in good + list(bad[:2]):
    pk = (int.from_bytes(bytes(pk_host[i][:48]), "big"), int.from_bytes(bytes(pk_host[i][48:]), "big"), False)
    s = sig_host[i]
    sig = ((int.from_bytes(bytes(s[:48]), "big"), int.from_bytes(bytes(s[48:96]), "big")),
           (int.from_bytes(bytes(s[96:144]), "big"), int.from_bytes(bytes(s[144:]), "big")), False)
    oracle_agree = oracle_agree and (O.verify(pk, bytes(hs[i]), sig) == bool(res[i]))
out["config5_batch_verify"] = {"n": n5, "corrupted": int(len(bad)), "ms": ms,
                               "signatures_per_s": n5 / (ms * 1e-3),
                               "all_booleans_match_ground_truth": bool(np.array_equal(res, want)),
                               "oracle_sample_agrees": bool(oracle_agree)}
print(json.dumps(out["config5_batch_verify"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as fh:
    json.dump(out, fh, indent=1)
