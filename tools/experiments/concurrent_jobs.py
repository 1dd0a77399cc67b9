import sys, time, os
import numpy as np
sys.path.insert(0, "python-bls_b200")
from bls_b200 import _lib, engine, synth
from bls_b200._lib import check, lib
from bls_b200.programs.curve import G1_GEN
_lib.init(0)
g1 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in G1_GEN), dtype=np.uint8)
n4 = 10000
sks = synth.scalars(synth.SEED_AGG_VERIFY, n4)
hs = synth.message_hashes(synth.SEED_AGG_VERIFY, n4)
H = engine.hash_to_g2(hs)
sigs = engine.scalar_mul(H, sks, True)
agg = engine.point_sum(sigs, True)
pks = engine.scalar_mul(np.tile(g1, n4), sks, False)
print("single", engine.aggregate_verify(agg, pks, hs))
jobs = [(agg, pks, hs)] * 32
for shape in (0, 4, 0, 4, 0, 4):
    check(lib.b200bls_set_ctas_per_sm(shape))
    engine.aggregate_verify_many(jobs[:8])
    t0 = time.perf_counter()
    res = engine.aggregate_verify_many(jobs)
    dt = time.perf_counter() - t0
    print(shape, all(res), "%.3f s  %.1f jobs/s" % (dt, 32 / dt), flush=True)
