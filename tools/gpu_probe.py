#!/usr/bin/env python3
"""GPU box probe: integer-multiply peak (the roofline denominator) and first kernel timings.
Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200 import _lib, engine                           # noqa: E402

out = {}
_lib.init(0)
out["sm_count"] = _lib.lib.b200bls_sm_count()
names = {0: "imad_lo", 1: "imad_hi", 2: "imad_wide", 3: "imad_wide_x_chain"}
mb = {}
for variant in range(4):
    for bps, thr in ((8, 256), (1, 128), (1, 256), (3, 128)):
        ops, ms = engine.microbench_imad(variant, bps, thr, 100)
        mb["%s_b%d_t%d" % (names[variant], bps, thr)] = {"ops_per_s": ops, "ms": ms}
        print(names[variant], bps, thr, "%.3e ops/s" % ops, "%.3f ms" % ms, flush=True)
out["microbench"] = mb


def time_prog(name, n, bufs_spec, reps=3):
    bufs = [engine.DeviceBuffer(n * s) for s in bufs_spec]
    rng = np.random.default_rng(1)
    for b, s in zip(bufs[:-1], bufs_spec[:-1]):
        data = rng.integers(0, 256, size=n * s, dtype=np.uint8)
        data.reshape(n, -1)[:, ::48] &= 0x0f
        b.upload(data)
    best = 1e30
    for r in range(reps + 1):
        engine.timer_start()
        engine.run_program_dev(name, n, bufs, bufs_spec)
        ms = engine.timer_stop()
        if r > 0:
            best = min(best, ms)
    for b in bufs:
        b.free()
    return best


sm = out["sm_count"]
res = {}
for ctas in (1, 2):
  _lib.check(_lib.lib.b200bls_set_ctas_per_sm(ctas))
  for name, spec, n in (("fq2_mul_chain", [96, 96, 96], sm * 128 * 4),
                      ("f12_mul", [576, 576, 576], sm * 128 * 16),
                      ("f12_sqr", [576, 576, 576], sm * 128 * 16),
                      ("miller_loop", [96, 192, 576], sm * 128),
                      ("final_exp", [576, 576], sm * 128),
                      ("pairing", [96, 192, 576], sm * 128),
                      ("pairing", [96, 192, 576], sm * 256),
                      ("pairing", [96, 192, 576], 65536)):
    ms = time_prog(name, n, spec)
    res["%s_n%d_ctas%d" % (name, n, ctas)] = {"ms": ms, "items_per_s": n / (ms * 1e-3)}
    print("ctas", ctas, name, n, "%.3f ms" % ms, "%.1f items/s" % (n / (ms * 1e-3)), flush=True)
out["programs"] = res
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print("done")
