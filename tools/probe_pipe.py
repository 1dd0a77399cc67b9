#!/usr/bin/env python3
"""How many warps per scheduler does the IMAD.WIDE pipe need?  Microbenchmark variants 3 (one serial carry chain per
thread) and 6 (four independent chains, the parallelism of a Montgomery round) at 1..16 warps per scheduler, one CTA
per SM.  Writes gpurun_out/probe_pipe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200 import _lib, engine                           # noqa: E402

_lib.init(0)
out = {}
peak = 148 * 32 * 1.965e9
for variant in (3, 6):
    for blocks, threads in ((1, 128), (1, 256), (1, 384), (1, 512), (1, 768), (1, 1024), (2, 384), (3, 128), (2, 1024), (8, 256)):
        ops, ms = engine.microbench_imad(variant, blocks, threads, 200)
        key = "v%d_b%d_t%d" % (variant, blocks, threads)
        out[key] = {"warps_per_scheduler": blocks * threads / 128.0, "limb_products_per_s": ops, "of_nominal": ops / peak}
        print(key, "%.1f warps/sched" % (blocks * threads / 128.0), "%.3e" % ops, "%.3f" % (ops / peak), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_pipe.json"), "w"), indent=1)
