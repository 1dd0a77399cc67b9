#!/usr/bin/env python3
"""Probe for round-2 planning: FP64 FMA throughput next to the integer-multiply pipe (writes
gpurun_out/probe_fp64.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200 import _lib, engine                           # noqa: E402

_lib.init(0)
out = {}
for name, variant in (("imad_wide_x_chain", 3), ("dfma", 4), ("mixed_even_imad_odd_dfma", 5)):
    for bps, thr in ((8, 256), (3, 128), (1, 128)):
        ops, ms = engine.microbench_imad(variant, bps, thr, 100)
        out["%s_b%d_t%d" % (name, bps, thr)] = {"instructions_per_s": ops, "ms": ms}
        print(name, bps, thr, "%.3e /s" % ops, "%.3f ms" % ms, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_fp64.json"), "w"), indent=1)
