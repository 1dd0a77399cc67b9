#!/usr/bin/env python3
"""Experiment: pairing throughput for each launch shape (CTAs per SM) at a batch size that is a
whole number of waves for all of them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200 import _lib, engine                           # noqa: E402

_lib.init(0)
sm = _lib.lib.b200bls_sm_count()
n = sm * 128 * 12
rng = np.random.default_rng(7)
P = rng.integers(0, 256, size=(n, 96), dtype=np.uint8)
Q = rng.integers(0, 256, size=(n, 192), dtype=np.uint8)
P[:, ::48] &= 0x0f
Q[:, ::48] &= 0x0f
dP, dQ, dO = engine.DeviceBuffer(96 * n).upload(P), engine.DeviceBuffer(192 * n).upload(Q), engine.DeviceBuffer(576 * n)
progs = sys.argv[1:] or ["pairing"]
for prog in progs:
    for ctas in (1, 2, 3, 4):
        _lib.check(_lib.lib.b200bls_set_ctas_per_sm(ctas))
        best = 1e9
        for _ in range(3):
            engine.timer_start()
            engine.run_program_dev(prog, n, [dP, dQ, dO], [96, 192, 576])
            best = min(best, engine.timer_stop())
        print("%s ctas/SM=%d n=%d: %.3f ms  %.0f items/s" % (prog, ctas, n, best, n / best * 1e3), flush=True)
