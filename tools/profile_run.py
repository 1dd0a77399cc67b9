#!/usr/bin/env python3
"""Small fixed workload for ncu: three launches of the pairing program over one full wave
(default: the wide shape, one 384-thread CTA per SM), inputs resident on the device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200 import _lib, engine                           # noqa: E402

_lib.init(0)
if _lib.lib.b200bls_get_ctas_per_sm() == 0:
    _lib.check(_lib.lib.b200bls_set_ctas_per_sm(4))
shape = _lib.lib.b200bls_get_ctas_per_sm()
n = _lib.lib.b200bls_sm_count() * 128 * min(shape, 3)      # one full wave (shape 4 = one 384-thread CTA per SM)
n = int(os.environ.get("B200BLS_PROFILE_N", n))            # e.g. 65536 = bench.py's batch
rng = np.random.default_rng(7)
P = rng.integers(0, 256, size=(n, 96), dtype=np.uint8)
Q = rng.integers(0, 256, size=(n, 192), dtype=np.uint8)
P[:, ::48] &= 0x0f
Q[:, ::48] &= 0x0f
dP, dQ, dO = engine.DeviceBuffer(96 * n).upload(P), engine.DeviceBuffer(192 * n).upload(Q), engine.DeviceBuffer(576 * n)
for _ in range(3):
    engine.timer_start()
    _lib.check(_lib.lib.b200bls_pairing_batch_dev(dP.ptr, dQ.ptr, dO.ptr, n))
    print("pairing n=%d: %.3f ms" % (n, engine.timer_stop()))
