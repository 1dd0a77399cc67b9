import sys, time
sys.path.insert(0, "python-bls_b200")
import numpy as np
from bls_b200 import _lib, engine, synth, workloads as W
_lib.init(0)
for g2 in (True, False):
    w = 192 if g2 else 96
    d = W.dev_scalar_mul(synth.scalars(5, 1024), g2)
    pts = d.download()
    for n in (1, 2, 3, 6, 7, 8, 128, 1024, 1025, 4096):
        sub = pts[:w * n].copy()
        engine.point_sum(sub, g2)
        t0 = time.perf_counter()
        for _ in range(20):
            r = engine.point_sum(sub, g2)
        dt = (time.perf_counter() - t0) / 20
        print("g2" if g2 else "g1", n, "%.3f ms per host-to-host sum" % (dt * 1e3), flush=True)
