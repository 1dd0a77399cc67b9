#!/usr/bin/env python3
"""Runs one embedded program a few times on random operands (timing / ncu target).

  python tools/prof_run.py pairing 56832 4 [reps]      # program, items, launch shape (ctas_per_sm id), repetitions
Under ncu:  ncu --set full --import-source on --clock-control none -k regex:vm -s 1 -c 1 -o gpurun_out/x python tools/prof_run.py ...
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200 import _lib, engine                           # noqa: E402

SPECS = {"pairing": [96, 192, 576], "miller_loop": [96, 192, 576], "final_exp": [576, 576], "f12_mul": [576, 576, 576],
         "fq2_mul_chain": [96, 96, 96], "f12_sqr": [576, 576, 576], "f2_mul": [96, 96, 96], "f2_sqr": [96, 96, 96],
         "g2_mul": [192, 32, 192], "g1_mul": [96, 32, 96], "hash_to_g2": [256, 192]}


def main():
    name = sys.argv[1]
    n = int(sys.argv[2])
    shape = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    _lib.init(0)
    _lib.check(_lib.lib.b200bls_set_ctas_per_sm(shape))
    spec = SPECS[name]
    bufs = [engine.DeviceBuffer(n * s) for s in spec]
    rng = np.random.default_rng(1)
    for b, s in zip(bufs[:-1], spec[:-1]):
        data = rng.integers(0, 256, size=n * s, dtype=np.uint8)
        if s % 48 == 0:
            data.reshape(n, -1)[:, ::48] &= 0x0f
        b.upload(data)
    best = 1e30
    for r in range(reps + 1):
        engine.timer_start()
        engine.run_program_dev(name, n, bufs, spec)
        ms = engine.timer_stop()
        if r > 0:
            best = min(best, ms)
    print("%s n=%d shape=%d kernel=%s: %.3f ms  %.4g items/s" % (name, n, shape, os.environ.get("B200BLS_KERNEL", "2"), best,
                                                                  n / (best * 1e-3)), flush=True)


if __name__ == "__main__":
    main()
