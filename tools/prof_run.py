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


def pipeline(name, n, reps):
    """multi-launch entry points on seeded device-resident inputs: g1_sum / g2_sum (config 3), g1_msm / g2_msm (secure
    aggregation), verify (config 5's kernel pair), aggregate_verify (config 4)"""
    from bls_b200 import synth, workloads as W
    from bls_b200._lib import check, lib
    g2 = name.startswith("g2")
    if name.endswith("_sum") or name.endswith("_msm"):
        d_pts, cnt, _ = W.config3_slice(n, g2)
        w = 192 if g2 else 96
        d_out = engine.DeviceBuffer(w)
        if name.endswith("_sum"):
            fn = lib.b200bls_g2_sum_dev if g2 else lib.b200bls_g1_sum_dev
            run = lambda: check(fn(d_pts.ptr, d_out.ptr, cnt))
        else:
            d_sc = engine.DeviceBuffer(32 * n).upload(synth.scalars(77, n))
            fn = lib.b200bls_g2_msm_dev if g2 else lib.b200bls_g1_msm_dev
            run = lambda: check(fn(d_pts.ptr, d_sc.ptr, d_out.ptr, cnt))
    elif name == "verify":
        d_pk, d_hs, d_sig, want, _ = W.config5_inputs(n)
        d_ok = engine.DeviceBuffer(n)
        run = lambda: check(lib.b200bls_verify_batch_dev(d_pk.ptr, d_hs.ptr, d_sig.ptr, d_ok.ptr, n))
    elif name == "aggregate_verify":
        agg, pks, hs, _ = W.config4_inputs(n)
        run = lambda: engine.aggregate_verify(agg, pks, hs)
    else:
        raise SystemExit("unknown pipeline " + name)
    best = 1e30
    for r in range(reps + 1):
        engine.timer_start()
        run()
        ms = engine.timer_stop()
        if r > 0:
            best = min(best, ms)
    print("%s n=%d: %.3f ms  %.4g items/s" % (name, n, best, n / (best * 1e-3)), flush=True)


def main():
    name = sys.argv[1]
    n = int(sys.argv[2])
    shape = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    _lib.init(0)
    _lib.check(_lib.lib.b200bls_set_ctas_per_sm(shape))
    if name not in SPECS:
        return pipeline(name, n, reps)
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
