#!/usr/bin/env python3
"""Small fixed workload for ncu: batch verification of one full wave of valid signatures (wide
shape).  Kernel launches in order: [setup: hash, ladder, ladder] then per verification:
sha_stage_kernel, vm_kernel(hash_to_g2), vm_kernel(verify_pair)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200 import _lib, engine, synth                    # noqa: E402
from bls_b200._lib import check, lib                        # noqa: E402
from bls_b200.programs.curve import G1_GEN                  # noqa: E402

_lib.init(0)
check(lib.b200bls_set_ctas_per_sm(4))
n = lib.b200bls_sm_count() * 384
sk = synth.scalars(5, n)
mh = synth.message_hashes(5, n)
g1 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in G1_GEN), dtype=np.uint8)
d_mh = engine.DeviceBuffer(32 * n).upload(mh)
d_sk = engine.DeviceBuffer(32 * n).upload(sk)
d_g = engine.DeviceBuffer(96 * n).upload(np.tile(g1, n))
d_h, d_sig, d_pk, d_ok = (engine.DeviceBuffer(192 * n), engine.DeviceBuffer(192 * n), engine.DeviceBuffer(96 * n),
                          engine.DeviceBuffer(n))
check(lib.b200bls_hash_to_g2_batch_dev(d_mh.ptr, d_h.ptr, n))
check(lib.b200bls_g2_scalar_mul_batch_dev(d_h.ptr, d_sk.ptr, d_sig.ptr, n))
check(lib.b200bls_g1_scalar_mul_batch_dev(d_g.ptr, d_sk.ptr, d_pk.ptr, n))
check(lib.b200bls_sync())
for _ in range(2):
    engine.timer_start()
    check(lib.b200bls_verify_batch_dev(d_pk.ptr, d_mh.ptr, d_sig.ptr, d_ok.ptr, n))
    print("verify n=%d: %.3f ms, all ok: %s" % (n, engine.timer_stop(), bool(d_ok.download().all())))
