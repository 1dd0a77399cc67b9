#!/usr/bin/env python3
"""Benchmark of the b200-bls hot path (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one batch of 65,536 independent full ate pairings (Miller loop + exact final
exponentiation) per GPU on synthetic seeded inputs; weak scaling: every rank owns its own
batch, no data-path collective.  Prints ONE JSON line on rank 0:
  value        pairings/s with inputs resident in HBM, CUDA events on the library stream
  e2e          the same through b200bls_pairing_batch() with pinned HOST buffers (H2D, kernel,
               D2H inside the timed region)
  roofline     integer-multiply roofline: algorithmic limb products / measured IMAD.WIDE peak
  cpu_baseline the oracle port of the reference's CPU algorithm on all host cores
`--impl reference` times that CPU path alone with the same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))

BATCH = 65536
M_PER_PAIRING = 15200            # SURVEY.md 8d, frozen: efficient-algorithm Fq products per pairing
LIMB_PRODUCTS_PER_M = 300        # 12x12 limbs: 144 + 144 + 12
METRIC = "pairings/s"
WORKLOAD = "config2: 65,536 independent ate pairings (Miller loop + final exponentiation) per GPU"


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm, one process per host core
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, count = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bls_oracle as O
    p, q = O.aff_mul(seed * 2 + 3, O.G1), O.aff_mul(seed * 2 + 5, O.G2)
    t = time.perf_counter()
    for _ in range(count):
        O.ate_pairing(p, q)
    return time.perf_counter() - t


def cpu_pairings_per_second(per_core=4, cores=None):
    """oracle/bls_oracle.py ate_pairing on every host core -> (pairings/s, cores, sample text)"""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, 1) for i in range(cores)])            # warm-up / import
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i, per_core) for i in range(cores)])
        dt = time.perf_counter() - t0
    total = per_core * cores
    return total / dt, cores, "%d pairings (%d per core on %d cores) of the same workload" % (total, per_core, cores)


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port: the reference is pure
    Python and its Cython extension does not build in this image)"""
    if rank != 0:
        return
    times = []
    cores = os.cpu_count() or 1
    per_core = 3
    for step in range(args.warmup + args.steps):
        v, cores, sample = cpu_pairings_per_second(per_core=per_core)
        if step >= args.warmup:
            times.append(v)
        if step == 0 and args.warmup + args.steps > 4:
            per_core = 2
    value = sum(times) / len(times)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * BATCH / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (python int)",
            "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit())
        mx = [int(float(s[2])) for s in self.samples if s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(s[3]) for s in self.samples if s[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def make_inputs(engine, synth, rank, n):
    """P_i = a_i G1, Q_i = b_i G2 produced ON DEVICE by the scalar-multiplication kernels from
    seeded scalars; returns device buffers (P, Q)"""
    import numpy as np
    from bls_b200.programs.curve import G1_GEN
    from bls_b200.programs.hashg2 import G2_GEN
    a = synth.scalars(synth.SEED_PAIRING + 2 * rank, n)
    b = synth.scalars(synth.SEED_PAIRING + 2 * rank + 1, n)
    g1 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in G1_GEN), dtype=np.uint8)
    g2 = np.frombuffer(b"".join(c.to_bytes(48, "big") for c in (G2_GEN[0] + G2_GEN[1])), dtype=np.uint8)
    dg1 = engine.DeviceBuffer(96 * n).upload(np.tile(g1, n))
    dg2 = engine.DeviceBuffer(192 * n).upload(np.tile(g2, n))
    da = engine.DeviceBuffer(32 * n).upload(a)
    db = engine.DeviceBuffer(32 * n).upload(b)
    dP, dQ = engine.DeviceBuffer(96 * n), engine.DeviceBuffer(192 * n)
    from bls_b200._lib import check, lib
    check(lib.b200bls_g1_scalar_mul_batch_dev(dg1.ptr, da.ptr, dP.ptr, n))
    check(lib.b200bls_g2_scalar_mul_batch_dev(dg2.ptr, db.ptr, dQ.ptr, n))
    check(lib.b200bls_sync())
    for d in (dg1, dg2, da, db):
        d.free()
    return dP, dQ, a, b


def run_gpu(args, rank, world, dist):
    import ctypes
    import numpy as np
    from bls_b200 import _lib, engine, synth
    from bls_b200._lib import check, lib
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.init(local)
    # throughput shape: 4 = one 384-thread CTA per SM (6 shared + 7 Tensor-Memory slots per thread);
    # batches overlap on two streams
    check(lib.b200bls_set_ctas_per_sm(int(os.environ.get("B200BLS_BENCH_SHAPE", "4"))))
    n = BATCH
    # --- inputs: NSETS rotating buffer sets so the working set (NSETS * 56.6 MB) exceeds the
    # 126 MB L2 between timed iterations
    NSETS = 4
    dP, dQ, a_sc, b_sc = make_inputs(engine, synth, rank, n)
    sets = [(dP, dQ, engine.DeviceBuffer(576 * n))]
    hP, hQ = dP.download(), dQ.download()
    for _ in range(NSETS - 1):
        p2, q2 = engine.DeviceBuffer(96 * n).upload(hP), engine.DeviceBuffer(192 * n).upload(hQ)
        sets.append((p2, q2, engine.DeviceBuffer(576 * n)))

    N_STREAMS = int(os.environ.get("B200BLS_BENCH_STREAMS", "2"))   # consecutive batches go to alternating library streams and overlap

    def step(i):
        p, q, o = sets[i % NSETS]
        check(lib.b200bls_set_stream(i % N_STREAMS))
        check(lib.b200bls_pairing_batch_dev(p.ptr, q.ptr, o.ptr, n))

    def barrier():
        check(lib.b200bls_sync())
        if dist is not None:
            dist.barrier()

    # --- integer-multiply peak, measured live (roofline denominator)
    peak_ops, _ = engine.microbench_imad(3, 8, 256, 200)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.b200bls_launch_count()
    engine.timer_start()
    for i in range(args.steps):
        step(args.warmup + i)
    ms = engine.timer_stop()
    launches = lib.b200bls_launch_count() - launches0
    barrier()
    clocks = sampler.stop()
    check(lib.b200bls_set_stream(0))

    # --- end to end through the host-buffer C ABI call, pinned host memory
    def pinned(nbytes):
        p = lib.b200bls_host_alloc(nbytes)
        if not p:
            raise RuntimeError("pinned allocation failed")
        return p, np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))
    # two pinned buffer sets, one per stream: consecutive batches overlap (copy of one with the
    # kernel of the other, and the tail wave of one with the head of the next)
    host_sets = []
    for _ in range(N_STREAMS):
        pP, aP = pinned(96 * n)
        pQ, aQ = pinned(192 * n)
        pO, aO = pinned(576 * n)
        aP[:] = hP
        aQ[:] = hQ
        host_sets.append((pP, pQ, pO, aO))
    e2e_steps = max(2, min(args.steps, 24))
    for k in range(N_STREAMS):                                   # warm-up (staging allocation)
        check(lib.b200bls_set_stream(k))
        check(lib.b200bls_pairing_batch_async(host_sets[k][0], host_sets[k][1], host_sets[k][2], n))
    barrier()
    engine.timer_start()
    for i in range(e2e_steps):
        k = i % N_STREAMS
        check(lib.b200bls_set_stream(k))
        check(lib.b200bls_pairing_batch_async(host_sets[k][0], host_sets[k][1], host_sets[k][2], n))
    e2e_ms = engine.timer_stop()
    barrier()
    check(lib.b200bls_set_stream(0))
    aO = host_sets[(e2e_steps - 1) % N_STREAMS][3]

    # --- parity spot check outside the timed region (rank 0): two outputs vs the oracle, and the
    # device-resident result equals the host-path result
    parity = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import bls_oracle as O
        dev_out = sets[(args.warmup + args.steps - 1) % NSETS][2].download()
        parity = bool(np.array_equal(dev_out, aO))
        for idx in (0, n - 1):
            p = O.aff_mul(int.from_bytes(bytes(a_sc[idx]), "big"), O.G1)
            q = O.aff_mul(int.from_bytes(bytes(b_sc[idx]), "big"), O.G2)
            parity = parity and bytes(aO[576 * idx:576 * (idx + 1)]) == O.f12_serialize(O.ate_pairing(p, q))

    # --- secondary metric: signatures verified/s (hash-to-G2 + 2 Miller loops + final exp) on one
    # full wave of VALID signatures sig_i = a_i H(m_i) for the public keys pk_i = a_i G1 = P_i
    nv = lib.b200bls_sm_count() * 384
    mh = synth.message_hashes(synth.SEED_BATCH_VERIFY + rank, nv)
    d_mh = engine.DeviceBuffer(32 * nv).upload(mh)
    d_pk = engine.DeviceBuffer(96 * nv).upload(hP[:96 * nv])
    d_sk = engine.DeviceBuffer(32 * nv).upload(a_sc[:nv])
    d_h = engine.DeviceBuffer(192 * nv)
    d_sig = engine.DeviceBuffer(192 * nv)
    d_ok = engine.DeviceBuffer(nv)
    check(lib.b200bls_hash_to_g2_batch_dev(d_mh.ptr, d_h.ptr, nv))
    check(lib.b200bls_g2_scalar_mul_batch_dev(d_h.ptr, d_sk.ptr, d_sig.ptr, nv))
    check(lib.b200bls_verify_batch_dev(d_pk.ptr, d_mh.ptr, d_sig.ptr, d_ok.ptr, nv))
    check(lib.b200bls_sync())
    engine.timer_start()
    for _ in range(3):
        check(lib.b200bls_verify_batch_dev(d_pk.ptr, d_mh.ptr, d_sig.ptr, d_ok.ptr, nv))
    verify_ms = engine.timer_stop() / 3
    verify_all_ok = bool(d_ok.download().all())
    # the same end to end from the wire formats: serialised keys (48 B) and signatures (96 B) in
    # host memory -> H2D, from_bytes on the device, verification, D2H of the result bytes
    pk48 = engine.compress(d_pk.download(), False)
    sig96 = engine.compress(d_sig.download(), True)
    ok_wire = np.empty(nv, dtype=np.uint8)
    check(lib.b200bls_verify_batch_wire(_lib.ptr(pk48), _lib.ptr(mh), _lib.ptr(sig96), _lib.ptr(ok_wire), nv))
    engine.timer_start()
    check(lib.b200bls_verify_batch_wire(_lib.ptr(pk48), _lib.ptr(mh), _lib.ptr(sig96), _lib.ptr(ok_wire), nv))
    verify_wire_ms = engine.timer_stop()
    verify_all_ok = verify_all_ok and bool(ok_wire.all())
    # pairings on a batch that is a whole number of waves (nv = one wave): the headline batch of
    # 65,536 is 1.15 waves, so K of them end with a partly filled round (K = 5: 96 %)
    d_wo = engine.DeviceBuffer(576 * nv)
    check(lib.b200bls_pairing_batch_dev(d_pk.ptr, d_sig.ptr, d_wo.ptr, nv))
    check(lib.b200bls_sync())
    engine.timer_start()
    for _ in range(3):
        check(lib.b200bls_pairing_batch_dev(d_pk.ptr, d_sig.ptr, d_wo.ptr, nv))
    wave_ms = engine.timer_stop() / 3

    # --- reduce over ranks: max time
    t = [ms, e2e_ms, verify_ms, verify_wire_ms, wave_ms]
    if dist is not None:
        import torch
        tt = torch.tensor(t, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = tt.tolist()
    ms, e2e_ms, verify_ms, verify_wire_ms, wave_ms = t
    if rank != 0:
        return
    value = world * n * args.steps / (ms * 1e-3)
    e2e = world * n * e2e_steps / (e2e_ms * 1e-3)
    per_gpu = value / world
    achieved = per_gpu * M_PER_PAIRING * LIMB_PRODUCTS_PER_M
    hbm_bytes = n * (96 + 192 + 576)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cpu_v, cpu_cores, cpu_sample = cpu_pairings_per_second(per_core=3)
    # DRAM bytes of one launch of the pairing kernel at this batch size: dram__bytes_read.sum +
    # dram__bytes_write.sum of the committed `ncu --set full` capture (tools/ncu_summary.py)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_pairing_final_ncu.json")) as fh:
            cap = json.load(fh)
        if "n = 65,536" in cap["description"] and n == 65536 and lib.b200bls_get_ctas_per_sm() == 4:
            traffic = cap["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    from bls_b200.programs import registry
    executed_m = registry.executed_mults("pairing", 4)
    executed_m_verify = registry.executed_mults("verify_full", 4)
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (381-bit Montgomery)",
        "data": "synthetic", "gpu_launches": int(launches),
        "config": {"workload": WORKLOAD, "batch_per_gpu": n, "l2": "inputs rotate over %d buffer sets (%d MB > L2)"
                   % (NSETS, NSETS * hbm_bytes // 2 ** 20), "ctas_per_sm": lib.b200bls_get_ctas_per_sm(), "streams": N_STREAMS,
                   "parity_spot_check": parity},
        "e2e": {"value": e2e, "unit": METRIC, "h2d_bytes_per_step": n * 288, "d2h_bytes_per_step": n * 576,
                "steps": e2e_steps, "api": "b200bls_pairing_batch_async + b200bls_sync (pinned host buffers, %d streams)" % N_STREAMS},
        "roofline": {"bound": "int32_mul", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
                     "unit": "T limb-products/s", "frac": achieved / peak_ops, "traffic": traffic,
                     "executed_M_per_unit": executed_m,
                     "frac_executed": per_gpu * executed_m * LIMB_PRODUCTS_PER_M / peak_ops,
                     "note": "achieved = pairings/s/GPU x 15,200 M x 300 limb products (SURVEY 8d, frozen efficient-"
                             "algorithm count); frac_executed uses the Montgomery products the pairing program really "
                             "executes (executed_M_per_unit, counted on the assembled code; inversions are ALU-pipe loops); peak = "
                             "IMAD.WIDE.U32.X carry-chain microbenchmark measured in this run; per-launch "
                             "algorithmic HBM bytes %d (%.4f of measured HBM peak at this rate); traffic = DRAM bytes of "
                             "one launch from profiles/r1_pairing_final_ncu.json (workspace spills to the L2-backed cold area)"
                             % (hbm_bytes, (per_gpu * 864 / 1e9) / peaks.get("hbm_gbs", 6553.3))},
        "cpu_baseline": {"value": cpu_v, "unit": METRIC, "cores": cpu_cores, "kind": "port", "sample": cpu_sample},
        "clocks": clocks,
        "extra": {"verify_signatures_per_s": world * nv / (verify_ms * 1e-3), "verify_batch_per_gpu": nv, "verify_all_accepted": verify_all_ok,
                  "verify_e2e_from_wire_bytes_per_s": world * nv / (verify_wire_ms * 1e-3),
                  "pairings_per_s_whole_wave_batches": world * nv / (wave_ms * 1e-3),
                  "whole_wave_roofline_frac": (nv / (wave_ms * 1e-3)) * M_PER_PAIRING * LIMB_PRODUCTS_PER_M / peak_ops,
                  "verify_roofline_frac": (nv / (verify_ms * 1e-3)) * 30400 * 300 / peak_ops,
                  "verify_executed_M_per_unit": executed_m_verify,
                  "verify_roofline_frac_executed": (nv / (verify_ms * 1e-3)) * executed_m_verify * 300 / peak_ops},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # default 12: K batches of 65,536 are K x 170.7 item blocks on 148 SMs, and the timed region ends with
    # a partly filled round (K = 5: 5.77 rounds -> 96 % of the steady-state rate, K = 12: 13.84 -> 98.9 %)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist_mod.init_process_group("gloo", rank=rank, world_size=world)
        dist = dist_mod
    run_gpu(args, rank, world, dist)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
