#!/usr/bin/env python3
"""Offline estimate for the 'two threads per item' idea: list-schedule the
virtual instruction stream of a program (before workspace allocation: no spills, true data
dependencies only) on L lanes that share the item's workspace, with a cost per instruction of
(Montgomery products x 300 limb products) + a fixed overhead, and report makespan / total work.
No GPU needed.

  python tools/experiments/two_lane_schedule.py pairing verify_full
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200.programs import registry                     # noqa: E402
from bls_b200.vm.builder import Half, Val                  # noqa: E402

M_OF = {"MUL2": 3, "SQR2": 2, "MULFP2": 2, "MUL1": 1, "SQR1": 1, "LDBE48": 1, "LDBE32": 1, "STBE48": 1, "FGTHALF": 1}
LOOP_COST = {"INV1": 50.0, "FSQR1": 40.0}       # ALU loops, in units of one Montgomery product (ncu: fp_inv 1.8 % of samples / 5 calls)
OVERHEAD = 0.35                                 # decode + operand traffic + additions per instruction, in products (35.6 % of samples / 20 k instructions)
SYNC = 0.05                                     # what a hand-over between the lanes might cost


def root_id(x):
    while isinstance(x, Half):
        x = x.parent
    return x.id if isinstance(x, Val) else None


def analyse(name, lanes=2):
    prog = registry.PROGRAMS[name]()
    ops = prog.ops
    last_writer = {}
    deps, cost = [], []
    for i, op in enumerate(ops):
        d = set()
        srcs = []
        for x in (op.a, op.b):
            srcs += list(x) if isinstance(x, (list, tuple)) else [x]
        if op.name in ("CSEL2", "CSEL1", "STFLAG", "SKIPZ") or op.name.startswith("F"):
            srcs.append(op.aux)
        for x in srcs:
            r = root_id(x) if x is not None and not isinstance(x, int) else None
            if r is not None and r in last_writer:
                d.add(last_writer[r])
        dst = op.d if isinstance(op.d, (list, tuple)) else [op.d]
        for x in dst:
            r = root_id(x) if x is not None and not isinstance(x, int) else None
            if r is not None:
                if r in last_writer:
                    d.add(last_writer[r])       # partial writes / in-place updates keep their order
                last_writer[r] = i
        deps.append(d)
        cost.append(M_OF.get(op.name, 0) + LOOP_COST.get(op.name, 0.0) + OVERHEAD)
    total = sum(cost)
    # critical path
    depth = [0.0] * len(ops)
    for i in range(len(ops)):
        depth[i] = cost[i] + max((depth[j] for j in deps[i]), default=0.0)
    crit = max(depth)
    # list scheduling in program order priority (what a simple two-issue assembler would do)
    finish = [0.0] * len(ops)
    lane_free = [0.0] * lanes
    lane_of = [0] * len(ops)
    for i in range(len(ops)):
        ready = max((finish[j] + (SYNC if lane_of[j] != -1 else 0) for j in deps[i]), default=0.0)
        # pick the lane that lets the instruction start first; same-lane producers cost no hand-over
        best, best_t = 0, None
        for ln in range(lanes):
            r = max((finish[j] + (0.0 if lane_of[j] == ln else SYNC) for j in deps[i]), default=0.0)
            t = max(r, lane_free[ln])
            if best_t is None or t < best_t:
                best, best_t = ln, t
        lane_of[i] = best
        finish[i] = best_t + cost[i]
        lane_free[best] = finish[i]
    makespan = max(finish)
    print("%-14s %6d instructions  work %8.0f  critical path %8.0f (parallelism %.1f)  %d-lane makespan %8.0f  speed-up %.2f of %d"
          % (name, len(ops), total, crit, total / crit, lanes, makespan, total / makespan, lanes))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["pairing", "verify_full"]):
        analyse(nm, 2)
        analyse(nm, 4)
