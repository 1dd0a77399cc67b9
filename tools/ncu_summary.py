#!/usr/bin/env python3
"""Summarise one `ncu --set full` capture of vm_kernel (development container; needs the ncu CLI).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_pairing_wide "description"

writes <out>_ncu_summary.txt (the metrics DESIGN.md quotes, a per-opcode-class breakdown of
the warp stall samples) and <out>_ncu.json (read by bench.py for roofline.traffic)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "sm__icc_request_hit_rate.pct", "smsp__inst_executed.sum", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def build_id():
    """same identity bench.py computes: kernel sources + embedded programs of the tree the capture was taken from (run
    this script before changing them)"""
    import hashlib
    import os
    base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-bls_b200", "csrc")
    h = hashlib.sha256()
    for nm in sorted(os.listdir(base)):
        if nm.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(base, nm), "rb").read())
    h.update(open(os.path.join(base, "gen", "programs.bin"), "rb").read())
    return h.hexdigest()[:16]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE,
                          stderr=subprocess.DEVNULL, text=True, check=True).stdout


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value) * scale


def main(rep, out, desc, index=0):
    """index: which captured launch of the report (0 = first)"""
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2 + index]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = [desc, "kernel: %s" % m["Kernel Name"][0], ""]
    for k in KEEP:
        if k in m:
            lines.append("%s [%s] = %s" % (k, m[k][1], m[k][0]))
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(m[h][0] or 0) >= 0.05:
            lines.append("%s = %s" % (h, m[h][0]))
    traffic = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    # warp stall samples by SASS opcode class
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    # one block per captured launch: a "Kernel Name" row, a header row, then the instructions
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    lo = starts[index]
    hi = starts[index + 1] if index + 1 < len(starts) else len(src)
    shdr, data = src[lo + 1], [r for r in src[lo + 2:hi] if len(r) == len(src[lo + 1])]
    ix = {h: i for i, h in enumerate(shdr)}
    S, E = ix["# Samples"], ix["Instructions Executed"]
    tot = sum(int(r[S]) for r in data) or 1
    tote = sum(int(r[E]) for r in data) or 1
    cls, clse = collections.Counter(), collections.Counter()
    for r in data:
        mm = re.match(r"\s*(@!?U?P\d+\s+)?(\S+)", r[ix["Source"]])
        op = mm.group(2)
        key = op.split(".")[0] + (".WIDE" if ".WIDE" in op else "") + (".MOV" if ".MOV" in op else "")
        cls[key] += int(r[S])
        clse[key] += int(r[E])
    lines += ["", "SASS instructions in the kernel: %d" % len(data),
              "warp stall samples / executed instructions by SASS opcode class:"]
    for key, c in cls.most_common(14):
        lines.append("  %-12s samples %5.1f %%   executed %5.1f %%" % (key, 100.0 * c / tot, 100.0 * clse[key] / tote))
    with open(out + "_ncu_summary.txt", "w") as fh:
        fh.write("\n".join(lines) + "\n")
    js = {"description": desc, "kernel": m["Kernel Name"][0],
          "duration_ms": float(m["gpu__time_duration.sum"][0]) * {"ms": 1, "us": 1e-3, "s": 1e3}[m["gpu__time_duration.sum"][1]],
          "dram_bytes_per_launch": traffic,
          "fmaheavy_pipe_pct": float(m["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
          "issue_active_pct": float(m["sm__issue_active.avg.pct_of_peak_sustained_elapsed"][0]),
          "registers_per_thread": int(m["launch__registers_per_thread"][0]),
          "local_memory_instructions": int(float(m["sass__inst_executed_local_loads"][0])) + int(float(m["sass__inst_executed_local_stores"][0])),
          "build_id": build_id()}
    with open(out + "_ncu.json", "w") as fh:
        json.dump(js, fh, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "", int(sys.argv[4]) if len(sys.argv) > 4 else 0)
