#!/usr/bin/env python3
"""Assemble every VM program and write csrc/gen/programs.bin (embedded into libb200bls.so).

Blob layout (little endian):
  char magic[8] = "B2BLSPRG"; u32 version; u32 n_programs;
  n_programs x { char name[32]; u32 n_ins, body_start, epi_start, n_consts, n_slots (shared),
                 n_cold, n_tmem (Tensor Memory slots), ctas_per_sm; u64 code_off, consts_off; }
  code:   (n_ins + 1) x 8 bytes (one trailing NOP of padding)
  consts: n_consts x 12 x u32 Montgomery-form limbs
"""
import os
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-bls_b200"))
from bls_b200.programs import registry                        # noqa: E402
from bls_b200.vm import isa                                   # noqa: E402


def main(out_path=None):
    out_path = out_path or os.path.join(ROOT, "python-bls_b200", "csrc", "gen", "programs.bin")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    progs = []
    for base, builder in registry.PROGRAMS.items():
        for ctas, (n_slots, n_tmem) in registry.SHAPES.items():
            if ctas != 1 and base in getattr(registry, "SHAPE1_ONLY", ()):
                continue
            name = "%s@%d" % (base, ctas)
            t = time.time()
            asm = None
            for tm in ((n_tmem,) if ctas in registry.WIDE_SHAPES else (n_tmem, 0) if n_tmem else (0,)):
                try:
                    asm = builder().assemble(n_slots, n_cold=4096, n_tmem=tm)
                    break
                except RuntimeError as e:
                    err = e
            if asm is not None and ctas in registry.WIDE_SHAPES:
                # block-level reductions are written for 128-thread CTAs
                ops = set(int(o) & 0xff for o in asm.code[:, 0])
                if ops & {isa.OPCODE["SYNC"], isa.OPCODE["XMOV2"], isa.OPCODE["STRAWB2"]}:
                    asm, err = None, "cross-thread program (128-thread CTAs only)"
            if asm is None:
                if ctas == 1:
                    raise err
                print("%-24s not available: %s" % (name, err))
                continue
            print("%-24s %6d ins  slots %2d+%2d  spills %5d fills %5d cold %3d  (%.1fs)" % (
                name, asm.stats["n_ins"], asm.n_slots, asm.n_tmem, asm.stats["spills"], asm.stats["fills"],
                asm.stats["max_cold"], time.time() - t))
            progs.append((name, ctas, asm))
    entry = "<32s8I2Q"
    head = 16 + len(progs) * struct.calcsize(entry)
    blobs = []
    off = head
    table = b""
    for name, ctas, asm in progs:
        code = np.concatenate([asm.code, np.zeros((1, 4), dtype=np.uint16)]).astype("<u2").tobytes()
        consts = asm.const_limbs().astype("<u4").tobytes()
        n_consts = max(1, len(asm.consts))
        off = (off + 15) & ~15
        code_off = off
        off += len(code)
        off = (off + 15) & ~15
        consts_off = off
        off += len(consts)
        table += struct.pack(entry, name.encode(), len(asm.code), asm.body_start, asm.epilogue_start, n_consts,
                             asm.n_slots, max(1, asm.stats["max_cold"]), asm.n_tmem, ctas, code_off, consts_off)
        blobs.append((code_off, code, consts_off, consts))
    data = bytearray(off)
    data[0:16] = struct.pack("<8sII", b"B2BLSPRG", 2, len(progs))
    data[16:16 + len(table)] = table
    for code_off, code, consts_off, consts in blobs:
        data[code_off:code_off + len(code)] = code
        data[consts_off:consts_off + len(consts)] = consts
    with open(out_path, "wb") as fh:
        fh.write(data)
    print("wrote", out_path, len(data), "bytes")


if __name__ == "__main__":
    main(*sys.argv[1:])
