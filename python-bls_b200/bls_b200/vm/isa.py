"""Instruction set of the b200-bls field VM (mirrored by csrc/vm_isa.h -- generated).

One instruction is 8 bytes: ``u16 op | aux << 8``, ``u16 d``, ``u16 a``, ``u16 b``.
Operands are *fp cell* indices in the thread's shared-memory workspace (an Fp2 value
occupies an even/odd cell pair), constant-table indices, flag numbers, input/output
buffer numbers or small immediates, depending on the opcode.
"""

OPS = [
    # name        operand kinds (d, a, b)   -- c2/c1 = Fp2/Fp cell, k = const idx,
    #                                           f = flag, u = buffer, i = immediate
    # the first nine opcodes after NOP are the hot ones of the pairing programs, in the order the
    # interpreter's range-compare dispatch expects (vm_exec.cuh static_asserts the numbering)
    ("NOP", "", ""),
    ("ADD2", "c2 c2 c2", ""),
    ("SUB2", "c2 c2 c2", ""),
    ("SQR2", "c2 c2 -", "d = a^2"),
    ("MUL2", "c2 c2 c2", "d = a * b in Fq2"),
    ("MULXI2", "c2 c2 -", "d = a * (1 + u)"),
    ("TRI2", "c2 c2 c2", "d = 3a - 2b (aux = 0) or 3a + 2b (aux = 1)"),
    ("FILL2", "c2 g -", "d <- cold[a]; aux = 0x80 | g: then drop the cache lines of the dead cold copy in slot g (g = a: this was its last use)"),
    ("SPILL2", "g c2 -", "cold[d] <- a   (global-memory spill area); aux = 0x80 | g: first drop the lines of the dead cold copy in slot g"),
    ("DBL2", "c2 c2 -", "d = 2a"),
    ("NEG2", "c2 c2 -", ""),
    ("CONJ2", "c2 c2 -", "d = (a.c0, -a.c1)"),
    ("MOV2", "c2 c2 -", ""),
    ("MULFP2", "c2 c2 c1", "d = a * b, b in Fq"),
    ("MUL1", "c1 c1 c1", ""),
    ("SQR1", "c1 c1 -", ""),
    ("ADD1", "c1 c1 c1", ""),
    ("SUB1", "c1 c1 c1", ""),
    ("NEG1", "c1 c1 -", ""),
    ("DBL1", "c1 c1 -", ""),
    ("MOV1", "c1 c1 -", ""),
    ("LDC1", "c1 k -", "d = const[a]"),
    ("LDC2", "c2 k -", "d = (const[a], const[a+1])"),
    ("FZERO1", "f c1 -", "flag[d] = (a == 0)"),
    ("FZERO2", "f c2 -", ""),
    ("FGTHALF", "f c1 -", "flag[d] = standard-form(a) > (q-1)/2"),
    ("FEQ1", "f c1 c1", ""),
    ("FEQ2", "f c2 c2", ""),
    ("FAND", "f f f", ""),
    ("FOR", "f f f", ""),
    ("FXOR", "f f f", ""),
    ("FNOT", "f f -", ""),
    ("FSET", "f i -", "flag[d] = a & 1"),
    ("FBIT", "f u i", "flag[d] = bit b of the item's big-endian scalar in buffer a (aux + 1 bytes; aux = 0: 32 bytes)"),
    ("FACTIVE", "f - -", "flag[d] = item index < n_items"),
    ("CSEL2", "c2 c2 c2", "d = flag[aux] ? a : b"),
    ("CSEL1", "c1 c1 c1", ""),
    ("LDBE48", "c1 u i", "d = to_mont(48 big-endian bytes at buffer a, byte offset 16*b)"),
    ("LDBE32", "c1 u i", "d = to_mont(32 big-endian bytes ...)"),
    ("STBE48", "u c1 i", "buffer d, byte offset 16*b <- 48 big-endian bytes of from_mont(a)"),
    ("STFLAG", "u f i", "buffer d, byte offset b <- flag[a] as one byte"),
    ("LDRAW2", "c2 u i", "d = Montgomery limbs from internal SoA buffer a, element b"),
    ("STRAW2", "u c2 i", "internal SoA buffer d, element b <- a"),
    ("STRAWB2", "u c2 i", "as STRAW2 but only thread 0 of the block, item = block index"),
    ("SYNC", "- - -", "block barrier"),
    ("XMOV2", "c2 c2 i", "d = cell a of thread (tid + b) mod block size"),
    ("SKIPZ", "f i -", "if no thread of the warp has flag[d]: skip the next a instructions"),
    ("FLDB", "f u i", "flag[d] = (byte b of the item's record in buffer a) != 0"),
    ("INV1", "c1 c1 -", "d = 1 / a (0 -> 0): binary almost-inverse on the ALU pipe + two products"),
    ("FSQR1", "f c1 -", "flag[d] = a is a nonzero square mod q (Legendre symbol by the binary algorithm: ALU pipe only)"),
    ("DISCARD2", "g - -", "the cold copy in slot d is dead: drop its cache lines from the L2 (no write-back to DRAM)"),
    ("END", "- - -", "end of a program section (prologue / body / epilogue): the paired kernel's interpreter loop "
                     "stops here instead of comparing its program counter with a bound it would have to keep in a register"),
]

OPCODE = {name: i for i, (name, _, _) in enumerate(OPS)}
OPNAME = {i: name for name, i in OPCODE.items()}
OPSIG = {name: sig.split() for name, sig, _ in OPS}

# Post-operations: the hot Fq2 producers can apply one more add-like step to their result z before
# storing it (builder.fuse_pairs).  Code in bits 12..14 of the destination operand (cells and
# cold slots are < 4096), third source cell c in aux.
POST_NONE, POST_ADD, POST_SUB, POST_RSUB, POST_XI, POST_DBL = 0, 1, 2, 3, 4, 5   # z+c, z-c, c-z, xi*z, 2z
POST_SHIFT = 12
POST_PRIMARY = ("ADD2", "SUB2", "SQR2", "MUL2", "MULXI2")
POST_SECONDARY = ("ADD2", "SUB2", "MULXI2", "DBL2")

N_FLAGS = 32
INS_BYTES = 8


# operand handling done once by the interpreter, outside the per-opcode switch
INFO_A1, INFO_A2, INFO_B1, INFO_B2, INFO_D1, INFO_D2 = 1, 2, 4, 8, 16, 32
_CUSTOM_OPERANDS = {"XMOV2": INFO_D2, "SPILL2": INFO_A2, "FILL2": INFO_D2, "CSEL2": INFO_A2 | INFO_B2 | INFO_D2,
                    "CSEL1": INFO_A1 | INFO_B1 | INFO_D1}


def op_info(name):
    if name in _CUSTOM_OPERANDS:
        return _CUSTOM_OPERANDS[name]
    sig = OPSIG[name] + ["-"] * 3
    bits = 0
    bits |= {"c1": INFO_D1, "c2": INFO_D2}.get(sig[0], 0)
    bits |= {"c1": INFO_A1, "c2": INFO_A2}.get(sig[1], 0)
    bits |= {"c1": INFO_B1, "c2": INFO_B2}.get(sig[2], 0)
    return bits


def c_header():
    lines = ["// generated from bls_b200/vm/isa.py -- do not edit", "#pragma once", "enum VmOp : int {"]
    for i, (name, sig, doc) in enumerate(OPS):
        lines.append("  OP_%s = %d,%s" % (name, i, ("  // " + (sig + "  " + doc).strip()) if (sig or doc) else ""))
    lines.append("  OP__COUNT = %d" % len(OPS))
    lines.append("};")
    lines.append("// operand handling per opcode: bit0/1 load a as Fq/Fq2, bit2/3 load b as Fq/Fq2,")
    lines.append("// bit4/5 store the result to d as Fq/Fq2")
    lines.append("enum VmOpInfo : unsigned { VM_A1 = 1, VM_A2 = 2, VM_B1 = 4, VM_B2 = 8, VM_D1 = 16, VM_D2 = 32 };")
    lines.append("#define VM_OP_INFO_TABLE {%s}" % ", ".join(str(op_info(n)) for n, _, _ in OPS))
    lines.append("// post-operation of the hot Fq2 producers: bits 12..14 of the destination operand, third source in aux")
    lines.append("enum VmPost : int { POST_NONE = %d, POST_ADD = %d, POST_SUB = %d, POST_RSUB = %d, POST_XI = %d, "
                 "POST_DBL = %d, POST_SHIFT = %d };" % (POST_NONE, POST_ADD, POST_SUB, POST_RSUB, POST_XI, POST_DBL, POST_SHIFT))
    return "\n".join(lines) + "\n"
