// generated from bls_b200/vm/isa.py -- do not edit
#pragma once
enum VmOp : int {
  OP_NOP = 0,
  OP_ADD2 = 1,  // c2 c2 c2
  OP_SUB2 = 2,  // c2 c2 c2
  OP_SQR2 = 3,  // c2 c2 -  d = a^2
  OP_MUL2 = 4,  // c2 c2 c2  d = a * b in Fq2
  OP_MULXI2 = 5,  // c2 c2 -  d = a * (1 + u)
  OP_TRI2 = 6,  // c2 c2 c2  d = 3a - 2b (aux = 0) or 3a + 2b (aux = 1)
  OP_FILL2 = 7,  // c2 g -  d <- cold[a]; aux = 0x80 | g: then drop the cache lines of the dead cold copy in slot g (g = a: this was its last use)
  OP_SPILL2 = 8,  // g c2 -  cold[d] <- a   (global-memory spill area); aux = 0x80 | g: first drop the lines of the dead cold copy in slot g
  OP_DBL2 = 9,  // c2 c2 -  d = 2a
  OP_NEG2 = 10,  // c2 c2 -
  OP_CONJ2 = 11,  // c2 c2 -  d = (a.c0, -a.c1)
  OP_MOV2 = 12,  // c2 c2 -
  OP_MULFP2 = 13,  // c2 c2 c1  d = a * b, b in Fq
  OP_MUL1 = 14,  // c1 c1 c1
  OP_SQR1 = 15,  // c1 c1 -
  OP_ADD1 = 16,  // c1 c1 c1
  OP_SUB1 = 17,  // c1 c1 c1
  OP_NEG1 = 18,  // c1 c1 -
  OP_DBL1 = 19,  // c1 c1 -
  OP_MOV1 = 20,  // c1 c1 -
  OP_LDC1 = 21,  // c1 k -  d = const[a]
  OP_LDC2 = 22,  // c2 k -  d = (const[a], const[a+1])
  OP_FZERO1 = 23,  // f c1 -  flag[d] = (a == 0)
  OP_FZERO2 = 24,  // f c2 -
  OP_FGTHALF = 25,  // f c1 -  flag[d] = standard-form(a) > (q-1)/2
  OP_FEQ1 = 26,  // f c1 c1
  OP_FEQ2 = 27,  // f c2 c2
  OP_FAND = 28,  // f f f
  OP_FOR = 29,  // f f f
  OP_FXOR = 30,  // f f f
  OP_FNOT = 31,  // f f -
  OP_FSET = 32,  // f i -  flag[d] = a & 1
  OP_FBIT = 33,  // f u i  flag[d] = bit b of the item's big-endian scalar in buffer a (aux + 1 bytes; aux = 0: 32 bytes)
  OP_FACTIVE = 34,  // f - -  flag[d] = item index < n_items
  OP_CSEL2 = 35,  // c2 c2 c2  d = flag[aux] ? a : b
  OP_CSEL1 = 36,  // c1 c1 c1
  OP_LDBE48 = 37,  // c1 u i  d = to_mont(48 big-endian bytes at buffer a, byte offset 16*b)
  OP_LDBE32 = 38,  // c1 u i  d = to_mont(32 big-endian bytes ...)
  OP_STBE48 = 39,  // u c1 i  buffer d, byte offset 16*b <- 48 big-endian bytes of from_mont(a)
  OP_STFLAG = 40,  // u f i  buffer d, byte offset b <- flag[a] as one byte
  OP_LDRAW2 = 41,  // c2 u i  d = Montgomery limbs from internal SoA buffer a, element b
  OP_STRAW2 = 42,  // u c2 i  internal SoA buffer d, element b <- a
  OP_STRAWB2 = 43,  // u c2 i  as STRAW2 but only thread 0 of the block, item = block index
  OP_SYNC = 44,  // - - -  block barrier
  OP_XMOV2 = 45,  // c2 c2 i  d = cell a of thread (tid + b) mod block size
  OP_SKIPZ = 46,  // f i -  if no thread of the warp has flag[d]: skip the next a instructions
  OP_FLDB = 47,  // f u i  flag[d] = (byte b of the item's record in buffer a) != 0
  OP_INV1 = 48,  // c1 c1 -  d = 1 / a (0 -> 0): binary almost-inverse on the ALU pipe + two products
  OP_FSQR1 = 49,  // f c1 -  flag[d] = a is a nonzero square mod q (Legendre symbol by the binary algorithm: ALU pipe only)
  OP_DISCARD2 = 50,  // g - -  the cold copy in slot d is dead: drop its cache lines from the L2 (no write-back to DRAM)
  OP_END = 51,  // - - -  end of a program section (prologue / body / epilogue): the paired kernel's interpreter loop stops here instead of comparing its program counter with a bound it would have to keep in a register
  OP__COUNT = 52
};
// operand handling per opcode: bit0/1 load a as Fq/Fq2, bit2/3 load b as Fq/Fq2,
// bit4/5 store the result to d as Fq/Fq2
enum VmOpInfo : unsigned { VM_A1 = 1, VM_A2 = 2, VM_B1 = 4, VM_B2 = 8, VM_D1 = 16, VM_D2 = 32 };
#define VM_OP_INFO_TABLE {0, 42, 42, 34, 42, 34, 42, 32, 2, 34, 34, 34, 34, 38, 21, 17, 21, 21, 17, 17, 17, 16, 32, 1, 2, 1, 5, 10, 0, 0, 0, 0, 0, 0, 0, 42, 21, 16, 16, 1, 0, 32, 2, 2, 0, 32, 0, 0, 17, 1, 0, 0}
// post-operation of the hot Fq2 producers: bits 12..14 of the destination operand, third source in aux
enum VmPost : int { POST_NONE = 0, POST_ADD = 1, POST_SUB = 2, POST_RSUB = 3, POST_XI = 4, POST_DBL = 5, POST_SHIFT = 12 };
