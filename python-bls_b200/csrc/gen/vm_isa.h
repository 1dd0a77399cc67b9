// generated from bls_b200/vm/isa.py -- do not edit
#pragma once
enum VmOp : int {
  OP_NOP = 0,
  OP_MUL2 = 1,  // c2 c2 c2  d = a * b in Fq2
  OP_SQR2 = 2,  // c2 c2 -  d = a^2
  OP_ADD2 = 3,  // c2 c2 c2
  OP_SUB2 = 4,  // c2 c2 c2
  OP_NEG2 = 5,  // c2 c2 -
  OP_DBL2 = 6,  // c2 c2 -  d = 2a
  OP_MULXI2 = 7,  // c2 c2 -  d = a * (1 + u)
  OP_CONJ2 = 8,  // c2 c2 -  d = (a.c0, -a.c1)
  OP_MOV2 = 9,  // c2 c2 -
  OP_MULFP2 = 10,  // c2 c2 c1  d = a * b, b in Fq
  OP_MUL1 = 11,  // c1 c1 c1
  OP_SQR1 = 12,  // c1 c1 -
  OP_ADD1 = 13,  // c1 c1 c1
  OP_SUB1 = 14,  // c1 c1 c1
  OP_NEG1 = 15,  // c1 c1 -
  OP_DBL1 = 16,  // c1 c1 -
  OP_MOV1 = 17,  // c1 c1 -
  OP_LDC1 = 18,  // c1 k -  d = const[a]
  OP_LDC2 = 19,  // c2 k -  d = (const[a], const[a+1])
  OP_FZERO1 = 20,  // f c1 -  flag[d] = (a == 0)
  OP_FZERO2 = 21,  // f c2 -
  OP_FGTHALF = 22,  // f c1 -  flag[d] = standard-form(a) > (q-1)/2
  OP_FEQ1 = 23,  // f c1 c1
  OP_FEQ2 = 24,  // f c2 c2
  OP_FAND = 25,  // f f f
  OP_FOR = 26,  // f f f
  OP_FXOR = 27,  // f f f
  OP_FNOT = 28,  // f f -
  OP_FSET = 29,  // f i -  flag[d] = a & 1
  OP_FBIT = 30,  // f u i  flag[d] = bit b of the item's 32-byte big-endian scalar in buffer a
  OP_FACTIVE = 31,  // f - -  flag[d] = item index < n_items
  OP_CSEL2 = 32,  // c2 c2 c2  d = flag[aux] ? a : b
  OP_CSEL1 = 33,  // c1 c1 c1
  OP_LDBE48 = 34,  // c1 u i  d = to_mont(48 big-endian bytes at buffer a, byte offset 16*b)
  OP_LDBE32 = 35,  // c1 u i  d = to_mont(32 big-endian bytes ...)
  OP_STBE48 = 36,  // u c1 i  buffer d, byte offset 16*b <- 48 big-endian bytes of from_mont(a)
  OP_STFLAG = 37,  // u f i  buffer d, byte offset b <- flag[a] as one byte
  OP_LDRAW2 = 38,  // c2 u i  d = Montgomery limbs from internal SoA buffer a, element b
  OP_STRAW2 = 39,  // u c2 i  internal SoA buffer d, element b <- a
  OP_STRAWB2 = 40,  // u c2 i  as STRAW2 but only thread 0 of the block, item = block index
  OP_SPILL2 = 41,  // g c2 -  cold[d] <- a   (global-memory spill area)
  OP_FILL2 = 42,  // c2 g -  d <- cold[a]
  OP_SYNC = 43,  // - - -  block barrier
  OP_XMOV2 = 44,  // c2 c2 i  d = cell a of thread (tid + b) mod block size
  OP_SKIPZ = 45,  // f i -  if no thread of the warp has flag[d]: skip the next a instructions
  OP_FLDB = 46,  // f u i  flag[d] = (byte b of the item's record in buffer a) != 0
  OP_TRI2 = 47,  // c2 c2 c2  d = 3a - 2b (aux = 0) or 3a + 2b (aux = 1)
  OP__COUNT = 48
};
// operand handling per opcode: bit0/1 load a as Fq/Fq2, bit2/3 load b as Fq/Fq2,
// bit4/5 store the result to d as Fq/Fq2
enum VmOpInfo : unsigned { VM_A1 = 1, VM_A2 = 2, VM_B1 = 4, VM_B2 = 8, VM_D1 = 16, VM_D2 = 32 };
#define VM_OP_INFO_TABLE {0, 42, 34, 42, 42, 34, 34, 34, 34, 34, 38, 21, 17, 21, 21, 17, 17, 17, 16, 32, 1, 2, 1, 5, 10, 0, 0, 0, 0, 0, 0, 0, 42, 21, 16, 16, 1, 0, 32, 2, 2, 2, 32, 0, 32, 0, 0, 42}
