// fp.cuh -- BLS12-381 base field Fq on 12 x 32-bit limbs, Montgomery form (R = 2^384).
//
// Device code is hand-written PTX carry chains: every 32x32->64 limb product is a
// mad.lo.cc / madc.hi.cc pair that ptxas fuses into one IMAD.WIDE.U32 with a
// predicate carry, accumulated into two interleaved ("even" / "odd" aligned)
// 64-bit column sets so that no carry ever has to be rippled by hand.
//
// Replaces (reference, /root/reference): the Python big-int arithmetic behind
// bls_py/fields.py:35-243 (class Fq) and the `% Q` reductions all over
// bls_py/fields_t.py.  Values are always fully reduced to [0, q) so that equality and
// zero tests are limb comparisons.
//
// The same header compiles for the host (B200BLS_HOSTSIM) with the PTX instructions
// emulated one by one, carry flag included.  That build exists ONLY so that tests can
// exercise the exact instruction sequences on the CPU-only development box
// (tests/hostsim); the product library never contains or calls it.
#pragma once
#include <stdint.h>

#include "gen/fp_consts.h"

#ifdef B200BLS_HOSTSIM
#define FP_DEV inline
#define FP_CONST static const
#else
#define FP_DEV __device__ __forceinline__
#define FP_CONST __device__ __constant__ const
#endif

namespace b200bls {

constexpr int NL = 12;  // limbs

struct fp {
  uint32_t v[NL];
};
struct fp2 {
  fp c0, c1;
};

// q, little-endian limbs
#define B200BLS_Q_LIMBS                                                                        \
  {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, \
   0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
constexpr uint32_t Q_INV_NEG = 0xfffcfffdu;  // -q^-1 mod 2^32

#ifdef B200BLS_HOSTSIM
static const uint32_t kQ[NL] = B200BLS_Q_LIMBS;
static const uint32_t kQ2[NL] = B200BLS_2Q_LIMBS;
static const uint32_t kQQ4[2 * NL] = B200BLS_4QQ_LIMBS;
#define QL(i) kQ[i]
#define Q2L(i) kQ2[i]
#define QQ4L(i) kQQ4[i]
#else
// Compile-time immediates: the modulus limbs become instruction immediates / constant
// bank operands instead of live registers.
__device__ __forceinline__ constexpr uint32_t q_limb(int i) {
  constexpr uint32_t t[NL] = B200BLS_Q_LIMBS;
  return t[i];
}
__device__ __forceinline__ constexpr uint32_t q2_limb(int i) {
  constexpr uint32_t t[NL] = B200BLS_2Q_LIMBS;
  return t[i];
}
__device__ __forceinline__ constexpr uint32_t qq4_limb(int i) {
  constexpr uint32_t t[2 * NL] = B200BLS_4QQ_LIMBS;
  return t[i];
}
#define QL(i) q_limb(i)
#define Q2L(i) q2_limb(i)
#define QQ4L(i) qq4_limb(i)
#endif

// ---------------------------------------------------------------------------------------
// PTX primitives (and their host emulation)
// ---------------------------------------------------------------------------------------
#ifdef B200BLS_HOSTSIM
static thread_local uint32_t g_cf = 0;  // emulated CC.CF
inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b;
  g_cf = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b + g_cf;
  g_cf = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a - b;
  g_cf = (uint32_t)(s >> 63);  // borrow
  return (uint32_t)s;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a - b - g_cf;
  g_cf = (uint32_t)(s >> 63);
  return (uint32_t)s;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + g_cf; }
#else
// NB: borrow semantics of sub.cc/subc on the GPU: CC.CF holds the borrow.
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
#endif

// ---------------------------------------------------------------------------------------
// Value discipline: every stored Fq value x is "weakly reduced", 0 <= x < 2q (zero is 0 or
// q).  With R = 2^384 and q ~ 0.1016 R this lets a Montgomery product skip its final
// subtraction (inputs < 2q -> output < 1.41 q) and lets Fq2 products be reduced once per
// coefficient from unreduced 768-bit sums (output < 2.63 q, one conditional subtraction of
// 2q).  Canonical form [0, q) is produced only where it is observable: comparisons, the
// sign rule (FGTHALF) and serialisation.
// ---------------------------------------------------------------------------------------
FP_DEV void fp_set_zero(fp& r) {
#pragma unroll
  for (int i = 0; i < NL; i++) r.v[i] = 0;
}

// x == 0 (mod q) for weakly reduced x: x is 0 or q
FP_DEV bool fp_is_zero(const fp& a) {
  uint32_t t = 0, u = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    t |= a.v[i];
    u |= a.v[i] ^ QL(i);
  }
  return t == 0 || u == 0;
}

// r = x - 2q if x >= 2q else x, for x < 4q
FP_DEV void fp_cond_sub_2q(fp& r, const fp& x) {
  fp t;
  t.v[0] = sub_cc(x.v[0], Q2L(0));
#pragma unroll
  for (int i = 1; i < NL; i++) t.v[i] = subc_cc(x.v[i], Q2L(i));
  uint32_t borrow = subc(0, 0);  // 0xffffffff when x < 2q
#pragma unroll
  for (int i = 0; i < NL; i++) r.v[i] = borrow ? x.v[i] : t.v[i];
}

// canonical form: weakly reduced x -> [0, q)
FP_DEV void fp_canonical(fp& r, const fp& x) {
  fp t;
  t.v[0] = sub_cc(x.v[0], QL(0));
#pragma unroll
  for (int i = 1; i < NL; i++) t.v[i] = subc_cc(x.v[i], QL(i));
  uint32_t borrow = subc(0, 0);  // 0xffffffff when x < q
#pragma unroll
  for (int i = 0; i < NL; i++) r.v[i] = borrow ? x.v[i] : t.v[i];
}

// plain 384-bit sum, no reduction (callers guarantee a + b < 2^384)
FP_DEV void fp_add_raw(fp& r, const fp& a, const fp& b) {
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[NL - 1] = addc(a.v[NL - 1], b.v[NL - 1]);
}

FP_DEV void fp_add(fp& r, const fp& a, const fp& b) {
  fp s;
  fp_add_raw(s, a, b);  // < 4q < 2^384
  fp_cond_sub_2q(r, s);
}

// a - b + 2q, no reduction: in (0, 4q) for weakly reduced a, b
FP_DEV void fp_sub_raw_2q(fp& r, const fp& a, const fp& b) {
  fp t;
  t.v[0] = add_cc(a.v[0], Q2L(0));
#pragma unroll
  for (int i = 1; i < NL - 1; i++) t.v[i] = addc_cc(a.v[i], Q2L(i));
  t.v[NL - 1] = addc(a.v[NL - 1], Q2L(NL - 1));
  r.v[0] = sub_cc(t.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.v[i] = subc_cc(t.v[i], b.v[i]);
  r.v[NL - 1] = subc(t.v[NL - 1], b.v[NL - 1]);
}

FP_DEV void fp_sub(fp& r, const fp& a, const fp& b) {
  fp d;
  d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < NL; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t mask = subc(0, 0);  // all ones when a < b
  r.v[0] = add_cc(d.v[0], Q2L(0) & mask);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.v[i] = addc_cc(d.v[i], Q2L(i) & mask);
  r.v[NL - 1] = addc(d.v[NL - 1], Q2L(NL - 1) & mask);
}

// a == b (mod q)
FP_DEV bool fp_eq(const fp& a, const fp& b) {
  fp d;
  fp_sub(d, a, b);
  return fp_is_zero(d);
}

FP_DEV void fp_neg(fp& r, const fp& a) {
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) nz |= a.v[i];
  uint32_t mask = nz ? 0xffffffffu : 0u;
  fp d;
  d.v[0] = sub_cc(Q2L(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) d.v[i] = subc_cc(Q2L(i), a.v[i]);
  d.v[NL - 1] = subc(Q2L(NL - 1), a.v[NL - 1]);
#pragma unroll
  for (int i = 0; i < NL; i++) r.v[i] = d.v[i] & mask;
}

FP_DEV void fp_dbl(fp& r, const fp& a) { fp_add(r, a, a); }

// a > b as plain 384-bit integers (used on canonical values taken out of Montgomery form)
FP_DEV bool fp_raw_gt(const fp& a, const fp& b) {
  sub_cc(b.v[0], a.v[0]);
#pragma unroll
  for (int i = 1; i < NL; i++) subc_cc(b.v[i], a.v[i]);
  return subc(0, 0) != 0;  // borrow <=> b < a
}

// ---------------------------------------------------------------------------------------
// Multiplication = unreduced 768-bit product (fp_mul_wide) + Montgomery reduction (fp_redc).
//
// Both are "row" algorithms on two interleaved 24-limb accumulators: T = E + O * 2^32, E
// holds the 64-bit columns at even word positions, O those at odd positions (O[k] has weight
// 2^(32(k+1))).  A row adds x * y_i * 2^(32 i) for a 12-limb x: its even limbs form one carry
// chain of six lo/hi pairs into E or O (by the parity of i), its odd limbs a second,
// independent chain into the other set.  ptxas fuses every lo/hi pair into one
// IMAD.WIDE.U32.X with a predicate carry and interleaves the chains (3-4 in flight).
// Splitting product and reduction lets Fq2 arithmetic add / subtract unreduced products and
// reduce once per coefficient: 3 products + 2 reductions = 744 limb products per Fq2
// multiplication instead of 900.
// ---------------------------------------------------------------------------------------
struct fpw {
  uint32_t v[2 * NL];
};

// acc[s .. s+11] += {x[off], x[off+2], ...} * y as one carry chain; carry out -> acc[s+12]
template <int OFF>
FP_DEV void row_chain(uint32_t* acc, int s, const uint32_t* x, uint32_t y) {
#pragma unroll
  for (int j = 0; j < NL; j += 2) {
    if (j == 0)
      acc[s + j] = mad_lo_cc(x[j + OFF], y, acc[s + j]);
    else
      acc[s + j] = madc_lo_cc(x[j + OFF], y, acc[s + j]);
    acc[s + j + 1] = madc_hi_cc(x[j + OFF], y, acc[s + j + 1]);
  }
  if (s + NL < 2 * NL) acc[s + NL] = addc(acc[s + NL], 0);
}

template <int OFF>
FP_DEV void row_chain_q(uint32_t* acc, int s, uint32_t y) {
#pragma unroll
  for (int j = 0; j < NL; j += 2) {
    if (j == 0)
      acc[s + j] = mad_lo_cc(QL(j + OFF), y, acc[s + j]);
    else
      acc[s + j] = madc_lo_cc(QL(j + OFF), y, acc[s + j]);
    acc[s + j + 1] = madc_hi_cc(QL(j + OFF), y, acc[s + j + 1]);
  }
  if (s + NL < 2 * NL) acc[s + NL] = addc(acc[s + NL], 0);
}

// t = a * b, any 384-bit a, b
FP_DEV void fp_mul_wide_inline(fpw& t, const fp& a, const fp& b) {
  uint32_t E[2 * NL], O[2 * NL];
#pragma unroll
  for (int k = 0; k < 2 * NL; k++) E[k] = O[k] = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    if ((i & 1) == 0) {
      row_chain<0>(E, i, a.v, b.v[i]);      // even limbs -> even positions i + j
      row_chain<1>(O, i, a.v, b.v[i]);      // odd limbs  -> odd positions, O index = pos - 1
    } else {
      row_chain<0>(O, i - 1, a.v, b.v[i]);  // even limbs -> odd positions i + j
      row_chain<1>(E, i + 1, a.v, b.v[i]);  // odd limbs  -> even positions i + j
    }
  }
  t.v[0] = E[0];
  t.v[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 2 * NL - 1; k++) t.v[k] = addc_cc(E[k], O[k - 1]);
  t.v[2 * NL - 1] = addc(E[2 * NL - 1], O[2 * NL - 2]);
}

// r = t / R mod q (Montgomery reduction), r < t / R + q.  Callers keep t < 16 q^2, so that
// r < 2.63 q fits 12 limbs and no intermediate exceeds 768 bits.
FP_DEV void fp_redc_inline(fp& r, const fpw& t) {
  uint32_t E[2 * NL], O[2 * NL];
#pragma unroll
  for (int k = 0; k < 2 * NL; k++) {
    E[k] = t.v[k];
    O[k] = 0;
  }
  uint32_t c = 0;  // carry out of the (zeroed) merged limbs below the current one
#pragma unroll
  for (int i = 0; i < NL; i++) {
    uint32_t lo = E[i] + c;
    if (i > 0) lo += O[i - 1];
    const uint32_t m = mul_lo(lo, Q_INV_NEG);
    if ((i & 1) == 0) {
      row_chain_q<0>(E, i, m);
      row_chain_q<1>(O, i, m);
    } else {
      row_chain_q<0>(O, i - 1, m);
      row_chain_q<1>(E, i + 1, m);
    }
    // merged limb i is now 0 mod 2^32; its carry moves up
    uint32_t s1 = add_cc(E[i], c);
    uint32_t k1 = addc(0, 0);
    uint32_t k2 = 0;
    if (i > 0) {
      s1 = add_cc(s1, O[i - 1]);
      k2 = addc(0, 0);
    }
    (void)s1;
    c = k1 + k2;
  }
  fp u;
  u.v[0] = add_cc(E[NL], c);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) u.v[k] = addc_cc(E[NL + k], 0);
  u.v[NL - 1] = addc(E[2 * NL - 1], 0);
  r.v[0] = add_cc(u.v[0], O[NL - 1]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) r.v[k] = addc_cc(u.v[k], O[NL - 1 + k]);
  r.v[NL - 1] = addc(u.v[NL - 1], O[2 * NL - 2]);
}

FP_DEV void fpw_add(fpw& r, const fpw& a, const fpw& b) {
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 2 * NL - 1; k++) r.v[k] = addc_cc(a.v[k], b.v[k]);
  r.v[2 * NL - 1] = addc(a.v[2 * NL - 1], b.v[2 * NL - 1]);
}

FP_DEV void fpw_sub(fpw& r, const fpw& a, const fpw& b) {
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int k = 1; k < 2 * NL - 1; k++) r.v[k] = subc_cc(a.v[k], b.v[k]);
  r.v[2 * NL - 1] = subc(a.v[2 * NL - 1], b.v[2 * NL - 1]);
}

// r = a + 4 q^2 (keeps a following subtraction of a product < 4 q^2 non-negative)
FP_DEV void fpw_add_4qq(fpw& r, const fpw& a) {
  r.v[0] = add_cc(a.v[0], QQ4L(0));
#pragma unroll
  for (int k = 1; k < 2 * NL - 1; k++) r.v[k] = addc_cc(a.v[k], QQ4L(k));
  r.v[2 * NL - 1] = addc(a.v[2 * NL - 1], QQ4L(2 * NL - 1));
}

#if defined(B200BLS_HOSTSIM) || !defined(B200BLS_MUL_CALL)
FP_DEV void fp_mul_wide(fpw& t, const fp& a, const fp& b) { fp_mul_wide_inline(t, a, b); }
FP_DEV void fp_redc(fp& r, const fpw& t) { fp_redc_inline(r, t); }
#else
// ONE copy of the product and ONE of the reduction in the whole kernel: operands and results
// travel in registers (ptxas: 0 bytes stack).  The interpreter's hot code then fits the
// instruction cache, which is what limits the number of co-resident warps (icc hit rate 81%
// and `no_instruction` stalls with 12 warps/SM when every opcode body inlines its own copies).
__device__ __noinline__ fpw fp_mul_wide_call(fp a, fp b) {
  fpw t;
  fp_mul_wide_inline(t, a, b);
  return t;
}
__device__ __noinline__ fp fp_redc_call(fpw t) {
  fp r;
  fp_redc_inline(r, t);
  return r;
}
__device__ __forceinline__ void fp_mul_wide(fpw& t, const fp& a, const fp& b) { t = fp_mul_wide_call(a, b); }
__device__ __forceinline__ void fp_redc(fp& r, const fpw& t) { r = fp_redc_call(t); }
#endif

// r = a * b / R mod q: weakly reduced in -> r < 1.41 q (weakly reduced), no final subtraction.
// b may be ANY 384-bit value when a < q (used by the byte loaders): r < 2q still.
FP_DEV void fp_mul(fp& r, const fp& a, const fp& b) {
  fpw t;
  fp_mul_wide(t, a, b);
  fp_redc(r, t);
}

FP_DEV void fp_sqr(fp& r, const fp& a) { fp_mul(r, a, a); }

// ---------------------------------------------------------------------------------------
// Fq2 = Fq[u]/(u^2+1)  (reference: bls_py/fields.py:321-482, fields_t.py:75-161)
// ---------------------------------------------------------------------------------------
FP_DEV void fp2_add(fp2& r, const fp2& a, const fp2& b) {
  fp_add(r.c0, a.c0, b.c0);
  fp_add(r.c1, a.c1, b.c1);
}
FP_DEV void fp2_sub(fp2& r, const fp2& a, const fp2& b) {
  fp_sub(r.c0, a.c0, b.c0);
  fp_sub(r.c1, a.c1, b.c1);
}
FP_DEV void fp2_neg(fp2& r, const fp2& a) {
  fp_neg(r.c0, a.c0);
  fp_neg(r.c1, a.c1);
}
// Karatsuba on unreduced products, one reduction per coefficient (fields_t.py:157-161 uses 4
// full multiplications): c0 = a0 b0 - a1 b1, c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1.
// Bounds for weakly reduced inputs: products < 4 q^2, the sum product < 16 q^2;
// c0 + 4 q^2 < 8 q^2 -> r.c0 < 1.81 q;  c1 < 16 q^2 -> < 2.63 q -> one conditional - 2q.
FP_DEV void fp2_mul(fp2& r, const fp2& a, const fp2& b) {
  fp sa, sb;
  fpw t0, t1, t2;
  fp_add_raw(sa, a.c0, a.c1);
  fp_add_raw(sb, b.c0, b.c1);
  fp_mul_wide(t0, a.c0, b.c0);
  fp_mul_wide(t1, a.c1, b.c1);
  fp_mul_wide(t2, sa, sb);
  fpw_sub(t2, t2, t0);
  fpw_sub(t2, t2, t1);          // a0 b1 + a1 b0 >= 0
  fpw_add_4qq(t0, t0);
  fpw_sub(t0, t0, t1);
  fp_redc(r.c0, t0);
  fp c1;
  fp_redc(c1, t2);
  fp_cond_sub_2q(r.c1, c1);
}
// (a0 + a1)(a0 - a1 + 2q), (2 a0) a1
FP_DEV void fp2_sqr(fp2& r, const fp2& a) {
  fp s, d, e;
  fpw t;
  fp_add_raw(s, a.c0, a.c1);        // < 4q
  fp_sub_raw_2q(d, a.c0, a.c1);     // in (0, 4q)
  fp_add_raw(e, a.c0, a.c0);        // < 4q
  fp_mul_wide(t, s, d);             // < 16 q^2
  fp c0;
  fp_redc(c0, t);
  fp_mul_wide(t, e, a.c1);          // < 8 q^2
  fp_redc(r.c1, t);
  fp_cond_sub_2q(r.c0, c0);
}
// times xi = 1 + u  (fields_t.py:113-116)
FP_DEV void fp2_mul_xi(fp2& r, const fp2& a) {
  fp t;
  fp_sub(t, a.c0, a.c1);
  fp_add(r.c1, a.c0, a.c1);
  r.c0 = t;
}

}  // namespace b200bls
