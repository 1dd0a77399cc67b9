// fp_dfma.cuh -- the SECOND multiplier: the same Montgomery product as fp_mul_inline (R = 2^384, identical
// result bit for bit), computed on the FP64 pipe instead of the integer-multiply pipe.
//
// Why: on the B200 the FP64 pipe (64 DFMA / clk / SM) is independent of the pipe that executes IMAD.WIDE
// (32 limb products / clk / SM), and the pairing kernel leaves it idle.  A warp that multiplies here does not
// compete with the warps that multiply there (profiles/r1_probe_fp64.json: both chains finish in the time of
// the slower one).
//
// How: the operands are re-cut into eight 48-bit limbs held as doubles (exact integers).  A 48 x 48 -> 96 bit
// limb product is split into two 48-bit halves by two fused multiply-adds in round-towards-zero mode
//     hi = fma_rz(a, b, 2^100)                  = 2^100 + floor(ab / 2^48) 2^48     (ulp = 2^48)
//     lo = fma_rz(a, b, (2^100 + 2^52) - hi)    = 2^52  + (ab mod 2^48)             (exact)
// so that the low 52 bits of the two bit patterns ARE the halves; the patterns are summed as 64-bit integers
// into columns whose start values cancel the exponent fields of everything they will receive.  (The splitting
// trick is the one of Emmart, Zheng, Weems, "Faster modular exponentiation using double precision floating
// point arithmetic on the GPU", ARITH 2018; the column layout, the 48-bit cut that keeps R = 2^384 and the
// word-serial reduction below are this library's.)  Eight rounds of operand scanning; round i adds a * b_i,
// reads the digit d = column_i mod 2^48, adds m_i * q with m_i = d * (-1/q) mod 2^48, and carries column_i /
// 2^48 into the next column.  The sum of the m_i 2^(48 i) is the same 384-bit m the 32-bit-limb product
// determines word by word, hence T = (ab + mq) / R is the same integer: no final subtraction, same bounds.
//
// Cost per product: 8 x (16 + 1) limb splits = 408 DFMA/DADD-class instructions + 16 + 8 conversions on the
// FP64 pipe (2 cycles each per scheduler: ~ 900 cycles against 1,200 on the integer-multiply pipe), ~ 400
// integer additions / byte permutes on the ALU pipe, no IMAD at all.
#pragma once
#include <stdint.h>

#ifdef B200BLS_HOSTSIM
#include <fenv.h>
#include <math.h>
#include <string.h>
#endif

namespace b200bls {

constexpr int DL = 8;  // 48-bit limbs
constexpr uint64_t DF_MASK48 = (1ull << 48) - 1;
constexpr uint64_t DF_OFF_LO = 0x433ull << 52;  // bit pattern of 2^52
constexpr uint64_t DF_OFF_HI = 0x463ull << 52;  // bit pattern of 2^100

// q in 48-bit limbs and -1/q mod 2^48 (checked against the 32-bit constants by tests/hostsim)
#define B200BLS_Q48_LIMBS                                                                                         \
  {0xffffffffaaabull, 0xb153ffffb9feull, 0xf6241eabfffeull, 0x6730d2a0f6b0ull, 0x4b84f38512bfull, 0x434bacd76477ull, \
   0xe69a4b1ba7b6ull, 0x1a0111ea397full}
constexpr uint64_t DF_QINV_NEG48 = 0xfffcfffcfffdull;

#ifdef B200BLS_HOSTSIM
inline double df_fma_rz(double a, double b, double c) {
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  volatile double va = a, vb = b, vc = c;
  volatile double r = fma(va, vb, vc);
  fesetround(old);
  return r;
}
inline double df_add(double a, double b) {
  volatile double r = a + b;
  return r;
}
inline uint64_t df_bits(double x) {
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
}
inline double df_from_bits(uint64_t u) {
  double x;
  memcpy(&x, &u, 8);
  return x;
}
inline double df_from_hilo(uint32_t hi, uint32_t lo) { return df_from_bits(((uint64_t)hi << 32) | lo); }
inline uint32_t df_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint8_t src[8];
  for (int i = 0; i < 4; i++) {
    src[i] = (uint8_t)(a >> (8 * i));
    src[4 + i] = (uint8_t)(b >> (8 * i));
  }
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)src[(sel >> (4 * i)) & 7] << (8 * i);
  return r;
}
static const double kQ48[DL] = {(double)0xffffffffaaabull, (double)0xb153ffffb9feull, (double)0xf6241eabfffeull,
                                (double)0x6730d2a0f6b0ull, (double)0x4b84f38512bfull, (double)0x434bacd76477ull,
                                (double)0xe69a4b1ba7b6ull, (double)0x1a0111ea397full};
#define Q48D(i) kQ48[i]
#else
__device__ __forceinline__ double df_fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
__device__ __forceinline__ double df_add(double a, double b) { return __dadd_rn(a, b); }  // always exact here
__device__ __forceinline__ uint64_t df_bits(double x) { return (uint64_t)__double_as_longlong(x); }
__device__ __forceinline__ double df_from_bits(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ double df_from_hilo(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ uint32_t df_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
__device__ __forceinline__ constexpr double q48_limb(int i) {
  constexpr uint64_t t[DL] = B200BLS_Q48_LIMBS;
  return (double)t[i];
}
#define Q48D(i) q48_limb(i)
#endif

// start value of column k: minus the exponent fields of every pattern it will receive (two products per
// (i, j) pair: a_j b_i and m_i q_j; the low half goes to column i + j, the high half to column i + j + 1)
FP_DEV constexpr uint64_t df_col_init(int k) {
  const int n_lo = (k <= 14) ? 2 * ((k < 14 - k ? k : 14 - k) + 1) : 0;
  const int n_hi = (k >= 1) ? 2 * ((k - 1 < 15 - k ? k - 1 : 15 - k) + 1) : 0;
  return 0ull - ((uint64_t)n_lo * DF_OFF_LO + (uint64_t)n_hi * DF_OFF_HI);
}

// twelve 32-bit words -> eight exact doubles below 2^48: limb 2e = word 3e + low half of word 3e+1,
// limb 2e+1 = high half of word 3e+1 + word 3e+2; each is placed in the mantissa of 2^52 and the 2^52 taken off
FP_DEV void df_to_limbs(double* L, const uint32_t* w) {
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const uint32_t w0 = w[3 * e], w1 = w[3 * e + 1], w2 = w[3 * e + 2];
    L[2 * e] = df_add(df_from_hilo(df_prmt(w1, 0x43300000u, 0x7610), w0), -0x1p52);
    L[2 * e + 1] = df_add(df_from_hilo(df_prmt(w2, 0x43300000u, 0x7632), df_prmt(w1, w2, 0x5432)), -0x1p52);
  }
}

// the two halves of a 96-bit limb product as bit patterns (exponent fields included)
FP_DEV void df_split(double a, double b, uint64_t& hi, uint64_t& lo) {
  const double h = df_fma_rz(a, b, 0x1p100);
  const double t = df_add(0x1p100 + 0x1p52, -h);
  const double l = df_fma_rz(a, b, t);
  hi = df_bits(h);
  lo = df_bits(l);
}

// c[i .. i+8] += x * y[0..7]
template <class Y>
FP_DEV void df_row(uint64_t* c, int i, double x, Y y) {
  uint64_t hi[DL], lo[DL];
#pragma unroll
  for (int j = 0; j < DL; j++) df_split(x, y(j), hi[j], lo[j]);
  c[i] += lo[0];
#pragma unroll
  for (int j = 1; j < DL; j++) c[i + j] += lo[j] + hi[j - 1];
  c[i + DL] += hi[DL - 1];
}

struct DfQ {
  FP_DEV double operator()(int j) const { return Q48D(j); }
};
struct DfArr {
  const double* p;
  FP_DEV double operator()(int j) const { return p[j]; }
};

FP_DEV void fp_mul_dfma(fp& r, const fp& a, const fp& b) {
  double A[DL], B[DL];
  df_to_limbs(A, a.v);
  df_to_limbs(B, b.v);
  uint64_t c[2 * DL];
#pragma unroll
  for (int k = 0; k < 2 * DL; k++) c[k] = df_col_init(k);
#pragma unroll
  for (int i = 0; i < DL; i++) {
    df_row(c, i, B[i], DfArr{A});
    // digit of column i (everything but the low half of m q_0 has arrived; its exponent field is a multiple of
    // 2^52) and m = d * (-1/q) mod 2^48, again by the splitting trick
    const uint64_t d = c[i] & DF_MASK48;
    const double D = df_add(df_from_bits(DF_OFF_LO | d), -0x1p52);
    const double h = df_fma_rz(D, (double)DF_QINV_NEG48, 0x1p100);
    const double t = df_add(0x1p100 + 0x1p52, -h);
    const double M = df_add(df_fma_rz(D, (double)DF_QINV_NEG48, t), -0x1p52);
    df_row(c, i, M, DfQ{});
    c[i + 1] += c[i] >> 48;  // column i is a multiple of 2^48 now
  }
  // T / R = sum of c[8 + k] 2^(48 k), columns below 2^64: even columns sit on word boundaries (3e), odd ones 16 bits
  // into word 3e + 1; the two 384-bit numbers are added with one carry chain
  uint32_t E[NL], O[NL + 1];
  O[0] = 0;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const uint64_t ve = c[DL + 2 * e], vo = c[DL + 2 * e + 1];
    E[3 * e] = (uint32_t)ve;
    E[3 * e + 1] = (uint32_t)(ve >> 32);
    E[3 * e + 2] = 0;
    O[3 * e + 1] = (uint32_t)vo << 16;
    O[3 * e + 2] = (uint32_t)(vo >> 16);
    O[3 * e + 3] = (uint32_t)(vo >> 48);  // joins the next group's even word; zero for the last group (T < 2^384)
  }
  // O[3e+3] shares word 3e+3 with nothing else in O, E[3e+2] is zero: fold them so that one chain suffices
  r.v[0] = add_cc(E[0], O[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) r.v[i] = addc_cc(E[i], O[i]);
  r.v[NL - 1] = addc(E[NL - 1], O[NL - 1]);
}

}  // namespace b200bls
