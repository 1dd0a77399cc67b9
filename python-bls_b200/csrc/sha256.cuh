// sha256.cuh -- the SHA-256 stage of hash_to_point_prehashed_Fq2 (bls_py/ec.py:528-537).
//
// For a 32-byte message hash h the reference derives two Fq2 field elements from
//   hash512(h || "G2_j_ck") = sha256(h || label || 0x00) || sha256(h || label || 0x01)
// (bls_py/util.py:7-16), j, k in {0, 1}: eight single-block SHA-256 compressions per message
// (40 bytes of data each).  This stage writes the 4 x 64 digest bytes per message; the field
// VM then reduces each 512-bit big-endian integer mod q (hash_to_g2 program).
#pragma once
#include <stdint.h>

#ifdef B200BLS_HOSTSIM
#define SHA_HD inline
#else
#define SHA_HD __host__ __device__ __forceinline__
#endif

namespace b200bls {

#define B200BLS_SHA_K                                                                              \
  {0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, \
   0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, \
   0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, \
   0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, \
   0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, \
   0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, \
   0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, \
   0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2}

#ifdef B200BLS_HOSTSIM
static const uint32_t kShaK[64] = B200BLS_SHA_K;
#define SHA_K(i) kShaK[i]
#else
// compile-time immediates after full unrolling: no table in memory, no local array
__host__ __device__ __forceinline__ constexpr uint32_t sha_k(int i) {
  constexpr uint32_t t[64] = B200BLS_SHA_K;
  return t[i];
}
#define SHA_K(i) sha_k(i)
#endif

SHA_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

// SHA-256 of ONE already padded 64-byte block given as 16 big-endian words (overwritten: the
// message schedule is a rolling 16-word window); digest words to r[0..7].  Every loop is fully
// unrolled, so everything lives in registers.
SHA_HD void sha256_single_block(uint32_t* w, uint32_t* r) {
  uint32_t a = 0x6a09e667, b = 0xbb67ae85, c = 0x3c6ef372, d = 0xa54ff53a;
  uint32_t e = 0x510e527f, f = 0x9b05688c, g = 0x1f83d9ab, hh = 0x5be0cd19;
#pragma unroll
  for (int i = 0; i < 64; i++) {
    if (i >= 16) {
      uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
      uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
      w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
    }
    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + SHA_K(i) + w[i & 15];
    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  r[0] = 0x6a09e667 + a; r[1] = 0xbb67ae85 + b; r[2] = 0x3c6ef372 + c; r[3] = 0xa54ff53a + d;
  r[4] = 0x510e527f + e; r[5] = 0x9b05688c + f; r[6] = 0x1f83d9ab + g; r[7] = 0x5be0cd19 + hh;
}

// one compression of a single padded block holding `h` (32 bytes), a 7-byte label and one
// suffix byte; digest written big-endian to out[0..31].
SHA_HD void sha256_h_label(const uint8_t* h, int j, int k, int suffix, uint8_t* out) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 8; i++)
    w[i] = ((uint32_t)h[4 * i] << 24) | ((uint32_t)h[4 * i + 1] << 16) | ((uint32_t)h[4 * i + 2] << 8) | h[4 * i + 3];
  // label "G2_j_ck" then the suffix byte, then 0x80 padding and the bit length (40 bytes)
  w[8] = ((uint32_t)'G' << 24) | ((uint32_t)'2' << 16) | ((uint32_t)'_' << 8) | (uint32_t)('0' + j);
  w[9] = ((uint32_t)'_' << 24) | ((uint32_t)'c' << 16) | ((uint32_t)('0' + k) << 8) | (uint32_t)suffix;
  w[10] = 0x80000000u;
  w[11] = w[12] = w[13] = w[14] = 0;
  w[15] = 40 * 8;
  uint32_t r[8];
  sha256_single_block(w, r);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(r[i] >> 24);
    out[4 * i + 1] = (uint8_t)(r[i] >> 16);
    out[4 * i + 2] = (uint8_t)(r[i] >> 8);
    out[4 * i + 3] = (uint8_t)r[i];
  }
}

// group order n of BLS12-381, big-endian words
#define B200BLS_N_WORDS \
  {0x73eda753u, 0x299d7d48u, 0x3339d808u, 0x09a1d805u, 0x53bda402u, 0xfffe5bfeu, 0xffffffffu, 0x00000001u}

// Aggregation exponent T_i = SHA256(i as 4 big-endian bytes || pk_hash) mod n (bls_py/util.py:
// 46-49): one 36-byte single-block hash, then at most two subtractions of n (2^256 < 2.21 n).
// pk_hash as 8 big-endian words; result as 32 big-endian bytes, the scalar format of the
// scalar-multiplication programs.
SHA_HD void hash_pks_exponent(uint32_t index, const uint32_t* pk_hash, uint8_t* out) {
  const uint32_t nw[8] = B200BLS_N_WORDS;
  uint32_t w[16], r[8];
  w[0] = index;
#pragma unroll
  for (int i = 0; i < 8; i++) w[1 + i] = pk_hash[i];
  w[9] = 0x80000000u;
  w[10] = w[11] = w[12] = w[13] = w[14] = 0;
  w[15] = 36 * 8;
  sha256_single_block(w, r);
#pragma unroll
  for (int round = 0; round < 2; round++) {
    uint32_t d[8];
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 7; i >= 0; i--) {  // r - n, least significant word last in the array
      uint64_t t = (uint64_t)r[i] - nw[i] - borrow;
      d[i] = (uint32_t)t;
      borrow = (uint32_t)(t >> 63);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : d[i];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(r[i] >> 24);
    out[4 * i + 1] = (uint8_t)(r[i] >> 16);
    out[4 * i + 2] = (uint8_t)(r[i] >> 8);
    out[4 * i + 3] = (uint8_t)r[i];
  }
}

#ifndef B200BLS_HOSTSIM
// one thread per (message, label, suffix): 8 threads per message.
// out layout per message: [t0.c0 | t0.c1 | t1.c0 | t1.c1], 64 bytes each
__global__ void sha_stage_kernel(const uint8_t* __restrict__ hashes, uint8_t* __restrict__ out, long long n) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 8) return;
  long long item = t >> 3;
  int sub = (int)(t & 7);
  int j = sub >> 2, k = (sub >> 1) & 1, suffix = sub & 1;
  sha256_h_label(hashes + item * 32, j, k, suffix, out + item * 256 + (j * 2 + k) * 64 + suffix * 32);
}

// T_first .. T_(first + n - 1), one thread each; out: n x 32 bytes
__global__ void hash_pks_kernel(const uint8_t* __restrict__ pk_hash, uint32_t first, uint8_t* __restrict__ out,
                                long long n) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; i++)
    h[i] = ((uint32_t)pk_hash[4 * i] << 24) | ((uint32_t)pk_hash[4 * i + 1] << 16) | ((uint32_t)pk_hash[4 * i + 2] << 8) |
           pk_hash[4 * i + 3];
  hash_pks_exponent(first + (uint32_t)t, h, out + t * 32);
}
#endif

}  // namespace b200bls
