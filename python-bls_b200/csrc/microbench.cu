// microbench.cu -- measures the B200's integer-multiply peak, the roofline denominator of
// this library (MEASURED_PEAKS.json has no INT32 figure).
//
// Register-only kernels, 8 independent accumulator chains per thread, full occupancy:
//   variant 0: IMAD      (mad.lo.u32)           -> 32-bit multiply-adds / s
//   variant 1: IMAD.HI   (mad.hi.u32)
//   variant 2: IMAD.WIDE (mad.wide.u32)         -> 32x32+64 limb products / s
//   variant 3: IMAD.WIDE.U32.X carry chains (mad.lo.cc/madc.hi.cc pairs, the instruction
//              mix of fp_mul)                  -> limb products / s
//   variant 4: DFMA (fma.rn.f64), 8 independent chains -> FP64 FMAs / s (the FP64 pipe is separate
//              from the integer-multiply pipe: a candidate second multiplier for 52-bit limbs)
//   variant 5: even warps run variant 3, odd warps variant 4, same instruction count each ->
//              total instructions / s; tells whether the two pipes overlap
// "limb-product peak" = max(variant2, variant3, min(variant0, variant1)/2).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200bls.h"

namespace {

constexpr int CHAINS = 8;
constexpr int INNER = 64;

__device__ __forceinline__ uint32_t dfma_body(int iters, uint32_t seed) {
  double x[CHAINS];
  const double m = 1.0 + (double)(seed & 7) * 1e-9, c = 1e-3;
#pragma unroll
  for (int k = 0; k < CHAINS; k++) x[k] = 1.0 + k * 1e-3 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < INNER; k++) {
#pragma unroll
      for (int ch = 0; ch < CHAINS; ch++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[ch]) : "d"(m), "d"(c));
    }
  }
  double r = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; k++) r += x[k];
  return (uint32_t)__double2int_rn(r * 1e-6);
}

template <int VARIANT>
__global__ void __launch_bounds__(1024) imad_kernel(uint32_t* out, int iters, uint32_t seed) {
  if (VARIANT == 4 || (VARIANT == 5 && ((threadIdx.x >> 5) & 1))) {
    out[blockIdx.x * blockDim.x + threadIdx.x] = dfma_body(iters, seed);
    return;
  }
  uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
  uint32_t lo[CHAINS], hi[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) {
    lo[c] = a + c;
    hi[c] = b ^ c;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < INNER; k++) {
      if (VARIANT == 0) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(b), "r"(a));
      } else if (VARIANT == 1) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(b), "r"(a));
      } else if (VARIANT == 2) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
          uint64_t acc = ((uint64_t)hi[c] << 32) | lo[c];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(lo[c]), "r"(b));
          lo[c] = (uint32_t)acc;
          hi[c] = (uint32_t)(acc >> 32);
        }
      } else if (VARIANT == 6) {
        // FOUR independent carry chains of two column pairs each, interleaved by the compiler: the instruction-level
        // parallelism a Montgomery round really has (variant 3 is one serial chain)
#pragma unroll
        for (int c = 0; c < CHAINS; c += 2) {
          asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(a), "r"(b));
          asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c + 1]) : "r"(a), "r"(b));
          asm volatile("madc.hi.u32 %0, %1, %2, %0;" : "+r"(hi[c + 1]) : "r"(a), "r"(b));
        }
      } else {
        // one carry chain across the 8 column pairs, exactly like a row of fp_mul
        asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[0]) : "r"(a), "r"(b));
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[0]) : "r"(a), "r"(b));
#pragma unroll
        for (int c = 1; c < CHAINS; c++) {
          asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(a), "r"(b));
        }
      }
    }
    a += lo[0];
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) r ^= lo[c] ^ hi[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace

extern "C" int b200bls_microbench_imad(int variant, int blocks_per_sm, int threads, int iters,
                                       double* ops_per_second, float* ms_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B200BLS_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return B200BLS_E_CUDA;
  if (threads < 32 || threads > 1024 || blocks_per_sm < 1) return B200BLS_E_ARG;
  const int blocks = prop.multiProcessorCount * blocks_per_sm;
  uint32_t* out = nullptr;
  if (cudaMalloc(&out, (size_t)threads * blocks * 4) != cudaSuccess) return B200BLS_E_NOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    switch (variant) {
      case 0: imad_kernel<0><<<blocks, threads>>>(out, iters, 12345u + rep); break;
      case 1: imad_kernel<1><<<blocks, threads>>>(out, iters, 12345u + rep); break;
      case 2: imad_kernel<2><<<blocks, threads>>>(out, iters, 12345u + rep); break;
      case 4: imad_kernel<4><<<blocks, threads>>>(out, iters, 12345u + rep); break;
      case 5: imad_kernel<5><<<blocks, threads>>>(out, iters, 12345u + rep); break;
      case 6: imad_kernel<6><<<blocks, threads>>>(out, iters, 12345u + rep); break;
      default: imad_kernel<3><<<blocks, threads>>>(out, iters, 12345u + rep); break;
    }
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      cudaFree(out);
      return B200BLS_E_CUDA;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  double ops = (double)threads * blocks * (double)iters * INNER * CHAINS;
  if (ops_per_second) *ops_per_second = ops / (best * 1e-3);
  if (ms_out) *ms_out = best;
  return 0;
}
