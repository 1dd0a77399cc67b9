// issue_probe.cu -- what does an ALU instruction cost next to IMAD.WIDE carry chains?  Stand-alone experiment:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/experiments/bin/issue_probe tools/experiments/issue_probe.cu
// Each thread runs four independent carry chains of mad.lo.cc / madc.hi.cc pairs (-> IMAD.WIDE.U32.X, the
// instruction of fp_mul) and, per IMAD.WIDE, K additions with carry on four other independent chains (-> IADD3 /
// IADD3.X, the instructions of fp_add).  K = 0, 1, 2, 4.  Reported per scheduler: cycles per IMAD.WIDE.  If the two
// pipes overlapped freely the figure would stay at 4.0 until K reaches 2 (ALU: 2 cycles per warp instruction).
// Variant "split": two warps of three run only IMAD.WIDE chains, the third only additions (same totals as K = 1).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int INNER = 32;

template <int K>
__device__ __forceinline__ void body(uint32_t (&lo)[8], uint32_t (&hi)[8], uint32_t (&s)[4][2], uint32_t a, uint32_t b) {
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a), "r"(b));
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi[c]) : "r"(a), "r"(b));
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo[c + 1]) : "r"(a), "r"(b));
    asm volatile("madc.hi.u32 %0, %1, %2, %0;" : "+r"(hi[c + 1]) : "r"(a), "r"(b));
  }
  // 8 IMAD.WIDE above (4 chains x 2); K additions per IMAD.WIDE = 8 K additions, on four chains
#pragma unroll
  for (int k = 0; k < K; k++) {
#pragma unroll
    for (int c = 0; c < 4; c++)
      asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(s[c][0]), "+r"(s[c][1]) : "r"(a), "r"(b));
  }
}

template <int K, bool SPLIT>
__global__ void __launch_bounds__(384, 1) probe(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
  uint32_t lo[8], hi[8], s[4][2];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    lo[c] = a + c;
    hi[c] = b ^ c;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) {
    s[c][0] = a * (c + 3);
    s[c][1] = b + c;
  }
  const int warp = threadIdx.x >> 5;
  if (SPLIT) {
    // warps 8..11 (one per scheduler) only add: 3 x the additions of a K = 1 warp; the others only multiply: 1.5 x
    if (warp >= 8) {
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int k = 0; k < INNER; k++)
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
              asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(s[c][0]), "+r"(s[c][1]) : "r"(a), "r"(b));
            }
    } else {
      for (int it = 0; it < iters; it++)
#pragma unroll
        for (int k = 0; k < INNER; k++) {
          body<0>(lo, hi, s, a, b);
          if (k & 1) body<0>(lo, hi, s, a, b);
        }
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < INNER; k++) body<K>(lo, hi, s, a, b);
      a += lo[0];
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < 4; c++) r ^= lo[c] ^ hi[c] ^ lo[c + 4] ^ hi[c + 4] ^ s[c][0] ^ s[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int K, bool SPLIT>
void run(const char* name, int n_sm, uint32_t* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    probe<K, SPLIT><<<n_sm, 384>>>(out, iters, 777u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  // per scheduler: 3 warps x iters x INNER x 8 IMAD.WIDE
  const double wide = 3.0 * iters * INNER * 8;
  const double cycles = best * 1e-3 * clk_khz * 1e3;
  printf("%-28s %8.3f ms  %.2f cycles per IMAD.WIDE per scheduler (nominal clock %d MHz)  %s\n", name, best, cycles / wide,
         clk_khz / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out;
  cudaMalloc(&out, (size_t)n_sm * 384 * 4);
  const int iters = 2000;
  run<0, false>("K=0 (IMAD.WIDE only)", n_sm, out, iters);
  run<1, false>("K=1 addition per IMAD.WIDE", n_sm, out, iters);
  run<2, false>("K=2", n_sm, out, iters);
  run<4, false>("K=4", n_sm, out, iters);
  run<1, true>("split warps, totals of K=1", n_sm, out, iters);
  return 0;
}
