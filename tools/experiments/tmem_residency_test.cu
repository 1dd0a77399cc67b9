// Does the presence / execution of tcgen05.alloc change how many CTAs share an SM?
// Each CTA: 128 threads, 100 KB dynamic smem, a fixed-length dependent ALU loop.  Grid = 2 x SMs:
// co-resident CTAs finish in ~T, serialised ones in ~2T.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>   // 0: no tcgen05 code, 1: code present, not executed, 2: alloc 128 cols, 3: alloc 256 cols
__global__ void __launch_bounds__(128, 2) k(uint32_t* out, int iters, int do_alloc) {
  extern __shared__ uint32_t sm[];
  __shared__ uint32_t s_base;
  if (MODE >= 1 && do_alloc) {
    if (threadIdx.x < 32) {
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_base);
      if (MODE == 3)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(dst));
      else
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(dst));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  uint32_t x = threadIdx.x + blockIdx.x;
  for (int i = 0; i < iters; i++) x = x * 1664525u + 1013904223u;
  sm[threadIdx.x] = x;
  out[blockIdx.x * 128 + threadIdx.x] = sm[threadIdx.x];
  if (MODE >= 1 && do_alloc) {
    __syncthreads();
    if (threadIdx.x < 32) {
      if (MODE == 3)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(s_base));
      else
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(s_base));
    }
  }
}

template <int MODE>
float run(uint32_t* d, int grid, int do_alloc) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    k<MODE><<<grid, 128, 100 * 1024>>>(d, 2000000, do_alloc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 148 * 4 * 128 * 4);
  for (int mult = 1; mult <= 2; mult++) {
    int grid = 148 * mult;
    printf("grid %d: no-tcgen05 %.2f ms | present-not-executed %.2f ms | alloc128 %.2f ms | alloc256 %.2f ms\n", grid,
           run<0>(d, grid, 0), run<1>(d, grid, 0), run<2>(d, grid, 1), run<3>(d, grid, 1));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
