// Stand-alone check that Tensor Memory can serve as per-thread scratch for a non-tensor
// kernel: every thread of a 128-thread CTA owns one TMEM lane (512 B .. 1 KB of columns),
// written with tcgen05.st and read back with tcgen05.ld at *dynamic* column offsets.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_scratch_test tmem_scratch_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int NCOLS>
__global__ void __launch_bounds__(128) tmem_test(uint32_t* mismatches, int rounds) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);   // this warp's lane quadrant
  const int n_slots = NCOLS / 24;                                 // 24 words = one Fq2 value
  uint32_t bad = 0;
  for (int r = 0; r < rounds; r++) {
    for (int s = 0; s < n_slots; s++) {
      uint32_t v[24];
#pragma unroll
      for (int k = 0; k < 24; k++) v[k] = (threadIdx.x * 1000003u) ^ (blockIdx.x * 7919u) ^ (s * 131u + k + r * 17u);
      uint32_t addr = base + s * 24;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]),
                   "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr + 8), "r"(v[8]),
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr + 16), "r"(v[16]),
                   "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    for (int s = n_slots - 1; s >= 0; s--) {
      uint32_t v[24];
      uint32_t addr = base + s * 24;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(addr));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(addr + 8));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23])
                   : "r"(addr + 16));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int k = 0; k < 24; k++)
        bad += v[k] != ((threadIdx.x * 1000003u) ^ (blockIdx.x * 7919u) ^ (s * 131u + k + r * 17u));
    }
  }
  if (bad) atomicAdd(mismatches, bad);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "n"(NCOLS));
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4);
  for (int cfg = 0; cfg < 3; cfg++) {
    cudaMemset(d, 0, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int grid = 148 * (cfg == 0 ? 2 : (cfg == 1 ? 3 : 4)) * 4;     // several waves
    const int rounds = 2000;
    cudaEventRecord(e0);
    if (cfg == 0) tmem_test<256><<<grid, 128>>>(d, rounds);
    if (cfg == 1) tmem_test<128><<<grid, 128>>>(d, rounds);
    if (cfg == 2) tmem_test<128><<<grid, 128>>>(d, rounds);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    uint32_t h = 123;
    cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    int cols = cfg == 0 ? 256 : 128;
    double bytes = (double)grid * 128 * rounds * (cols / 24) * 96 * 2;
    printf("cfg %d cols %d grid %d: %s, mismatches %u, %.3f ms, %.1f GB/s ld+st\n", cfg, cols, grid,
           cudaGetErrorString(err), h, ms, bytes / ms / 1e6);
  }
  return 0;
}
