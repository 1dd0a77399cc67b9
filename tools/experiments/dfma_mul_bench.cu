// dfma_mul_bench.cu -- does a second Montgomery multiplier on the FP64 pipe (csrc/fp_dfma.cuh) add throughput next
// to the integer one?  Stand-alone experiment (not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DB200BLS_MUL_CALL -Xptxas -v \
//        -o gpurun_out/dfma_mul_bench tools/experiments/dfma_mul_bench.cu
// One CTA of 384 threads per SM (the pairing kernel's shape).  Every warp takes chunks of 64 dependent products
// x <- x * y (+ `nadd` modular additions after each, standing in for the interpreter's non-multiplying work) from a
// global counter until the work is gone.  mode 0: all products on the integer pipe; 1: all on the FP64 pipe;
// 2: warps with (warp % 3 == 0) use the FP64 pipe (one per scheduler), the others the integer pipe; 3 / 4: every warp
// sends every third / second product to the FP64 pipe; 5: pairs of independent products through ONE function with both
// products' carry chains interleaved (fp_mul_dual_call); 6: the same pairs as two single calls (its control).  Prints products / s and the share each multiplier did, and
// checks that every thread's result equals the all-integer result.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../python-bls_b200/csrc/fp.cuh"
#include "fp_dfma.cuh"

using namespace b200bls;

static __device__ __noinline__ fp fp_mul_dfma_call(fp a, fp b) {
  fp r;
  fp_mul_dfma(r, a, b);
  return r;
}

// TWO independent products in one instruction stream (twice the carry chains in flight per warp): the candidate for
// the t0 / t1 pair of a Karatsuba Fq2 product
struct fp_pair {
  fp a, b;
};
static __device__ __noinline__ fp_pair fp_mul_dual_call(fp a0, fp b0, fp a1, fp b1) {
  uint32_t ev0[NL], od0[NL], ev1[NL], od1[NL];
  mont_round_first(ev0, od0, a0.v, b0.v[0]);
  mont_round_first(ev1, od1, a1.v, b1.v[0]);
#pragma unroll
  for (int i = 1; i < NL; i += 2) {
    mont_round(od0, ev0, a0.v, b0.v[i]);
    mont_round(od1, ev1, a1.v, b1.v[i]);
    if (i + 1 < NL) {
      mont_round(ev0, od0, a0.v, b0.v[i + 1]);
      mont_round(ev1, od1, a1.v, b1.v[i + 1]);
    }
  }
  fp_pair r;
  mont_finish(r.a, ev0, od0);
  mont_finish(r.b, ev1, od1);
  return r;
}

constexpr int CHUNK = 64;

__global__ void __launch_bounds__(384, 1) bench_kernel(int mode, int nadd, int n_chunks, int* counter,
                                                       unsigned long long* done, uint32_t* out) {
  // jitter: the additions per product vary per warp and iteration (uniform in 0 .. 2 nadd), which takes the warps of a
  // scheduler out of lock step -- the interpreter's situation, where instructions differ in length
  const bool jitter = (mode & 8) != 0;
  mode &= 7;
  uint32_t rnd = 12345u + 7919u * (threadIdx.x >> 5) + 104729u * blockIdx.x;
  fp x, y, x2;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    x.v[i] = 0x01234567u * (i + 1) + threadIdx.x * 977u + blockIdx.x * 131071u;
    y.v[i] = 0x89abcdefu * (i + 3) ^ (threadIdx.x * 7919u);
    x2.v[i] = x.v[i] ^ 0x5a5a5a5au;
  }
  x.v[NL - 1] &= 0x0fffffffu;  // < 2^380 < q
  y.v[NL - 1] &= 0x0fffffffu;
  x2.v[NL - 1] &= 0x0fffffffu;
  if (mode >= 5) {
    // mode 5: pairs of independent products through the dual function; mode 6: the same pairs as two single calls
    // (the control: same work, same additions).  A chunk is CHUNK / 2 pairs.
    unsigned long long n_int = 0;
    for (;;) {
      int c = 0;
      if ((threadIdx.x & 31) == 0) c = atomicAdd(counter, 1);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= n_chunks) break;
      for (int k = 0; k < CHUNK / 2; k++) {
        if (mode == 5) {
          fp_pair r = fp_mul_dual_call(x, y, x2, y);
          x = r.a;
          x2 = r.b;
        } else {
          x = fp_mul_call(x, y);
          x2 = fp_mul_call(x2, y);
        }
        n_int += 2;
        rnd = rnd * 1664525u + 1013904223u;
        const int na = jitter ? (int)((rnd >> 16) % (2 * nadd + 1)) : nadd;
        for (int j = 0; j < na; j++) {
          fp_add(x, x, y);
          fp_add(x2, x2, y);
        }
      }
    }
    if ((threadIdx.x & 31) == 0) atomicAdd(done, n_int);
    uint32_t h = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) h ^= x.v[i] ^ x2.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = h;
    return;
  }
  const int warp = threadIdx.x >> 5;
  unsigned long long n_int = 0, n_fp = 0;
  int phase = warp;
  for (;;) {
    int c = 0;
    if ((threadIdx.x & 31) == 0) c = atomicAdd(counter, 1);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= n_chunks) break;
    for (int k = 0; k < CHUNK; k++) {
      bool use_fp;
      if (mode == 0) use_fp = false;
      else if (mode == 1) use_fp = true;
      else if (mode == 2) use_fp = (warp % 3) == 0;
      else if (mode == 3) use_fp = (phase % 3) == 0;
      else use_fp = (phase & 1) == 0;
      phase++;
      if (use_fp) {
        x = fp_mul_dfma_call(x, y);
        n_fp++;
      } else {
        x = fp_mul_call(x, y);
        n_int++;
      }
      rnd = rnd * 1664525u + 1013904223u;
      const int na = jitter ? (int)((rnd >> 16) % (2 * nadd + 1)) : nadd;
      for (int j = 0; j < na; j++) fp_add(x, x, y);
    }
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(done, n_int);
    atomicAdd(done + 1, n_fp);
  }
  uint32_t h = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) h ^= x.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = h;
}

// parity on the device: every thread multiplies the same pairs both ways
__global__ void parity_kernel(int iters, int* bad) {
  fp x, y;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    x.v[i] = 0x9e3779b9u * (i + 1) + threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    y.v[i] = 0x7f4a7c15u * (i + 5) ^ (threadIdx.x * 69069u + blockIdx.x);
  }
  x.v[NL - 1] &= 0x3fffffffu;
  y.v[NL - 1] &= 0x3fffffffu;
  for (int k = 0; k < iters; k++) {
    fp u = fp_mul_call(x, y);
    fp w = fp_mul_dfma_call(x, y);
    bool ne = false;
#pragma unroll
    for (int i = 0; i < NL; i++) ne |= u.v[i] != w.v[i];
    if (ne) atomicAdd(bad, 1);
    y = x;
    x = u;
  }
}

int main(int argc, char** argv) {
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  int* counter;
  unsigned long long* done;
  uint32_t* out;
  int* bad;
  cudaMalloc(&counter, 4);
  cudaMalloc(&done, 16);
  cudaMalloc(&bad, 4);
  cudaMalloc(&out, (size_t)n_sm * 384 * 4);
  cudaMemset(bad, 0, 4);
  parity_kernel<<<n_sm, 128>>>(200, bad);
  int hbad = -1;
  cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
  printf("parity: %d mismatches in %d products (%s)\n", hbad, n_sm * 128 * 200, cudaGetErrorString(cudaGetLastError()));
  const int chunks_per_warp = argc > 1 ? atoi(argv[1]) : 24;
  const int n_chunks = n_sm * 12 * chunks_per_warp;
  uint32_t hout5[64], hout6[64];
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int nadd = 0; nadd <= 8; nadd += 4) {
    for (int mi = 0; mi < 14; mi++) {
      const int mode = (mi % 7) | (mi >= 7 ? 8 : 0);
      if (mi >= 7 && nadd == 0) continue;
      float best = 1e30f;
      unsigned long long h[2] = {0, 0};
      for (int rep = 0; rep < 3; rep++) {
        cudaMemset(counter, 0, 4);
        cudaMemset(done, 0, 16);
        cudaEventRecord(e0);
        bench_kernel<<<n_sm, 384>>>(mode, nadd, n_chunks, counter, done, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        cudaMemcpy(h, done, 16, cudaMemcpyDeviceToHost);
      }
      const double prods = (double)n_chunks * CHUNK * 32;
      if ((mode & 7) == 5) cudaMemcpy(hout5, out, 4 * 64, cudaMemcpyDeviceToHost);
      if (mode == 6) {
        cudaMemcpy(hout6, out, 4 * 64, cudaMemcpyDeviceToHost);
        int same = 1;
        for (int i = 0; i < 64; i++) same &= hout5[i] == hout6[i];
        printf("dual == two singles on 64 threads: %s\n", same ? "yes" : "NO");
      }
      printf("nadd %d mode %d: %8.3f ms  %.4e products/s  (x300 = %.3e limb-product equivalents/s)  int %.1f %%  fp64 %.1f %%\n",
             nadd, mode, best, prods / (best * 1e-3), 300 * prods / (best * 1e-3),
             100.0 * h[0] / (double)(h[0] + h[1]), 100.0 * h[1] / (double)(h[0] + h[1]));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
