// Register-budget probe for a 24-warp design ("two threads per item"): an Fq
// product whose operands STREAM from shared memory -- a in registers (3 LDS.128), b one uint4 per four
// rounds, result stored with 3 STS.128 -- inside a minimal Fq-granular interpreter loop.  Build only
// (no GPU needed) to read the register count:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -DB200BLS_MUL_CALL \
//        -I python-bls_b200/csrc -c tools/experiments/lean_mul_regs.cu -o /dev/null
// 24 warps per SM leave 65536 / 768 = 85 registers per thread.
#include <cuda_runtime.h>

#include "fp.cuh"

using namespace b200bls;

constexpr int NT = 768;

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// cell c of this thread: chunks at base + (c * 3 + k) * NT * 16
__device__ __noinline__ void fp_mul_stream(uint32_t base, int d, int ca, int cb) {
  fp a;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    uint4 v = lds128(base + (ca * 3 + k) * NT * 16);
    a.v[4 * k] = v.x, a.v[4 * k + 1] = v.y, a.v[4 * k + 2] = v.z, a.v[4 * k + 3] = v.w;
  }
  uint32_t ev[NL], od[NL];
  uint4 b = lds128(base + (cb * 3) * NT * 16);
  mont_round_first(ev, od, a.v, b.x);
  mont_round(od, ev, a.v, b.y);
  mont_round(ev, od, a.v, b.z);
  mont_round(od, ev, a.v, b.w);
#pragma unroll
  for (int k = 1; k < 3; k++) {
    b = lds128(base + (cb * 3 + k) * NT * 16);
    mont_round(ev, od, a.v, b.x);
    mont_round(od, ev, a.v, b.y);
    mont_round(ev, od, a.v, b.z);
    mont_round(od, ev, a.v, b.w);
  }
  fp t;
  t.v[0] = add_cc(ev[0], od[1]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) t.v[i] = addc_cc(ev[i], od[i + 1]);
  t.v[NL - 1] = addc(ev[NL - 1], 0);
#pragma unroll
  for (int k = 0; k < 3; k++)
    sts128(base + (d * 3 + k) * NT * 16, make_uint4(t.v[4 * k], t.v[4 * k + 1], t.v[4 * k + 2], t.v[4 * k + 3]));
}

__device__ __noinline__ void fp_addsub_stream(uint32_t base, int d, int ca, int cb, bool sub) {
  fp a, b, r;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    uint4 v = lds128(base + (ca * 3 + k) * NT * 16);
    a.v[4 * k] = v.x, a.v[4 * k + 1] = v.y, a.v[4 * k + 2] = v.z, a.v[4 * k + 3] = v.w;
    v = lds128(base + (cb * 3 + k) * NT * 16);
    b.v[4 * k] = v.x, b.v[4 * k + 1] = v.y, b.v[4 * k + 2] = v.z, b.v[4 * k + 3] = v.w;
  }
  if (sub)
    fp_sub(r, a, b);
  else
    fp_add(r, a, b);
#pragma unroll
  for (int k = 0; k < 3; k++)
    sts128(base + (d * 3 + k) * NT * 16, make_uint4(r.v[4 * k], r.v[4 * k + 1], r.v[4 * k + 2], r.v[4 * k + 3]));
}

extern "C" __global__ void __launch_bounds__(NT, 1) lean_vm(const uint2* code, int n_ins, uint4* io) {
  extern __shared__ uint4 ws[];
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(ws) + threadIdx.x * 16;
  for (int k = 0; k < 9; k++) ws[k * NT + threadIdx.x] = io[(blockIdx.x * 9 + k) * NT + threadIdx.x];
  for (int pc = 0; pc < n_ins; pc++) {
    const uint2 w = code[pc];
    const int op = w.x & 0xff, d = (w.x >> 16) & 0xfff, a = w.y & 0xffff, b = w.y >> 16;
    if (op == 0)
      fp_mul_stream(base, d, a, b);
    else
      fp_addsub_stream(base, d, a, b, op == 2);
  }
  for (int k = 0; k < 9; k++) io[(blockIdx.x * 9 + k) * NT + threadIdx.x] = ws[k * NT + threadIdx.x];
}
