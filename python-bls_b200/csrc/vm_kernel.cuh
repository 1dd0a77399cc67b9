// vm_kernel.cuh -- the sm_100a kernel that interprets b200-bls field programs.
//
// Launch shape: one CTA of VM_NT threads per SM (persistent over the item loop), one batch
// item (pairing, signature, point ...) per thread.  Each thread owns `n_slots` Fq2 slots
// (2 fp cells each) of shared memory laid out [cell][16-byte chunk][thread], so a warp's
// LDS.128/STS.128 touches 512 contiguous bytes: conflict free.  With 18 slots x 128 threads
// the CTA uses 216 KB of the 227 KB carve-out; values that do not fit are spilled by the
// program itself (SPILL2/FILL2) to a global-memory cold area laid out [slot][chunk][thread]
// (fully coalesced, L2 resident).  The integer-multiply pipe is the roofline;
// shared memory and L2 traffic are an order of magnitude below their limits.
#pragma once
#include <cuda_runtime.h>

#include "vm_exec.cuh"
#include "vm_params.h"

namespace b200bls {

// ---- Tensor Memory as per-thread scratch --------------------------------------------------------
// TMEM is 512 columns x 128 lanes x 32 bit per SM.  With the 32x32b access shape, thread i of
// warp w reads/writes lane 32*(w%4)+i: a 128-thread CTA therefore owns one private lane per
// thread, `tmem_cols` words deep -- a second on-chip workspace next to shared memory (no tensor
// core is involved; tcgen05.ld/st only).  An fp cell is 12 columns, an Fq2 slot 24.
__device__ __forceinline__ void tm_ld8(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr));
}
__device__ __forceinline__ void tm_ld4(uint32_t addr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(addr));
}
__device__ __forceinline__ void tm_st8(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tm_st4(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-state-space accesses with a 32-bit address computed once per kernel: the generic
// pointer form makes ptxas rebuild the shared window base (S2UR CgaCtaId + ULEA) per access
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <bool USE_TMEM, int NT>
struct DevEnv {
  uint32_t sm;             // shared workspace (shared-space byte address), already offset by threadIdx.x
  uint32_t tm_base;        // TMEM address of column 0 in this warp's lane quadrant
  int smem_cells;
  const VmParams* p;
  uint4* cold;             // cold area, already offset by the global thread id
  long long total;
  long long item_raw, item;
  uint32_t flags;

  __device__ __forceinline__ static void unpack(fp& x, int k, uint4 v) {
    x.v[4 * k] = v.x;
    x.v[4 * k + 1] = v.y;
    x.v[4 * k + 2] = v.z;
    x.v[4 * k + 3] = v.w;
  }
  __device__ __forceinline__ static uint4 pack(const fp& x, int k) {
    return make_uint4(x.v[4 * k], x.v[4 * k + 1], x.v[4 * k + 2], x.v[4 * k + 3]);
  }
  // the cell index comes from the instruction word: the shared/tensor-memory branch is
  // warp-uniform, as tcgen05.ld/st (.sync.aligned) require
#define VM_IN_SMEM(c) (!USE_TMEM || (c) < smem_cells)
  __device__ __forceinline__ void ld1(int c, fp& x) {
    if (VM_IN_SMEM(c)) {
      const uint32_t q = sm + c * (3 * NT * 16);
#pragma unroll
      for (int k = 0; k < 3; k++) unpack(x, k, lds128(q + k * (NT * 16)));
    } else {
      const uint32_t t = tm_base + (uint32_t)(c - smem_cells) * 12;
      tm_wait_st();
      tm_ld8(t, x.v);
      tm_ld4(t + 8, x.v + 8);
      tm_wait_ld();
    }
  }
  __device__ __forceinline__ void st1(int c, const fp& x) {
    if (VM_IN_SMEM(c)) {
      const uint32_t q = sm + c * (3 * NT * 16);
#pragma unroll
      for (int k = 0; k < 3; k++) sts128(q + k * (NT * 16), pack(x, k));
    } else {
      const uint32_t t = tm_base + (uint32_t)(c - smem_cells) * 12;
      tm_st8(t, x.v);
      tm_st4(t + 8, x.v + 8);
    }
  }
  __device__ __forceinline__ void ld2(int c, fp2& x) {
    if (VM_IN_SMEM(c)) {
      const uint32_t q = sm + c * (3 * NT * 16);
#pragma unroll
      for (int k = 0; k < 3; k++) unpack(x.c0, k, lds128(q + k * (NT * 16)));
#pragma unroll
      for (int k = 0; k < 3; k++) unpack(x.c1, k, lds128(q + (3 + k) * (NT * 16)));
    } else {
      const uint32_t t = tm_base + (uint32_t)(c - smem_cells) * 12;
      tm_wait_st();
      tm_ld8(t, x.c0.v);
      tm_ld4(t + 8, x.c0.v + 8);
      tm_ld8(t + 12, x.c1.v);
      tm_ld4(t + 20, x.c1.v + 8);
      tm_wait_ld();
    }
  }
  __device__ __forceinline__ void st2(int c, const fp2& x) {
    if (VM_IN_SMEM(c)) {
      const uint32_t q = sm + c * (3 * NT * 16);
#pragma unroll
      for (int k = 0; k < 3; k++) sts128(q + k * (NT * 16), pack(x.c0, k));
#pragma unroll
      for (int k = 0; k < 3; k++) sts128(q + (3 + k) * (NT * 16), pack(x.c1, k));
    } else {
      const uint32_t t = tm_base + (uint32_t)(c - smem_cells) * 12;
      tm_st8(t, x.c0.v);
      tm_st4(t + 8, x.c0.v + 8);
      tm_st8(t + 12, x.c1.v);
      tm_st4(t + 20, x.c1.v + 8);
    }
  }
  __device__ __forceinline__ void ld2_lane(int c, int off, fp2& x) {
    int t = (threadIdx.x + off) % NT;
    const uint32_t q = sm + (t - (int)threadIdx.x) * 16 + c * (3 * NT * 16);
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x.c0, k, lds128(q + k * (NT * 16)));
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x.c1, k, lds128(q + (3 + k) * (NT * 16)));
  }
  __device__ __forceinline__ void ldc(int idx, fp& x) {
    const uint4* q = p->consts + idx * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x, k, __ldg(q + k));
  }
  __device__ __forceinline__ void set_flag(int f, bool v) {
    flags = (flags & ~(1u << f)) | ((v ? 1u : 0u) << f);
  }
  __device__ __forceinline__ bool get_flag(int f) { return (flags >> f) & 1u; }
  __device__ __forceinline__ bool any_flag(int f) { return __any_sync(0xffffffffu, (flags >> f) & 1u); }
  __device__ __forceinline__ bool active() { return item_raw < p->n_items; }
  __device__ __forceinline__ uint32_t ld_byte(int buf, int off) {
    return p->bufs[buf].ptr[item * p->bufs[buf].stride + off];
  }
  // block_only: thread 0 of the CTA writes the record of item = CTA index (reduction results)
  __device__ __forceinline__ void st_byte(int buf, int off, unsigned char v, bool block_only) {
    if (block_only) {
      if (threadIdx.x == 0) p->bufs[buf].ptr[(long long)blockIdx.x * p->bufs[buf].stride + off] = v;
    } else if (active()) {
      p->bufs[buf].ptr[item * p->bufs[buf].stride + off] = v;
    }
  }
  __device__ __forceinline__ void ld_be(int buf, int off, int nwords, fp& x) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p->bufs[buf].ptr + item * p->bufs[buf].stride + off);
#pragma unroll
    for (int i = 0; i < NL; i++) x.v[i] = 0;
    if (nwords == 12) {
#pragma unroll
      for (int i = 0; i < 12; i++) x.v[i] = __byte_perm(__ldg(w + 11 - i), 0, 0x0123);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) x.v[i] = __byte_perm(__ldg(w + 7 - i), 0, 0x0123);
    }
  }
  __device__ __forceinline__ void st_be48(int buf, int off, const fp& x, bool block_only) {
    long long it = item;
    if (block_only) {
      if (threadIdx.x != 0) return;
      it = blockIdx.x;
    } else if (!active()) {
      return;
    }
    uint32_t* w = reinterpret_cast<uint32_t*>(p->bufs[buf].ptr + it * p->bufs[buf].stride + off);
#pragma unroll
    for (int i = 0; i < NL; i++) w[11 - i] = __byte_perm(x.v[i], 0, 0x0123);
  }
  __device__ __forceinline__ void ld_raw2(int buf, int elem, fp2& x) {
    const uint4* q = reinterpret_cast<const uint4*>(p->bufs[buf].ptr) + (long long)elem * 6 * p->bufs[buf].stride + item;
    long long s = p->bufs[buf].stride;
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x.c0, k, q[k * s]);
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x.c1, k, q[(3 + k) * s]);
  }
  __device__ __forceinline__ void st_raw2(int buf, int elem, const fp2& x, bool block_only) {
    long long it;
    if (block_only) {
      if (threadIdx.x != 0) return;
      it = blockIdx.x;
    } else {
      if (!active()) return;
      it = item_raw;
    }
    long long s = p->bufs[buf].stride;
    uint4* q = reinterpret_cast<uint4*>(p->bufs[buf].ptr) + (long long)elem * 6 * s + it;
#pragma unroll
    for (int k = 0; k < 3; k++) q[k * s] = pack(x.c0, k);
#pragma unroll
    for (int k = 0; k < 3; k++) q[(3 + k) * s] = pack(x.c1, k);
  }
  __device__ __forceinline__ void st_cold(int g, const fp2& x) {
    uint4* q = cold + (long long)g * 6 * total;
#pragma unroll
    for (int k = 0; k < 3; k++) __stcg(q + k * total, pack(x.c0, k));
#pragma unroll
    for (int k = 0; k < 3; k++) __stcg(q + (3 + k) * total, pack(x.c1, k));
  }
  __device__ __forceinline__ void ld_cold(int g, fp2& x) {
    const uint4* q = cold + (long long)g * 6 * total;
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x.c0, k, __ldcg(q + k * total));
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x.c1, k, __ldcg(q + (3 + k) * total));
  }
  // The cold copy in slot g is dead (builder: FILL2 with aux = 1).  Without this every spilled value is written
  // back to DRAM when its lines leave the L2, although nobody will read them again: discard.L2 drops the
  // lines instead.  A 128-byte line holds the same chunk of 8 consecutive threads, all of the same warp and all
  // at the same instruction, so one lane in eight discards it.
  __device__ __forceinline__ void discard_cold(int g) {
    if ((threadIdx.x & 7) == 0) {
      const uint4* q = cold + (long long)g * 6 * total;
#pragma unroll
      for (int k = 0; k < 6; k++) asm volatile("discard.global.L2 [%0], 128;" ::"l"(q + k * total) : "memory");
    }
  }
  __device__ __forceinline__ void sync() { __syncthreads(); }
};

// The next instruction word is loaded before the current one executes and is only consumed
// after it: the register move below is volatile, so the compiler cannot hoist it (and with it
// the wait for the load) in front of the opcode switch -- measured: that wait was 4.7 % of
// all warp stall samples.  One L1 line holds 16 instructions; the line after next is
// prefetched when a line boundary is crossed.
template <class Env>
__device__ __forceinline__ void vm_run_section(Env& env, const uint2* code, int lo, int hi) {
  int pc = lo;
  if (pc >= hi) return;
  uint2 ins = __ldg(code + pc);
  while (pc < hi) {
    uint2 nxt = __ldg(code + pc + 1);  // the program is padded: always readable
    if ((pc & 15) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(code + pc + 32));
    int skip = vm_exec(env, ins.x, ins.y);
    if (skip) {
      pc += 1 + skip;
      if (pc < hi) nxt = __ldg(code + pc);
    } else {
      pc += 1;
    }
    asm volatile("mov.b32 %0, %2;\n\tmov.b32 %1, %3;" : "=r"(ins.x), "=r"(ins.y) : "r"(nxt.x), "r"(nxt.y));
  }
}

// USE_TMEM = false: no tcgen05 instruction in the kernel at all.  (Measured on B200: a kernel
// that merely CONTAINS tcgen05.alloc holds the SM's allocation permit from CTA launch until it
// executes relinquish_alloc_permit or exits, which serialises co-resident CTAs -- see
// tools/experiments/tmem_residency_test.cu.  Hence two instantiations, and the TMEM one always
// allocates and relinquishes first thing.)
// WARP_FETCH: item blocks of 32 per warp (isolated batches, see launch_program) instead of CTA-wide blocks.  A
// template parameter, not a run-time branch: the mere presence of the second fetch path in the section loop made
// ptxas if-convert the shared / Tensor-Memory operand loads of every opcode body (192 instead of 60 predicated
// LDS, +5 % executed instructions, 33.1 -> 34.4 ms per wave of pairings).
template <bool USE_TMEM, int MIN_CTAS, int NT = VM_NT, bool WARP_FETCH = false>
__global__ void __launch_bounds__(NT, MIN_CTAS) vm_kernel(const __grid_constant__ VmParams p) {
  extern __shared__ uint4 vm_smem[];
  __shared__ uint32_t s_tmem;
  __shared__ int s_blk;
  DevEnv<USE_TMEM, NT> env;
  env.smem_cells = p.smem_cells;
  env.tm_base = 0;
  if (USE_TMEM) {
    // warp 0 allocates this CTA's columns; co-resident CTAs share the SM's 512
    if (threadIdx.x < 32) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_tmem);
      if (p.tmem_cols == 128)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(dst));
      else if (p.tmem_cols == 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(dst));
      else
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // lane quadrant of this warp (warp % 4) in the upper half-word, column base of its group below
    env.tm_base = s_tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) +
                  (uint32_t)(threadIdx.x >> 7) * (uint32_t)p.tmem_group_cols;
  }
  env.sm = (uint32_t)__cvta_generic_to_shared(vm_smem) + threadIdx.x * 16;
  env.p = &p;
  env.total = (long long)gridDim.x * NT;
  const long long gtid = (long long)blockIdx.x * NT + threadIdx.x;
  env.cold = p.cold + gtid;
  env.flags = 0;
  const long long last = p.n_items > 0 ? p.n_items - 1 : 0;
  env.item_raw = gtid;
  env.item = gtid < last ? gtid : last;
  // prologue once; then item blocks of NT items are fetched from a global counter until the
  // batch is exhausted (CTAs that find no work left exit early, so the CTAs of the next launch
  // on another stream can move in: no tail-wave quantisation across back-to-back batches);
  // epilogue once.  One copy of the interpreter loop serves all three sections.
  // segmented mode: this thread's segment and the longest segment of the CTA
  const bool seg_mode = p.seg_start != nullptr;
  unsigned seg_s0 = 0, seg_len = 0, seg_iters = 0, seg_k = 0;
  if (seg_mode) {
    __shared__ unsigned s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (gtid < p.n_items) {
      seg_s0 = __ldg(p.seg_start + gtid);
      seg_len = __ldg(p.seg_start + gtid + 1) - seg_s0;
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, seg_len);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    seg_iters = s_max;
  }
  // programs without cross-thread reads: item blocks of 32 per WARP, by at most active_warps warps per CTA
  // (no block-wide barrier in the item loop; see launch_program for the balanced-waves policy)
  constexpr bool warp_mode = WARP_FETCH;
  const bool warp_idle = warp_mode && (int)(threadIdx.x >> 5) >= p.active_warps;
  for (int phase = warp_idle ? 3 : 0; phase < 3;) {
    int lo, hi;
    if (phase == 0) {
      lo = 0;
      hi = p.body_start;
      phase = 1;
    } else if (phase == 1 && seg_mode) {
      // body iteration k handles the k-th record of this thread's segment
      if (seg_k >= seg_iters) {
        phase = 2;
        env.item_raw = gtid;
        env.item = gtid < last ? gtid : last;
        continue;
      }
      const bool act = seg_k < seg_len;
      env.item = act ? (long long)__ldg(p.seg_idx + seg_s0 + seg_k) : 0;
      env.item_raw = act ? 0 : p.n_items;  // FACTIVE reads item_raw < n_items
      seg_k++;
      lo = p.body_start;
      hi = p.epi_start;
    } else if (phase == 1 && warp_mode) {
      int blk = 0;
      if ((threadIdx.x & 31) == 0) blk = atomicAdd(p.counter, 1);
      blk = __shfl_sync(0xffffffffu, blk, 0);
      if (blk >= p.n_blocks) {
        phase = 2;
        continue;
      }
      lo = p.body_start;
      hi = p.epi_start;
      env.item_raw = (long long)blk * 32 + (threadIdx.x & 31);
      env.item = env.item_raw < last ? env.item_raw : last;
    } else if (phase == 1) {
      // (tried: item blocks of 128 fetched per group of four warps through named barriers -- worse:
      // a wide CTA then stays resident until its slowest group is done and the next launch's CTA
      // cannot move in, 1.36 vs 1.48 M pairings/s)
      __syncthreads();
      if (threadIdx.x == 0) s_blk = atomicAdd(p.counter, 1);
      __syncthreads();
      const long long blk = s_blk;
      if (blk >= p.n_blocks) {
        phase = 2;
        continue;
      }
      lo = p.body_start;
      hi = p.epi_start;
      env.item_raw = blk * NT + threadIdx.x;
      env.item = env.item_raw < last ? env.item_raw : last;
    } else {
      lo = p.epi_start;
      hi = p.n_ins;
      phase = 3;
      // the epilogue addresses the thread's own record (per-thread partials of g?_sumf), as in segmented mode
      env.item_raw = gtid;
      env.item = gtid < last ? gtid : last;
    }
    vm_run_section(env, p.code, lo, hi);
  }
  if (USE_TMEM) {
    tm_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x < 32) {
      if (p.tmem_cols == 128)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(s_tmem));
      else if (p.tmem_cols == 256)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(s_tmem));
      else
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem));
    }
  }
}

}  // namespace b200bls
