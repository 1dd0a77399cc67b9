// vm_kernel2.cuh -- the sm_100a kernel that interprets b200-bls field programs with TWO threads per
// batch item (vm_exec2.cuh): thread pair (2i, 2i + 1) of a warp owns item i of the warp's block of
// 16 items, thread `role` owning coefficient `role` of every Fq2 slot.
//
// Per-item resources are those of the one-thread-per-item kernel (vm_kernel.cuh, the throughput
// kernel): `n_slots` Fq2 slots in shared memory, laid out
// [slot][16-byte chunk 0..2][thread] so that a warp's LDS.128 / STS.128 covers 512 contiguous bytes
// (own column) or the same 512 bytes permuted within pairs (partner's column): conflict free;
// `n_tmem` slots in Tensor Memory (12 columns per slot and thread, lanes private); spills to the
// L2-backed cold area [cold slot][chunk][global thread].  What doubles is the number of warps (2 CTAs x
// 384 threads = 24 warps per SM on <= 80 registers, against 12 warps on 146) and what halves is the
// work per thread.  Measured: 0.81x the throughput of the one-thread kernel --
// decode, dispatch and operand addressing are paid per thread -- but 0.70x the latency of an isolated
// pass, so the launcher sends small isolated batches here.
//
// Work distribution: programs without cross-thread reads hand out blocks of 16 items PER WARP from a
// global counter -- no block-wide barrier in the item loop, a warp that finds no work left is done,
// and the launcher can bound the number of active warps per CTA so that a batch of 1.15 waves runs as
// two equal waves at lower occupancy instead of a full one and a 15 % tail.  Programs with block
// reductions (SYNC / XMOV2) keep the CTA-wide item blocks of vm_kernel.cuh.
#pragma once
#include <cuda_runtime.h>

#include "vm_exec2.cuh"
#include "vm_kernel.cuh"

namespace b200bls {


__device__ __forceinline__ void shfl_fp(fp& x) {
#pragma unroll
  for (int i = 0; i < NL; i++) x.v[i] = __shfl_xor_sync(0xffffffffu, x.v[i], 1);
}

// The streaming two-row product (fp.cuh: fp_mul2_rounds4): a and b stay in registers, the other two
// operands -- the two coefficients of one workspace slot -- arrive one 16-byte chunk per four
// rounds: from the own and the partner's shared-memory column, or, for a Tensor-Memory slot, from
// the own lane plus a shuffle.  One copy in the kernel (the hot code has to fit the instruction
// cache); operands and result travel in registers.
// `src` is the shared-space address of the x operand's column (the y operand is the partner column,
// src ^ 16), or, with bit 0 set, the Tensor-Memory address of the slot shifted left by one and bit 1 =
// swap (x is the partner's coefficient).  A noinline call keeps the caller's live state in registers
// across it only if callee + caller fit the 80-register budget, hence the packed arguments.
template <bool USE_TMEM, int NT>
__device__ __forceinline__ fp fp_mul2_stream(const fp& a, const fp& b, uint32_t src) {
  Mul2State s;
#pragma unroll
  for (int i = 0; i < NL; i++) s.ev[i] = s.od[i] = 0;
  if (USE_TMEM && (src & 1u)) tm_wait_st();
  // three passes of four rounds: a loop, not 12 unrolled rounds -- the hot code has to fit the instruction cache
#pragma unroll 1
  for (int k = 0; k < 3; k++) {
    uint32_t x4[4], y4[4];
    if (!USE_TMEM || !(src & 1u)) {
      const uint4 xv = lds128(src);
      const uint4 yv = lds128(src ^ 16u);
      x4[0] = xv.x, x4[1] = xv.y, x4[2] = xv.z, x4[3] = xv.w;
      y4[0] = yv.x, y4[1] = yv.y, y4[2] = yv.z, y4[3] = yv.w;
      src += NT * 16;
    } else {
      uint32_t o4[4];
      tm_ld4(src >> 2, o4);
      tm_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t p = __shfl_xor_sync(0xffffffffu, o4[i], 1);
        x4[i] = (src & 2u) ? p : o4[i];
        y4[i] = (src & 2u) ? o4[i] : p;
      }
      src += 4u << 2;
    }
    fp_mul2_rounds4(s, a, b, x4, y4);
  }
  fp r;
  mont_finish(r, s.ev, s.od);
  return r;
}

template <bool USE_TMEM, int NT, bool SEG>
struct PairEnv {
  static constexpr uint32_t CHUNK = NT * 16;       // bytes between the chunks of one coefficient
  static constexpr uint32_t SLOT = 3 * NT * 16;    // bytes per slot row
  // live across every instruction, so kept small: what can be re-derived from the (constant-bank) kernel
  // parameters or from one of the other fields is not stored
  uint32_t sm_own;          // shared-space byte address of this thread's column; the partner's is sm_own ^ 16
  uint32_t tm_base;         // TMEM address of column 0 of this thread's strip
  const VmParams* p;
  int item_raw;             // batches are bounded by 2^31 - 1 items (checked by the launcher)
  int item;                 // segmented mode only: the record this iteration reads
  // the record this thread's loads address: inactive lanes read the last one
  __device__ __forceinline__ int item_idx() const {
    if (SEG) return item;
    const int last = (int)p->n_items - 1;
    return item_raw < last ? item_raw : (last < 0 ? 0 : last);
  }
  uint32_t flags;

  __device__ __forceinline__ int role() const { return (int)((sm_own >> 4) & 1u); }   // = threadIdx.x & 1
  __device__ __forceinline__ uint32_t sm_oth() const { return sm_own ^ 16u; }
  __device__ __forceinline__ int smem_slots() const { return p->smem_cells >> 1; }
  __device__ __forceinline__ static void unpack(fp& x, int k, uint4 v) {
    x.v[4 * k] = v.x;
    x.v[4 * k + 1] = v.y;
    x.v[4 * k + 2] = v.z;
    x.v[4 * k + 3] = v.w;
  }
  __device__ __forceinline__ static uint4 pack(const fp& x, int k) {
    return make_uint4(x.v[4 * k], x.v[4 * k + 1], x.v[4 * k + 2], x.v[4 * k + 3]);
  }
  __device__ __forceinline__ bool in_smem(int slot) const { return !USE_TMEM || slot < smem_slots(); }
  __device__ __forceinline__ uint32_t tm_addr(int slot) const { return tm_base + (uint32_t)(slot - smem_slots()) * 12; }
  __device__ __forceinline__ void ld_smem(uint32_t col, int slot, fp& x) {
    const uint32_t q = col + slot * SLOT;
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x, k, lds128(q + k * CHUNK));
  }
  __device__ __forceinline__ void ld_tm(int slot, fp& x) {
    const uint32_t t = tm_addr(slot);
    tm_wait_st();
    tm_ld8(t, x.v);
    tm_ld4(t + 8, x.v + 8);
    tm_wait_ld();
  }
  // the slot index comes from the instruction word: shared / tensor memory is a warp-uniform choice
  __device__ __forceinline__ void ld_own(int slot, fp& x) {
    if (in_smem(slot))
      ld_smem(sm_own, slot, x);
    else
      ld_tm(slot, x);
  }
  __device__ __forceinline__ void ld_oth(int slot, fp& x) {
    if (in_smem(slot)) {
      ld_smem(sm_oth(), slot, x);
    } else {
      ld_tm(slot, x);
      shfl_fp(x);
    }
  }
  // fp cell c = 2 * slot + half, whichever thread owns it
  __device__ __forceinline__ void ld_cell(int c, fp& x) {
    const int slot = c >> 1;
    const bool mine = (c & 1) == role();
    if (in_smem(slot)) {
      ld_smem(mine ? sm_own : sm_oth(), slot, x);
    } else {
      fp o;
      ld_tm(slot, o);
      x = o;
      shfl_fp(o);
      if (!mine) x = o;
    }
  }
  __device__ __forceinline__ void st_own(int slot, const fp& x) {
    if (in_smem(slot)) {
      const uint32_t q = sm_own + slot * SLOT;
#pragma unroll
      for (int k = 0; k < 3; k++) sts128(q + k * CHUNK, pack(x, k));
      __syncwarp();
    } else {
      const uint32_t t = tm_addr(slot);
      tm_st8(t, x.v);
      tm_st4(t + 8, x.v + 8);
    }
  }
  // only the owner of cell c stores; tcgen05.st is warp-collective, so the partner writes back what
  // its lane holds
  __device__ __forceinline__ void st_cell(int c, const fp& x) {
    const int slot = c >> 1;
    const bool mine = (c & 1) == role();
    if (in_smem(slot)) {
      if (mine) {
        const uint32_t q = sm_own + slot * SLOT;
#pragma unroll
        for (int k = 0; k < 3; k++) sts128(q + k * CHUNK, pack(x, k));
      }
      __syncwarp();
    } else {
      fp cur;
      ld_tm(slot, cur);
      const uint32_t t = tm_addr(slot);
#pragma unroll
      for (int i = 0; i < NL; i++) cur.v[i] = mine ? x.v[i] : cur.v[i];
      tm_st8(t, cur.v);
      tm_st4(t + 8, cur.v + 8);
    }
  }
  __device__ __forceinline__ void xchg(fp& x) { shfl_fp(x); }
  __device__ __forceinline__ void mul2(fp& r, const fp& a, const fp& b, int slot, bool swap) {
    // TMEM addresses are (lane << 16) | column with lane < 128 and column < 512: two spare bits on top
    if (in_smem(slot))
      r = fp_mul2_stream<USE_TMEM, NT>(a, b, (swap ? sm_oth() : sm_own) + slot * SLOT);
    else
      r = fp_mul2_stream<USE_TMEM, NT>(a, b, (tm_addr(slot) << 2) | (swap ? 3u : 1u));
  }
  // thread (tid + 2 * off) mod NT: the same coefficient of the item `off` places further in the CTA
  __device__ __forceinline__ void ld_lane_own(int slot, int off, fp& x) {
    const int t = ((int)threadIdx.x + 2 * off) % NT;
    ld_smem(sm_own + (t - (int)threadIdx.x) * 16, slot, x);
  }
  __device__ __forceinline__ void ldc(int idx, fp& x) {
    const uint4* q = p->consts + idx * 3;
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x, k, __ldg(q + k));
  }
  __device__ __forceinline__ void set_flag(int f, bool v) { flags = (flags & ~(1u << f)) | ((v ? 1u : 0u) << f); }
  __device__ __forceinline__ bool get_flag(int f) { return (flags >> f) & 1u; }
  __device__ __forceinline__ bool any_flag(int f) { return __any_sync(0xffffffffu, (flags >> f) & 1u); }
  __device__ __forceinline__ bool active() { return item_raw < (int)p->n_items; }
  __device__ __forceinline__ uint32_t ld_byte(int buf, int off) {
    return p->bufs[buf].ptr[(long long)item_idx() * p->bufs[buf].stride + off];
  }
  // block_only: the first pair of the CTA writes the record of item = CTA index (reduction results)
  __device__ __forceinline__ void st_byte(int buf, int off, unsigned char v, bool block_only) {
    if (role() != 0) return;
    if (block_only) {
      if (threadIdx.x == 0) p->bufs[buf].ptr[(long long)blockIdx.x * p->bufs[buf].stride + off] = v;
    } else if (active()) {
      p->bufs[buf].ptr[(long long)item_idx() * p->bufs[buf].stride + off] = v;
    }
  }
  __device__ __forceinline__ void ld_be(int buf, int off, int nwords, fp& x) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p->bufs[buf].ptr + (long long)item_idx() * p->bufs[buf].stride + off);
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
  // both threads hold the value; the thread whose role equals `parity` (the cell's owner) stores
  __device__ __forceinline__ void st_be48(int buf, int off, const fp& x, bool block_only, int parity) {
    if (role() != parity) return;
    int it = item_idx();
    if (block_only) {
      if (threadIdx.x >= 2) return;
      it = blockIdx.x;
    } else if (!active()) {
      return;
    }
    uint32_t* w = reinterpret_cast<uint32_t*>(p->bufs[buf].ptr + (long long)it * p->bufs[buf].stride + off);
#pragma unroll
    for (int i = 0; i < NL; i++) w[11 - i] = __byte_perm(x.v[i], 0, 0x0123);
  }
  // raw SoA buffers [element][chunk 0..5][item]: chunks 3 * role .. 3 * role + 2 are this thread's
  __device__ __forceinline__ void ld_raw_own(int buf, int elem, fp& x) {
    const long long s = p->bufs[buf].stride;
    const uint4* q = reinterpret_cast<const uint4*>(p->bufs[buf].ptr) + ((long long)elem * 6 + 3 * role()) * s + item_idx();
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x, k, q[k * s]);
  }
  __device__ __forceinline__ void st_raw_own(int buf, int elem, const fp& x, bool block_only) {
    int it;
    if (block_only) {
      if (threadIdx.x >= 2) return;
      it = blockIdx.x;
    } else {
      if (!active()) return;
      it = item_raw;
    }
    const long long s = p->bufs[buf].stride;
    uint4* q = reinterpret_cast<uint4*>(p->bufs[buf].ptr) + ((long long)elem * 6 + 3 * role()) * s + it;
#pragma unroll
    for (int k = 0; k < 3; k++) q[k * s] = pack(x, k);
  }
  // cold area [cold slot][chunk 0..2][global thread]
  __device__ __forceinline__ uint4* cold_ptr(int g, int& total) const {
    total = (int)(gridDim.x * NT);
    return p->cold + (long long)g * 3 * total + (blockIdx.x * NT + threadIdx.x);
  }
  __device__ __forceinline__ void st_cold_own(int g, const fp& x) {
    int total;
    uint4* q = cold_ptr(g, total);
#pragma unroll
    for (int k = 0; k < 3; k++) __stcg(q + (long long)k * total, pack(x, k));
  }
  __device__ __forceinline__ void ld_cold_own(int g, fp& x) {
    int total;
    const uint4* q = cold_ptr(g, total);
#pragma unroll
    for (int k = 0; k < 3; k++) unpack(x, k, __ldcg(q + (long long)k * total));
  }
  // the cold copy in slot g is dead: drop its lines from the L2 instead of writing them back (vm_kernel.cuh)
  __device__ __forceinline__ void discard_cold_own(int g) {
    if ((threadIdx.x & 7) == 0) {
      int total;
      const uint4* q = cold_ptr(g, total);
#pragma unroll
      for (int k = 0; k < 3; k++) asm volatile("discard.global.L2 [%0], 128;" ::"l"(q + (long long)k * total) : "memory");
    }
  }
  __device__ __forceinline__ void sync() { __syncthreads(); }
};

// Runs one program section, from `lo` to its END instruction.  The next instruction word is loaded
// before the current one executes and consumed (through a volatile move) after it; one L1 line holds
// 16 instructions and the line after next is prefetched at every line boundary.
// USE_TMEM = false: no tcgen05 instruction in the kernel at all (a kernel that merely contains
// tcgen05.alloc holds the SM's allocation permit until it relinquishes it, vm_kernel.cuh).
// SEG: segmented launch mode (multi-scalar multiplication buckets, see VmParams); its per-thread
// segment cursor would otherwise cost every program four registers
template <bool USE_TMEM, int NT, int MIN_CTAS, bool SEG = false>
__global__ void __launch_bounds__(NT, MIN_CTAS) vm2_kernel(const __grid_constant__ VmParams p) {
  extern __shared__ __align__(128) uint4 vm_smem[];
  __shared__ uint32_t s_tmem;
  __shared__ int s_blk;
  constexpr int ITEMS = NT / 2;
  PairEnv<USE_TMEM, NT, SEG> env;
  env.tm_base = 0;
  if (USE_TMEM) {
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
    // lane quadrant of this warp (warp % 4) in the upper half-word; warps of the same quadrant (warp / 4 =
    // group) are told apart by column
    env.tm_base = s_tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) +
                  (uint32_t)(threadIdx.x >> 7) * (uint32_t)p.tmem_group_cols;
  }
  const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(vm_smem);
  env.sm_own = sm0 + threadIdx.x * 16;      // 128-byte aligned base: bit 4 is the thread's role
  env.p = &p;
  const int gtid = (int)(blockIdx.x * NT + threadIdx.x);
  const int gitem = gtid >> 1;
  const int n_items = (int)p.n_items;
  env.flags = 0;
  const int last = n_items > 0 ? n_items - 1 : 0;
  env.item_raw = gitem;
  env.item = gitem < last ? gitem : last;   // (dead unless SEG)
  constexpr bool seg_mode = SEG;
  unsigned seg_s0 = 0, seg_len = 0, seg_iters = 0, seg_k = 0;
  if (seg_mode) {
    __shared__ unsigned s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (gitem < n_items) {
      seg_s0 = __ldg(p.seg_start + gitem);
      seg_len = __ldg(p.seg_start + gitem + 1) - seg_s0;
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, seg_len);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    seg_iters = s_max;
  }
  // per-warp item blocks: warps beyond the launcher's bound do not take part in the item loop
  const bool warp_mode = p.warp_fetch != 0;
  const bool warp_idle = warp_mode && (int)(threadIdx.x >> 5) >= p.active_warps;
  // One interpreter loop serves the three sections.  Which section an END instruction closes is read
  // off the program counter, so no section state lives in a register across the instruction bodies:
  // after the prologue and after every body pass the next item block is fetched (body again, or on to
  // the epilogue when the batch is exhausted); the END of the epilogue leaves the loop.
  const uint2* pc = p.code;
  while (!warp_idle) {
#ifdef B200BLS_VM2_PREFETCH_INS
    if (((uint32_t)(uintptr_t)pc & 127u) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(pc + 32));
#endif
    const uint2 ins = __ldg(pc);
    const int skip = vm_exec2(env, ins.x, ins.y);
    if (skip >= 0) {
      pc += 1 + skip;
      continue;
    }
    if (pc - p.code >= p.epi_start) break;  // END of the epilogue
    bool more;
    if (seg_mode) {
      more = seg_k < seg_iters;
      if (more) {
        const bool act = seg_k < seg_len;
        env.item = act ? (int)__ldg(p.seg_idx + seg_s0 + seg_k) : 0;
        env.item_raw = act ? 0 : n_items;  // FACTIVE reads item_raw < n_items
        seg_k++;
      } else {
        env.item_raw = gitem;
        env.item = gitem < last ? gitem : last;
      }
    } else if (warp_mode) {
      int blk = 0;
      if ((threadIdx.x & 31) == 0) blk = atomicAdd(p.counter, 1);
      blk = __shfl_sync(0xffffffffu, blk, 0);
      more = blk < p.n_blocks;
      if (more) env.item_raw = blk * VM2_ITEMS_PER_WARP + (int)((threadIdx.x & 31) >> 1);
    } else {
      __syncthreads();
      if (threadIdx.x == 0) s_blk = atomicAdd(p.counter, 1);
      __syncthreads();
      const int blk = s_blk;
      more = blk < p.n_blocks;
      if (more) env.item_raw = blk * ITEMS + (int)(threadIdx.x >> 1);
    }
    pc = p.code + (more ? p.body_start : p.epi_start);
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
