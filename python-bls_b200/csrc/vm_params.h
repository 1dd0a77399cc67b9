// vm_params.h -- launch parameters shared by the host side and both interpreter kernels
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace b200bls {

constexpr int VM_NT = 128;      // threads per CTA (default shapes)
constexpr int VM_NT_WIDE = 384; // "wide" shape: ONE CTA of 12 warps per SM.  Three 128-thread CTAs can only
                                // allocate 128 TMEM columns each (power-of-two allocations, 512 per SM); one CTA
                                // owns all 512 and gives each group of four warps 168 columns = 7 Fq2 slots
constexpr int VM_NT_XWIDE = 512; // shape 5: ONE CTA of 16 warps per SM (128 registers per thread), 128 TMEM columns per group of four
                                 // warps = 5 Fq2 slots, 4 slots of shared memory
constexpr int VM_MAX_BUFS = 8;

struct VmBuf {
  unsigned char* ptr;
  long long stride;  // bytes per item (byte buffers) or item capacity (raw SoA buffers)
};

struct VmParams {
  const uint2* code;       // instructions, padded with one trailing NOP
  int body_start, epi_start, n_ins;
  const uint4* consts;     // Montgomery-form constants, 3 x uint4 each
  uint4* cold;             // [n_cold * 6][total threads]
  long long n_items;
  long long n_blocks;      // ceil(n_items / CTA threads): item blocks handed out dynamically
  int* counter;            // zeroed before the launch; next item block to process
  int smem_cells;          // cells [0, smem_cells) live in shared memory, the rest in Tensor Memory
  int tmem_cols;           // TMEM columns to allocate per CTA (0, 128, 256 or 512)
  int tmem_group_cols;     // columns owned by each group of four warps (CTAs wider than 128 threads)
  // Segmented mode (multi-scalar multiplication buckets): thread t owns segment t of `n_items`
  // segments; body iteration k processes record seg_idx[seg_start[t] + k] of the indexed buffers
  // (inactive once k reaches the segment length); prologue / epilogue address record t.
  const unsigned* seg_start;  // n_items + 1 offsets into seg_idx, or nullptr (normal mode)
  const unsigned* seg_idx;
  // paired kernel (vm_kernel2.cuh): item blocks of 16 handed out per warp (programs without
  // cross-thread reads), by at most active_warps warps of every CTA
  int warp_fetch;
  int active_warps;
  VmBuf bufs[VM_MAX_BUFS];
};

constexpr int VM2_NT_WIDE = 384;    // paired kernel, throughput shape: 2 CTAs of 384 threads (192 items) per SM
constexpr int VM2_NT = 256;         // paired kernel, narrow shapes: 128 items per CTA, 1-3 CTAs per SM (block reductions)
constexpr int VM2_ITEMS_PER_WARP = 16;

}  // namespace b200bls
