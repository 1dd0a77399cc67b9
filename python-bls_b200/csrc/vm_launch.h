// vm_launch.h -- host entry points of the two interpreter kernels.  Each kernel family is its own translation unit
// (kernel1.cu: one thread per item; kernel2.cu: two threads per item), so that their register budgets (148 / 80)
// never meet in one ptxas run: out-of-line device functions shared by both families would be allocated once, for
// the tighter budget, and that changed the code ptxas generated for the throughput kernel (measured: +5 %
// executed instructions, 33.1 -> 34.4 ms per wave of pairings).
#pragma once
#include <cuda_runtime.h>

#include "vm_params.h"

namespace b200bls {
cudaError_t vm1_configure();
void vm1_launch(bool use_tmem, bool wide, int grid, size_t smem, cudaStream_t stream, const VmParams& p);
cudaError_t vm3_configure();   // kernel3.cu: the one-thread kernel at 512 threads per CTA (its own 128-register budget)
void vm3_launch(int grid, size_t smem, cudaStream_t stream, const VmParams& p);
cudaError_t vm2_configure();
void vm2_launch(bool use_tmem, bool wide, bool seg, int grid, size_t smem, cudaStream_t stream, const VmParams& p);
}  // namespace b200bls
