// kernel1.cu -- the one-thread-per-item interpreter kernel (vm_kernel.cuh) as its own translation unit
#include "vm_kernel.cuh"
#include "vm_launch.h"

namespace b200bls {

cudaError_t vm1_configure() {
  cudaError_t e = cudaFuncSetAttribute(vm_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(vm_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(vm_kernel<true, 1, VM_NT_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(vm_kernel<false, 3, VM_NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(vm_kernel<true, 3, VM_NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(vm_kernel<true, 1, VM_NT_WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  return e;
}

void vm1_launch(bool use_tmem, bool wide, int grid, size_t smem, cudaStream_t stream, const VmParams& p) {
  if (p.warp_fetch) {   // isolated batches: item blocks per warp
    if (wide)
      vm_kernel<true, 1, VM_NT_WIDE, true><<<grid, VM_NT_WIDE, smem, stream>>>(p);
    else if (use_tmem)
      vm_kernel<true, 3, VM_NT, true><<<grid, VM_NT, smem, stream>>>(p);
    else
      vm_kernel<false, 3, VM_NT, true><<<grid, VM_NT, smem, stream>>>(p);
    return;
  }
  if (wide)
    vm_kernel<true, 1, VM_NT_WIDE><<<grid, VM_NT_WIDE, smem, stream>>>(p);
  else if (use_tmem)
    vm_kernel<true, 3><<<grid, VM_NT, smem, stream>>>(p);
  else
    vm_kernel<false, 3><<<grid, VM_NT, smem, stream>>>(p);
}

}  // namespace b200bls
