// kernel3.cu -- the one-thread-per-item interpreter kernel at 512 threads per CTA (shape 5), its own translation unit:
// 16 warps per SM leave 128 registers per thread, and the out-of-line field functions must be allocated for that
// budget here without touching the 168-register budget of kernel1.cu (see vm_launch.h).
#include "vm_kernel.cuh"
#include "vm_launch.h"

namespace b200bls {

cudaError_t vm3_configure() {
  cudaError_t e =
      cudaFuncSetAttribute(vm_kernel<true, 1, VM_NT_XWIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(vm_kernel<true, 1, VM_NT_XWIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             226 * 1024);
  return e;
}

void vm3_launch(int grid, size_t smem, cudaStream_t stream, const VmParams& p) {
  if (p.warp_fetch)
    vm_kernel<true, 1, VM_NT_XWIDE, true><<<grid, VM_NT_XWIDE, smem, stream>>>(p);
  else
    vm_kernel<true, 1, VM_NT_XWIDE><<<grid, VM_NT_XWIDE, smem, stream>>>(p);
}

}  // namespace b200bls
