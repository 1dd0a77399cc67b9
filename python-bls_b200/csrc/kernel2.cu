// kernel2.cu -- the two-threads-per-item interpreter kernel (vm_kernel2.cuh) as its own translation unit
#include "vm_kernel2.cuh"
#include "vm_launch.h"

namespace b200bls {

cudaError_t vm2_configure() {
  cudaError_t e = cudaSuccess;
#define B200BLS_ATTR(K, BYTES) \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES)
  B200BLS_ATTR((vm2_kernel<false, VM2_NT, 3>), 226 * 1024);
  B200BLS_ATTR((vm2_kernel<true, VM2_NT, 3>), 226 * 1024);
  B200BLS_ATTR((vm2_kernel<true, VM2_NT_WIDE, 2>), 113 * 1024);
  B200BLS_ATTR((vm2_kernel<false, VM2_NT, 3, true>), 226 * 1024);
  B200BLS_ATTR((vm2_kernel<true, VM2_NT, 3, true>), 226 * 1024);
  B200BLS_ATTR((vm2_kernel<true, VM2_NT_WIDE, 2, true>), 113 * 1024);
#undef B200BLS_ATTR
  return e;
}

void vm2_launch(bool use_tmem, bool wide, bool seg, int grid, size_t smem, cudaStream_t stream, const VmParams& p) {
  if (wide) {
    if (seg)
      vm2_kernel<true, VM2_NT_WIDE, 2, true><<<grid, VM2_NT_WIDE, smem, stream>>>(p);
    else
      vm2_kernel<true, VM2_NT_WIDE, 2><<<grid, VM2_NT_WIDE, smem, stream>>>(p);
  } else if (use_tmem) {
    if (seg)
      vm2_kernel<true, VM2_NT, 3, true><<<grid, VM2_NT, smem, stream>>>(p);
    else
      vm2_kernel<true, VM2_NT, 3><<<grid, VM2_NT, smem, stream>>>(p);
  } else {
    if (seg)
      vm2_kernel<false, VM2_NT, 3, true><<<grid, VM2_NT, smem, stream>>>(p);
    else
      vm2_kernel<false, VM2_NT, 3><<<grid, VM2_NT, smem, stream>>>(p);
  }
}

}  // namespace b200bls
