// cuda_ctx.cuh -- device-side helpers shared by the execution contexts of the stand-alone stage kernels (launch_impl.cuh:
// CudaCtx) and of the fused two-pass kernel (fused.cuh: FusedCtx).  CUDA only.
#pragma once

namespace accfft_b200 {

// G threads of one pencil: inside one warp -> __syncwarp; whole warps and at most 15 pencils per CTA -> one named barrier
// per pencil (ids 1..15; 0 is __syncthreads); anything else -> the CTA barrier
template <int G, int PC>
__device__ __forceinline__ void cuda_sync_group(int pcl) {
  if constexpr (G <= 32 && 32 % G == 0) {
    __syncwarp();
  } else if constexpr (G % 32 == 0 && PC <= 15) {
    asm volatile("bar.sync %0, %1;" ::"r"(pcl + 1), "n"(G) : "memory");
  } else {
    __syncthreads();
  }
}

}  // namespace accfft_b200
