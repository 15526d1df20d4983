// cuda_ctx.cuh -- device-side helpers shared by the execution contexts of the stand-alone stage kernels (launch_impl.cuh:
// CudaCtx) and of the fused two-pass kernel (fused.cuh: FusedCtx).  CUDA only.
#pragma once

namespace accfft_b200 {

// G threads of one pencil: inside one warp -> __syncwarp; whole warps and at most 15 pencils per CTA -> one named barrier
// per pencil (ids 1..15; 0 is __syncthreads); anything else -> the CTA barrier
// (the barrier id is an immediate: with the id in a register ptxas has to reserve all 16 barriers for the kernel)
template <int G, int PC, int I>
__device__ __forceinline__ void cuda_named_barrier(int pcl) {
  if constexpr (I + 1 >= PC) {
    asm volatile("bar.sync %0, %1;" ::"n"(I + 1), "n"(G) : "memory");
  } else {
    if (pcl == I) asm volatile("bar.sync %0, %1;" ::"n"(I + 1), "n"(G) : "memory");
    else cuda_named_barrier<G, PC, I + 1>(pcl);
  }
}
template <int G, int PC>
__device__ __forceinline__ void cuda_sync_group(int pcl) {
  if constexpr (G <= 32 && 32 % G == 0) {
    __syncwarp();
  } else if constexpr (PC == 1) {
    __syncthreads();
  } else if constexpr (G % 32 == 0 && PC <= 15) {
    cuda_named_barrier<G, PC, 0>(pcl);
  } else {
    __syncthreads();
  }
}

}  // namespace accfft_b200
