// launch_impl.cuh -- __global__ wrappers around the stage bodies (stage.cuh) and the per-schedule
// dispatch.  Included by the k_<kind>_<prec>.cu translation units, each of which instantiates ONE
// (kind, precision) pair for every line of schedules.def so that the kernels compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include "kernel_cfg.hpp"
#include "plan.hpp"

namespace accfft_b200 {

struct CudaCtx {
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ long long bid() const { return (long long)blockIdx.x; }
  __device__ __forceinline__ void* smem() const {
    extern __shared__ __align__(16) unsigned char accfft_smem_raw[];
    return accfft_smem_raw;
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

template <class T, class S>
__global__ void __launch_bounds__(Cfg<T, S>::STRIDED_THREADS) k_strided_c2c(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  strided_c2c_body<T, S, Cfg<T, S>::TC>(d, ctx);
}
template <class T, class S>
__global__ void __launch_bounds__(Cfg<T, S>::CONTIG_THREADS) k_contig_c2c(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  contig_c2c_body<T, S, Cfg<T, S>::PC>(d, ctx);
}
template <class T, class S>
__global__ void __launch_bounds__(Cfg<T, S>::CONTIG_THREADS) k_contig_r2c(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  contig_r2c_body<T, S, Cfg<T, S>::PC>(d, ctx);
}
template <class T, class S>
__global__ void __launch_bounds__(Cfg<T, S>::CONTIG_THREADS) k_contig_c2r(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  contig_c2r_body<T, S, Cfg<T, S>::PC>(d, ctx);
}

template <class K>
static int launch_kernel(K kernel, long long nblocks, int threads, int smem, const StageDesc& d, cudaStream_t st,
                         bool dry) {
  if (dry) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);   // forces the (lazy) load of the kernel
    if (e == cudaSuccess && smem > 48 * 1024)
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return (int)e;
  }
  if (nblocks <= 0) return 0;
  if (nblocks > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
  }
  kernel<<<(unsigned)nblocks, threads, smem, st>>>(d);
  return (int)cudaGetLastError();
}

template <class T, class S>
static int launch_one(const Launch& l, cudaStream_t st) {
  using C = Cfg<T, S>;
  const StageDesc& d = l.desc;
#if ACCFFT_KIND == 0
  const long long nb = (long long)d.nA * ((d.nB + d.colShift + C::TC - 1) / C::TC);
  return launch_kernel(k_strided_c2c<T, S>, nb, C::STRIDED_THREADS, C::STRIDED_SMEM, d, st, l.dry);
#elif ACCFFT_KIND == 1
  const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
  return launch_kernel(k_contig_c2c<T, S>, nb, C::CONTIG_THREADS, C::CONTIG_SMEM, d, st, l.dry);
#elif ACCFFT_KIND == 2
  const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
  return launch_kernel(k_contig_r2c<T, S>, nb, C::CONTIG_THREADS, C::CONTIG_SMEM, d, st, l.dry);
#else
  const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
  return launch_kernel(k_contig_c2r<T, S>, nb, C::CONTIG_THREADS, C::CONTIG_SMEM, d, st, l.dry);
#endif
}

template <class T>
static int launch_any(const Launch& l, cudaStream_t st) {
#define X(N_, G_, R0, R1, R2, R3) \
  if (l.fftN == N_) return launch_one<T, Sched<N_, G_, R0, R1, R2, R3>>(l, st);
#include ACCFFT_SCHEDULES_FILE
#undef X
  return -1;
}

}  // namespace accfft_b200
