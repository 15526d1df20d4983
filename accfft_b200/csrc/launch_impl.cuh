// launch_impl.cuh -- __global__ wrappers around the stage bodies (stage.cuh) and the per-schedule
// dispatch.  Included by the k_<kind>_<prec>.cu translation units, each of which instantiates ONE
// (kind, precision) pair for every line of schedules.def so that the kernels compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include "kernel_cfg.hpp"
#include "launch.hpp"
#include "cuda_ctx.cuh"
#if ACCFFT_KIND == 4
#include "stage_tma.cuh"
#endif

namespace accfft_b200 {

struct CudaCtx {
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ long long bid() const { return (long long)blockIdx.x; }
  __device__ __forceinline__ void* smem() const {
    extern __shared__ __align__(16) unsigned char accfft_smem_raw[];
    return accfft_smem_raw;
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  // barrier among the G threads of pencil `pcl` only (stage.cuh: group_sync)
  static constexpr bool kGroupSync = true;
  template <int G, int PC>
  __device__ __forceinline__ void sync_group(int pcl) const {
    cuda_sync_group<G, PC>(pcl);
  }
  // bulk (TMA) copy shared -> global; the destination may be a peer-mapped buffer (NVLink).  16-byte aligned
  // addresses and size.  Generic-proxy writes to the source must be fenced (bulk_fence) and CTA-synchronised first.
  __device__ __forceinline__ void bulk_fence() const { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) const {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  // the CTA may exit / reuse the staging buffer once the copies have READ their source
  __device__ __forceinline__ void bulk_wait() const { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
};

// resident CTAs the contiguous kernels are compiled for: 1024 threads per SM (<= 64 registers per thread)
constexpr int contig_minb(int threads) { return (1024 / threads) > 0 ? (1024 / threads) : 1; }

// Every stage kernel is instantiated twice: GEN = false for "simple" stages (one local destination, no spectral operator,
// no accumulation: everything a single-GPU transform launches) and GEN = true for the general case (peer segments, staged
// bulk stores, operators, accumulation).  The simple instantiation carries none of that code.
// (min blocks: 1024 resident threads per SM, i.e. <= 64 registers per thread, for both instantiations)
template <class T, class S, bool GEN>
__global__ void __launch_bounds__(Cfg<T, S>::STRIDED_THREADS, (1024 / Cfg<T, S>::STRIDED_THREADS) > 0 ? (1024 / Cfg<T, S>::STRIDED_THREADS) : 1)
    k_strided_c2c(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  strided_c2c_body<T, S, Cfg<T, S>::TC, GEN>(d, ctx);
}
// single precision: two adjacent columns per thread as one 16-byte quantity (stage.cuh: strided_c2c_pair_body)
template <class S, bool GEN>
__global__ void __launch_bounds__(PairCfg<S>::THREADS, PairCfg<S>::MINB)
    k_strided_c2c_pair(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  strided_c2c_pair_body<S, PairCfg<S>::TCP, GEN>(d, ctx);
}
// same body, wide CTA shape (kernel_cfg.hpp: CfgW); instantiated only where it differs from the default shape
template <class T, class S, bool GEN>
__global__ void __launch_bounds__(CfgW<T, S>::STRIDED_THREADS, 1) k_strided_c2c_wide(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  strided_c2c_body<T, S, CfgW<T, S>::TC, GEN>(d, ctx);
}
template <class T, class S, bool GEN>
__global__ void __launch_bounds__(Cfg<T, S>::CONTIG_THREADS, contig_minb(Cfg<T, S>::CONTIG_THREADS)) k_contig_c2c(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  contig_c2c_body<T, S, Cfg<T, S>::PC, GEN>(d, ctx);
}
template <class T, class S, bool GEN>
__global__ void __launch_bounds__(Cfg<T, S>::CONTIG_THREADS, contig_minb(Cfg<T, S>::CONTIG_THREADS)) k_contig_r2c(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  contig_r2c_body<T, S, Cfg<T, S>::PC, GEN>(d, ctx);
}
template <class T, class S, bool GEN>
__global__ void __launch_bounds__(Cfg<T, S>::CONTIG_THREADS, contig_minb(Cfg<T, S>::CONTIG_THREADS)) k_contig_c2r(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  contig_c2r_body<T, S, Cfg<T, S>::PC, GEN>(d, ctx);
}

#ifndef ACCFFT_REAL_MINB
#define ACCFFT_REAL_MINB 2   // two 512-thread CTAs per SM (<= 64 registers): measured 25 % faster than one (profiles/op_bench_r07.txt)
#endif
// real transforms along a strided axis (stage_real.cuh): CTA shape of k_strided_c2c, a tile = 2*TC real columns.
// ACCFFT_REAL_WIDE = 1 gives the lengths / precisions that have a wide shape (kernel_cfg.hpp: CfgW) that shape -- one shape
// per kernel, chosen at compile time.  Measured on one GPU (512^3 single, 1024^3 single, 768^3 double): 3-25 % SLOWER than
// the default shape for all three kernels (they are latency-, not segment-length-bound), hence off.
#ifndef ACCFFT_REAL_WIDE
#define ACCFFT_REAL_WIDE 0
#endif
template <class T, class S>
struct RealCfg {
  static constexpr bool WIDE = ACCFFT_REAL_WIDE != 0 && Cfg<T, S>::HAS_WIDE;
  static constexpr int TC = WIDE ? CfgW<T, S>::TC : Cfg<T, S>::TC;
  static constexpr int THREADS = WIDE ? CfgW<T, S>::STRIDED_THREADS : Cfg<T, S>::STRIDED_THREADS;
  static constexpr int SMEM = WIDE ? CfgW<T, S>::STRIDED_SMEM : Cfg<T, S>::STRIDED_SMEM;
  static constexpr int MINB = WIDE ? 1 : ACCFFT_REAL_MINB;
};
template <class T, class S>
__global__ void __launch_bounds__(RealCfg<T, S>::THREADS, RealCfg<T, S>::MINB) k_strided_r2c(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  strided_r2c_body<T, S, RealCfg<T, S>::TC>(d, ctx);
}
template <class T, class S>
__global__ void __launch_bounds__(RealCfg<T, S>::THREADS, RealCfg<T, S>::MINB) k_strided_c2r(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  strided_c2r_body<T, S, RealCfg<T, S>::TC>(d, ctx);
}
template <class T, class S>
__global__ void __launch_bounds__(RealCfg<T, S>::THREADS, RealCfg<T, S>::MINB) k_strided_deriv(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  strided_deriv_body<T, S, RealCfg<T, S>::TC>(d, ctx);
}

template <class T, class S>
__global__ void __launch_bounds__(Cfg<T, S>::CONTIG_THREADS) k_contig_deriv(const __grid_constant__ StageDesc d) {
  CudaCtx ctx;
  contig_deriv_body<T, S, Cfg<T, S>::PC>(d, ctx);
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

// launch the simple or the general instantiation of a stage kernel (dry runs load and configure both)
template <class KS, class KG>
static int launch_pair(KS simple_kernel, KG general_kernel, bool general, long long nblocks, int threads, int smem_simple,
                       int smem_general, const StageDesc& d, cudaStream_t st, bool dry) {
  if (dry) {
    const int rc = launch_kernel(simple_kernel, nblocks, threads, smem_simple, d, st, true);
    return rc ? rc : launch_kernel(general_kernel, nblocks, threads, smem_general, d, st, true);
  }
  return general ? launch_kernel(general_kernel, nblocks, threads, smem_general, d, st, false)
                 : launch_kernel(simple_kernel, nblocks, threads, smem_simple, d, st, false);
}

#if ACCFFT_KIND != 4
// staged stores can be switched off for A/B measurements: ACCFFT_B200_STAGED=0
static bool staged_enabled() {
  static const int on = [] {
    const char* e = getenv("ACCFFT_B200_STAGED");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on != 0;
}

template <class T, class S>
static int launch_one(const Launch& l, cudaStream_t st) {
  using C = Cfg<T, S>;
#if ACCFFT_KIND == 1 || ACCFFT_KIND == 2
  StageDesc d = l.desc;
  d.staged = (d.nseg > 1 && d.op == OP_NONE && d.accumulate == 0 && staged_enabled()) ? 1 : 0;
  const int contig_smem = l.dry || d.staged ? C::CONTIG_STAGED_SMEM : C::CONTIG_SMEM;
#else
  const StageDesc& d = l.desc;
#endif
  const bool gen = !stage_is_simple(d);
  (void)gen;
#if ACCFFT_KIND == 9
  // wide CTA shape of the plain strided kernel (its own translation unit: k9)
  if constexpr (C::HAS_WIDE) {
    using W = CfgW<T, S>;
    const long long nbw = (long long)d.nA * ((d.nB + d.colShift + W::TC - 1) / W::TC);
    return launch_pair(k_strided_c2c_wide<T, S, false>, k_strided_c2c_wide<T, S, true>, gen, nbw, W::STRIDED_THREADS,
                       W::STRIDED_SMEM, W::STRIDED_SMEM, d, st, l.dry);
  } else {
    return l.dry ? 0 : -2;   // no wide shape for this length / precision: the caller falls back to the default shape
  }
#elif ACCFFT_KIND == 0
  if constexpr (sizeof(T) == 4) {
    // single precision: the default shape is the pair kernel (every strided stage has unit column stride on its input side;
    // the 8-byte-element form of the plain kernel is not instantiated at all)
    using PC = PairCfg<S>;
    if (d.inStrideB != 1 && !l.dry) return (int)cudaErrorInvalidValue;
    const long long nbp = (long long)d.nA * ((d.nB + d.colShift + 2 * PC::TCP - 1) / (2 * PC::TCP));
    return launch_pair(k_strided_c2c_pair<S, false>, k_strided_c2c_pair<S, true>, gen, nbp, PC::THREADS, PC::SMEM, PC::SMEM, d, st,
                       l.dry);
  } else {
    const long long nb = (long long)d.nA * ((d.nB + d.colShift + C::TC - 1) / C::TC);
    return launch_pair(k_strided_c2c<T, S, false>, k_strided_c2c<T, S, true>, gen, nb, C::STRIDED_THREADS, C::STRIDED_SMEM,
                       C::STRIDED_SMEM, d, st, l.dry);
  }
#elif ACCFFT_KIND == 8
  const long long npairs = ((long long)d.nA * d.nB + 1) / 2;
  return launch_kernel(k_contig_deriv<T, S>, (npairs + C::PC - 1) / C::PC, C::CONTIG_THREADS, C::CONTIG_SMEM, d, st, l.dry);
#elif ACCFFT_KIND >= 5
  using RC = RealCfg<T, S>;
  const long long nb = (long long)d.nA * ((d.nB + 2 * RC::TC - 1) / (2 * RC::TC));
#if ACCFFT_KIND == 5
  return launch_kernel(k_strided_r2c<T, S>, nb, RC::THREADS, RC::SMEM, d, st, l.dry);
#elif ACCFFT_KIND == 6
  return launch_kernel(k_strided_c2r<T, S>, nb, RC::THREADS, RC::SMEM, d, st, l.dry);
#else
  return launch_kernel(k_strided_deriv<T, S>, nb, RC::THREADS, RC::SMEM, d, st, l.dry);
#endif
#elif ACCFFT_KIND == 1
  const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
  return launch_pair(k_contig_c2c<T, S, false>, k_contig_c2c<T, S, true>, gen, nb, C::CONTIG_THREADS, C::CONTIG_SMEM, contig_smem,
                     d, st, l.dry);
#elif ACCFFT_KIND == 2
  const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
  return launch_pair(k_contig_r2c<T, S, false>, k_contig_r2c<T, S, true>, gen, nb, C::CONTIG_THREADS, C::CONTIG_SMEM, contig_smem,
                     d, st, l.dry);
#else
  const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
  return launch_pair(k_contig_c2r<T, S, false>, k_contig_c2r<T, S, true>, gen, nb, C::CONTIG_THREADS, C::CONTIG_SMEM, C::CONTIG_SMEM,
                     d, st, l.dry);
#endif
}

#endif
#if ACCFFT_KIND == 9
// does this length / precision have a wide CTA shape (kernel_cfg.hpp: CfgW)?  Host-side query for the variant choice.
template <class T>
static int has_wide_any(int n) {
#define X(N_, G_, R0, R1, R2, R3) \
  if (n == N_) return Cfg<T, Sched<N_, G_, R0, R1, R2, R3>>::HAS_WIDE ? 1 : 0;
#include ACCFFT_SCHEDULES_FILE
#undef X
  return 0;
}
#endif
#if ACCFFT_KIND == 4
// ---- TMA-pipelined strided stage (stage_tma.cuh) ------------------------------------------------------------
// returns -2 when the launch does not qualify (caller falls back to k_strided_c2c)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (EncodeTiledFn)p;
  }();
  return fn;
}
static int sm_count() {
  static thread_local int dev_cached = -1, sms = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dev_cached = dev;
  }
  return sms > 0 ? sms : 148;
}

template <class T, class S>
static int launch_tma(const Launch& l, cudaStream_t st) {
  using C0 = Cfg<T, S>;
  using C = TmaCfg<T, S, C0::TC>;
  if constexpr (!C::OK) {
    return -2;
  } else {
    const StageDesc& d = l.desc;
    const bool gen = !stage_is_simple(d);
    auto kernel = gen ? k_strided_tma<T, S, C0::TC, true> : k_strided_tma<T, S, C0::TC, false>;
    if (l.dry) {
      cudaFuncAttributes fa;
      cudaError_t e = cudaSuccess;
      for (auto k : {k_strided_tma<T, S, C0::TC, false>, k_strided_tma<T, S, C0::TC, true>}) {
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, k);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
      }
      encode_tiled_fn();
      sm_count();
      return (int)e;
    }
    constexpr long long es = C::ES;
    EncodeTiledFn enc = encode_tiled_fn();
    // (every box must start on a 16-byte boundary in global memory: a column shift of an odd number of 8-byte
    // elements would start the boxes of a single-precision tile 8 bytes off -- the TMA unit faults on that)
    if (!enc || d.inStrideB != 1 || ((uintptr_t)d.in & 15) || (d.inStrideI * es) % 16 != 0 || (d.colShift * es) % 16 != 0 ||
        (d.nA > 1 && (d.inStrideA * es) % 16 != 0) || d.inStrideI * es >= (1LL << 40) || d.inStrideA * es >= (1LL << 40))
      return -2;
    const int tpa = (d.nB + d.colShift + C0::TC - 1) / C0::TC;
    const long long ntiles = (long long)d.nA * tpa;
    if (ntiles <= 0) return 0;
    alignas(64) CUtensorMap map;
    constexpr unsigned E = (unsigned)(es / 8);
    const cuuint64_t gdim[3] = {(cuuint64_t)d.nB * E, (cuuint64_t)S::N, (cuuint64_t)d.nA};
    const cuuint64_t gstr[2] = {(cuuint64_t)(d.inStrideI * es),
                                (cuuint64_t)((d.nA > 1 ? d.inStrideA : d.inStrideI * S::N) * es)};
    const cuuint32_t box[3] = {(cuuint32_t)(C0::TC * E), (cuuint32_t)C::BOXR, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(d.in), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -2;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    long long grid = (long long)sm_count() * C::CTAS;
    if (grid > ntiles) grid = ntiles;
    kernel<<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(d, map, ntiles);
    return (int)cudaGetLastError();
  }
}

template <class T>
static int launch_any(const Launch& l, cudaStream_t st) {
#define X(N_, G_, R0, R1, R2, R3) \
  if (l.fftN == N_) return launch_tma<T, Sched<N_, G_, R0, R1, R2, R3>>(l, st);
#include ACCFFT_SCHEDULES_FILE
#undef X
  return -1;
}
#else
template <class T>
static int launch_any(const Launch& l, cudaStream_t st) {
#define X(N_, G_, R0, R1, R2, R3) \
  if (l.fftN == N_) return launch_one<T, Sched<N_, G_, R0, R1, R2, R3>>(l, st);
#include ACCFFT_SCHEDULES_FILE
#undef X
  return -1;
}
#endif

}  // namespace accfft_b200
