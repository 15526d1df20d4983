// fused.cuh -- ONE persistent kernel for the two passes of a transform that are local to a rank on a 1 x 1 or slab
// (P1 == 1) grid: forward Z then Y, inverse Y then Z   (reference: the cufftExecD2Z + per-slab cufftExecZ2Z pair of
// src/accfft_gpu.cpp:477-516 and its mirror :560-602, with T_plan_1 between them degenerating to a local transpose).
//
// Run as two kernels, the array between the two passes costs one HBM write and one HBM read (2 C of the R + 5 C bytes of a
// transform).  Here both passes are tiles of one work list that the resident CTAs claim in order from a global counter:
//     step s:   [pass-1 tiles of chunk s]  [pass-2 tiles of chunk s - lag]          chunk = `cp` planes of the outer index a
// and the intermediate buffer is addressed as a RING of `ns * cp` planes (plane a lives at slot a % ring, StageDesc::inRing /
// outRing) -- a few tens of MB that stay in the 126 MB L2: pass 1 writes a plane, pass 2 reads it back a few microseconds
// later, and the slot is overwritten while still dirty in L2, so the intermediate never reaches HBM at all.
// Dependencies are per chunk through two counters (pass-1 tiles done, pass-2 tiles done):
//     pass-2 tile of chunk c   waits for  done1[c] == t1                 (its input planes are complete)
//     pass-1 tile of chunk c   waits for  done2[c - ns] == t2            (the ring slot it overwrites has been read)
// Both wait only on tiles that were claimed EARLIER in the work list (ns > lag), i.e. by CTAs that are already running, so the
// scheme cannot deadlock whatever the number of resident CTAs.  With lag * (t1 + t2) >= resident CTAs the waits are almost
// never taken.  Tiles run the same bodies as the stand-alone kernels (stage.cuh) under a context that enables the ring
// addressing; ring data is read with ld.global.cg (L2), everything else streams (evict-first).
#pragma once
#include <cuda_runtime.h>
#include "kernel_cfg.hpp"
#include "launch.hpp"

namespace accfft_b200 {

struct FusedArgs {
  int* ctr;          // [0] work counter, [1 .. nch] pass-1 tiles done per chunk, [1 + nch .. 2 nch] pass-2 tiles done per chunk
  int nch, cp;       // chunks, planes per chunk
  int lag, ns;       // pass 2 runs `lag` chunks behind pass 1; ring slots (in chunks), ns > lag
  int t1, t2;        // tiles per chunk of pass 1 / pass 2
  int tpa;           // strided pass: tiles per plane
  int tz;            // contiguous pass: tiles per chunk (= cp * pencils per plane / PCZ)
  long long total;   // work items = (nch + lag) * (t1 + t2)
};

struct FusedCtx {
  static constexpr bool kRing = true;
  long long bid_;
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ long long bid() const { return bid_; }
  __device__ __forceinline__ void* smem() const {
    extern __shared__ __align__(16) unsigned char accfft_smem_raw[];
    return accfft_smem_raw;
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ void bulk_fence() const {}
  __device__ __forceinline__ void bulk_store(void*, const void*, unsigned) const {}
  __device__ __forceinline__ void bulk_wait() const {}
};

// CTA shape: the strided pass decides the thread count (its tile = 128-byte row segments), the contiguous pass takes
// as many pencils per tile as those threads cover
template <class T, class SZ, class SY>
struct FusedCfg {
  static constexpr bool F = sizeof(T) == 4;
  static constexpr int THREADS = F ? PairCfg<SY>::THREADS : Cfg<T, SY>::STRIDED_THREADS;
  static constexpr int SMEM_Y = F ? PairCfg<SY>::SMEM : Cfg<T, SY>::STRIDED_SMEM;
  static constexpr int PCZ = THREADS / SZ::G;
  static constexpr int SMEM_Z = (PCZ > 0 ? PCZ : 1) * SZ::SMEM_ELEMS * (int)sizeof(cplx<T>);
  static constexpr int SMEM = (SMEM_Y > SMEM_Z ? SMEM_Y : SMEM_Z) + 16;
  static constexpr int minb() {
    int m = F ? PairCfg<SY>::MINB : ((1024 / THREADS) > 0 ? (1024 / THREADS) : 1);
    while (m > 1 && (long long)m * (SMEM + 1024) > 227 * 1024) --m;
    return m;
  }
  static constexpr int MINB = minb();
  static constexpr bool OK = PCZ >= 1 && THREADS % SZ::G == 0 && THREADS <= 1024 && SMEM <= 220 * 1024;
};

__device__ __forceinline__ int fused_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// MODE 0: forward r2c (contig r2c -> strided), 1: inverse c2r (strided -> contig c2r),
//      2: forward c2c (contig c2c -> strided), 3: inverse c2c (strided -> contig c2c).  GENY: the strided pass is the general
// instantiation (forward on a slab grid: it stores to peers).
template <class T, class SZ, class SY, int MODE, bool GENY>
__global__ void __launch_bounds__(FusedCfg<T, SZ, SY>::THREADS, FusedCfg<T, SZ, SY>::MINB)
    k_fused(const __grid_constant__ StageDesc dz, const __grid_constant__ StageDesc dy, const FusedArgs fa) {
  using C = FusedCfg<T, SZ, SY>;
  constexpr bool Z_FIRST = (MODE == 0 || MODE == 2);
  __shared__ int s_item;
  const int tid = (int)threadIdx.x;
  int* const done1 = fa.ctr + 1;
  int* const done2 = fa.ctr + 1 + fa.nch;
  const int per = fa.t1 + fa.t2;
#pragma unroll 1
  for (;;) {
    if (tid == 0) s_item = atomicAdd(fa.ctr, 1);
    __syncthreads();
    const long long w = s_item;
    if (w >= fa.total) break;
    const int s = (int)(w / per), i = (int)(w % per);
    const bool ph1 = i < fa.t1;
    const int c = ph1 ? s : s - fa.lag;
    const int t = ph1 ? i : i - fa.t1;
    const bool live = c >= 0 && c < fa.nch;
    if (live && tid == 0) {
      if (ph1) {
        if (c >= fa.ns)
          while (fused_ld_acquire(done2 + c - fa.ns) < fa.t2) __nanosleep(64);
      } else {
        while (fused_ld_acquire(done1 + c) < fa.t1) __nanosleep(64);
      }
    }
    __syncthreads();   // (also: everyone has read s_item before thread 0 claims the next item)
    if (!live) continue;
    const bool ztype = ph1 == Z_FIRST;
    FusedCtx ctx;
    if (ztype) {
      ctx.bid_ = (long long)c * fa.tz + t;
      if (ctx.bid_ * C::PCZ < (long long)dz.nA * dz.nB) {
        if constexpr (MODE == 0) contig_r2c_body<T, SZ, C::PCZ, false>(dz, ctx);
        else if constexpr (MODE == 1) contig_c2r_body<T, SZ, C::PCZ, false>(dz, ctx);
        else contig_c2c_body<T, SZ, C::PCZ, false>(dz, ctx);
      }
    } else {
      ctx.bid_ = (long long)c * fa.cp * fa.tpa + t;
      if (ctx.bid_ / fa.tpa < dy.nA) {
        if constexpr (C::F) strided_c2c_pair_body<SY, PairCfg<SY>::TCP, GENY>(dy, ctx);
        else strided_c2c_body<T, SY, Cfg<T, SY>::TC, GENY>(dy, ctx);
      }
    }
    __syncthreads();   // every thread's stores are issued (and the tile's shared memory is free)
    if (tid == 0) {
      __threadfence();
      atomicAdd(ph1 ? done1 + c : done2 + c, 1);
    }
  }
}

}  // namespace accfft_b200
