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
#include "cuda_ctx.cuh"
#include "kernel_cfg.hpp"
#include "launch.hpp"
#include "fused_sched.hpp"

namespace accfft_b200 {

__device__ __forceinline__ int fused_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int fused_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// scheduler state of a CTA (shared memory; only thread 0 touches anything but `next`)
struct FusedSched {
  FusedArgs fa;
  int pending;     // ctr index of the previous tile's completion counter, not yet signalled (0: none)
  int slot;        // slot of `next` the next item is published to
  int hooked;      // the scheduler has run for the current tile
  FusedSlot next[2];
};
__device__ __forceinline__ FusedSched& fused_sched() {
  __shared__ FusedSched ss;
  return ss;
}

// Execution context of a tile inside the fused kernel.  Besides the ring addressing (RIN / ROUT: the stage's input / output is
// the ring) it carries the tile scheduler, which runs in thread 0 from the `after_loads` hook the stage bodies call once their
// global loads are issued -- i.e. in the shadow of the tile's own memory latency instead of between two tiles:
//   * the completion signal of the PREVIOUS tile (fence + counter increment) is sent there,
//   * the NEXT work item (claimed from the global counter when this tile started) is decoded, its dependency counter is read
//     once, and the decoded item + "dependency already satisfied" are published to shared memory for the next loop iteration.
// The loop itself then costs one CTA barrier per tile; it polls only when that early check failed (after flushing the pending
// signal, so that a CTA never waits while it holds back a signal).
struct FusedSchedOps {
  static __device__ __forceinline__ void flush() {
    FusedSched& ss = fused_sched();
    if (ss.pending) {
      __threadfence();
      atomicAdd(ss.fa.ctr + ss.pending, 1);
      ss.pending = 0;
    }
  }
  static __device__ __noinline__ void sched(int claim) {
    FusedSched& ss = fused_sched();
    ss.hooked = 1;
    FusedSlot it = fused_decode(ss.fa, claim);
    if (ss.fa.flags & 4) {   // timing experiment: no synchronisation at all
      it.ready = 1;
      it.sig = 0;
      ss.next[ss.slot] = it;
      return;
    }
    int depv = 0;
    if (it.dep && !(ss.fa.flags & 2)) depv = fused_ld_relaxed(ss.fa.ctr + it.dep);
    __threadfence();   // releases this CTA's earlier stores for the signal below AND makes the relaxed load above an acquire
    if (ss.pending) {
      atomicAdd(ss.fa.ctr + ss.pending, 1);
      ss.pending = 0;
    }
    it.ready = (it.dep == 0 || (depv >= it.target && !(ss.fa.flags & 2))) ? 1 : 0;
    ss.next[ss.slot] = it;
  }
};

template <bool RIN, bool ROUT>
struct FusedCtx {
  static constexpr bool kRing = true;
  static constexpr bool kRingIn = RIN;
  static constexpr bool kRingOut = ROUT;
  static constexpr bool kGroupSync = true;
  long long bid_;
  int claim;      // thread 0: the next work item (in flight from the global counter while the tile loads)
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ long long bid() const { return bid_; }
  __device__ __forceinline__ void* smem() const {
    extern __shared__ __align__(16) unsigned char accfft_smem_raw[];
    return accfft_smem_raw;
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  template <int G, int PC>
  __device__ __forceinline__ void sync_group(int pcl) const {
    cuda_sync_group<G, PC>(pcl);
  }
  __device__ __forceinline__ void bulk_fence() const {}
  __device__ __forceinline__ void bulk_store(void*, const void*, unsigned) const {}
  __device__ __forceinline__ void bulk_wait() const {}
  __device__ __forceinline__ void after_loads() const {
    if (threadIdx.x == 0 && !fused_sched().hooked) FusedSchedOps::sched(claim);
  }
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
  // (shapes that leave an SM with fewer than 768 resident threads lose more in the contiguous pass than the ring saves:
  // measured 25 ms vs 13 ms at 1024^3 single, whose strided tile holds 16 lane pairs per thread at one CTA per SM)
  static constexpr bool OK = PCZ >= 1 && THREADS % SZ::G == 0 && THREADS <= 1024 && SMEM <= 220 * 1024 && THREADS * MINB >= 768;
};

// MODE 0: forward r2c (contig r2c -> strided), 1: inverse c2r (strided -> contig c2r),
//      2: forward c2c (contig c2c -> strided), 3: inverse c2c (strided -> contig c2c).  GENY: the strided pass is the general
// instantiation (forward on a slab grid: it stores to peers).
template <class T, class SZ, class SY, int MODE, bool GENY>
__global__ void __launch_bounds__(FusedCfg<T, SZ, SY>::THREADS, FusedCfg<T, SZ, SY>::MINB)
    k_fused(const __grid_constant__ StageDesc dz, const __grid_constant__ StageDesc dy, const FusedArgs fa) {
  using C = FusedCfg<T, SZ, SY>;
  constexpr bool Z_FIRST = (MODE == 0 || MODE == 2);
  // the ring is the output of the first pass and the input of the second
  using CtxZ = FusedCtx<!Z_FIRST, Z_FIRST>;
  using CtxY = FusedCtx<Z_FIRST, !Z_FIRST>;
  FusedSched& ss = fused_sched();
  const int tid = (int)threadIdx.x;
  int claim = 0;
  if (tid == 0) {
    ss.fa = fa;
    ss.pending = 0;
    ss.slot = 0;
    ss.hooked = 0;
    FusedSchedOps::sched((fa.flags & 8) ? (int)blockIdx.x : atomicAdd(fa.ctr, 1));   // publishes the first item to slot 0
  }
  __syncthreads();
#pragma unroll 1
  for (int iter = 0;; ++iter) {
    const int slot = iter & 1;
    const int kind = ss.next[slot].kind;
    const int bid = ss.next[slot].bid;
    if (kind == FK_END) break;
    if (kind != FK_DEAD && !ss.next[slot].ready) {   // rare: the early check in sched() found the dependency open
      if (tid == 0) {
        FusedSchedOps::flush();
        const int* dep = fa.ctr + ss.next[slot].dep;
        const int target = ss.next[slot].target;
        while (fused_ld_acquire(dep) < target) __nanosleep(64);
      }
      __syncthreads();
    }
    if (tid == 0) {
      claim = (fa.flags & 8) ? (int)blockIdx.x + (iter + 1) * (int)gridDim.x
                             : atomicAdd(fa.ctr, 1);   // consumed in sched(), a memory latency from now
      ss.slot = slot ^ 1;
      ss.hooked = 0;
    }
    if (kind == FK_Z) {
      CtxZ ctx;
      ctx.bid_ = bid;
      ctx.claim = claim;
      if constexpr (MODE == 0) contig_r2c_body<T, SZ, C::PCZ, false>(dz, ctx);
      else if constexpr (MODE == 1) contig_c2r_body<T, SZ, C::PCZ, false>(dz, ctx);
      else contig_c2c_body<T, SZ, C::PCZ, false>(dz, ctx);
    } else if (kind == FK_Y) {
      CtxY ctx;
      ctx.bid_ = bid;
      ctx.claim = claim;
      if constexpr (C::F) strided_c2c_pair_body<SY, PairCfg<SY>::TCP, GENY>(dy, ctx);
      else strided_c2c_body<T, SY, Cfg<T, SY>::TC, GENY>(dy, ctx);
    }
    if (tid == 0 && !ss.hooked) FusedSchedOps::sched(claim);   // (no tile, or thread 0 had nothing to load)
    __syncthreads();   // every thread's stores are issued, the tile's shared memory is free, the next item is published
    if (tid == 0 && kind != FK_DEAD && !(fa.flags & 4)) {
      ss.pending = ss.next[slot].sig;
      if (fa.flags & 1) FusedSchedOps::flush();
    }
  }
  if (tid == 0) FusedSchedOps::flush();
}

}  // namespace accfft_b200
