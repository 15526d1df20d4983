// fused_tu.cu -- compiled once per (ACCFFT_FUSED_MODE, ACCFFT_PREC): the fused two-pass persistent kernel (fused.cuh) for the
// length pairs of fused_combos.def, and its launcher.
//   ACCFFT_FUSED_MODE: 0 forward r2c (Z then Y), 1 inverse c2r (Y then Z), 2 forward c2c, 3 inverse c2c
//   ACCFFT_PREC: 0 double, 1 float
#include <stdlib.h>
#include "fused.cuh"
#include "fused_launch.hpp"

#ifndef ACCFFT_SCHEDULES_FILE
#define ACCFFT_SCHEDULES_FILE "schedules.def"
#endif

namespace accfft_b200 {

template <int N> struct SchedOf;
#define X(N_, G_, R0, R1, R2, R3) \
  template <> struct SchedOf<N_> { using type = Sched<N_, G_, R0, R1, R2, R3>; };
#include ACCFFT_SCHEDULES_FILE
#undef X

static int fused_sm_count() {
  static thread_local int dev_cached = -1, sms = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dev_cached = dev;
  }
  return sms > 0 ? sms : 148;
}

template <class T, class SZ, class SY, int MODE>
static int fused_launch_one(const Launch& lz, const Launch& ly, const FusedTune& tune, int* ctr, int ctr_ints, cudaStream_t st, bool dry,
                            FusedInfo* info) {
  using C = FusedCfg<T, SZ, SY>;
  if constexpr (!C::OK) {
    return -2;
  } else {
    constexpr bool FWD = MODE == 0 || MODE == 2;
    auto ks = k_fused<T, SZ, SY, MODE, false>;
    auto kg = k_fused<T, SZ, SY, MODE, FWD>;   // the general strided pass exists for the forward direction only
    if (dry) {
      cudaFuncAttributes fa;
      cudaError_t e = cudaFuncGetAttributes(&fa, ks);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
      if (FWD && e == cudaSuccess) e = cudaFuncGetAttributes(&fa, kg);
      if (FWD && e == cudaSuccess) e = cudaFuncSetAttribute(kg, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
      fused_sm_count();
      return (int)e;
    }
    StageDesc dz = lz.desc, dy = ly.desc;
    // the contiguous pass must be a simple stage; the strided pass may store to peers in the forward direction
    if (!stage_is_simple(dz) || dz.nA != dy.nA || dz.nA <= 0 || dz.nB <= 0 || dy.nB <= 0) return -2;
    const bool geny = !stage_is_simple(dy);
    if (geny && !FWD) return -2;
    if (C::F && dy.inStrideB != 1) return -2;
    const int nA = dz.nA;
    constexpr int TCE = C::F ? 2 * PairCfg<SY>::TCP : Cfg<T, SY>::TC;
    long long grid = (long long)fused_sm_count() * C::MINB;
    FusedArgs fa;
    FusedShape sh;
    if (!fused_make_args(nA, dz.nB, C::PCZ, (dy.nB + dy.colShift + TCE - 1) / TCE, grid, tune.cp, tune.lag, tune.ns, tune.flags, FWD,
                         ctr_ints, fa, sh))
      return -2;
    fa.ctr = ctr;
    const int cp = sh.cp, lag = sh.lag, ns = sh.ns, ringA = sh.ring_planes;
    const int mask = ringA ? ringA - 1 : 0x7fffffff;   // a ring as long as the array is plain addressing (mask = all ones)
    if (FWD) dz.outRing = mask, dy.inRing = mask;
    else dy.outRing = mask, dz.inRing = mask;
    if (info) {
      info->cp = cp, info->lag = lag, info->ns = ns, info->nch = fa.nch, info->ring_planes = ringA, info->grid = (int)grid;
      info->threads = C::THREADS, info->smem = C::SMEM;
    }
    if (grid > fa.total) grid = fa.total;
    cudaError_t e = cudaMemsetAsync(ctr, 0, sizeof(int) * (size_t)(1 + 2 * fa.nch), st);
    if (e != cudaSuccess) return (int)e;
    auto k = geny ? kg : ks;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    k<<<(unsigned)grid, C::THREADS, C::SMEM, st>>>(dz, dy, fa);
    return (int)cudaGetLastError();
  }
}

#define ACCFFT_FCAT(a, b, c, d) a##b##c##d
#define ACCFFT_FNAME(m, p) ACCFFT_FCAT(accfft_launch_fused_m, m, _p, p)

#if ACCFFT_PREC == 0
typedef double fused_real;
#else
typedef float fused_real;
#endif

// lz: the contiguous (Z) stage, ly: the strided (Y) stage, whichever runs first.  -2: no fused kernel for this pair / shape.
int ACCFFT_FNAME(ACCFFT_FUSED_MODE, ACCFFT_PREC)(const Launch& lz, const Launch& ly, const FusedTune& tune, int* ctr, int ctr_ints,
                                                  cudaStream_t st, bool dry, FusedInfo* info) {
#define F(NY, MZ)                                                                                                                  \
  if (ly.fftN == NY && lz.fftN == MZ)                                                                                              \
    return fused_launch_one<fused_real, SchedOf<MZ>::type, SchedOf<NY>::type, ACCFFT_FUSED_MODE>(lz, ly, tune, ctr, ctr_ints, st, dry, \
                                                                                                 info);
#include "fused_combos.def"
#undef F
  return -2;
}

}  // namespace accfft_b200
