// kernel_cfg.hpp -- CTA shapes derived from a schedule (shared by the CUDA backend and tests/emul).
#pragma once
#include "fft_engine.cuh"

#ifndef ACCFFT_SCHEDULES_FILE
#define ACCFFT_SCHEDULES_FILE "schedules.def"
#endif

namespace accfft_b200 {

#ifndef ACCFFT_MAX_STRIDED_THREADS
#define ACCFFT_MAX_STRIDED_THREADS 512
#endif
#ifndef ACCFFT_CONTIG_THREADS
#define ACCFFT_CONTIG_THREADS 128
#endif
#ifndef ACCFFT_MAX_STRIDED_SMEM
#define ACCFFT_MAX_STRIDED_SMEM (100 * 1024)
#endif

// strided kernels: TC adjacent columns per tile; 128-byte row segments when threads / smem allow
template <class T, class S>
constexpr int strided_tc_pick(int max_threads, long long max_smem) {
  int tc = 128 / (int)sizeof(cplx<T>);
  while (tc > 1 && (S::G * tc > max_threads || (long long)S::SMEM_ELEMS * tc * (int)sizeof(cplx<T>) > max_smem)) tc /= 2;
  return tc;
}

// "wide" CTA shape of the plain strided kernel: up to 1024 threads and 200 KB of shared memory, ONE CTA per SM.  It exists
// for the lengths / precisions whose default shape cannot reach 128-byte row segments (threads per column G > 64 in
// double, > 32 in single: 512 single, 768, 1024, ... -- e.g. 1024 single has 32-byte segments by default), which costs most
// where the segments cross NVLink (profiles/mgpu8_r11.log: 290 GB/s out per GPU at 1024^3 single vs ~600 at 512^3 double).
template <class T, class S>
struct CfgW {
  static constexpr int TC = strided_tc_pick<T, S>(1024, 200 * 1024);
  static constexpr int STRIDED_THREADS = S::G * TC;
  static constexpr int STRIDED_SMEM = S::SMEM_ELEMS * TC * (int)sizeof(cplx<T>);
};

// single-precision pair kernel (stage.cuh: strided_c2c_pair_body): TCP pair-columns (= 2 * TCP columns) per tile
template <class S>
struct PairCfg {
  static constexpr int pick() {
    int tc = 8;   // 16 columns = 128-byte rows
    while (tc > 1 && (S::G * tc > ACCFFT_MAX_STRIDED_THREADS || (long long)S::SMEM_ELEMS * tc * 16 > ACCFFT_MAX_STRIDED_SMEM)) tc /= 2;
    return tc;
  }
  static constexpr int TCP = pick();
  static constexpr int THREADS = S::G * TCP;
  static constexpr int SMEM = S::SMEM_ELEMS * TCP * 16;
  // resident CTAs the kernel is compiled for: 1024 threads per SM (<= 64 registers) when a thread holds up to 8 lane pairs;
  // schedules with 16 values per thread (e.g. the single-precision 1024-point one) get the full register file instead
  static constexpr int MINB = S::NREG > 8 ? 1 : ((1024 / THREADS) > 0 ? (1024 / THREADS) : 1);
};

template <class T, class S>
struct Cfg {
  static constexpr int TC = strided_tc_pick<T, S>(ACCFFT_MAX_STRIDED_THREADS, ACCFFT_MAX_STRIDED_SMEM);
  // true when the wide shape has longer row segments than this one
  static constexpr bool HAS_WIDE = CfgW<T, S>::TC > TC;
  static constexpr int STRIDED_THREADS = S::G * TC;
  static constexpr int STRIDED_SMEM = S::SMEM_ELEMS * TC * (int)sizeof(cplx<T>);
  // contiguous kernels: PC pencils per CTA, about ACCFFT_CONTIG_THREADS threads
  static constexpr int PC = (S::G >= ACCFFT_CONTIG_THREADS) ? 1 : ACCFFT_CONTIG_THREADS / S::G;
  static constexpr int CONTIG_THREADS = S::G * PC;
  static constexpr int CONTIG_SMEM = S::SMEM_ELEMS * PC * (int)sizeof(cplx<T>);
  // staged stores (stage.cuh: staged_flush): room for the PC x NOUT outputs behind the transform scratch,
  // NOUT = N (c2c) or N + 1 (r2c); N + 2 keeps the region a multiple of 16 bytes for both
  // (the region starts at an even element index so that it is 16-byte aligned for float as well)
  static constexpr int CONTIG_STAGED_SMEM = (((S::SMEM_ELEMS * PC + 1) & ~1) + (S::N + 2) * PC) * (int)sizeof(cplx<T>);
};

}  // namespace accfft_b200
