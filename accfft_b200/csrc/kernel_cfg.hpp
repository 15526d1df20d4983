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

template <class T, class S>
struct Cfg {
  // strided kernels: TC adjacent columns per tile; 128-byte row segments when threads / smem allow
  static constexpr int tc_pick() {
    int tc = 128 / (int)sizeof(cplx<T>);
    while (tc > 1 && (S::G * tc > ACCFFT_MAX_STRIDED_THREADS ||
                      (long long)S::SMEM_ELEMS * tc * (int)sizeof(cplx<T>) > ACCFFT_MAX_STRIDED_SMEM))
      tc /= 2;
    return tc;
  }
  static constexpr int TC = tc_pick();
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
