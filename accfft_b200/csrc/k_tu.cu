// k_tu.cu -- compiled once per (ACCFFT_KIND, ACCFFT_PREC) pair (see Makefile): instantiates the stage
// kernels of that kind / precision for every schedule and exports one launcher.
//   ACCFFT_KIND: 0 strided c2c (Y and X passes), 1 contiguous c2c (Z pass of c2c plans),
//                2 contiguous r2c, 3 contiguous c2r (Z pass of r2c plans), 4 TMA-pipelined strided c2c,
//                5 strided r2c, 6 strided c2r, 7 strided spectral derivative (single-axis real paths, stage_real.cuh),
//                8 contiguous (Z-axis) spectral derivative, 9 strided c2c in the wide CTA shape (kernel_cfg.hpp: CfgW)
//   ACCFFT_PREC: 0 double, 1 float
#include "launch_impl.cuh"

#define ACCFFT_CAT3(a, b, c, d) a##b##c##d
#define ACCFFT_NAME(k, p) ACCFFT_CAT3(accfft_launch_k, k, _p, p)

namespace accfft_b200 {
#if ACCFFT_PREC == 0
int ACCFFT_NAME(ACCFFT_KIND, ACCFFT_PREC)(const Launch& l, cudaStream_t st) { return launch_any<double>(l, st); }
#else
int ACCFFT_NAME(ACCFFT_KIND, ACCFFT_PREC)(const Launch& l, cudaStream_t st) { return launch_any<float>(l, st); }
#endif
#if ACCFFT_KIND == 9
#if ACCFFT_PREC == 0
int accfft_has_wide_p0(int n) { return has_wide_any<double>(n); }
#else
int accfft_has_wide_p1(int n) { return has_wide_any<float>(n); }
#endif
#endif
}  // namespace accfft_b200
