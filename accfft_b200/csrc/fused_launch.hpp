// fused_launch.hpp -- host-side interface of the fused two-pass kernel (fused.cuh / fused_tu.cu).
#pragma once
#include <cuda_runtime.h>
#include "launch.hpp"

namespace accfft_b200 {

struct FusedTune {
  int cp;      // planes of the outer index per chunk (0: default)
  int lag;     // chunks the second pass runs behind the first (0: derived from the number of resident CTAs)
  int ns;      // ring slots in chunks (0: lag + 2)
  int flags;   // debugging (FusedArgs::flags): 1 = signal tiles right away, 2 = no early dependency check
};
struct FusedInfo {
  int cp, lag, ns, nch, ring_planes, grid, threads, smem;
};

#define ACCFFT_FUSED_DECL(m, p)                                                                                                   \
  int accfft_launch_fused_m##m##_p##p(const Launch& lz, const Launch& ly, const FusedTune& tune, int* ctr, int ctr_ints, cudaStream_t st, \
                                       bool dry, FusedInfo* info);
ACCFFT_FUSED_DECL(0, 0)
ACCFFT_FUSED_DECL(0, 1)
ACCFFT_FUSED_DECL(1, 0)
ACCFFT_FUSED_DECL(1, 1)
ACCFFT_FUSED_DECL(2, 0)
ACCFFT_FUSED_DECL(2, 1)
ACCFFT_FUSED_DECL(3, 0)
ACCFFT_FUSED_DECL(3, 1)
#undef ACCFFT_FUSED_DECL

}  // namespace accfft_b200
