// launch.hpp -- what a stage launch is: kernel kind, precision, length and the stage descriptor.  This is all the kernel
// translation units (k_tu.cu / launch_impl.cuh) need to know of the plan; plan.hpp (geometry, Backend, Plan) builds on it.
#pragma once
#include "stage_real.cuh"

namespace accfft_b200 {

enum Kind : int {
  K_STRIDED_C2C = 0, K_CONTIG_C2C = 1, K_CONTIG_R2C = 2, K_CONTIG_C2R = 3, K_COPY = 4,
  // real transforms along a strided axis (stage_real.cuh): the single-axis executes and the per-axis derivative
  K_STRIDED_R2C = 5, K_STRIDED_C2R = 6, K_STRIDED_DERIV = 7, K_COPY_REAL = 8,
  K_CONTIG_DERIV = 9   // fused derivative along the contiguous (Z) axis: launcher translation unit k8
};
enum Prec : int { PREC_DOUBLE = 0, PREC_FLOAT = 1 };

struct Launch {
  Kind kind;
  Prec prec;
  int fftN;        // length of the complex engine transform (M = N/2 for r2c / c2r)
  bool inverse;
  bool dry;        // true: only load / configure the kernel (plan time), do not launch
  int slot;        // profiling slot: 0 fwd Z, 1 fwd Y, 2 fwd X, 3 inv X, 4 inv Y, 5 inv Z  (6 = barrier),
                   // 7 = real copy / exchange stages, 8 = single-axis real transform / derivative stages
  long long copyN; // K_COPY_REAL: extent of the fastest input index k (a flat plane index on the x path)
  int lane;        // 0 = the plan's stream, 1 = the plan's auxiliary stream (software-pipelined stages, see Backend::fork)
  int variant;     // strided stages: 0 = backend's choice (plan-time tuning / heuristic), 1 = plain kernel, 2 = TMA-pipelined kernel
  StageDesc desc;
};

}  // namespace accfft_b200
