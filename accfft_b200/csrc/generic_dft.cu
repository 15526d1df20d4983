// generic_dft.cu -- fallback stage kernel for transform lengths without a compiled Stockham schedule (schedules.def) and for
// r2c / c2r plans with an ODD last dimension.  The reference hands any length to cuFFT / FFTW (src/accfft_gpu.cpp:173-253,
// src/accfft.cpp:179-197), so a drop-in must plan any length too; here such an axis is transformed by the O(N^2) definition,
//     y[k] = sum_j x[j] W_N^(+-jk),
// one column per CTA, inputs and the root table staged in shared memory, double-precision accumulation.  Slow next to the
// scheduled kernels (it is meant for the odd axis of an awkward grid, not for 1021^3), but it takes exactly the same stage
// descriptor -- segments / peer destinations, spectral operators, accumulation, in-place padded rows -- so every plan-level
// feature works for every length.  accfftb_length_supported() tells the two classes of lengths apart.
#include <cuda_runtime.h>

#include "launch.hpp"

namespace accfft_b200 {

constexpr int kGenericThreads = 256;
constexpr int kGenericMaxSmem = 224 * 1024;   // inputs + roots of one column: N <= 7168 (double), 14336 (single)

// MODE 0: complex -> complex (K_STRIDED_C2C, K_CONTIG_C2C), 1: real -> half spectrum (K_CONTIG_R2C), 2: half spectrum -> real
// (K_CONTIG_C2R).  N = transform length (the REAL length for modes 1 and 2).  Stride conventions are those of the scheduled
// kernels of the same kind (stage.cuh): real sides count in real elements.
template <class T, int MODE>
__global__ void __launch_bounds__(kGenericThreads) k_generic_dft(const __grid_constant__ StageDesc d, int N) {
  extern __shared__ __align__(16) unsigned char generic_smem_raw[];
  cplx<T>* x = (cplx<T>*)generic_smem_raw;
  cplx<T>* w = x + N;
  const long long col = blockIdx.x;
  const int a = (int)(col / d.nB), b = (int)(col % d.nB);
  const int tid = (int)threadIdx.x;
  const bool inv = MODE == 2 || (MODE == 0 && d.inv != 0);
  const cplx<T>* tw = (const cplx<T>*)d.tw;
  for (int i = tid; i < N; i += kGenericThreads) {
    cplx<T> v;
    if (MODE == 1) {
      const T* in = (const T*)d.in + (long long)a * d.inStrideA + (long long)b * d.inStrideB;
      v = {in[(long long)i * d.inStrideI], T(0)};
    } else if (MODE == 2) {
      const cplx<T>* in = (const cplx<T>*)d.in + (long long)a * d.inStrideA + (long long)b * d.inStrideB;
      const bool lower = 2 * i <= N;
      const int m = lower ? i : N - i;
      v = in[(long long)m * d.inStrideI];
      if (d.op != OP_NONE) v = apply_op(d, v, m, a, b);
      if (m == 0 || 2 * m == N) v.y = T(0);   // FFTW / cuFFT ignore these
      if (!lower) v.y = -v.y;                 // Hermitian extension
    } else {
      const cplx<T>* in = (const cplx<T>*)d.in + (long long)a * d.inStrideA + (long long)b * d.inStrideB;
      v = in[(long long)i * d.inStrideI];
      if (d.opOnLoad) v = apply_op(d, v, i, a, b);
    }
    x[i] = v;
    w[i] = tw[i];
  }
  __syncthreads();
  const int nout = MODE == 1 ? N / 2 + 1 : N;
  for (int k = tid; k < nout; k += kGenericThreads) {
    double sr = 0.0, si = 0.0;
    int idx = 0;
    for (int j = 0; j < N; ++j) {
      const cplx<T> wv = w[idx];
      const double wr = (double)wv.x, wi = inv ? -(double)wv.y : (double)wv.y;
      const cplx<T> xv = x[j];
      sr += (double)xv.x * wr - (double)xv.y * wi;
      si += (double)xv.x * wi + (double)xv.y * wr;
      idx += k;
      if (idx >= N) idx -= N;
    }
    if (MODE == 2) {
      const SegDst& sd = d.seg[0];
      T* out = (T*)sd.base + (long long)a * sd.strideA + (long long)b * sd.strideB + (long long)k * sd.strideI;
      *out = d.accumulate ? T(*out + (T)sr) : (T)sr;
    } else {
      store_generic(d, cplx<T>{(T)sr, (T)si}, k, a, b);
    }
  }
}

template <class T, int MODE>
static int launch_generic_mode(const Launch& l, int N, cudaStream_t st) {
  auto kernel = k_generic_dft<T, MODE>;
  const int smem = 2 * N * (int)sizeof(cplx<T>);
  if (smem > kGenericMaxSmem) return -1;
  cudaError_t e = cudaSuccess;
  if (l.dry || smem > 48 * 1024) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 48 * 1024 ? smem : 48 * 1024);
  if (e != cudaSuccess) return (int)e;
  if (l.dry) {
    cudaFuncAttributes fa;
    return (int)cudaFuncGetAttributes(&fa, kernel);
  }
  const long long cols = (long long)l.desc.nA * l.desc.nB;
  if (cols <= 0) return 0;
  if (cols > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  kernel<<<(unsigned)cols, kGenericThreads, smem, st>>>(l.desc, N);
  return (int)cudaGetLastError();
}

// realN: the real length of the Z axis for K_CONTIG_R2C / K_CONTIG_C2R (their Launch::fftN is the half length of the
// scheduled kernels); ignored for the complex kinds.  Returns -1 for kinds / lengths this kernel does not serve.
// (Column shifts of the descriptor only re-align tile origins of the scheduled strided kernels; columns are addressed
// directly here.)
int accfft_launch_generic(const Launch& l, int realN, cudaStream_t st) {
  const bool real_kind = l.kind == K_CONTIG_R2C || l.kind == K_CONTIG_C2R;
  const int N = real_kind ? realN : l.fftN;
  if (N < 1) return -1;
  const bool dbl = l.prec == PREC_DOUBLE;
  switch (l.kind) {
    case K_STRIDED_C2C:
    case K_CONTIG_C2C: return dbl ? launch_generic_mode<double, 0>(l, N, st) : launch_generic_mode<float, 0>(l, N, st);
    case K_CONTIG_R2C: return dbl ? launch_generic_mode<double, 1>(l, N, st) : launch_generic_mode<float, 1>(l, N, st);
    case K_CONTIG_C2R: return dbl ? launch_generic_mode<double, 2>(l, N, st) : launch_generic_mode<float, 2>(l, N, st);
    default: return -1;
  }
}
int accfft_generic_max_length(Prec prec) { return kGenericMaxSmem / (2 * (prec == PREC_DOUBLE ? 16 : 8)); }

}  // namespace accfft_b200
