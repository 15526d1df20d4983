/* accfft_operators_gpu.h -- AccFFT's GPU spectral operators (reference: include/accfft_operators_gpu.h:44-67;
 * drivers src/operators_gpu.txx:72-609, kernels src/operators_cuda.cu:15-215) on the B200-native engine.
 * Real field in the input layout (isize) in, real field(s) in the same layout out, normalised, Nyquist modes
 * zeroed.  The plan must be an out-of-place r2c plan, as in the reference (src/operators_gpu.cpp:78-79).
 * timer: NULL or double[7] (+=).  pXYZ selects gradient components: bit 0 = x, bit 1 = y, bit 2 = z; NULL = all.
 */
#ifndef ACCFFT_OPERATORS_GPU_H
#define ACCFFT_OPERATORS_GPU_H
#include <bitset>
#include "accfft_gpu.h"
#include "accfft_gpuf.h"

void accfft_grad_gpu(double* A_x, double* A_y, double* A_z, double* A, accfft_plan_gpu* plan,
                     std::bitset<3>* pXYZ = NULL, double* timer = NULL);
void accfft_laplace_gpu(double* LA, double* A, accfft_plan_gpu* plan, double* timer = NULL);
void accfft_divergence_gpu(double* divA, double* A_x, double* A_y, double* A_z, accfft_plan_gpu* plan,
                           double* timer = NULL);
void accfft_divergence_slow_gpu(double* divA, double* A_x, double* A_y, double* A_z, accfft_plan_gpu* plan,
                                double* timer = NULL);
void accfft_biharmonic_gpu(double* LA, double* A, accfft_plan_gpu* plan, double* timer = NULL);

void accfft_grad_gpuf(float* A_x, float* A_y, float* A_z, float* A, accfft_plan_gpuf* plan,
                      std::bitset<3>* pXYZ = NULL, double* timer = NULL);
void accfft_laplace_gpuf(float* LA, float* A, accfft_plan_gpuf* plan, double* timer = NULL);
void accfft_divergence_gpuf(float* divA, float* A_x, float* A_y, float* A_z, accfft_plan_gpuf* plan,
                            double* timer = NULL);
void accfft_divergence_slow_gpuf(float* divA, float* A_x, float* A_y, float* A_z, accfft_plan_gpuf* plan,
                                 double* timer = NULL);
void accfft_biharmonic_gpuf(float* LA, float* A, accfft_plan_gpuf* plan, double* timer = NULL);

/* GPU twins of the reference's CPU-only inverse operators (include/accfft_operators.h:57-60; src/operators.txx:275-392)
 * and the general (alpha + beta |k|^2)^(+-1) symbol that covers the Poisson filter of steps/step4 and the heat steps of
 * steps/step6 -- extensions: these names do not exist in the reference's GPU header. */
void accfft_inv_laplace_gpu(double* invLA, double* A, accfft_plan_gpu* plan, double* timer = NULL);
void accfft_inv_biharmonic_gpu(double* invBA, double* A, accfft_plan_gpu* plan, double* timer = NULL);
void accfft_symbol_solve_gpu(double* out, double* A, double alpha, double beta, bool inverse, accfft_plan_gpu* plan,
                             double* timer = NULL);
void accfft_inv_laplace_gpuf(float* invLA, float* A, accfft_plan_gpuf* plan, double* timer = NULL);
void accfft_inv_biharmonic_gpuf(float* invBA, float* A, accfft_plan_gpuf* plan, double* timer = NULL);
void accfft_symbol_solve_gpuf(float* out, float* A, double alpha, double beta, bool inverse, accfft_plan_gpuf* plan,
                              double* timer = NULL);

template <typename T, typename Tp>
void accfft_grad_gpu_t(T* A_x, T* A_y, T* A_z, T* A, Tp* plan, std::bitset<3>* pXYZ = NULL, double* timer = NULL);
template <typename T, typename Tp>
void accfft_grad_gpu_slow_t(T* A_x, T* A_y, T* A_z, T* A, Tp* plan, std::bitset<3>* pXYZ = NULL, double* timer = NULL);
template <typename T, typename Tp>
void accfft_laplace_gpu_t(T* LA, T* A, Tp* plan, double* timer = NULL);
template <typename T, typename Tp>
void accfft_divergence_gpu_t(T* divA, T* A_x, T* A_y, T* A_z, Tp* plan, double* timer = NULL);
template <typename T, typename Tp>
void accfft_divergence_gpu_slow_t(T* divA, T* A_x, T* A_y, T* A_z, Tp* plan, double* timer = NULL);
template <typename T, typename Tp>
void accfft_biharmonic_gpu_t(T* LA, T* A, Tp* plan, double* timer = NULL);

#endif
