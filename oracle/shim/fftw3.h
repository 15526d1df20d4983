/* fftw3.h -- TEST INFRASTRUCTURE (oracle/): a stand-in for the FFTW 3 header so that the reference's own CPU library
 * (/root/reference/src/{accfft,accfft_common,transpose,operators}.cpp) compiles UNMODIFIED in an image that has no FFTW.
 * Only the entry points those files use are declared.  The arithmetic behind them (fftw3_shim.cpp) is the unnormalised
 * DFT exactly as FFTW defines it (sign -1 forward, +1 backward; r2c keeps n/2+1 outputs; c2r ignores the imaginary
 * parts of the DC and Nyquist inputs), evaluated in long double by a plain mixed-radix recursion that falls back to
 * the O(n^2) definition for prime factors -- slow, but independent of every FFT this repo ships.  Plans carry the
 * full stride / batch geometry of fftw_plan_many_* and the guru interface, and the new-array execute functions
 * apply it to whatever pointers they are given, as FFTW's do.
 * Nothing outside tests/ and oracle/ may include or link this.
 */
#ifndef ACCFFT_ORACLE_FFTW3_SHIM_H
#define ACCFFT_ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int n, is, os; } fftw_iodim;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_UNALIGNED (1U << 1)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

#define ACCFFT_FFTW_SHIM_API(X, R)                                                                                          \
  typedef R X(complex)[2];                                                                                                \
  typedef struct fftw_shim_plan_s* X(plan);                                                                               \
  X(plan) X(plan_many_dft)(int rank, const int* n, int howmany, X(complex)* in, const int* inembed, int istride, int idist, \
                           X(complex)* out, const int* onembed, int ostride, int odist, int sign, unsigned flags);        \
  X(plan) X(plan_many_dft_r2c)(int rank, const int* n, int howmany, R* in, const int* inembed, int istride, int idist,     \
                               X(complex)* out, const int* onembed, int ostride, int odist, unsigned flags);              \
  X(plan) X(plan_many_dft_c2r)(int rank, const int* n, int howmany, X(complex)* in, const int* inembed, int istride,       \
                               int idist, R* out, const int* onembed, int ostride, int odist, unsigned flags);            \
  X(plan) X(plan_guru_dft)(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims,             \
                           X(complex)* in, X(complex)* out, int sign, unsigned flags);                                    \
  X(plan) X(plan_guru_dft_r2c)(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims, R* in,  \
                               X(complex)* out, unsigned flags);                                                          \
  X(plan) X(plan_guru_dft_c2r)(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims,         \
                               X(complex)* in, R* out, unsigned flags);                                                   \
  void X(execute)(const X(plan) p);                                                                                       \
  void X(execute_dft)(const X(plan) p, X(complex)* in, X(complex)* out);                                                  \
  void X(execute_dft_r2c)(const X(plan) p, R* in, X(complex)* out);                                                       \
  void X(execute_dft_c2r)(const X(plan) p, X(complex)* in, R* out);                                                       \
  void X(destroy_plan)(X(plan) p);                                                                                        \
  void* X(malloc)(size_t n);                                                                                              \
  void X(free)(void* p);                                                                                                  \
  int X(init_threads)(void);                                                                                              \
  void X(plan_with_nthreads)(int nthreads);                                                                               \
  void X(cleanup_threads)(void);                                                                                          \
  void X(cleanup)(void);

#define ACCFFT_FFTW_D(name) fftw_##name
#define ACCFFT_FFTW_F(name) fftwf_##name
ACCFFT_FFTW_SHIM_API(ACCFFT_FFTW_D, double)
ACCFFT_FFTW_SHIM_API(ACCFFT_FFTW_F, float)
typedef fftw_iodim fftwf_iodim;

#ifdef __cplusplus
}
#endif
#endif
