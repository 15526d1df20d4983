/* accfft_common.h -- AccFFT-compatible common declarations on top of libaccfft_b200 (B200-native engine).
 *
 * Same names, argument meaning and return conventions as the reference's include/accfft_common.h (:34-58), so that a
 * program written against AccFFT compiles unchanged.  Implemented in accfft_b200/csrc/accfft_gpu_api.cpp; the
 * arithmetic lives behind the C ABI of include/accfft_b200.h.
 */
#ifndef ACCFFT_COMMON_H
#define ACCFFT_COMMON_H
#include <mpi.h>
#include <omp.h>   /* the reference's headers pull it in (include/accfft_common.h:25) and its step programs rely on that */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <cstddef>
#include <iostream>

#define PCOUT if (procid == 0) std::cout

typedef double Complex[2];  /* re, im -- the FFTW / cuFFT layout */
typedef float Complexf[2];

#define ACCFFT_FORWARD -1
#define ACCFFT_BACKWARD +1
#define ACCFFT_INVERSE +1
#define ACCFFT_ESTIMATE 1
#define ACCFFT_MEASURE 2
#define ACCFFT_PATIENT 4

/* P0 x P1 Cartesian communicator.  If c_dims[0] * c_dims[1] does not match the size of in_comm, c_dims is
 * OVERWRITTEN with MPI_Dims_create's balanced choice, like the reference (src/accfft_common.cpp:275-305). */
void accfft_create_comm(MPI_Comm in_comm, int* c_dims, MPI_Comm* c_comm);
/* No-op kept for source compatibility (src/accfft_common.cpp:310-312). */
int accfft_init();
/* Host buffers for staging data in and out of the GPU.  The reference hands out fftw_malloc memory
 * (src/accfft_common.cpp:35-46); here the memory is page-locked so that cudaMemcpy runs at full PCIe speed. */
void* accfft_alloc(ptrdiff_t size);
void accfft_free(void* ptr);

/* Block decomposition of the first two axes over the process grid (src/accfft_common.cpp:49-91): ceil blocks, the
 * last one short or empty.  Returns the local byte count. */
template <typename T>
int dfft_get_local_size_t(int N0, int N1, int N2, int* isize, int* istart, MPI_Comm c_comm);
/* Local input / output blocks and the allocation size (bytes) every data_out / in-place buffer must have
 * (src/accfft_common.cpp:98-192, :203-256). */
template <typename T>
int accfft_local_size_dft_r2c_t(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm);
template <typename T>
int accfft_local_size_dft_c2c_t(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm);

#endif
