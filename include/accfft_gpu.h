/* accfft_gpu.h -- AccFFT's double-precision GPU API (reference: include/accfft_gpu.h:37-155) served by the
 * B200-native engine.  Source compatible: same function names, argument order, default arguments and plan fields
 * that callers read.  What differs is behind the plan: no cuFFT handles and no host-staged transpose plans -- `impl`
 * is an accfftb_plan (include/accfft_b200.h) whose stage kernels exchange pencils over NVLink.
 *
 * Deliberate differences (tested): plan creation never writes to data / data_out (ACCFFT_MEASURE
 * tunes on plan-owned scratch); c2r and backward c2c leave their input intact; accfft_destroy_plan_gpu frees
 * everything including the plan object.
 */
#ifndef ACCFFT_GPU_H
#define ACCFFT_GPU_H
#include <string.h>
#include <bitset>
#include <cstddef>
#include "accfft_common.h"

struct accfftb_plan;

/* Fields shared by accfft_plan_gpu and accfft_plan_gpuf (the reference declares them twice). */
template <typename Real>
struct accfft_b200_plan_fields {
  typedef Real ComplexT[2];
  int N[3];
  int alloc_max;                       /* bytes, as returned by accfft_local_size_dft_*_gpu */
  int coord[2], np[2], periods[2];
  MPI_Comm c_comm, row_comm, col_comm;
  /* layouts after the Z pass (0), after the first exchange (1) and after the second (2 = output) */
  int osize_0[3], ostart_0[3];
  int osize_1[3], ostart_1[3];
  int osize_2[3], ostart_2[3];
  int osize_1i[3], ostart_1i[3];
  int osize_2i[3], ostart_2i[3];
  int isize[3], istart[3];
  int osize[3], ostart[3];
  /* layouts of the single-axis (y-only / x-only) real transforms */
  int osize_y[3], osize_yi[3], ostart_y[3];
  int osize_x[3], osize_xi[3], ostart_x[3];
  Real* data;
  Real* data_out;
  ComplexT* data_c;
  ComplexT* data_out_c;
  int procid;
  bool inplace;
  bool oneD;
  bool r2c_plan_baked;
  bool c2c_plan_baked;
  accfftb_plan* impl;                  /* the engine's plan */
  accfft_b200_plan_fields() { memset(this, 0, sizeof(*this)); }
};

struct accfft_plan_gpu : accfft_b200_plan_fields<double> {};

/* ---- sizes ------------------------------------------------------------------------------------------------ */
int dfft_get_local_size_gpu(int N0, int N1, int N2, int* isize, int* istart, MPI_Comm c_comm);
int accfft_local_size_dft_r2c_gpu(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm);
int accfft_local_size_dft_c2c_gpu(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm);
template <typename T>
int accfft_local_size_dft_r2c_gpu_t(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm);

/* ---- plans (collective over c_comm; device pointers; data_d == data_out_d selects the in-place layout) ------- */
accfft_plan_gpu* accfft_plan_dft_3d_r2c_gpu(int* n, double* data_d, double* data_out_d, MPI_Comm c_comm,
                                            unsigned flags = ACCFFT_MEASURE);
accfft_plan_gpu* accfft_plan_dft_3d_c2c_gpu(int* n, Complex* data_d, Complex* data_out_d, MPI_Comm c_comm,
                                            unsigned flags = ACCFFT_MEASURE);
void accfft_destroy_plan(accfft_plan_gpu* plan);
void accfft_destroy_plan_gpu(accfft_plan_gpu* plan);
void accfft_cleanup_gpu();

/* ---- execution (blocking; NULL data pointers mean "the pointers given at plan time"; timer: double[5], +=) --- */
void accfft_execute_r2c_gpu(accfft_plan_gpu* plan, double* data, Complex* data_out, double* timer = NULL,
                            std::bitset<3> xyz = 111);
void accfft_execute_c2r_gpu(accfft_plan_gpu* plan, Complex* data, double* data_out, double* timer = NULL,
                            std::bitset<3> xyz = 111);
void accfft_execute_gpu(accfft_plan_gpu* plan, int direction, double* data_d, double* data_out_d,
                        double* timer = NULL, std::bitset<3> xyz = 111);
void accfft_execute_c2c_gpu(accfft_plan_gpu* plan, int direction, Complex* data_d, Complex* data_out_d,
                            double* timer = NULL, std::bitset<3> xyz = 111);

template <typename T, typename Tc>
void accfft_execute_r2c_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer = NULL,
                              std::bitset<3> XYZ = 111);
template <typename Tc, typename T>
void accfft_execute_c2r_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer = NULL,
                              std::bitset<3> XYZ = 111);

// single-axis real transforms (reference: include/accfft_gpu.h:137-156; src/accfft_gpu.cpp:1212-1604).  Complex side layouts:
// z [isize0][isize1][N2/2+1] (osize_0), y [isize0][N1/2+1][N2 block of the row group] (osize_yi),
// x [N0/2+1][block of the flattened isize1*N2 plane over the column group] (osize_xi).
template <typename T, typename Tc>
void accfft_execute_r2c_z_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer = NULL);
template <typename Tc, typename T>
void accfft_execute_c2r_z_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer = NULL);
template <typename T, typename Tc>
void accfft_execute_r2c_y_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer = NULL);
template <typename Tc, typename T>
void accfft_execute_c2r_y_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer = NULL);
template <typename T, typename Tc>
void accfft_execute_r2c_x_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer = NULL);
template <typename Tc, typename T>
void accfft_execute_c2r_x_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer = NULL);

#endif
