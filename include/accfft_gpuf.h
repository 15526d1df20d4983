/* accfft_gpuf.h -- single-precision twin of accfft_gpu.h (reference: include/accfft_gpuf.h:36-125). */
#ifndef ACCFFT_GPUF_H
#define ACCFFT_GPUF_H
#include "accfft_gpu.h"

struct accfft_plan_gpuf : accfft_b200_plan_fields<float> {};

int dfft_get_local_size_gpuf(int N0, int N1, int N2, int* isize, int* istart, MPI_Comm c_comm);
int accfft_local_size_dft_r2c_gpuf(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm);
int accfft_local_size_dft_c2c_gpuf(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm);

accfft_plan_gpuf* accfft_plan_dft_3d_r2c_gpuf(int* n, float* data_d, float* data_out_d, MPI_Comm c_comm,
                                              unsigned flags = ACCFFT_MEASURE);
accfft_plan_gpuf* accfft_plan_dft_3d_c2c_gpuf(int* n, Complexf* data_d, Complexf* data_out_d, MPI_Comm c_comm,
                                              unsigned flags = ACCFFT_MEASURE);
void accfft_destroy_plan(accfft_plan_gpuf* plan);
void accfft_destroy_plan_gpu(accfft_plan_gpuf* plan);
void accfft_cleanup_gpuf();

void accfft_execute_r2c_gpuf(accfft_plan_gpuf* plan, float* data = NULL, Complexf* data_out = NULL,
                             double* timer = NULL, std::bitset<3> xyz = 111);
void accfft_execute_c2r_gpuf(accfft_plan_gpuf* plan, Complexf* data = NULL, float* data_out = NULL,
                             double* timer = NULL, std::bitset<3> xyz = 111);
void accfft_execute_gpuf(accfft_plan_gpuf* plan, int direction, float* data_d = NULL, float* data_out_d = NULL,
                         double* timer = NULL, std::bitset<3> xyz = 111);
void accfft_execute_c2c_gpuf(accfft_plan_gpuf* plan, int direction, Complexf* data_d = NULL,
                             Complexf* data_out_d = NULL, double* timer = NULL, std::bitset<3> xyz = 111);

template <typename T, typename Tc>
void accfft_execute_r2c_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer = NULL,
                              std::bitset<3> XYZ = 111);
template <typename Tc, typename T>
void accfft_execute_c2r_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer = NULL,
                              std::bitset<3> XYZ = 111);

// single-axis real transforms (reference: include/accfft_gpuf.h:137-156; src/accfft_gpu.cpp:1212-1604).  Complex side layouts:
// z [isize0][isize1][N2/2+1] (osize_0), y [isize0][N1/2+1][N2 block of the row group] (osize_yi),
// x [N0/2+1][block of the flattened isize1*N2 plane over the column group] (osize_xi).
template <typename T, typename Tc>
void accfft_execute_r2c_z_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer = NULL);
template <typename Tc, typename T>
void accfft_execute_c2r_z_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer = NULL);
template <typename T, typename Tc>
void accfft_execute_r2c_y_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer = NULL);
template <typename Tc, typename T>
void accfft_execute_c2r_y_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer = NULL);
template <typename T, typename Tc>
void accfft_execute_r2c_x_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer = NULL);
template <typename Tc, typename T>
void accfft_execute_c2r_x_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer = NULL);

#endif
