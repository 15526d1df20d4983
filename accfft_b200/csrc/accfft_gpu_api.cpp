// accfft_gpu_api.cpp -- the AccFFT-compatible C++ layer (include/accfft_common.h, accfft_gpu.h, accfft_gpuf.h,
// accfft_operators_gpu.h) over the C ABI of libaccfft_b200 (include/accfft_b200.h).
//
// Reference entry points served here (paths in the AccFFT tree):
//   accfft_create_comm / accfft_init / accfft_alloc / accfft_free      src/accfft_common.cpp:35-46, :275-312
//   accfft_local_size_dft_{r2c,c2c}_gpu[f], dfft_get_local_size_gpu[f]  src/accfft_gpu.cpp:49-65, :689-692
//   accfft_plan_dft_3d_{r2c,c2c}_gpu[f]                                  src/accfft_gpu.cpp:77-442, :704-931
//   accfft_execute_{r2c,c2r,c2c}_gpu[f], accfft_execute_gpu[f]           src/accfft_gpu.cpp:444-675, :943-1136
//   accfft_destroy_plan[_gpu], accfft_cleanup_gpu[f]                     src/accfft_gpu.cpp:1142-1192
//   accfft_{grad,laplace,divergence,biharmonic}_gpu[f] (+ _slow, _t)     src/operators_gpu.cpp:85-135
// MPI is used for what it is used for in the reference's plan set-up -- the Cartesian grid -- plus ONE
// MPI_Allgather of a 128-byte CUDA-IPC blob per rank at plan time; no data ever travels through MPI.
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>

#include "../../include/accfft_b200.h"
#include "../../include/accfft_operators_gpu.h"

namespace {

struct Grid {
  int dims[2], coords[2], rank, nprocs;
};

Grid grid_of(MPI_Comm c_comm) {
  Grid g;
  int periods[2];
  MPI_Cart_get(c_comm, 2, g.dims, periods, g.coords);
  MPI_Comm_size(c_comm, &g.nprocs);
  g.rank = g.coords[0] * g.dims[1] + g.coords[1];   // row-major Cartesian rank (src/accfft_common.cpp:298)
  return g;
}

template <typename Real>
int local_size(bool r2c, int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  const Grid g = grid_of(c_comm);
  const long long a = r2c ? accfftb_local_size_dft_r2c(n, g.dims, g.coords, (int)sizeof(Real), isize, istart, osize, ostart)
                          : accfftb_local_size_dft_c2c(n, g.dims, g.coords, (int)sizeof(Real), isize, istart, osize, ostart);
  if (a < 0) return -1;   // local block too large: the reference's overflow guard (src/accfft_common.cpp:179-184)
  if (a > 2147483647LL) {
    int procid;
    MPI_Comm_rank(c_comm, &procid);
    if (procid == 0)
      fprintf(stderr, "accfft: alloc_max does not fit the reference's int return type; use accfftb_local_size_dft_* or more ranks\n");
    return -1;
  }
  return (int)a;
}

template <typename Plan, typename Real>
Plan* make_plan(bool r2c, int* n, void* data, void* data_out, MPI_Comm c_comm, unsigned flags) {
  const Grid g = grid_of(c_comm);
  Plan* p = new Plan();
  for (int i = 0; i < 3; ++i) p->N[i] = n[i];
  p->c_comm = c_comm;
  p->np[0] = g.dims[0];
  p->np[1] = g.dims[1];
  p->coord[0] = g.coords[0];
  p->coord[1] = g.coords[1];
  MPI_Comm_rank(c_comm, &p->procid);
  MPI_Comm_split(c_comm, g.coords[0], g.coords[1], &p->row_comm);   // ranks sharing coord 0: first exchange
  MPI_Comm_split(c_comm, g.coords[1], g.coords[0], &p->col_comm);   // ranks sharing coord 1: second exchange
  const int prec = sizeof(Real) == 8 ? ACCFFTB_DOUBLE : ACCFFTB_FLOAT;
  const int kind = r2c ? ACCFFTB_R2C : ACCFFTB_C2C;
  p->alloc_max = local_size<Real>(r2c, n, p->isize, p->istart, p->osize, p->ostart, c_comm);
  int* sizes[7] = {p->osize_0, p->osize_1, p->osize_2, p->osize_y, p->osize_yi, p->osize_x, p->osize_xi};
  int* starts[7] = {p->ostart_0, p->ostart_1, p->ostart_2, p->ostart_y, NULL, NULL, NULL};
  for (int w = 0; w < 7; ++w) accfftb_stage_size(n, g.dims, g.coords, kind, (int)sizeof(Real), w, sizes[w], starts[w]);
  for (int i = 0; i < 3; ++i) {   // the inverse path walks the same layouts backwards
    p->osize_1i[i] = p->osize_1[i];
    p->ostart_1i[i] = p->ostart_1[i];
    p->osize_2i[i] = p->osize_2[i];
    p->ostart_2i[i] = p->ostart_2[i];
  }
  p->data = (Real*)data;
  p->data_out = (Real*)data_out;
  p->data_c = (typename Plan::ComplexT*)data;
  p->data_out_c = (typename Plan::ComplexT*)data_out;
  p->inplace = data == data_out;
  p->oneD = g.dims[0] == 1 || g.dims[1] == 1;
  p->impl = accfftb_plan_create(n, g.dims, g.rank, kind, prec, flags);
  // every rank must learn whether ALL ranks got a plan before anybody enters the collective wiring
  int ok = p->impl ? 1 : 0, all_ok = 0;
  MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, c_comm);
  if (all_ok && g.nprocs > 1) {
    std::vector<char> blobs((size_t)ACCFFTB_IPC_BLOB_BYTES * g.nprocs);
    char mine[ACCFFTB_IPC_BLOB_BYTES];
    ok = accfftb_plan_ipc_export(p->impl, mine) == 0;
    MPI_Allgather(mine, ACCFFTB_IPC_BLOB_BYTES, MPI_BYTE, blobs.data(), ACCFFTB_IPC_BLOB_BYTES, MPI_BYTE, c_comm);
    if (ok) ok = accfftb_plan_ipc_import(p->impl, blobs.data(), g.nprocs) == 0;
    MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, c_comm);
  }
  if (!all_ok) {
    // the reference prints and returns NULL when FFT plan creation fails (src/accfft_gpu.cpp:187-191)
    if (p->procid == 0) fprintf(stderr, "accfft: plan creation failed: %s\n", accfftb_last_error());
    accfftb_plan_destroy(p->impl);
    MPI_Comm_free(&p->row_comm);
    MPI_Comm_free(&p->col_comm);
    delete p;
    return NULL;
  }
  if (r2c) p->r2c_plan_baked = true;
  else p->c2c_plan_baked = true;
  return p;
}

template <typename Plan>
void destroy(Plan* p) {
  if (!p) return;
  cudaDeviceSynchronize();
  if (p->np[0] * p->np[1] > 1) MPI_Barrier(p->c_comm);   // nobody unmaps a slab a peer may still be writing
  accfftb_plan_destroy(p->impl);
  MPI_Comm_free(&p->row_comm);
  MPI_Comm_free(&p->col_comm);
  delete p;   // (the reference leaks the plan object, src/accfft_gpu.cpp:1150-1192)
}

void report(int procid, const char* what, int rc) {
  if (rc) fprintf(stderr, "accfft[%d]: %s failed (rc = %d): %s\n", procid, what, rc, accfftb_last_error());
}

template <typename Plan>
bool baked(Plan* p, bool want_r2c, const char* what) {
  if (p && (want_r2c ? p->r2c_plan_baked : p->c2c_plan_baked)) return true;
  // same wording idea as the reference (src/accfft_gpu.cpp:641-646): rank 0 reports, the call returns
  if (!p || p->procid == 0)
    fprintf(stderr, "accfft: %s needs a %s plan; nothing was executed\n", what, want_r2c ? "r2c" : "c2c");
  return false;
}

int bits(std::bitset<3> xyz) { return (int)(xyz.to_ulong() & 7ul); }

template <typename Plan, typename Real, typename Cplx>
void exec_r2c(Plan* p, Real* data, Cplx* out, double* timer, std::bitset<3> xyz) {
  if (!baked(p, true, "accfft_execute_r2c_gpu")) return;
  if (!data) data = p->data;
  if (!out) out = (Cplx*)p->data_out;
  report(p->procid, "accfft_execute_r2c_gpu", accfftb_execute_r2c(p->impl, data, out, timer, bits(xyz)));
}
template <typename Plan, typename Real, typename Cplx>
void exec_c2r(Plan* p, Cplx* data, Real* out, double* timer, std::bitset<3> xyz) {
  if (!baked(p, true, "accfft_execute_c2r_gpu")) return;
  if (!data) data = (Cplx*)p->data_out;   // the spectrum lives in data_out (src/accfft_gpu.cpp:664-668)
  if (!out) out = p->data;
  report(p->procid, "accfft_execute_c2r_gpu", accfftb_execute_c2r(p->impl, data, out, timer, bits(xyz)));
}
template <typename Plan, typename Cplx>
void exec_c2c(Plan* p, int direction, Cplx* data, Cplx* out, double* timer, std::bitset<3> xyz) {
  if (!baked(p, false, "accfft_execute_c2c_gpu")) return;
  if (!data) data = (Cplx*)p->data_c;
  if (!out) out = (Cplx*)p->data_out_c;
  report(p->procid, "accfft_execute_c2c_gpu", accfftb_execute_c2c(p->impl, direction, data, out, timer, bits(xyz)));
}

// single-axis real transforms: accfft_execute_{r2c,c2r}_{x,y,z}_gpu_t (src/accfft_gpu.cpp:1212-1604)
template <typename Plan, typename Real, typename Cplx>
void exec_r2c_axis(Plan* p, int axis, Real* data, Cplx* out, double* timer) {
  if (!baked(p, true, "accfft_execute_r2c_{x,y,z}_gpu_t")) return;
  if (!data) data = p->data;
  if (!out) out = (Cplx*)p->data_out;
  report(p->procid, "accfft_execute_r2c_{x,y,z}_gpu_t", accfftb_execute_r2c_axis(p->impl, axis, data, out, timer));
}
template <typename Plan, typename Real, typename Cplx>
void exec_c2r_axis(Plan* p, int axis, Cplx* data, Real* out, double* timer) {
  if (!baked(p, true, "accfft_execute_c2r_{x,y,z}_gpu_t")) return;
  if (!data) data = (Cplx*)p->data_out;
  if (!out) out = p->data;
  report(p->procid, "accfft_execute_c2r_{x,y,z}_gpu_t", accfftb_execute_c2r_axis(p->impl, axis, data, out, timer));
}

}  // namespace

// ---- common ---------------------------------------------------------------------------------------------------
void accfft_create_comm(MPI_Comm in_comm, int* c_dims, MPI_Comm* c_comm) {
  int nprocs, procid;
  MPI_Comm_rank(in_comm, &procid);
  MPI_Comm_size(in_comm, &nprocs);
  if (accfftb_create_comm_dims(nprocs, c_dims) && procid == 0)
    printf("accfft: c_dims did not match the communicator size; switched to %d x %d\n", c_dims[0], c_dims[1]);
  int period[2] = {0, 0};
  MPI_Cart_create(in_comm, 2, c_dims, period, 0, c_comm);
}
int accfft_init() { return 0; }
void* accfft_alloc(ptrdiff_t size) {
  void* p = NULL;
  if (size <= 0 || cudaMallocHost(&p, (size_t)size) != cudaSuccess) {
    cudaGetLastError();
    return NULL;
  }
  return p;
}
void accfft_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

template <typename T>
int dfft_get_local_size_t(int N0, int N1, int N2, int* isize, int* istart, MPI_Comm c_comm) {
  int n[3] = {N0, N1, N2}, osize[3], ostart[3];
  const Grid g = grid_of(c_comm);
  accfftb_local_size_dft_c2c(n, g.dims, g.coords, 8, isize, istart, osize, ostart);   // only the block shapes are used
  return isize[0] * isize[1] * isize[2] * (int)sizeof(T);
}
template int dfft_get_local_size_t<double>(int, int, int, int*, int*, MPI_Comm);
template int dfft_get_local_size_t<float>(int, int, int, int*, int*, MPI_Comm);
template int dfft_get_local_size_t<Complex>(int, int, int, int*, int*, MPI_Comm);
template int dfft_get_local_size_t<Complexf>(int, int, int, int*, int*, MPI_Comm);

template <typename T>
int accfft_local_size_dft_r2c_t(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  return local_size<T>(true, n, isize, istart, osize, ostart, c_comm);
}
template <typename T>
int accfft_local_size_dft_c2c_t(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  return local_size<T>(false, n, isize, istart, osize, ostart, c_comm);
}
template int accfft_local_size_dft_r2c_t<double>(int*, int*, int*, int*, int*, MPI_Comm);
template int accfft_local_size_dft_r2c_t<float>(int*, int*, int*, int*, int*, MPI_Comm);
template int accfft_local_size_dft_c2c_t<double>(int*, int*, int*, int*, int*, MPI_Comm);
template int accfft_local_size_dft_c2c_t<float>(int*, int*, int*, int*, int*, MPI_Comm);

// ---- double precision -----------------------------------------------------------------------------------------
int dfft_get_local_size_gpu(int N0, int N1, int N2, int* isize, int* istart, MPI_Comm c_comm) {
  return dfft_get_local_size_t<double>(N0, N1, N2, isize, istart, c_comm);
}
int accfft_local_size_dft_r2c_gpu(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  return local_size<double>(true, n, isize, istart, osize, ostart, c_comm);
}
int accfft_local_size_dft_c2c_gpu(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  return local_size<double>(false, n, isize, istart, osize, ostart, c_comm);
}
template <typename T>
int accfft_local_size_dft_r2c_gpu_t(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  return local_size<T>(true, n, isize, istart, osize, ostart, c_comm);
}
template int accfft_local_size_dft_r2c_gpu_t<double>(int*, int*, int*, int*, int*, MPI_Comm);
template int accfft_local_size_dft_r2c_gpu_t<float>(int*, int*, int*, int*, int*, MPI_Comm);

accfft_plan_gpu* accfft_plan_dft_3d_r2c_gpu(int* n, double* data_d, double* data_out_d, MPI_Comm c_comm, unsigned flags) {
  return make_plan<accfft_plan_gpu, double>(true, n, data_d, data_out_d, c_comm, flags);
}
accfft_plan_gpu* accfft_plan_dft_3d_c2c_gpu(int* n, Complex* data_d, Complex* data_out_d, MPI_Comm c_comm, unsigned flags) {
  return make_plan<accfft_plan_gpu, double>(false, n, data_d, data_out_d, c_comm, flags);
}
void accfft_destroy_plan(accfft_plan_gpu* plan) { destroy(plan); }
void accfft_destroy_plan_gpu(accfft_plan_gpu* plan) { destroy(plan); }
void accfft_cleanup_gpu() {}

void accfft_execute_r2c_gpu(accfft_plan_gpu* plan, double* data, Complex* data_out, double* timer, std::bitset<3> xyz) {
  exec_r2c(plan, data, data_out, timer, xyz);
}
void accfft_execute_c2r_gpu(accfft_plan_gpu* plan, Complex* data, double* data_out, double* timer, std::bitset<3> xyz) {
  exec_c2r(plan, data, data_out, timer, xyz);
}
void accfft_execute_gpu(accfft_plan_gpu* plan, int direction, double* data_d, double* data_out_d, double* timer,
                        std::bitset<3> xyz) {
  if (direction == ACCFFT_FORWARD) exec_r2c(plan, data_d, (Complex*)data_out_d, timer, xyz);
  else exec_c2r(plan, (Complex*)data_d, data_out_d, timer, xyz);
}
void accfft_execute_c2c_gpu(accfft_plan_gpu* plan, int direction, Complex* data_d, Complex* data_out_d, double* timer,
                            std::bitset<3> xyz) {
  exec_c2c(plan, direction, data_d, data_out_d, timer, xyz);
}
template <typename T, typename Tc>
void accfft_execute_r2c_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer, std::bitset<3> XYZ) {
  exec_r2c(plan, data, data_out, timer, XYZ);
}
template <typename Tc, typename T>
void accfft_execute_c2r_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer, std::bitset<3> XYZ) {
  exec_c2r(plan, data, data_out, timer, XYZ);
}
template void accfft_execute_r2c_gpu_t<double, Complex>(accfft_plan_gpu*, double*, Complex*, double*, std::bitset<3>);
template void accfft_execute_c2r_gpu_t<Complex, double>(accfft_plan_gpu*, Complex*, double*, double*, std::bitset<3>);
template <typename T, typename Tc>
void accfft_execute_r2c_z_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer) {
  exec_r2c_axis(plan, 2, data, data_out, timer);
}
template <typename Tc, typename T>
void accfft_execute_c2r_z_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer) {
  exec_c2r_axis(plan, 2, data, data_out, timer);
}
template void accfft_execute_r2c_z_gpu_t<double, Complex>(accfft_plan_gpu*, double*, Complex*, double*);
template void accfft_execute_c2r_z_gpu_t<Complex, double>(accfft_plan_gpu*, Complex*, double*, double*);
template <typename T, typename Tc>
void accfft_execute_r2c_y_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer) {
  exec_r2c_axis(plan, 1, data, data_out, timer);
}
template <typename Tc, typename T>
void accfft_execute_c2r_y_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer) {
  exec_c2r_axis(plan, 1, data, data_out, timer);
}
template void accfft_execute_r2c_y_gpu_t<double, Complex>(accfft_plan_gpu*, double*, Complex*, double*);
template void accfft_execute_c2r_y_gpu_t<Complex, double>(accfft_plan_gpu*, Complex*, double*, double*);
template <typename T, typename Tc>
void accfft_execute_r2c_x_gpu_t(accfft_plan_gpu* plan, T* data, Tc* data_out, double* timer) {
  exec_r2c_axis(plan, 0, data, data_out, timer);
}
template <typename Tc, typename T>
void accfft_execute_c2r_x_gpu_t(accfft_plan_gpu* plan, Tc* data, T* data_out, double* timer) {
  exec_c2r_axis(plan, 0, data, data_out, timer);
}
template void accfft_execute_r2c_x_gpu_t<double, Complex>(accfft_plan_gpu*, double*, Complex*, double*);
template void accfft_execute_c2r_x_gpu_t<Complex, double>(accfft_plan_gpu*, Complex*, double*, double*);

// ---- single precision -----------------------------------------------------------------------------------------
int dfft_get_local_size_gpuf(int N0, int N1, int N2, int* isize, int* istart, MPI_Comm c_comm) {
  return dfft_get_local_size_t<float>(N0, N1, N2, isize, istart, c_comm);
}
int accfft_local_size_dft_r2c_gpuf(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  return local_size<float>(true, n, isize, istart, osize, ostart, c_comm);
}
int accfft_local_size_dft_c2c_gpuf(int* n, int* isize, int* istart, int* osize, int* ostart, MPI_Comm c_comm) {
  return local_size<float>(false, n, isize, istart, osize, ostart, c_comm);
}
accfft_plan_gpuf* accfft_plan_dft_3d_r2c_gpuf(int* n, float* data_d, float* data_out_d, MPI_Comm c_comm, unsigned flags) {
  return make_plan<accfft_plan_gpuf, float>(true, n, data_d, data_out_d, c_comm, flags);
}
accfft_plan_gpuf* accfft_plan_dft_3d_c2c_gpuf(int* n, Complexf* data_d, Complexf* data_out_d, MPI_Comm c_comm, unsigned flags) {
  return make_plan<accfft_plan_gpuf, float>(false, n, data_d, data_out_d, c_comm, flags);
}
void accfft_destroy_plan(accfft_plan_gpuf* plan) { destroy(plan); }
void accfft_destroy_plan_gpu(accfft_plan_gpuf* plan) { destroy(plan); }
void accfft_cleanup_gpuf() {}

void accfft_execute_r2c_gpuf(accfft_plan_gpuf* plan, float* data, Complexf* data_out, double* timer, std::bitset<3> xyz) {
  exec_r2c(plan, data, data_out, timer, xyz);
}
void accfft_execute_c2r_gpuf(accfft_plan_gpuf* plan, Complexf* data, float* data_out, double* timer, std::bitset<3> xyz) {
  exec_c2r(plan, data, data_out, timer, xyz);
}
void accfft_execute_gpuf(accfft_plan_gpuf* plan, int direction, float* data_d, float* data_out_d, double* timer,
                         std::bitset<3> xyz) {
  if (direction == ACCFFT_FORWARD) exec_r2c(plan, data_d, (Complexf*)data_out_d, timer, xyz);
  else exec_c2r(plan, (Complexf*)data_d, data_out_d, timer, xyz);
}
void accfft_execute_c2c_gpuf(accfft_plan_gpuf* plan, int direction, Complexf* data_d, Complexf* data_out_d, double* timer,
                             std::bitset<3> xyz) {
  exec_c2c(plan, direction, data_d, data_out_d, timer, xyz);
}
template <typename T, typename Tc>
void accfft_execute_r2c_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer, std::bitset<3> XYZ) {
  exec_r2c(plan, data, data_out, timer, XYZ);
}
template <typename Tc, typename T>
void accfft_execute_c2r_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer, std::bitset<3> XYZ) {
  exec_c2r(plan, data, data_out, timer, XYZ);
}
template void accfft_execute_r2c_gpu_t<float, Complexf>(accfft_plan_gpuf*, float*, Complexf*, double*, std::bitset<3>);
template void accfft_execute_c2r_gpu_t<Complexf, float>(accfft_plan_gpuf*, Complexf*, float*, double*, std::bitset<3>);
template <typename T, typename Tc>
void accfft_execute_r2c_z_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer) {
  exec_r2c_axis(plan, 2, data, data_out, timer);
}
template <typename Tc, typename T>
void accfft_execute_c2r_z_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer) {
  exec_c2r_axis(plan, 2, data, data_out, timer);
}
template void accfft_execute_r2c_z_gpu_t<float, Complexf>(accfft_plan_gpuf*, float*, Complexf*, double*);
template void accfft_execute_c2r_z_gpu_t<Complexf, float>(accfft_plan_gpuf*, Complexf*, float*, double*);
template <typename T, typename Tc>
void accfft_execute_r2c_y_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer) {
  exec_r2c_axis(plan, 1, data, data_out, timer);
}
template <typename Tc, typename T>
void accfft_execute_c2r_y_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer) {
  exec_c2r_axis(plan, 1, data, data_out, timer);
}
template void accfft_execute_r2c_y_gpu_t<float, Complexf>(accfft_plan_gpuf*, float*, Complexf*, double*);
template void accfft_execute_c2r_y_gpu_t<Complexf, float>(accfft_plan_gpuf*, Complexf*, float*, double*);
template <typename T, typename Tc>
void accfft_execute_r2c_x_gpu_t(accfft_plan_gpuf* plan, T* data, Tc* data_out, double* timer) {
  exec_r2c_axis(plan, 0, data, data_out, timer);
}
template <typename Tc, typename T>
void accfft_execute_c2r_x_gpu_t(accfft_plan_gpuf* plan, Tc* data, T* data_out, double* timer) {
  exec_c2r_axis(plan, 0, data, data_out, timer);
}
template void accfft_execute_r2c_x_gpu_t<float, Complexf>(accfft_plan_gpuf*, float*, Complexf*, double*);
template void accfft_execute_c2r_x_gpu_t<Complexf, float>(accfft_plan_gpuf*, Complexf*, float*, double*);

// ---- spectral operators ------------------------------------------------------------------------------------------
// The reference offers a "fast" (per-axis transforms, src/operators_gpu.txx:172-278) and a "slow" (full 3-D transforms,
// :72-170) variant of grad and divergence with identical results; both exist here (accfftb_grad / accfftb_grad_slow).
template <typename T, typename Tp>
void accfft_grad_gpu_t(T* A_x, T* A_y, T* A_z, T* A, Tp* plan, std::bitset<3>* pXYZ, double* timer) {
  if (!baked(plan, true, "accfft_grad_gpu")) return;
  const int comps = pXYZ ? bits(*pXYZ) : 7;
  report(plan->procid, "accfft_grad_gpu", accfftb_grad(plan->impl, A, A_x, A_y, A_z, comps, timer));
}
template <typename T, typename Tp>
void accfft_grad_gpu_slow_t(T* A_x, T* A_y, T* A_z, T* A, Tp* plan, std::bitset<3>* pXYZ, double* timer) {
  if (!baked(plan, true, "accfft_grad_gpu_slow")) return;
  const int comps = pXYZ ? bits(*pXYZ) : 7;
  report(plan->procid, "accfft_grad_gpu_slow", accfftb_grad_slow(plan->impl, A, A_x, A_y, A_z, comps, timer));
}
template <typename T, typename Tp>
void accfft_laplace_gpu_t(T* LA, T* A, Tp* plan, double* timer) {
  if (!baked(plan, true, "accfft_laplace_gpu")) return;
  report(plan->procid, "accfft_laplace_gpu", accfftb_laplace(plan->impl, A, LA, timer));
}
template <typename T, typename Tp>
void accfft_divergence_gpu_t(T* divA, T* A_x, T* A_y, T* A_z, Tp* plan, double* timer) {
  if (!baked(plan, true, "accfft_divergence_gpu")) return;
  report(plan->procid, "accfft_divergence_gpu", accfftb_divergence(plan->impl, A_x, A_y, A_z, divA, timer));
}
template <typename T, typename Tp>
void accfft_divergence_gpu_slow_t(T* divA, T* A_x, T* A_y, T* A_z, Tp* plan, double* timer) {
  if (!baked(plan, true, "accfft_divergence_slow_gpu")) return;
  report(plan->procid, "accfft_divergence_slow_gpu", accfftb_divergence_slow(plan->impl, A_x, A_y, A_z, divA, timer));
}
template <typename T, typename Tp>
void accfft_biharmonic_gpu_t(T* LA, T* A, Tp* plan, double* timer) {
  if (!baked(plan, true, "accfft_biharmonic_gpu")) return;
  report(plan->procid, "accfft_biharmonic_gpu", accfftb_biharmonic(plan->impl, A, LA, timer));
}
#define ACCFFT_B200_INSTANTIATE_OPS(T, Tp)                                                        \
  template void accfft_grad_gpu_t<T, Tp>(T*, T*, T*, T*, Tp*, std::bitset<3>*, double*);          \
  template void accfft_grad_gpu_slow_t<T, Tp>(T*, T*, T*, T*, Tp*, std::bitset<3>*, double*);     \
  template void accfft_laplace_gpu_t<T, Tp>(T*, T*, Tp*, double*);                                \
  template void accfft_divergence_gpu_t<T, Tp>(T*, T*, T*, T*, Tp*, double*);                     \
  template void accfft_divergence_gpu_slow_t<T, Tp>(T*, T*, T*, T*, Tp*, double*);                \
  template void accfft_biharmonic_gpu_t<T, Tp>(T*, T*, Tp*, double*);
ACCFFT_B200_INSTANTIATE_OPS(double, accfft_plan_gpu)
ACCFFT_B200_INSTANTIATE_OPS(float, accfft_plan_gpuf)

void accfft_grad_gpu(double* A_x, double* A_y, double* A_z, double* A, accfft_plan_gpu* plan, std::bitset<3>* pXYZ,
                     double* timer) {
  accfft_grad_gpu_t(A_x, A_y, A_z, A, plan, pXYZ, timer);
}
void accfft_laplace_gpu(double* LA, double* A, accfft_plan_gpu* plan, double* timer) { accfft_laplace_gpu_t(LA, A, plan, timer); }
void accfft_divergence_gpu(double* divA, double* A_x, double* A_y, double* A_z, accfft_plan_gpu* plan, double* timer) {
  accfft_divergence_gpu_t(divA, A_x, A_y, A_z, plan, timer);
}
void accfft_divergence_slow_gpu(double* divA, double* A_x, double* A_y, double* A_z, accfft_plan_gpu* plan, double* timer) {
  accfft_divergence_gpu_slow_t(divA, A_x, A_y, A_z, plan, timer);   // the full 3-D scheme (src/operators_gpu.txx:336-446)
}
void accfft_biharmonic_gpu(double* LA, double* A, accfft_plan_gpu* plan, double* timer) {
  accfft_biharmonic_gpu_t(LA, A, plan, timer);
}
void accfft_grad_gpuf(float* A_x, float* A_y, float* A_z, float* A, accfft_plan_gpuf* plan, std::bitset<3>* pXYZ,
                      double* timer) {
  accfft_grad_gpu_t(A_x, A_y, A_z, A, plan, pXYZ, timer);
}
void accfft_laplace_gpuf(float* LA, float* A, accfft_plan_gpuf* plan, double* timer) { accfft_laplace_gpu_t(LA, A, plan, timer); }
void accfft_divergence_gpuf(float* divA, float* A_x, float* A_y, float* A_z, accfft_plan_gpuf* plan, double* timer) {
  accfft_divergence_gpu_t(divA, A_x, A_y, A_z, plan, timer);
}
void accfft_divergence_slow_gpuf(float* divA, float* A_x, float* A_y, float* A_z, accfft_plan_gpuf* plan, double* timer) {
  accfft_divergence_gpu_slow_t(divA, A_x, A_y, A_z, plan, timer);
}
void accfft_biharmonic_gpuf(float* LA, float* A, accfft_plan_gpuf* plan, double* timer) {
  accfft_biharmonic_gpu_t(LA, A, plan, timer);
}

// ---- solver-style operators (extensions; GPU twins of accfft_inv_laplace / accfft_inv_biharmonic) ------------------------
namespace {
template <typename T, typename Tp>
void symbol_t(const char* what, int kind, T* out, T* A, double alpha, double beta, bool inverse, Tp* plan, double* timer) {
  if (!baked(plan, true, what)) return;
  int rc = kind == 0   ? accfftb_inv_laplace(plan->impl, A, out, timer)
           : kind == 1 ? accfftb_inv_biharmonic(plan->impl, A, out, timer)
                       : accfftb_symbol_solve(plan->impl, A, out, alpha, beta, inverse ? 1 : 0, timer);
  report(plan->procid, what, rc);
}
}  // namespace
void accfft_inv_laplace_gpu(double* invLA, double* A, accfft_plan_gpu* plan, double* timer) {
  symbol_t("accfft_inv_laplace_gpu", 0, invLA, A, 0, 0, false, plan, timer);
}
void accfft_inv_biharmonic_gpu(double* invBA, double* A, accfft_plan_gpu* plan, double* timer) {
  symbol_t("accfft_inv_biharmonic_gpu", 1, invBA, A, 0, 0, false, plan, timer);
}
void accfft_symbol_solve_gpu(double* out, double* A, double alpha, double beta, bool inverse, accfft_plan_gpu* plan, double* timer) {
  symbol_t("accfft_symbol_solve_gpu", 2, out, A, alpha, beta, inverse, plan, timer);
}
void accfft_inv_laplace_gpuf(float* invLA, float* A, accfft_plan_gpuf* plan, double* timer) {
  symbol_t("accfft_inv_laplace_gpuf", 0, invLA, A, 0, 0, false, plan, timer);
}
void accfft_inv_biharmonic_gpuf(float* invBA, float* A, accfft_plan_gpuf* plan, double* timer) {
  symbol_t("accfft_inv_biharmonic_gpuf", 1, invBA, A, 0, 0, false, plan, timer);
}
void accfft_symbol_solve_gpuf(float* out, float* A, double alpha, double beta, bool inverse, accfft_plan_gpuf* plan, double* timer) {
  symbol_t("accfft_symbol_solve_gpuf", 2, out, A, alpha, beta, inverse, plan, timer);
}
