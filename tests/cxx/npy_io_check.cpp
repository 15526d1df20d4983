// npy_io_check.cpp -- write_npy / read_npy of include/accfft_utils.h (the .npy stand-in for the reference's
// read_pnetcdf / write_pnetcdf) with the pencil decomposition of the transform path (CPU only):
//     accfft_mpirun -n 4 --gpus-per-node 0 ./npy_io_check out.npy 12 10 14
// Every rank writes the input block [isize] of a 2-D process grid, filled with the global linear index; then reads the
// block of ANOTHER decomposition (its output-layout pencil [N0][osize1][osize2]) back and checks it.  The pytest side
// loads the file with numpy and compares it with arange(N0*N1*N2).
#include <accfft_gpu.h>
#include <accfft_utils.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank, size;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (argc < 5) {
    if (rank == 0) printf("usage: npy_io_check file N0 N1 N2\n");
    MPI_Finalize();
    return 2;
  }
  int n[3] = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4])};
  int c_dims[2] = {0, 0};
  MPI_Comm c_comm;
  accfft_create_comm(MPI_COMM_WORLD, c_dims, &c_comm);
  int isize[3], istart[3], osize[3], ostart[3];
  accfft_local_size_dft_r2c_gpu(n, isize, istart, osize, ostart, c_comm);
  std::vector<double> mine((size_t)isize[0] * isize[1] * isize[2] + 1);
  for (int i = 0; i < isize[0]; ++i)
    for (int j = 0; j < isize[1]; ++j)
      for (int k = 0; k < isize[2]; ++k)
        mine[((size_t)i * isize[1] + j) * isize[2] + k] = (double)((((long long)(i + istart[0])) * n[1] + (j + istart[1])) * n[2] + k + istart[2]);
  MPI_Offset st[3] = {istart[0], istart[1], istart[2]}, ct[3] = {isize[0], isize[1], isize[2]};
  write_npy(argv[1], st, ct, c_comm, n, mine.data());
  int fail = accfft_npy_last_error();
  // read back with a different decomposition: all of N0, block of N1 over P0, block of N2 over P1
  const int b1 = (n[1] + c_dims[0] - 1) / c_dims[0], b2 = (n[2] + c_dims[1] - 1) / c_dims[1];
  int coords[2] = {rank / c_dims[1], rank % c_dims[1]};
  MPI_Offset rs[3] = {0, (MPI_Offset)b1 * coords[0], (MPI_Offset)b2 * coords[1]};
  MPI_Offset rc[3] = {n[0], b1, b2};
  for (int d = 1; d < 3; ++d) {
    if (rs[d] > n[d]) rs[d] = n[d];
    if (rs[d] + rc[d] > n[d]) rc[d] = n[d] - rs[d];
  }
  std::vector<double> back((size_t)rc[0] * rc[1] * rc[2] + 1, -1.0);
  read_npy(argv[1], rs, rc, c_comm, n, back.data());
  fail |= accfft_npy_last_error();
  for (MPI_Offset i = 0; i < rc[0]; ++i)
    for (MPI_Offset j = 0; j < rc[1]; ++j)
      for (MPI_Offset k = 0; k < rc[2]; ++k) {
        const double want = (double)(((long long)(i + rs[0]) * n[1] + (j + rs[1])) * n[2] + k + rs[2]);
        if (back[((size_t)i * rc[1] + j) * rc[2] + k] != want) fail = 1;
      }
  // a wrong shape must be refused, not read
  int wrong[3] = {n[0], n[1], n[2] + 1};
  MPI_Offset zs[3] = {0, 0, 0}, zc[3] = {0, 0, 0};
  read_npy(argv[1], zs, zc, c_comm, wrong, back.data());
  if (!accfft_npy_last_error()) fail = 1;
  int all = 0;
  MPI_Allreduce(&fail, &all, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  if (rank == 0) printf(all ? "npy_io_check: FAILED\n" : "npy_io_check: OK (%d x %d process grid)\n", c_dims[0], c_dims[1]);
  MPI_Comm_free(&c_comm);
  MPI_Finalize();
  return all;
}
