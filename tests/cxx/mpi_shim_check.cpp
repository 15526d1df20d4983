// mpi_shim_check.cpp -- exercises the calls of include/mpi_shim/mpi.h the AccFFT layer relies on (CPU only):
//     accfft_mpirun -n 4 --gpus-per-node 0 ./mpi_shim_check
#include <mpi.h>
#include <stdio.h>
#include <vector>

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      printf("rank %d: check failed at line %d: %s\n", rank, __LINE__, #cond); \
      fail = 1;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank, size, fail = 0;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  int dims[2] = {0, 0};
  MPI_Dims_create(size, 2, dims);
  CHECK(dims[0] * dims[1] == size && dims[0] >= dims[1]);
  int periods[2] = {0, 0}, coords[2], d2[2], p2[2];
  MPI_Comm cart;
  CHECK(MPI_Cart_create(MPI_COMM_WORLD, 2, dims, periods, 0, &cart) == MPI_SUCCESS);
  MPI_Cart_get(cart, 2, d2, p2, coords);
  CHECK(d2[0] == dims[0] && d2[1] == dims[1] && coords[0] == rank / dims[1] && coords[1] == rank % dims[1]);
  MPI_Comm row, col;
  MPI_Comm_split(cart, coords[0], coords[1], &row);
  MPI_Comm_split(cart, coords[1], coords[0], &col);
  int rrank, rsize, crank, csize;
  MPI_Comm_rank(row, &rrank);
  MPI_Comm_size(row, &rsize);
  MPI_Comm_rank(col, &crank);
  MPI_Comm_size(col, &csize);
  CHECK(rsize == dims[1] && rrank == coords[1] && csize == dims[0] && crank == coords[0]);
  for (int it = 0; it < 50; ++it) {   // repeated collectives on overlapping communicators
    double v = rank + 1.0, s = 0, m = 0;
    MPI_Allreduce(&v, &s, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    CHECK(s == size * (size + 1) / 2.0);
    MPI_Allreduce(&v, &m, 1, MPI_DOUBLE, MPI_MAX, row);
    CHECK(m == coords[0] * dims[1] + dims[1]);
    int token = rank == 1 % size ? 4242 + it : -1;
    MPI_Bcast(&token, 1, MPI_INT, 1 % size, MPI_COMM_WORLD);
    CHECK(token == 4242 + it);
    long long mine = 1000LL * it + crank, got = -1;
    MPI_Reduce(&mine, &got, 1, MPI_LONG_LONG, MPI_MIN, 0, col);
    if (crank == 0) CHECK(got == 1000LL * it);
  }
  // payload larger than one 64 KB slot
  std::vector<double> big(20000, (double)rank), all((size_t)20000 * size);
  MPI_Allgather(big.data(), 20000, MPI_DOUBLE, all.data(), 20000, MPI_DOUBLE, MPI_COMM_WORLD);
  for (int r = 0; r < size; ++r) CHECK(all[(size_t)r * 20000] == r && all[(size_t)r * 20000 + 19999] == r);
  MPI_Barrier(cart);
  MPI_Comm_free(&row);
  MPI_Comm_free(&col);
  MPI_Comm_free(&cart);
  int anyfail = 0;
  MPI_Allreduce(&fail, &anyfail, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
  if (rank == 0) printf(anyfail ? "mpi_shim_check: FAILED\n" : "mpi_shim_check: OK (%d ranks, %d x %d)\n", size, dims[0], dims[1]);
  MPI_Finalize();
  return anyfail;
}
