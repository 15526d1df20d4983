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
  // ---- point-to-point, all-to-all and derived datatypes (what the reference's CPU transposes use) ----
  {
    // ring Sendrecv with a tag, then unordered Isend / Irecv / Waitall to everybody
    int next = (rank + 1) % size, prev = (rank + size - 1) % size, got = -1, mine = 100 + rank;
    MPI_Status st;
    MPI_Sendrecv(&mine, 1, MPI_INT, next, 7, &got, 1, MPI_INT, prev, 7, MPI_COMM_WORLD, &st);
    CHECK(got == 100 + prev && st.MPI_SOURCE == prev && st.MPI_TAG == 7);
    std::vector<double> sb(size * 3), rb(size * 3, -1.0);
    std::vector<MPI_Request> rq(2 * size);
    for (int p = 0; p < size; ++p)
      for (int k = 0; k < 3; ++k) sb[p * 3 + k] = rank * 1000 + p * 10 + k;
    for (int p = 0; p < size; ++p) MPI_Irecv(&rb[p * 3], 3, MPI_DOUBLE, p, 11, MPI_COMM_WORLD, &rq[p]);
    for (int p = size - 1; p >= 0; --p) MPI_Isend(&sb[p * 3], 3, MPI_DOUBLE, p, 11, MPI_COMM_WORLD, &rq[size + p]);
    MPI_Waitall(2 * size, rq.data(), MPI_STATUSES_IGNORE);
    for (int p = 0; p < size; ++p)
      for (int k = 0; k < 3; ++k) CHECK(rb[p * 3 + k] == p * 1000 + rank * 10 + k);
    // tag matching: two messages to the same peer, received in the opposite order
    int a = 1, b = 2, ra = 0, rb2 = 0;
    MPI_Request q2[4];
    MPI_Isend(&a, 1, MPI_INT, next, 21, MPI_COMM_WORLD, &q2[0]);
    MPI_Isend(&b, 1, MPI_INT, next, 22, MPI_COMM_WORLD, &q2[1]);
    MPI_Irecv(&rb2, 1, MPI_INT, prev, 22, MPI_COMM_WORLD, &q2[2]);
    MPI_Wait(&q2[2], MPI_STATUS_IGNORE);
    MPI_Irecv(&ra, 1, MPI_INT, prev, 21, MPI_COMM_WORLD, &q2[3]);
    MPI_Waitall(4, q2, MPI_STATUSES_IGNORE);
    CHECK(ra == 1 && rb2 == 2);
    // Alltoall with a contiguous derived type (the reference's MPI_T = sizeof(T) bytes), on the row communicator
    MPI_Datatype T8;
    MPI_Type_contiguous(8, MPI_BYTE, &T8);
    MPI_Type_commit(&T8);
    std::vector<double> as(rsize * 2), ar(rsize * 2, -1.0);
    for (int p = 0; p < rsize; ++p) as[2 * p] = rrank * 100 + p, as[2 * p + 1] = -(rrank * 100 + p);
    MPI_Alltoall(as.data(), 2, T8, ar.data(), 2, T8, row);
    for (int p = 0; p < rsize; ++p) CHECK(ar[2 * p] == p * 100 + rrank && ar[2 * p + 1] == -(p * 100 + rrank));
    // Alltoallv with uneven counts (rank r sends p+1 values to rank p), on the column communicator
    std::vector<int> sc(csize), sd(csize), rc(csize), rd(csize);
    int tot = 0;
    for (int p = 0; p < csize; ++p) sc[p] = p + 1, sd[p] = tot, tot += sc[p];
    for (int p = 0; p < csize; ++p) rc[p] = crank + 1, rd[p] = p * (crank + 1);
    std::vector<double> vs(tot), vr(csize * (crank + 1), -1.0);
    for (int p = 0; p < csize; ++p)
      for (int k = 0; k < sc[p]; ++k) vs[sd[p] + k] = crank * 10000 + p * 100 + k;
    MPI_Alltoallv(vs.data(), sc.data(), sd.data(), T8, vr.data(), rc.data(), rd.data(), T8, col);
    for (int p = 0; p < csize; ++p)
      for (int k = 0; k <= crank; ++k) CHECK(vr[rd[p] + k] == p * 10000 + crank * 100 + k);
    // vector type: send column 1 of a 4 x 3 row-major matrix, receive it contiguously
    MPI_Datatype colT;
    MPI_Type_vector(4, 1, 3, MPI_DOUBLE, &colT);
    MPI_Type_commit(&colT);
    double mat[12], colv[4] = {0, 0, 0, 0};
    for (int i = 0; i < 12; ++i) mat[i] = rank * 100 + i;
    MPI_Sendrecv(&mat[1], 1, colT, next, 31, colv, 4, MPI_DOUBLE, prev, 31, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    for (int i = 0; i < 4; ++i) CHECK(colv[i] == prev * 100 + 3 * i + 1);
    MPI_Type_free(&colT);
    MPI_Type_free(&T8);
    // a payload of several MB through the eager path
    std::vector<double> bigs(1 << 19, (double)rank), bigr(1 << 19, -1.0);
    MPI_Sendrecv(bigs.data(), 1 << 19, MPI_DOUBLE, next, 41, bigr.data(), 1 << 19, MPI_DOUBLE, prev, 41, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    CHECK(bigr[0] == prev && bigr[(1 << 19) - 1] == prev);
  }
  // communicator ids are recycled: far more split / free cycles than the shared pool has ids
  for (int it = 0; it < 600; ++it) {
    MPI_Comm t;
    CHECK(MPI_Comm_split(MPI_COMM_WORLD, rank % 2, rank, &t) == MPI_SUCCESS);
    int one = 1, cnt = 0;
    MPI_Allreduce(&one, &cnt, 1, MPI_INT, MPI_SUM, t);
    CHECK(cnt == (size + (rank % 2 == 0 ? 1 : 0)) / 2);
    MPI_Comm_free(&t);
  }
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
