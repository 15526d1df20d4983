// ref_driver.cpp -- TEST INFRASTRUCTURE (oracle/): drives the REFERENCE's own CPU library, compiled unmodified from
// /root/reference/src/{accfft,accfftf,accfft_common,transpose,operators,operatorsf}.cpp against the fftw3.h and mpi.h
// stand-ins (oracle/shim/fftw3.h, include/mpi_shim/mpi.h), and dumps what every rank holds:
//     isize / istart / osize / ostart / alloc_max,  the forward spectrum block,  the c2r round trip,
//     grad (3 components), laplace, divergence(grad), biharmonic, inv_laplace, inv_biharmonic        (r2c cases)
//     forward and backward c2c blocks                                                                 (c2c cases)
// i.e. the outputs of the reference's decomposition (src/accfft_common.cpp:49-192), transposes (src/transpose.cpp), execute
// drivers (src/accfft.cpp:1352-1551) and Hadamard loops (src/operators.txx) on seeded inputs.  tests/golden/make_ref_golden.py
// runs it under accfft_mpirun and packs the per-rank files into tests/golden/ref_*.npz.
//     accfft_mpirun -n P ref_driver <outdir> <case> <r2c|r2c_inplace|c2c> N0 N1 N2 P0 P1 seed [measure]
#include <mpi.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include <accfft.h>
#include <accfft_operators.h>

// the same generator as oracle.accfft_oracle.hashed_field: splitmix64 of (seed, linear global index) -> [-1, 1)
static double hashed(uint64_t seed, uint64_t idx) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ULL + idx + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

static std::string g_prefix;
static void dump(const char* name, const void* p, size_t bytes) {
  const std::string f = g_prefix + "." + name + ".bin";
  FILE* fp = fopen(f.c_str(), "wb");
  if (!fp) {
    perror(f.c_str());
    MPI_Abort(MPI_COMM_WORLD, 5);
  }
  if (bytes) fwrite(p, 1, bytes, fp);
  fclose(fp);
}

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank, nprocs;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
  if (argc < 10) {
    if (!rank) fprintf(stderr, "usage: ref_driver outdir case r2c|r2c_inplace|c2c N0 N1 N2 P0 P1 seed [measure]\n");
    MPI_Finalize();
    return 2;
  }
  const std::string outdir = argv[1], cname = argv[2], kind = argv[3];
  int n[3] = {atoi(argv[4]), atoi(argv[5]), atoi(argv[6])};
  int c_dims[2] = {atoi(argv[7]), atoi(argv[8])};
  const uint64_t seed = strtoull(argv[9], nullptr, 10);
  const unsigned flags = (argc > 10 && !strcmp(argv[10], "measure")) ? ACCFFT_MEASURE : ACCFFT_ESTIMATE;
  g_prefix = outdir + "/" + cname + ".r" + std::to_string(rank);

  MPI_Comm c_comm;
  accfft_create_comm(MPI_COMM_WORLD, c_dims, &c_comm);
  accfft_init(1);
  int isize[3], istart[3], osize[3], ostart[3];
  const bool c2c = kind == "c2c", inplace = kind == "r2c_inplace";
  const int alloc_max = c2c ? accfft_local_size_dft_c2c(n, isize, istart, osize, ostart, c_comm)
                            : accfft_local_size_dft_r2c(n, isize, istart, osize, ostart, c_comm);
  int meta[16] = {alloc_max, isize[0], isize[1], isize[2], istart[0], istart[1], istart[2], osize[0], osize[1], osize[2],
                  ostart[0], ostart[1], ostart[2], c_dims[0], c_dims[1], nprocs};
  dump("meta", meta, sizeof(meta));
  const size_t nin = (size_t)isize[0] * isize[1] * isize[2], nout = (size_t)osize[0] * osize[1] * osize[2];

  if (c2c) {
    Complex* x = (Complex*)accfft_alloc(alloc_max);
    Complex* X = (Complex*)accfft_alloc(alloc_max);
    Complex* y = (Complex*)accfft_alloc(alloc_max);
    accfft_plan* plan = accfft_plan_dft_3d_c2c(n, x, X, c_comm, flags);
    for (int i = 0; i < isize[0]; ++i)
      for (int j = 0; j < isize[1]; ++j)
        for (int k = 0; k < isize[2]; ++k) {
          const uint64_t g = ((uint64_t)(i + istart[0]) * n[1] + (j + istart[1])) * n[2] + (k + istart[2]);
          const size_t l = ((size_t)i * isize[1] + j) * isize[2] + k;
          x[l][0] = hashed(seed, 2 * g);
          x[l][1] = hashed(seed, 2 * g + 1);
        }
    dump("input", x, nin * sizeof(Complex));
    accfft_execute_c2c(plan, ACCFFT_FORWARD, x, X);
    dump("forward", X, nout * sizeof(Complex));
    accfft_execute_c2c(plan, ACCFFT_BACKWARD, X, y);
    dump("backward", y, nin * sizeof(Complex));
    accfft_destroy_plan(plan);
  } else {
    // in place: one alloc_max buffer, real rows padded to 2 * (N2/2 + 1) (steps/step2/step2.cpp)
    const int n2p = inplace ? 2 * (n[2] / 2 + 1) : n[2];
    double* x = (double*)accfft_alloc(alloc_max);
    Complex* X = inplace ? (Complex*)x : (Complex*)accfft_alloc(alloc_max);
    accfft_plan* plan = accfft_plan_dft_3d_r2c(n, x, (double*)X, c_comm, flags);
    std::vector<double> dense(nin);
    for (int i = 0; i < isize[0]; ++i)
      for (int j = 0; j < isize[1]; ++j)
        for (int k = 0; k < isize[2]; ++k) {
          const uint64_t g = ((uint64_t)(i + istart[0]) * n[1] + (j + istart[1])) * n[2] + (k + istart[2]);
          const double v = hashed(seed, g);
          x[((size_t)i * isize[1] + j) * n2p + k] = v;
          dense[((size_t)i * isize[1] + j) * isize[2] + k] = v;
        }
    dump("input", dense.data(), nin * sizeof(double));
    accfft_execute_r2c(plan, x, X);
    dump("forward", X, nout * sizeof(Complex));
    double* y = inplace ? x : (double*)accfft_alloc(alloc_max);
    accfft_execute_c2r(plan, X, y);
    for (int i = 0; i < isize[0]; ++i)
      for (int j = 0; j < isize[1]; ++j)
        for (int k = 0; k < isize[2]; ++k) dense[((size_t)i * isize[1] + j) * isize[2] + k] = y[((size_t)i * isize[1] + j) * n2p + k];
    dump("roundtrip", dense.data(), nin * sizeof(double));
    if (!inplace) {
      // operators (out of place plan, as steps/step5/step5.cpp uses them)
      for (size_t l = 0; l < nin; ++l) {
        const int k = (int)(l % isize[2]), j = (int)((l / isize[2]) % isize[1]), i = (int)(l / ((size_t)isize[2] * isize[1]));
        x[l] = hashed(seed, ((uint64_t)(i + istart[0]) * n[1] + (j + istart[1])) * n[2] + (k + istart[2]));
      }
      const size_t rb = (nin ? nin : 1) * sizeof(double);
      double* g[3] = {(double*)accfft_alloc(rb), (double*)accfft_alloc(rb), (double*)accfft_alloc(rb)};
      double* w = (double*)accfft_alloc(rb);
      std::bitset<3> XYZ;
      XYZ[0] = XYZ[1] = XYZ[2] = 1;
      double timer[8] = {0};
      accfft_grad(g[0], g[1], g[2], x, plan, &XYZ, timer);
      dump("grad_x", g[0], nin * sizeof(double));
      dump("grad_y", g[1], nin * sizeof(double));
      dump("grad_z", g[2], nin * sizeof(double));
      accfft_divergence(w, g[0], g[1], g[2], plan, timer);
      dump("div_grad", w, nin * sizeof(double));
      accfft_laplace(w, x, plan, timer);
      dump("laplace", w, nin * sizeof(double));
      accfft_biharmonic(w, x, plan, timer);
      dump("biharmonic", w, nin * sizeof(double));
      accfft_inv_laplace(w, x, plan, timer);
      dump("inv_laplace", w, nin * sizeof(double));
      accfft_inv_biharmonic(w, x, plan, timer);
      dump("inv_biharmonic", w, nin * sizeof(double));
      // the full-3-D ("slow") schemes of grad / divergence
      accfft_grad_slow_t<double, accfft_plan>(g[0], g[1], g[2], x, plan, &XYZ, timer);
      dump("grad_slow_x", g[0], nin * sizeof(double));
      dump("grad_slow_y", g[1], nin * sizeof(double));
      dump("grad_slow_z", g[2], nin * sizeof(double));
      accfft_divergence_slow_t<double, accfft_plan>(w, g[0], g[1], g[2], plan, timer);
      dump("div_slow_grad_slow", w, nin * sizeof(double));
    }
    accfft_destroy_plan(plan);
  }
  MPI_Barrier(MPI_COMM_WORLD);
  MPI_Finalize();
  return 0;
}
