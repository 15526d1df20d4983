// decomp.hpp -- pencil decomposition arithmetic (host only, no CUDA).
// Restates, with 64-bit sizes, src/accfft_common.cpp:49-91 (dfft_get_local_size_t),
// :98-192 (accfft_local_size_dft_r2c_t), :203-256 (accfft_local_size_dft_c2c_t) and the
// MPI_Dims_create / MPI_Cart_create conventions of :275-305 of the reference.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <utility>

namespace accfft_b200 {

struct Blk {
  int size, start;
};

// ceil-block c of N over P; short or empty (size 0, start N) last blocks  (accfft_common.cpp:58-74)
inline Blk block_of(int N, int P, int c) {
  int size = (N + P - 1) / P;
  int start = size * c;
  if (N - size * c < size) {
    size = N - size * c;
    if (size < 0) size = 0;
    start = N - size;
  }
  return {size, start};
}

inline int ceil_blk(int N, int P) { return (N + P - 1) / P; }

// MPI_Dims_create(nprocs, 2, {0,0}): balanced, non-increasing  (2 -> 2x1, 4 -> 2x2, 8 -> 4x2)
inline void dims_create(int nprocs, int dims[2]) {
  int best0 = nprocs, best1 = 1;
  for (int a = 1; (long long)a * a <= nprocs; ++a)
    if (nprocs % a == 0) {
      best0 = nprocs / a;
      best1 = a;
    }
  dims[0] = best0;
  dims[1] = best1;
}

struct LocalSizes {
  int isize[3], istart[3], osize[3], ostart[3];
  int osize_0[3], ostart_0[3], osize_1[3], ostart_1[3], osize_2[3], ostart_2[3];
  int osize_y[3], ostart_y[3], osize_yi[3], osize_x[3], osize_xi[3];
  long long alloc_max;   // bytes
};

// dfft_get_local_size_t: returns alloc_local bytes
inline long long dfft_local(int N0, int N1, int tuple, const int np[2], const int coords[2], int sizeofT, int* isize,
                            int* istart) {
  Blk b0 = block_of(N0, np[0], coords[0]);
  Blk b1 = block_of(N1, np[1], coords[1]);
  isize[0] = b0.size;
  isize[1] = b1.size;
  isize[2] = tuple;
  istart[0] = b0.start;
  istart[1] = b1.start;
  istart[2] = 0;
  return (long long)b0.size * b1.size * tuple * sizeofT;
}

// sizeofT: 8 for double plans, 4 for float plans (the reference instantiates the template on the REAL type)
inline LocalSizes local_size_r2c(const int n[3], const int np[2], const int coords[2], int sizeofT) {
  LocalSizes L{};
  long long alloc_max = 0, a;
  const int n_tuples = (n[2] / 2 + 1) * 2;
  a = dfft_local(n[0], n[1], n_tuples, np, coords, sizeofT, L.osize_0, L.ostart_0);
  alloc_max = std::max(alloc_max, a);
  a = dfft_local(n[0], n_tuples / 2, n[1], np, coords, sizeofT, L.osize_1, L.ostart_1);
  alloc_max = std::max(alloc_max, a * 2);
  a = dfft_local(n[1], n_tuples / 2, n[0], np, coords, sizeofT, L.osize_2, L.ostart_2);
  alloc_max = std::max(alloc_max, a * 2);
  std::swap(L.osize_1[1], L.osize_1[2]);
  std::swap(L.ostart_1[1], L.ostart_1[2]);
  std::swap(L.ostart_2[1], L.ostart_2[2]);
  std::swap(L.ostart_2[0], L.ostart_2[1]);
  std::swap(L.osize_2[1], L.osize_2[2]);
  std::swap(L.osize_2[0], L.osize_2[1]);
  dfft_local(n[0], n[1], n[2], np, coords, sizeofT, L.isize, L.istart);
  for (int i = 0; i < 3; ++i) {
    L.osize[i] = L.osize_2[i];
    L.ostart[i] = L.ostart_2[i];
  }
  L.osize_0[2] = n_tuples / 2;   // as stored by the plan (src/accfft_gpu.cpp:127)
  // y-only path
  int tmp_s[3], tmp_st[3];
  a = dfft_local(n[0], n[2], n[1], np, coords, sizeofT, L.osize_y, L.ostart_y);
  alloc_max = std::max(alloc_max, a);
  a = dfft_local(n[0], n[2], n[1] / 2 + 1, np, coords, sizeofT, tmp_s, tmp_st);
  alloc_max = std::max(alloc_max, 2 * a);
  std::swap(L.osize_y[1], L.osize_y[2]);
  std::swap(L.ostart_y[1], L.ostart_y[2]);
  L.osize_yi[0] = L.osize_y[0];
  L.osize_yi[1] = L.osize_y[1] / 2 + 1;
  L.osize_yi[2] = L.osize_y[2];
  // x-only path: (N1/P1 * N2) split completely over P0
  dfft_local(L.isize[1] * n[2], n[2], n[0], np, coords, sizeofT, tmp_s, tmp_st);
  L.osize_x[0] = n[0];
  L.osize_x[1] = tmp_s[0];
  L.osize_x[2] = 1;
  alloc_max = std::max(alloc_max, (long long)tmp_s[0] * n[0] * sizeofT);
  dfft_local(L.isize[1] * n[2], n[2], n[0] / 2 + 1, np, coords, sizeofT, tmp_s, tmp_st);
  L.osize_xi[0] = n[0] / 2 + 1;
  L.osize_xi[1] = tmp_s[0];
  L.osize_xi[2] = 1;
  alloc_max = std::max(alloc_max, 2LL * tmp_s[0] * (n[0] / 2 + 1) * sizeofT);
  L.alloc_max = alloc_max;
  return L;
}

inline LocalSizes local_size_c2c(const int n[3], const int np[2], const int coords[2], int sizeofT) {
  LocalSizes L{};
  long long alloc_max = 0, a;
  a = dfft_local(n[0], n[1], n[2], np, coords, sizeofT, L.osize_0, L.ostart_0);
  alloc_max = std::max(alloc_max, a);
  a = dfft_local(n[0], n[2], n[1], np, coords, sizeofT, L.osize_1, L.ostart_1);
  alloc_max = std::max(alloc_max, a);
  a = dfft_local(n[1], n[2], n[0], np, coords, sizeofT, L.osize_2, L.ostart_2);
  alloc_max = std::max(alloc_max, a);
  alloc_max *= 2;
  std::swap(L.osize_1[1], L.osize_1[2]);
  std::swap(L.ostart_1[1], L.ostart_1[2]);
  std::swap(L.ostart_2[1], L.ostart_2[2]);
  std::swap(L.ostart_2[0], L.ostart_2[1]);
  std::swap(L.osize_2[1], L.osize_2[2]);
  std::swap(L.osize_2[0], L.osize_2[1]);
  dfft_local(n[0], n[1], n[2], np, coords, sizeofT, L.isize, L.istart);
  for (int i = 0; i < 3; ++i) {
    L.osize[i] = L.osize_2[i];
    L.ostart[i] = L.ostart_2[i];
  }
  L.alloc_max = alloc_max;
  return L;
}

}  // namespace accfft_b200
