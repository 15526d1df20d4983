// accfft_utils_npy.cpp -- read_npy / write_npy (include/accfft_utils.h): the `.npy` stand-in for the reference's
// read_pnetcdf / write_pnetcdf (include/accfft_utils.h:24-38 of the reference; its implementation needs PnetCDF).
// Host code only.  Part of libaccfft_gpu.so, which plays the role of libaccfft_gpu + libaccfft_utils_gpu.
#include <accfft_utils.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string>

namespace {
int g_npy_err = 0;
const size_t kHeaderBytes = 128;   // magic + version + HEADER_LEN + dict, padded with spaces, ends with '\n' (npy format 1.0)

std::string make_header(const int g[3]) {
  char dict[96];
  snprintf(dict, sizeof(dict), "{'descr': '<f8', 'fortran_order': False, 'shape': (%d, %d, %d), }", g[0], g[1], g[2]);
  std::string h("\x93NUMPY", 6);
  h += '\x01';
  h += '\x00';
  const unsigned short hlen = (unsigned short)(kHeaderBytes - 10);
  h += (char)(hlen & 0xff);
  h += (char)(hlen >> 8);
  h += dict;
  h.resize(kHeaderBytes - 1, ' ');
  h += '\n';
  return h;
}

// parses the header of an .npy file of doubles in C order; returns the data offset or 0
size_t parse_header(int fd, int shape[3]) {
  unsigned char pre[10];
  if (pread(fd, pre, 10, 0) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) return 0;
  size_t hlen, off;
  if (pre[6] == 1) {
    hlen = pre[8] | (pre[9] << 8);
    off = 10;
  } else {
    unsigned char more[2];
    if (pread(fd, more, 2, 10) != 2) return 0;
    hlen = pre[8] | (pre[9] << 8) | ((size_t)more[0] << 16) | ((size_t)more[1] << 24);
    off = 12;
  }
  std::string dict(hlen, '\0');
  if (pread(fd, &dict[0], hlen, (off_t)off) != (ssize_t)hlen) return 0;
  if (dict.find("'<f8'") == std::string::npos || dict.find("'fortran_order': False") == std::string::npos) return 0;
  const size_t sp = dict.find("'shape': (");
  if (sp == std::string::npos || sscanf(dict.c_str() + sp, "'shape': (%d, %d, %d", &shape[0], &shape[1], &shape[2]) != 3) return 0;
  return off + hlen;
}

void fail(const char* what, const std::string& file) {
  fprintf(stderr, "accfft npy i/o: %s (%s)\n", what, file.c_str());
  g_npy_err = 1;
}

// rows of counts[2] doubles of this rank's block <-> their place in the global C-order array
template <bool kWrite>
bool move_rows(int fd, size_t data_off, const MPI_Offset starts[3], const MPI_Offset counts[3], const int g[3], double* local) {
  const size_t row = (size_t)counts[2] * sizeof(double);
  for (MPI_Offset i = 0; i < counts[0]; ++i)
    for (MPI_Offset j = 0; j < counts[1]; ++j) {
      const long long at = ((long long)(starts[0] + i) * g[1] + (starts[1] + j)) * g[2] + starts[2];
      double* p = local + ((size_t)i * counts[1] + j) * counts[2];
      const off_t off = (off_t)(data_off + (size_t)at * sizeof(double));
      const ssize_t n = kWrite ? pwrite(fd, p, row, off) : pread(fd, p, row, off);
      if (n != (ssize_t)row) return false;
    }
  return true;
}

bool block_ok(const MPI_Offset s[3], const MPI_Offset c[3], const int g[3]) {
  for (int d = 0; d < 3; ++d)
    if (s[d] < 0 || c[d] < 0 || s[d] + c[d] > g[d]) return false;
  return true;
}
}  // namespace

int accfft_npy_last_error(void) { return g_npy_err; }

void write_npy(const std::string& filename, MPI_Offset starts[3], MPI_Offset counts[3], MPI_Comm c_comm, int gsizes[3],
               double* localData) {
  g_npy_err = 0;
  int rank = 0;
  MPI_Comm_rank(c_comm, &rank);
  if (!block_ok(starts, counts, gsizes)) fail("block outside the global array", filename);
  if (rank == 0) {   // header + full length first, so that every rank can pwrite into an existing file
    const int fd = open(filename.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
    const std::string h = make_header(gsizes);
    const off_t total = (off_t)(kHeaderBytes + (size_t)gsizes[0] * gsizes[1] * gsizes[2] * sizeof(double));
    if (fd < 0 || pwrite(fd, h.data(), h.size(), 0) != (ssize_t)h.size() || ftruncate(fd, total) != 0) fail("cannot create", filename);
    if (fd >= 0) close(fd);
  }
  MPI_Barrier(c_comm);
  if (!g_npy_err && counts[0] * counts[1] * counts[2] > 0) {
    const int fd = open(filename.c_str(), O_WRONLY);
    if (fd < 0 || !move_rows<true>(fd, kHeaderBytes, starts, counts, gsizes, localData)) fail("write failed", filename);
    if (fd >= 0) close(fd);
  }
  MPI_Barrier(c_comm);
}

void read_npy(const std::string& filename, MPI_Offset starts[3], MPI_Offset counts[3], MPI_Comm c_comm, int gsizes[3],
              double* localData) {
  g_npy_err = 0;
  if (!block_ok(starts, counts, gsizes)) fail("block outside the global array", filename);
  const int fd = open(filename.c_str(), O_RDONLY);
  int shape[3] = {0, 0, 0};
  const size_t off = fd >= 0 ? parse_header(fd, shape) : 0;
  if (!off) fail("not a C-order float64 .npy file", filename);
  else if (shape[0] != gsizes[0] || shape[1] != gsizes[1] || shape[2] != gsizes[2]) fail("shape differs from gsizes", filename);
  else if (!g_npy_err && counts[0] * counts[1] * counts[2] > 0 && !move_rows<false>(fd, off, starts, counts, gsizes, localData))
    fail("read failed", filename);
  if (fd >= 0) close(fd);
  MPI_Barrier(c_comm);
}
