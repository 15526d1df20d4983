// fftw3_shim.cpp -- TEST INFRASTRUCTURE: arithmetic behind oracle/shim/fftw3.h (see that header).
#include "fftw3.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <map>
#include <vector>

namespace {
typedef long double R;
struct C { R re, im; };

// W_n^k = exp(-2 pi i k / n), cached per n, evaluated with octant symmetry in long double
const std::vector<C>& roots(int n) {
  static std::map<int, std::vector<C>> cache;
  auto it = cache.find(n);
  if (it != cache.end()) return it->second;
  std::vector<C> w(n);
  const R two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < n; ++k) {
    // exact octant reduction in integers: theta = 2 pi k / n = q * pi/2 + b with |b| <= pi/4, so that the symmetries
    // (W^(n/4) = -i, W^(n/8) = (1 - i)/sqrt 2, ...) hold to the last bit
    const long long k8 = 8LL * k;
    const int oct = (int)(k8 / n);
    const long long rem = k8 - (long long)oct * n;   // theta = (oct + rem / n) * pi/4
    int q;
    R c, s;
    if (oct % 2 == 0) {
      q = oct / 2;
      const R b = two_pi * (R)rem / (8.0L * (R)n);
      c = cosl(b), s = sinl(b);
    } else {
      q = (oct + 1) / 2;
      const R b = two_pi * (R)(n - rem) / (8.0L * (R)n);   // theta = q * pi/2 - b
      c = cosl(b), s = -sinl(b);
    }
    R cr, sr;   // cos / sin of q * pi/2 + b
    switch (q & 3) {
      case 0: cr = c, sr = s; break;
      case 1: cr = -s, sr = c; break;
      case 2: cr = -c, sr = -s; break;
      default: cr = s, sr = -c; break;
    }
    w[k] = {cr, -sr};
  }
  return cache.emplace(n, std::move(w)).first->second;
}

int smallest_factor(int n) {
  for (int p = 2; (long long)p * p <= n; ++p)
    if (n % p == 0) return p;
  return n;
}

// out[k] = sum_j in[j * stride] W_N^{sign-adjusted j k}; `w` = roots(N0) of the top-level length, `wstep` = N0 / n
void dft_rec(int n, const C* in, int stride, C* out, const std::vector<C>& w, int wstep, int sign) {
  if (n == 1) {
    out[0] = in[0];
    return;
  }
  const int p = smallest_factor(n);
  const int N0 = (int)w.size();
  if (p == n) {   // prime length: the definition
    for (int k = 0; k < n; ++k) {
      R sr = 0, si = 0;
      for (int j = 0; j < n; ++j) {
        const C t = w[(size_t)(((long long)j * k) % n) * wstep % N0];
        const R ti = sign < 0 ? t.im : -t.im;
        const C x = in[(size_t)j * stride];
        sr += x.re * t.re - x.im * ti;
        si += x.re * ti + x.im * t.re;
      }
      out[k] = {sr, si};
    }
    return;
  }
  const int m = n / p;
  std::vector<C> sub((size_t)n);
  for (int r = 0; r < p; ++r) dft_rec(m, in + (size_t)r * stride, stride * p, sub.data() + (size_t)r * m, w, wstep * p, sign);
  for (int k = 0; k < m; ++k)
    for (int q = 0; q < p; ++q) {
      const int kk = k + m * q;
      R sr = 0, si = 0;
      for (int r = 0; r < p; ++r) {
        const C t = w[(size_t)(((long long)r * kk) % n) * wstep % N0];
        const R ti = sign < 0 ? t.im : -t.im;
        const C x = sub[(size_t)r * m + k];
        sr += x.re * t.re - x.im * ti;
        si += x.re * ti + x.im * t.re;
      }
      out[kk] = {sr, si};
    }
}

enum Kind { K_C2C, K_R2C, K_C2R };
struct Dim { long n, is, os; };
}  // namespace

struct fftw_shim_plan_s {
  bool single;             // float (fftwf_*) or double data
  Kind kind;
  int sign;
  Dim d;                   // the (single) transform dimension
  std::vector<Dim> how;    // batch dimensions
  void* in;
  void* out;
};

namespace {
fftw_plan make(bool single, Kind kind, int sign, int rank, const Dim* dims, int hrank, const Dim* hdims, void* in, void* out) {
  if (rank != 1) {
    fprintf(stderr, "fftw3 shim: only rank-1 transforms are implemented (the reference plans nothing else)\n");
    return nullptr;
  }
  fftw_plan p = new fftw_shim_plan_s();
  p->single = single;
  p->kind = kind;
  p->sign = sign;
  p->d = dims[0];
  for (int i = 0; i < hrank; ++i) p->how.push_back(hdims[i]);
  p->in = in;
  p->out = out;
  return p;
}

template <class F>
void run_one(const fftw_shim_plan_s* p, const char* in, char* out) {
  const int n = (int)p->d.n;
  const std::vector<C>& w = roots(n);
  std::vector<C> x((size_t)n), y((size_t)n);
  if (p->kind == K_C2C) {
    const F* s = (const F*)in;
    for (int j = 0; j < n; ++j) x[j] = {(R)s[2 * j * p->d.is], (R)s[2 * j * p->d.is + 1]};
    dft_rec(n, x.data(), 1, y.data(), w, 1, p->sign);
    F* d = (F*)out;
    for (int k = 0; k < n; ++k) d[2 * k * p->d.os] = (F)y[k].re, d[2 * k * p->d.os + 1] = (F)y[k].im;
  } else if (p->kind == K_R2C) {
    const F* s = (const F*)in;
    for (int j = 0; j < n; ++j) x[j] = {(R)s[j * p->d.is], 0.0L};
    dft_rec(n, x.data(), 1, y.data(), w, 1, -1);
    F* d = (F*)out;
    for (int k = 0; k <= n / 2; ++k) d[2 * k * p->d.os] = (F)y[k].re, d[2 * k * p->d.os + 1] = (F)y[k].im;
  } else {
    const F* s = (const F*)in;
    for (int k = 0; k <= n / 2; ++k) x[k] = {(R)s[2 * k * p->d.is], (R)s[2 * k * p->d.is + 1]};
    x[0].im = 0;                         // FFTW's c2r ignores these
    if (n % 2 == 0) x[n / 2].im = 0;
    for (int k = n / 2 + 1; k < n; ++k) x[k] = {x[n - k].re, -x[n - k].im};
    dft_rec(n, x.data(), 1, y.data(), w, 1, +1);
    F* d = (F*)out;
    for (int j = 0; j < n; ++j) d[j * p->d.os] = (F)y[j].re;
  }
}

void run(const fftw_shim_plan_s* p, void* in, void* out) {
  if (!p) return;
  // element sizes of the two sides (strides are in elements of that side's type)
  const size_t rs = p->single ? sizeof(float) : sizeof(double);
  const size_t ie = p->kind == K_R2C ? rs : 2 * rs;
  const size_t oe = p->kind == K_C2R ? rs : 2 * rs;
  std::vector<long> idx(p->how.size(), 0);
  for (;;) {
    long ioff = 0, ooff = 0;
    bool empty = false;
    for (size_t i = 0; i < p->how.size(); ++i) {
      if (p->how[i].n <= 0) empty = true;
      ioff += idx[i] * p->how[i].is;
      ooff += idx[i] * p->how[i].os;
    }
    if (empty || p->d.n <= 0) return;
    if (p->single) run_one<float>(p, (const char*)in + ioff * (long)ie, (char*)out + ooff * (long)oe);
    else run_one<double>(p, (const char*)in + ioff * (long)ie, (char*)out + ooff * (long)oe);
    size_t i = 0;
    for (; i < idx.size(); ++i) {
      if (++idx[i] < p->how[i].n) break;
      idx[i] = 0;
    }
    if (i == idx.size()) return;
  }
}
}  // namespace

static fftw_plan guru(bool single, Kind kind, int sign, int rank, const fftw_iodim* dims, int hrank, const fftw_iodim* hdims, void* in,
                      void* out) {
  std::vector<Dim> d, h;
  for (int i = 0; i < rank; ++i) d.push_back({dims[i].n, dims[i].is, dims[i].os});
  for (int i = 0; i < hrank; ++i) h.push_back({hdims[i].n, hdims[i].is, hdims[i].os});
  return make(single, kind, sign, rank, d.data(), hrank, h.data(), in, out);
}

extern "C" {

#define ACCFFT_FFTW_SHIM_IMPL(X, R, SINGLE)                                                                                  \
  X(plan) X(plan_many_dft)(int rank, const int* n, int howmany, X(complex)* in, const int*, int istride, int idist,          \
                           X(complex)* out, const int*, int ostride, int odist, int sign, unsigned) {                       \
    Dim d = {n[0], istride, ostride}, h = {howmany, idist, odist};                                                          \
    return make(SINGLE, K_C2C, sign, rank, &d, 1, &h, in, out);                                                             \
  }                                                                                                                         \
  X(plan) X(plan_many_dft_r2c)(int rank, const int* n, int howmany, R* in, const int*, int istride, int idist,               \
                               X(complex)* out, const int*, int ostride, int odist, unsigned) {                             \
    Dim d = {n[0], istride, ostride}, h = {howmany, idist, odist};                                                          \
    return make(SINGLE, K_R2C, -1, rank, &d, 1, &h, in, out);                                                               \
  }                                                                                                                         \
  X(plan) X(plan_many_dft_c2r)(int rank, const int* n, int howmany, X(complex)* in, const int*, int istride, int idist,      \
                               R* out, const int*, int ostride, int odist, unsigned) {                                      \
    Dim d = {n[0], istride, ostride}, h = {howmany, idist, odist};                                                          \
    return make(SINGLE, K_C2R, +1, rank, &d, 1, &h, in, out);                                                               \
  }                                                                                                                         \
  X(plan) X(plan_guru_dft)(int rank, const fftw_iodim* dims, int hrank, const fftw_iodim* hdims, X(complex)* in,             \
                           X(complex)* out, int sign, unsigned) {                                                           \
    return guru(SINGLE, K_C2C, sign, rank, dims, hrank, hdims, in, out);                                                    \
  }                                                                                                                         \
  X(plan) X(plan_guru_dft_r2c)(int rank, const fftw_iodim* dims, int hrank, const fftw_iodim* hdims, R* in,                  \
                               X(complex)* out, unsigned) {                                                                 \
    return guru(SINGLE, K_R2C, -1, rank, dims, hrank, hdims, in, out);                                                      \
  }                                                                                                                         \
  X(plan) X(plan_guru_dft_c2r)(int rank, const fftw_iodim* dims, int hrank, const fftw_iodim* hdims, X(complex)* in,         \
                               R* out, unsigned) {                                                                          \
    return guru(SINGLE, K_C2R, +1, rank, dims, hrank, hdims, in, out);                                                      \
  }                                                                                                                         \
  void X(execute)(const X(plan) p) { run(p, p->in, p->out); }                                                               \
  void X(execute_dft)(const X(plan) p, X(complex)* in, X(complex)* out) { run(p, in, out); }                                \
  void X(execute_dft_r2c)(const X(plan) p, R* in, X(complex)* out) { run(p, in, out); }                                     \
  void X(execute_dft_c2r)(const X(plan) p, X(complex)* in, R* out) { run(p, in, out); }                                     \
  void X(destroy_plan)(X(plan) p) { delete p; }                                                                             \
  void* X(malloc)(size_t n) {                                                                                               \
    void* p = nullptr;                                                                                                      \
    return posix_memalign(&p, 64, n ? n : 64) == 0 ? p : nullptr;                                                           \
  }                                                                                                                         \
  void X(free)(void* p) { free(p); }                                                                                        \
  int X(init_threads)(void) { return 1; }                                                                                   \
  void X(plan_with_nthreads)(int) {}                                                                                        \
  void X(cleanup_threads)(void) {}                                                                                          \
  void X(cleanup)(void) {}

ACCFFT_FFTW_SHIM_IMPL(ACCFFT_FFTW_D, double, false)
ACCFFT_FFTW_SHIM_IMPL(ACCFFT_FFTW_F, float, true)

}  // extern "C"
