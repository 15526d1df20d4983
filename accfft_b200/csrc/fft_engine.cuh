// fft_engine.cuh -- register-resident mixed-radix Stockham engine (compile-time N).
//
// Replaces the cuFFT plans of the reference (src/accfft_gpu.cpp:173-323: fplan_0/1/2, iplan_0, ...).
// Everything in this header is __host__ __device__ so that tests/emul can run the *same* index math
// under a CPU fiber emulator (test infrastructure only; the product has no CPU path).
//
// Algorithm: Stockham autosort, pass p with radix R_p and Ns_p = prod_{q<p} R_q:
//     butterfly j in [0, N/R):  v[r] = x[j + r*N/R] * W_N^{ r * (j % Ns) * N/(Ns*R) }
//                               v    = DFT_R(v)
//                               y[(j/Ns)*Ns*R + (j%Ns) + r*Ns] = v[r]
// A column (one 1-D transform) is worked on by G threads; thread t owns butterflies t, t+G, ...
// The first pass reads its operands straight from global memory and the last pass writes straight
// to global memory (both are "thread-consecutive, register-strided by N/R" => coalesced);
// only the exchanges between passes go through shared memory.
#pragma once
#include <stdint.h>
#include <type_traits>

#if defined(__CUDACC__)
#define ACC_HD __host__ __device__ __forceinline__
#else
#define ACC_HD inline __attribute__((always_inline))
#endif

// out-of-line device functions (the spectral-operator multiply): ONE copy per translation unit instead of one inlined
// copy per transform element and kernel -- the stage kernels shrank ~5x when apply_op left their bodies
#if defined(__CUDACC__)
#define ACC_HD_NOINLINE static __host__ __device__ __noinline__   // (static: one private copy per translation unit)
#else
#define ACC_HD_NOINLINE inline
#endif

#if defined(__CUDA_ARCH__)
#define ACC_UNROLL _Pragma("unroll")
#else
#define ACC_UNROLL
#endif

namespace accfft_b200 {

template <class T>
struct alignas(2 * sizeof(T)) cplx {
  T x, y;
};

// Two single-precision lanes handled as ONE scalar: cplx<f2> = {re0, re1, im0, im1} is a pair of complex floats (two
// adjacent columns of a strided stage), 16 bytes like a complex double.  Every butterfly below is written with + - * on its
// scalar type only, so the whole engine runs on cplx<f2> unchanged: the single-precision strided kernels move 128-bit
// quantities through registers and shared memory exactly like the double-precision ones (stage.cuh: strided_c2c_pair_body).
struct f2 {
  float a, b;
  f2() = default;
  ACC_HD f2(float s) : a(s), b(s) {}
  ACC_HD f2(double s) : a((float)s), b((float)s) {}
  ACC_HD f2(int s) : a((float)s), b((float)s) {}
  ACC_HD f2(float a_, float b_) : a(a_), b(b_) {}
};
ACC_HD f2 operator+(f2 u, f2 v) { return {u.a + v.a, u.b + v.b}; }
ACC_HD f2 operator-(f2 u, f2 v) { return {u.a - v.a, u.b - v.b}; }
ACC_HD f2 operator*(f2 u, f2 v) { return {u.a * v.a, u.b * v.b}; }
ACC_HD f2 operator*(f2 u, float s) { return {u.a * s, u.b * s}; }
ACC_HD f2 operator*(float s, f2 u) { return {u.a * s, u.b * s}; }
ACC_HD f2 operator-(f2 u) { return {-u.a, -u.b}; }

// scalar type of the twiddle factors that go with a value type
template <class T> struct twiddle_of { using type = T; };
template <> struct twiddle_of<f2> { using type = float; };

template <class T> ACC_HD cplx<T> cadd(cplx<T> a, cplx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <class T> ACC_HD cplx<T> csub(cplx<T> a, cplx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <class T> ACC_HD cplx<T> cmul(cplx<T> a, cplx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
// value times twiddle (the twiddle may be a plain float while the value is a lane pair)
template <class T, class TW> ACC_HD cplx<T> cmul_tw(cplx<T> a, cplx<TW> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <class T> ACC_HD cplx<T> cconj(cplx<T> a) { return {a.x, -a.y}; }
template <class T> ACC_HD cplx<T> cswap(cplx<T> a) { return {a.y, a.x}; }
template <class T> ACC_HD cplx<T> cmul_mi(cplx<T> a) { return {a.y, -a.x}; }   // a * (-i)
template <class T> ACC_HD cplx<T> cmul_pi(cplx<T> a) { return {-a.y, a.x}; }   // a * (+i)
template <class T> ACC_HD cplx<T> cscale(cplx<T> a, T s) { return {a.x * s, a.y * s}; }

// ---- streaming global accesses ---------------------------------------------------------------------------------
// Every stage reads its input once and writes its output once; nothing is re-read from L2 by the same kernel.
// ACCFFT_CACHE_HINTS bit 0: loads with evict-first (ld.global.cs), bit 1: stores with evict-first (st.global.cs).
#ifndef ACCFFT_CACHE_HINTS
#define ACCFFT_CACHE_HINTS 0
#endif
template <class T>
ACC_HD cplx<T> ld_stream(const cplx<T>* p) {
#if defined(__CUDA_ARCH__) && (ACCFFT_CACHE_HINTS & 1)
  if constexpr (sizeof(T) == 8) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return {v.x, v.y};
  } else {
    const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
    return {v.x, v.y};
  }
#else
  return *p;
#endif
}
template <class T>
ACC_HD void st_stream(cplx<T>* p, cplx<T> v) {
#if defined(__CUDA_ARCH__) && (ACCFFT_CACHE_HINTS & 2)
  if constexpr (sizeof(T) == 8) __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
  else __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
#else
  *p = v;
#endif
}

// ---- compile-time roots of unity W_R^k = exp(-2 pi i k / R) for the in-register composites ----
// (literal constants, 21 significant digits)
// generic accessor through a constexpr switch so device code folds it into immediates
constexpr double kC8 = 0.707106781186547524401;      // cos(pi/4)
constexpr double kC16 = 0.923879532511286756128;     // cos(pi/8)
constexpr double kS16 = 0.382683432365089771728;     // sin(pi/8)
constexpr double kS3 = 0.866025403784438646764;      // sin(pi/3)
constexpr double kC5a = 0.309016994374947424102;     // cos(2pi/5)
constexpr double kS5a = 0.951056516295153572116;     // sin(2pi/5)
constexpr double kC5b = -0.809016994374947424102;    // cos(4pi/5)
constexpr double kS5b = 0.587785252292473129169;     // sin(4pi/5)

// generic compile-time cos/sin(2 pi k / R): exact octant reduction in integers, then a Taylor series on
// |alpha| <= pi/4 (converges to < 1 ulp in 12 terms).  Only used for the composites that have no literal above.
constexpr double kPi4 = 0.785398163397448309615660845819875721;
constexpr double taylor_cos(double a) {
  double a2 = a * a, term = 1.0, sum = 1.0;
  for (int i = 1; i <= 12; ++i) {
    term *= -a2 / ((2.0 * i - 1.0) * (2.0 * i));
    sum += term;
  }
  return sum;
}
constexpr double taylor_sin(double a) {
  double a2 = a * a, term = a, sum = a;
  for (int i = 1; i <= 12; ++i) {
    term *= -a2 / ((2.0 * i) * (2.0 * i + 1.0));
    sum += term;
  }
  return sum;
}
// returns cos (want_sin = false) or sin (want_sin = true) of theta = 2 pi k / R
constexpr double generic_trig(int R, int k, bool want_sin) {
  k = ((k % R) + R) % R;
  const int o = (8 * k) / R;            // octant
  const int rem = 8 * k - o * R;        // theta = (o + rem/R) * pi/4
  int q = 0;
  double alpha = 0.0;
  if (o % 2 == 0) {
    q = o / 2;
    alpha = kPi4 * (double)rem / (double)R;
  } else {
    q = (o + 1) / 2;
    alpha = -kPi4 * (double)(R - rem) / (double)R;
  }
  const double c = taylor_cos(alpha), sn = taylor_sin(alpha);
  q %= 4;                               // theta = q * pi/2 + alpha
  const double ct = q == 0 ? c : q == 1 ? -sn : q == 2 ? -c : sn;
  const double st = q == 0 ? sn : q == 1 ? c : q == 2 ? -sn : -c;
  return want_sin ? st : ct;
}

constexpr double root_re(int R, int k) {
  // cos(2 pi k / R) for the (R,k) pairs the composites need
  k %= R;
  if (k == 0) return 1.0;
  if (2 * k == R) return -1.0;
  if (4 * k == R || 4 * k == 3 * R) return 0.0;
  if (8 * k == R || 8 * k == 7 * R) return kC8;
  if (8 * k == 3 * R || 8 * k == 5 * R) return -kC8;
  if (16 * k == R || 16 * k == 15 * R) return kC16;
  if (16 * k == 7 * R || 16 * k == 9 * R) return -kC16;
  if (16 * k == 3 * R || 16 * k == 13 * R) return kS16;
  if (16 * k == 5 * R || 16 * k == 11 * R) return -kS16;
  if (6 * k == R || 6 * k == 5 * R) return 0.5;
  if (3 * k == R || 3 * k == 2 * R) return -0.5;
  if (12 * k == R || 12 * k == 11 * R) return kS3;
  if (12 * k == 5 * R || 12 * k == 7 * R) return -kS3;
  return generic_trig(R, k, false);
}
constexpr double root_im(int R, int k) {
  // -sin(2 pi k / R)
  k %= R;
  if (k == 0 || 2 * k == R) return 0.0;
  if (4 * k == R) return -1.0;
  if (4 * k == 3 * R) return 1.0;
  if (8 * k == R || 8 * k == 3 * R) return -kC8;
  if (8 * k == 5 * R || 8 * k == 7 * R) return kC8;
  if (16 * k == R || 16 * k == 7 * R) return -kS16;
  if (16 * k == 9 * R || 16 * k == 15 * R) return kS16;
  if (16 * k == 3 * R || 16 * k == 5 * R) return -kC16;
  if (16 * k == 11 * R || 16 * k == 13 * R) return kC16;
  if (6 * k == R || 3 * k == R) return -kS3;
  if (3 * k == 2 * R || 6 * k == 5 * R) return kS3;
  if (12 * k == R || 12 * k == 5 * R) return -0.5;
  if (12 * k == 7 * R || 12 * k == 11 * R) return 0.5;
  return -generic_trig(R, k, true);
}

// multiply by the compile-time root W_R^K, using the cheap forms for multiples of 1/8 turn
template <int R, int K, class T>
ACC_HD cplx<T> mul_root(cplx<T> a) {
  constexpr int k = ((K % R) + R) % R;
  if constexpr (k == 0) return a;
  else if constexpr (2 * k == R) return {-a.x, -a.y};
  else if constexpr (4 * k == R) return cmul_mi(a);
  else if constexpr (4 * k == 3 * R) return cmul_pi(a);
  else if constexpr (8 * k == R) return {(a.x + a.y) * T(kC8), (a.y - a.x) * T(kC8)};          // (1-i)/sqrt2
  else if constexpr (8 * k == 3 * R) return {(a.y - a.x) * T(kC8), -(a.x + a.y) * T(kC8)};     // (-1-i)/sqrt2
  else if constexpr (8 * k == 5 * R) return {-(a.x + a.y) * T(kC8), (a.x - a.y) * T(kC8)};     // (-1+i)/sqrt2
  else if constexpr (8 * k == 7 * R) return {(a.x - a.y) * T(kC8), (a.x + a.y) * T(kC8)};      // (1+i)/sqrt2
  else {
    constexpr double wr = root_re(R, k), wi = root_im(R, k);
    return {a.x * T(wr) - a.y * T(wi), a.x * T(wi) + a.y * T(wr)};
  }
}

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, class F>
ACC_HD void static_for_idx(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for_idx<B + 1, E>(f);
  }
}

// ---- small forward DFTs (sign -1), in place on v[0], v[S], v[2S], ... --------------------------
template <int R, int S, class T> struct Dft;

template <int S, class T> struct Dft<1, S, T> {
  static ACC_HD void run(cplx<T>*) {}
};
template <int S, class T> struct Dft<2, S, T> {
  static ACC_HD void run(cplx<T>* v) {
    cplx<T> a = v[0], b = v[S];
    v[0] = cadd(a, b);
    v[S] = csub(a, b);
  }
};
template <int S, class T> struct Dft<3, S, T> {
  static ACC_HD void run(cplx<T>* v) {
    cplx<T> a0 = v[0], a1 = v[S], a2 = v[2 * S];
    cplx<T> t1 = cadd(a1, a2);
    cplx<T> t2 = {a0.x - T(0.5) * t1.x, a0.y - T(0.5) * t1.y};
    cplx<T> t3 = cscale(csub(a1, a2), T(kS3));
    v[0] = cadd(a0, t1);
    v[S] = {t2.x + t3.y, t2.y - t3.x};
    v[2 * S] = {t2.x - t3.y, t2.y + t3.x};
  }
};
template <int S, class T> struct Dft<4, S, T> {
  static ACC_HD void run(cplx<T>* v) {
    cplx<T> a0 = v[0], a1 = v[S], a2 = v[2 * S], a3 = v[3 * S];
    cplx<T> t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
    v[0] = cadd(t0, t2);
    v[S] = cadd(t1, t3);
    v[2 * S] = csub(t0, t2);
    v[3 * S] = csub(t1, t3);
  }
};
template <int S, class T> struct Dft<5, S, T> {
  static ACC_HD void run(cplx<T>* v) {
    cplx<T> a0 = v[0], a1 = v[S], a2 = v[2 * S], a3 = v[3 * S], a4 = v[4 * S];
    cplx<T> s1 = cadd(a1, a4), d1 = csub(a1, a4), s2 = cadd(a2, a3), d2 = csub(a2, a3);
    v[0] = {a0.x + s1.x + s2.x, a0.y + s1.y + s2.y};
    cplx<T> p1 = {a0.x + T(kC5a) * s1.x + T(kC5b) * s2.x, a0.y + T(kC5a) * s1.y + T(kC5b) * s2.y};
    cplx<T> p2 = {a0.x + T(kC5b) * s1.x + T(kC5a) * s2.x, a0.y + T(kC5b) * s1.y + T(kC5a) * s2.y};
    // q = -i * (sin terms)
    cplx<T> q1 = {T(kS5a) * d1.x + T(kS5b) * d2.x, T(kS5a) * d1.y + T(kS5b) * d2.y};
    cplx<T> q2 = {T(kS5b) * d1.x - T(kS5a) * d2.x, T(kS5b) * d1.y - T(kS5a) * d2.y};
    v[S] = {p1.x + q1.y, p1.y - q1.x};
    v[4 * S] = {p1.x - q1.y, p1.y + q1.x};
    v[2 * S] = {p2.x + q2.y, p2.y - q2.x};
    v[3 * S] = {p2.x - q2.y, p2.y + q2.x};
  }
};

// generic odd prime P (7, 11, 13): pair the inputs j and P - j,
//   X[k] , X[P-k] = (x0 + sum_j cos(2 pi j k / P) (x_j + x_(P-j)))  -+  i (sum_j sin(2 pi j k / P) (x_j - x_(P-j))),   j = 1 .. (P-1)/2
// with compile-time cosines / sines: (P-1)^2 / 2 real FMAs per component instead of a full P x P complex product.
template <int P, int S, class T>
struct DftPrime {
  static constexpr int H = (P - 1) / 2;
  static ACC_HD void run(cplx<T>* v) {
    const cplx<T> a0 = v[0];
    cplx<T> s[H], d[H];
    cplx<T> sum = a0;
    static_for_idx<1, H + 1>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      s[j - 1] = cadd(v[j * S], v[(P - j) * S]);
      d[j - 1] = csub(v[j * S], v[(P - j) * S]);
      sum = cadd(sum, s[j - 1]);
    });
    v[0] = sum;
    static_for_idx<1, H + 1>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      cplx<T> a = a0, b = {T(0), T(0)};
      static_for_idx<1, H + 1>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr double c = root_re(P, j * k), sn = -root_im(P, j * k);
        a = {a.x + T(c) * s[j - 1].x, a.y + T(c) * s[j - 1].y};
        b = {b.x + T(sn) * d[j - 1].x, b.y + T(sn) * d[j - 1].y};
      });
      v[k * S] = {a.x + b.y, a.y - b.x};          // a - i b
      v[(P - k) * S] = {a.x - b.y, a.y + b.x};    // a + i b
    });
  }
};
template <int S, class T> struct Dft<7, S, T> : DftPrime<7, S, T> {};
template <int S, class T> struct Dft<11, S, T> : DftPrime<11, S, T> {};
template <int S, class T> struct Dft<13, S, T> : DftPrime<13, S, T> {};

// Cooley-Tukey composite R = R1*R2 entirely in registers:
//   n = R2*n1 + n2, k = k1 + R1*k2 :  X[k1 + R1 k2] = sum_n2 W_R2^{n2 k2} W_R^{n2 k1} (sum_n1 x[R2 n1 + n2] W_R1^{n1 k1})
// Output is written back in natural order (needs a register permutation, free after unrolling).
template <int R1, int R2, int S, class T>
struct DftCT {
  static constexpr int R = R1 * R2;
  template <int N2, int K1> static ACC_HD void tw_row(cplx<T>* v) {
    if constexpr (K1 < R1) {
      v[(R2 * K1 + N2) * S] = mul_root<R, N2 * K1, T>(v[(R2 * K1 + N2) * S]);
      tw_row<N2, K1 + 1>(v);
    }
  }
  template <int N2> static ACC_HD void inner(cplx<T>* v) {
    if constexpr (N2 < R2) {
      // DFT_R1 over n1 of x[R2 n1 + n2]  (stride R2*S), result A[n2][k1] stored at position R2*k1 + n2
      Dft<R1, R2 * S, T>::run(v + N2 * S);
      tw_row<N2, 0>(v);
      inner<N2 + 1>(v);
    }
  }
  template <int K1> static ACC_HD void outer(cplx<T>* v) {
    if constexpr (K1 < R1) {
      // DFT_R2 over n2 of A[n2][k1] (positions R2*k1 + n2, stride S) -> X[k1 + R1 k2] left at R2*k1 + k2
      Dft<R2, S, T>::run(v + R2 * K1 * S);
      outer<K1 + 1>(v);
    }
  }
  static ACC_HD void run(cplx<T>* v) {
    inner<0>(v);
    outer<0>(v);
    // permute: position R2*k1 + k2 holds X[k1 + R1*k2]
    cplx<T> tmp[R];
    ACC_UNROLL
    for (int k1 = 0; k1 < R1; ++k1)
      ACC_UNROLL
      for (int k2 = 0; k2 < R2; ++k2) tmp[k1 + R1 * k2] = v[(R2 * k1 + k2) * S];
    ACC_UNROLL
    for (int i = 0; i < R; ++i) v[i * S] = tmp[i];
  }
};
template <int S, class T> struct Dft<6, S, T> : DftCT<3, 2, S, T> {};
template <int S, class T> struct Dft<8, S, T> : DftCT<4, 2, S, T> {};
template <int S, class T> struct Dft<12, S, T> : DftCT<3, 4, S, T> {};
template <int S, class T> struct Dft<16, S, T> : DftCT<4, 4, S, T> {};
template <int S, class T> struct Dft<9, S, T> : DftCT<3, 3, S, T> {};
template <int S, class T> struct Dft<10, S, T> : DftCT<5, 2, S, T> {};
template <int S, class T> struct Dft<15, S, T> : DftCT<5, 3, S, T> {};
template <int S, class T> struct Dft<14, S, T> : DftCT<7, 2, S, T> {};

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, class F>
ACC_HD void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

// ---- schedule ------------------------------------------------------------------------------------
template <int N_, int G_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1>
struct Sched {
  static constexpr int N = N_;
  static constexpr int G = G_;   // threads per column
  static constexpr int R[4] = {R0_, R1_, R2_, R3_};
  static constexpr int NP = (R1_ == 1) ? 1 : (R2_ == 1) ? 2 : (R3_ == 1) ? 3 : 4;
  static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
  static constexpr int radix(int p) { return R[p]; }
  static constexpr int ns(int p) { return p == 0 ? 1 : p == 1 ? R0_ : p == 2 ? R0_ * R1_ : R0_ * R1_ * R2_; }
  static constexpr int nb(int p) { return N_ / R[p]; }
  static constexpr int cnt(int p) { return (nb(p) + G_ - 1) / G_; }
  static constexpr int regs(int p) { return cnt(p) * R[p]; }
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int NREG = cmax(cmax(regs(0), NP > 1 ? regs(1) : 0), cmax(NP > 2 ? regs(2) : 0, NP > 3 ? regs(3) : 0));
  // base twiddles a thread keeps in registers: one per (pass >= 1, butterfly slot u)
  static constexpr int tw_off(int p) { return p <= 1 ? 0 : p == 2 ? cnt(1) : cnt(1) + cnt(2); }
  static constexpr int NTW = NP == 1 ? 1 : NP == 2 ? cnt(1) : NP == 3 ? cnt(1) + cnt(2) : cnt(1) + cnt(2) + cnt(3);
  // shared-memory index padding: one extra slot every 8 (kills the stride-R conflicts of the
  // early scatter passes for 16-byte and 8-byte elements alike)
  static ACC_HD int phys(int i) { return i + (i >> 3); }
  static constexpr int SMEM_ELEMS = N_ + (N_ >> 3) + 1;
};

// ---- per-pass building blocks ------------------------------------------------------------------------
// Twiddles.  Butterfly j of pass P multiplies operand r by W^(r*q), q = (j % NS) * N/(NS*R).  Only the base
// W^q is fetched from the table tw (W_M[k] = exp(-2 pi i k / M), M = S::N * TWS; TWS = 2 lets the r2c kernels
// share the length-2M table of their split step) -- once per thread, up front, so that the fetch overlaps the
// data loads -- and the powers are built in registers:  W^r = W^(r/2) * W^(r - r/2)  (depth <= 4 for R <= 16).
template <class S, int P, int TWS, class T>
ACC_HD void load_base_twiddles_pass(cplx<T>* w1, int t, const cplx<T>* __restrict__ tw) {
  if constexpr (P < S::NP) {
    constexpr int R = S::radix(P), NS = S::ns(P), NB = S::nb(P), CNT = S::cnt(P);
    constexpr int TSTEP = (S::N / (NS * R)) * TWS;
    ACC_UNROLL
    for (int u = 0; u < CNT; ++u) {
      const int j = t + u * S::G;
      w1[S::tw_off(P) + u] = (CNT * S::G == NB || j < NB) ? tw[(j % NS) * TSTEP] : cplx<T>{T(1), T(0)};
    }
  }
}
template <class S, int TWS, class T>
ACC_HD void load_base_twiddles(cplx<T>* w1, int t, const cplx<T>* __restrict__ tw) {
  load_base_twiddles_pass<S, 1, TWS>(w1, t, tw);
  load_base_twiddles_pass<S, 2, TWS>(w1, t, tw);
  load_base_twiddles_pass<S, 3, TWS>(w1, t, tw);
}

template <int R, class T, class TW>
ACC_HD void apply_twiddle_powers(cplx<T>* v, cplx<TW> w1) {
  cplx<TW> w[R > 2 ? R : 2];
  w[1] = w1;
  v[1] = cmul_tw(v[1], w1);
  ACC_UNROLL
  for (int r = 2; r < R; ++r) {
    const int h = r >> 1;
    if (r == 2 * h) w[r] = {w[h].x * w[h].x - w[h].y * w[h].y, TW(2) * w[h].x * w[h].y};
    else w[r] = cmul(w[h], w[r - h]);
    v[r] = cmul_tw(v[r], w[r]);
  }
}

template <class S, int P, class T, class TW>
ACC_HD void pass_compute(cplx<T>* v, int t, const cplx<TW>* w1) {
  constexpr int R = S::radix(P), NS = S::ns(P), NB = S::nb(P), CNT = S::cnt(P);
  ACC_UNROLL
  for (int u = 0; u < CNT; ++u) {
    const int j = t + u * S::G;
    if (CNT * S::G == NB || j < NB) {
      if constexpr (NS > 1) apply_twiddle_powers<R, T, TW>(v + u * R, w1[S::tw_off(P) + u]);
      Dft<R, 1, T>::run(v + u * R);
    }
  }
}

// scatter pass-P results into a column's shared-memory image; addr(i) maps logical index -> pointer
template <class S, int P, class T, class Addr>
ACC_HD void pass_scatter(const cplx<T>* v, int t, Addr addr) {
  constexpr int R = S::radix(P), NS = S::ns(P), NB = S::nb(P), CNT = S::cnt(P);
  ACC_UNROLL
  for (int u = 0; u < CNT; ++u) {
    const int j = t + u * S::G;
    if (CNT * S::G == NB || j < NB) {
      const int j0 = (j / NS) * NS * R + (j % NS);
      ACC_UNROLL
      for (int r = 0; r < R; ++r) *addr(j0 + r * NS) = v[u * R + r];
    }
  }
}

// gather pass-P operands:  v[u*R + r] = x[j + r*N/R]
template <class S, int P, class T, class Addr>
ACC_HD void pass_gather(cplx<T>* v, int t, Addr addr) {
  constexpr int R = S::radix(P), NB = S::nb(P), CNT = S::cnt(P);
  ACC_UNROLL
  for (int u = 0; u < CNT; ++u) {
    const int j = t + u * S::G;
    if (CNT * S::G == NB || j < NB) {
      ACC_UNROLL
      for (int r = 0; r < R; ++r) v[u * R + r] = *addr(j + r * NB);
    }
  }
}

// visit the (logical index, register) pairs of pass P in gather order:  f(index, v[slot])
template <class S, int P, class T, class F>
ACC_HD void pass_foreach(cplx<T>* v, int t, F f) {
  constexpr int R = S::radix(P), NB = S::nb(P), CNT = S::cnt(P);
  ACC_UNROLL
  for (int u = 0; u < CNT; ++u) {
    const int j = t + u * S::G;
    if (CNT * S::G == NB || j < NB) {
      ACC_UNROLL
      for (int r = 0; r < R; ++r) f(j + r * NB, v[u * R + r]);
    }
  }
}

// Run passes 0..NP-1 on registers already holding pass-0 operands (gather order); on return the
// registers hold the final outputs in pass-(NP-1) gather order (index j + r*N/R_last).
// w1: base twiddles from load_base_twiddles.  smem_addr(i) -> pointer to this column's slot i;
// sync() is a CTA-wide barrier.
template <class S, class T, class TW, class Addr, class Sync>
ACC_HD void fft_passes(cplx<T>* v, int t, const cplx<TW>* w1, Addr smem_addr, Sync sync) {
  pass_compute<S, 0>(v, t, w1);
  if constexpr (S::NP > 1) {
    pass_scatter<S, 0>(v, t, smem_addr);
    sync();
    pass_gather<S, 1>(v, t, smem_addr);
    pass_compute<S, 1>(v, t, w1);
  }
  if constexpr (S::NP > 2) {
    sync();
    pass_scatter<S, 1>(v, t, smem_addr);
    sync();
    pass_gather<S, 2>(v, t, smem_addr);
    pass_compute<S, 2>(v, t, w1);
  }
  if constexpr (S::NP > 3) {
    sync();
    pass_scatter<S, 2>(v, t, smem_addr);
    sync();
    pass_gather<S, 3>(v, t, smem_addr);
    pass_compute<S, 3>(v, t, w1);
  }
}

}  // namespace accfft_b200
