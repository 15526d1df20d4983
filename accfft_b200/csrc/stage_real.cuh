// stage_real.cuh -- real transforms along a STRIDED axis (bodies only): the single-axis executes of the reference,
//   accfft_execute_{r2c,c2r}_{x,y}_gpu_t            (src/accfft_gpu.cpp:1330-1604)
// and the per-axis spectral derivative its operators build from them
//   accfft_grad_gpu_t / accfft_divergence_gpu_t      (src/operators_gpu.txx:172-278, :336-553).
//
// The reference runs cufftExecD2Z / Z2D with a stride (fplan_y, fplan_x) after a real-data transpose.  Here two
// ADJACENT real columns (b, b+1) -- 16 contiguous bytes in double -- are transformed as ONE complex column
// z = x_b + i x_(b+1) by the same register-resident Stockham passes as strided_c2c_body:
//   r2c:   Z = DFT(z);  X_b[k] = (Z[k] + conj Z[N-k]) / 2,  X_(b+1)[k] = (Z[k] - conj Z[N-k]) / 2i,  k = 0..N/2
//   c2r:   Z[k] = X_b[k] + i X_(b+1)[k] (Hermitian-extended to k > N/2);  z = IDFT(Z) = x_b + i x_(b+1)
//   deriv: a spectral multiplier with real physical-space action (i k) acts on z directly: no untangling at all,
//          forward passes -> multiply -> inverse passes inside one kernel = ONE read and ONE write of the field.
// All strides of the REAL side are in units of real elements; an odd column count, odd strides or an unaligned base
// fall back from 16-byte pair accesses to scalar ones.  Bodies are __host__ __device__ (tests/emul runs them on the CPU).
#pragma once
#include "stage.cuh"

namespace accfft_b200 {

template <class T>
ACC_HD bool pair_aligned(const void* p) { return ((unsigned long long)p & (2 * sizeof(T) - 1)) == 0; }

// segment of index k for the real copy stages: plain division (k can be a flat plane index of many millions,
// beyond the exact range of the magic multiply used by seg_of)
ACC_HD int seg_of_div(const StageDesc& d, long long k) { return d.nseg == 1 ? 0 : (int)(k / d.blk); }

// generic (segmented, optionally accumulating) store of ONE real element: output index k of column (a, b)
template <class T>
ACC_HD void store_real(const StageDesc& d, T y, long long k, int a, int b) {
  const int s = seg_of_div(d, k);
  const SegDst& sd = d.seg[s];
  T* dst = (T*)sd.base + (long long)a * sd.strideA + (long long)b * sd.strideB + (k - (long long)s * d.blk) * sd.strideI;
  *dst = d.accumulate ? T(*dst + y) : y;
}

// store the pair (x_b0[n], x_(b0+1)[n]) = (z.x, z.y) of output index n
template <class T>
ACC_HD void store_real_pair(const StageDesc& d, cplx<T> z, int n, int a, int b0, bool valid1) {
  const int s = seg_of(d, n);
  const SegDst& sd = d.seg[s];
  T* dst = (T*)sd.base + (long long)a * sd.strideA + (long long)b0 * sd.strideB + (long long)(n - s * d.blk) * sd.strideI;
  if (valid1 && sd.strideB == 1 && pair_aligned<T>(dst)) {
    cplx<T>* p = (cplx<T>*)dst;
    if (d.accumulate) {
      const cplx<T> o = *p;
      z = {o.x + z.x, o.y + z.y};
    }
    *p = z;
  } else {
    dst[0] = d.accumulate ? T(dst[0] + z.x) : z.x;
    if (valid1) dst[sd.strideB] = d.accumulate ? T(dst[sd.strideB] + z.y) : z.y;
  }
}

// pass-0 operands of a column pair straight from global memory (gather order), zero beyond the last column
template <class T, class S>
ACC_HD void load_real_pairs(const StageDesc& d, cplx<T>* v, int t, int a, int b0, bool valid1) {
  const T* in = (const T*)d.in + (long long)a * d.inStrideA + (long long)b0 * d.inStrideB;
  const long long si = d.inStrideI;
  const bool vec = valid1 && d.inStrideB == 1 && pair_aligned<T>(in) && !(si & 1);
  pass_foreach<S, 0>(v, t, [&](int i, cplx<T>& r) {
    const T* p = in + (long long)i * si;
    if (vec) r = *(const cplx<T>*)p;
    else r = {p[0], valid1 ? p[d.inStrideB] : T(0)};
  });
}

// ================================================================================================
// strided r2c: real [a][i][b] -> half spectrum complex [a][k = 0..N/2][b].   threads = S::G * TC, each thread
// column c of the tile owns the real columns 2c, 2c+1.  Output is local (one segment): the single-axis forward
// transforms of the reference end in the caller's buffer (src/accfft_gpu.cpp:1378-1386, :1524-1528).
// ================================================================================================
template <class T, class S, int TC, class Ctx>
ACC_HD void strided_r2c_body(const StageDesc& d, Ctx& ctx) {
  constexpr int N = S::N, G = S::G;
  const int tid = ctx.tid();
  const int c = tid % TC, t = tid / TC;
  const int tpa = (d.nB + 2 * TC - 1) / (2 * TC);
  const long long tile = ctx.bid();
  const int a = (int)(tile / tpa);
  const int b0 = 2 * ((int)(tile % tpa) * TC + c);
  const bool valid0 = b0 < d.nB, valid1 = b0 + 1 < d.nB;
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t, (const cplx<T>*)d.tw);
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  if (valid0) load_real_pairs<T, S>(d, v, t, a, b0, valid1);
  cplx<T>* sm = (cplx<T>*)ctx.smem();
  fft_passes<S>(v, t, w1, [&](int i) { return sm + S::phys(i) * TC + c; }, [&]() { ctx.sync(); });
  // publish Z so that every thread can read the mirror partner Z[N-k]
  if constexpr (S::NP > 1) ctx.sync();
  pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { sm[S::phys(k) * TC + c] = r; });
  ctx.sync();
  if (valid0) {
    const SegDst& sd = d.seg[0];
    cplx<T>* out = (cplx<T>*)sd.base + (long long)a * sd.strideA + (long long)b0 * sd.strideB;
    for (int k = t; 2 * k <= N; k += G) {
      const cplx<T> zk = sm[S::phys(k) * TC + c];
      const cplx<T> zm = sm[S::phys(k == 0 ? 0 : N - k) * TC + c];
      cplx<T>* p = out + (long long)k * sd.strideI;
      p[0] = {T(0.5) * (zk.x + zm.x), T(0.5) * (zk.y - zm.y)};                       // (Z[k] + conj Z[N-k]) / 2
      if (valid1) p[sd.strideB] = {T(0.5) * (zk.y + zm.y), T(0.5) * (zm.x - zk.x)};  // (Z[k] - conj Z[N-k]) / 2i
    }
  }
}

// ================================================================================================
// strided c2r (unnormalised, like cufftExecZ2D / fftw c2r: imaginary parts of X[0] and X[N/2] are ignored):
// half spectrum complex [a][k][b] -> real, every output element stored to the rank that owns it in the physical
// layout (segments over the output index n).  d.op (on load) makes it the second half of a spectral derivative.
// ================================================================================================
template <class T, class S, int TC, class Ctx>
ACC_HD void strided_c2r_body(const StageDesc& d, Ctx& ctx) {
  constexpr int N = S::N;
  const int tid = ctx.tid();
  const int c = tid % TC, t = tid / TC;
  const int tpa = (d.nB + 2 * TC - 1) / (2 * TC);
  const long long tile = ctx.bid();
  const int a = (int)(tile / tpa);
  const int b0 = 2 * ((int)(tile % tpa) * TC + c);
  const bool valid0 = b0 < d.nB, valid1 = b0 + 1 < d.nB;
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t, (const cplx<T>*)d.tw);
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  if (valid0) {
    const cplx<T>* in = (const cplx<T>*)d.in + (long long)a * d.inStrideA + (long long)b0 * d.inStrideB;
    const long long si = d.inStrideI;
    const OpCol cola = op_prepare_if(d, a, b0), colb = op_prepare_if(d, a, b0 + 1);
    pass_foreach<S, 0>(v, t, [&](int i, cplx<T>& r) {
      const bool lower = 2 * i <= N;
      const int m = lower ? i : N - i;
      const cplx<T>* p = in + (long long)m * si;
      cplx<T> xa = p[0];
      cplx<T> xb = valid1 ? p[d.inStrideB] : cplx<T>{T(0), T(0)};
      if (d.op != OP_NONE) {
        xa = apply_op(d, cola, xa, m);
        xb = apply_op(d, colb, xb, m);
      }
      if (m == 0 || 2 * m == N) xa.y = T(0), xb.y = T(0);
      // Z[i] = X_a + i X_b  (i <= N/2),  conj X_a + i conj X_b  (mirror);  stored swapped: inverse = swap.forward.swap
      const cplx<T> z = lower ? cplx<T>{xa.x - xb.y, xa.y + xb.x} : cplx<T>{xa.x + xb.y, xb.x - xa.y};
      r = cswap(z);
    });
  }
  cplx<T>* sm = (cplx<T>*)ctx.smem();
  fft_passes<S>(v, t, w1, [&](int i) { return sm + S::phys(i) * TC + c; }, [&]() { ctx.sync(); });
  if (valid0)
    pass_foreach<S, S::NP - 1>(v, t, [&](int n, cplx<T>& r) { store_real_pair(d, cswap(r), n, a, b0, valid1); });
}

// i * k * scale for transform index k of a derivative stage (OP_GRAD along the transform axis, per-axis scheme: the mode
// pair +-floor(N/2) vanishes as a pair) -- the inline special case of op_weight / apply_op (stage.cuh), in the kernel's own
// precision: no call and no double-precision arithmetic per element in the single-precision kernels
template <class T>
ACC_HD cplx<T> deriv_mul(const StageDesc& d, cplx<T> v, int k) {
  const int N = d.n[d.opAxis];
  const int X = k + d.off[d.opAxis];
  const int kw = (X == N / 2 || X == N - N / 2) ? 0 : (X > N / 2 ? X - N : X);
  const T w = T(d.scale) * T(kw);
  return {-(w * v.y), w * v.x};
}

// ================================================================================================
// strided spectral derivative: real -> forward passes -> multiply by the symbol (d.op, on the FULL axis) -> inverse
// passes -> real, one read and one write of the field.  Replaces r2c + grad_mult_wave_number{x,y} + c2r of
// src/operators_gpu.txx:224-262 (three kernels and four array passes in the reference).
// ================================================================================================
template <class T, class S, int TC, class Ctx>
ACC_HD void strided_deriv_body(const StageDesc& d, Ctx& ctx) {
  const int tid = ctx.tid();
  const int c = tid % TC, t = tid / TC;
  const int tpa = (d.nB + 2 * TC - 1) / (2 * TC);
  const long long tile = ctx.bid();
  const int a = (int)(tile / tpa);
  const int b0 = 2 * ((int)(tile % tpa) * TC + c);
  const bool valid0 = b0 < d.nB, valid1 = b0 + 1 < d.nB;
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t, (const cplx<T>*)d.tw);
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  if (valid0) load_real_pairs<T, S>(d, v, t, a, b0, valid1);
  cplx<T>* sm = (cplx<T>*)ctx.smem();
  auto addr = [&](int i) { return sm + S::phys(i) * TC + c; };
  fft_passes<S>(v, t, w1, addr, [&]() { ctx.sync(); });
  // symbol multiply in output order (index j + r*N/R_last).  When the first and the last radix agree that IS the pass-0
  // gather order of the inverse (index j + r*N/R_0), so the registers go straight into the second set of passes;
  // otherwise they are re-ordered through the tile.  (swapped: inverse = swap . forward . swap)
  constexpr bool kSameOrder = S::radix(0) == S::radix(S::NP - 1);
  if constexpr (kSameOrder) {
    pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { r = cswap(deriv_mul(d, r, k)); });
    if constexpr (S::NP > 1) ctx.sync();   // the forward's last gathers are done before the inverse scatters again
  } else {
    if constexpr (S::NP > 1) ctx.sync();
    pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { *addr(k) = cswap(deriv_mul(d, r, k)); });
    ctx.sync();
    pass_gather<S, 0>(v, t, addr);
    if constexpr (S::NP > 1) ctx.sync();
  }
  fft_passes<S>(v, t, w1, addr, [&]() { ctx.sync(); });
  if (valid0)
    pass_foreach<S, S::NP - 1>(v, t, [&](int n, cplx<T>& r) { store_real_pair(d, cswap(r), n, a, b0, valid1); });
}

// ================================================================================================
// contiguous spectral derivative (Z axis): two ADJACENT PENCILS (rows of N real values) form one complex row
// z[n] = row_p[n] + i row_(p+1)[n]; forward passes -> symbol -> inverse passes -> real parts back to row_p, imaginary parts to
// row_(p+1).  One read and one write of the field instead of r2c + multiply + c2r (src/operators_gpu.txx:264-276: three
// kernels, four array passes).  PC pencil PAIRS per CTA, S::G threads per pair; S is the schedule of the FULL length N2.
// Strides in real elements; loads / stores are scalar and coalesced along n.
// ================================================================================================
template <class T, class S, int PC, class Ctx>
ACC_HD void contig_deriv_body(const StageDesc& d, Ctx& ctx) {
  const int tid = ctx.tid();
  const int t = tid % S::G, pcl = tid / S::G;
  const long long total = (long long)d.nA * d.nB;                 // pencils
  const long long p0 = 2 * (ctx.bid() * PC + pcl), p1 = p0 + 1;   // this thread group's two pencils
  const bool valid0 = p0 < total, valid1 = p1 < total;
  const int a0 = valid0 ? (int)(p0 / d.nB) : 0, b0 = valid0 ? (int)(p0 % d.nB) : 0;
  const int a1 = valid1 ? (int)(p1 / d.nB) : 0, b1 = valid1 ? (int)(p1 % d.nB) : 0;
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t, (const cplx<T>*)d.tw);
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  if (valid0) {
    const T* r0 = (const T*)d.in + (long long)a0 * d.inStrideA + (long long)b0 * d.inStrideB;
    const T* r1 = (const T*)d.in + (long long)a1 * d.inStrideA + (long long)b1 * d.inStrideB;
    pass_foreach<S, 0>(v, t, [&](int i, cplx<T>& r) { r = {r0[i], valid1 ? r1[i] : T(0)}; });
  }
  cplx<T>* sm = (cplx<T>*)ctx.smem() + pcl * S::SMEM_ELEMS;
  auto addr = [&](int i) { return sm + S::phys(i); };
  fft_passes<S>(v, t, w1, addr, [&]() { group_sync<S::G, PC>(ctx, pcl); });
  constexpr bool kSameOrder = S::radix(0) == S::radix(S::NP - 1);
  if constexpr (kSameOrder) {
    pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { r = cswap(deriv_mul(d, r, k)); });
    if constexpr (S::NP > 1) group_sync<S::G, PC>(ctx, pcl);
  } else {
    if constexpr (S::NP > 1) group_sync<S::G, PC>(ctx, pcl);
    pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { *addr(k) = cswap(deriv_mul(d, r, k)); });
    group_sync<S::G, PC>(ctx, pcl);
    pass_gather<S, 0>(v, t, addr);
    if constexpr (S::NP > 1) group_sync<S::G, PC>(ctx, pcl);
  }
  fft_passes<S>(v, t, w1, addr, [&]() { group_sync<S::G, PC>(ctx, pcl); });
  if (valid0) {
    const SegDst& sd = d.seg[0];
    T* o0 = (T*)sd.base + (long long)a0 * sd.strideA + (long long)b0 * sd.strideB;
    T* o1 = (T*)sd.base + (long long)a1 * sd.strideA + (long long)b1 * sd.strideB;
    const bool acc = d.accumulate != 0;
    pass_foreach<S, S::NP - 1>(v, t, [&](int n, cplx<T>& r) {   // r holds the swapped result: (im, re)
      o0[n] = acc ? T(o0[n] + r.y) : r.y;
      if (valid1) o1[n] = acc ? T(o1[n] + r.x) : r.x;
    });
  }
}

// ================================================================================================
// real copy stage: element (a, b, k) of the input goes to segment k / blk (the real-data transposes T_plan_y / T_plan_x
// and their inverses, src/accfft_gpu.cpp:369-373, :426-438, and the D2D copies around them, :1357, :1407, :1509, :1543).
// One call = one element; the launchers walk the index space with k fastest.
// ================================================================================================
template <class T>
ACC_HD void copy_real_element(const StageDesc& d, long long idx, long long nK) {
  const long long k = idx % nK;
  const long long ab = idx / nK;
  const int b = (int)(ab % d.nB), a = (int)(ab / d.nB);
  const T* in = (const T*)d.in;
  store_real(d, in[(long long)a * d.inStrideA + (long long)b * d.inStrideB + k * d.inStrideI], k, a, b);
}

}  // namespace accfft_b200
