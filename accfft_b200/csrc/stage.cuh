// stage.cuh -- the per-axis stage kernels of the distributed 3-D transform (bodies only).
//
// One "stage" = batched 1-D transforms along one pencil axis, with the pencil transpose of the
// reference fused into the kernel's stores: every output element is written straight to the rank
// (peer-mapped buffer) that owns it in the NEXT stage's layout.  This replaces, per stage,
//   cufftExec*                       (src/accfft_gpu.cpp:477-539, :541-602)
// + pack kernel / memcpy_cuda        (src/transpose_cuda.cu:59-75, :97-204)
// + cudaMemcpy D2H / MPI_Alltoall[v] / cudaMemcpy H2D   (src/transpose_gpu.cpp:3940-3955)
// + unpack kernel                    (src/transpose_cuda.cu:13-21)
// of the reference.  Bodies are __host__ __device__ templates over an execution context so the CPU
// fiber emulator in tests/emul can execute the identical code (test infrastructure only).
#pragma once
#include "fft_engine.cuh"

namespace accfft_b200 {

constexpr int kMaxSeg = 16;   // max ranks in one row / column group

// destination of output indices [s*blk, (s+1)*blk) of every column: the buffer of group peer s
struct SegDst {
  void* base;            // peer-mapped (or local) buffer, offset already folded in
  long long strideA;     // elements per step of the column's outer index a
  long long strideB;     // elements per step of the column's inner index b
  long long strideI;     // elements per step of the transform index (k - s*blk)
};

enum SpectralOp : int {
  OP_NONE = 0,
  OP_SCALE = 1,       // multiply by scale
  OP_GRAD = 2,        // i * k_axis * scale          (src/operators_cuda.cu:15-112)
  OP_LAPLACE = 3,     // -(kx^2+ky^2+kz^2) * scale   (src/operators_cuda.cu:114-158)
  OP_BIHARMONIC = 4,  // (kx^2+ky^2+kz^2)^2 * scale  (src/operators_cuda.cu:160-205)
  // solver-style symbols.  The reference has them on the CPU only (accfft_inv_laplace / accfft_inv_biharmonic,
  // include/accfft_operators.h:57-60, src/operators.txx:275-392) or as user kernels between two transforms
  // (steps/step4/kernels.cu:22-112 Poisson filter, steps/step6 heat step); here they ride on the X stage like the others.
  OP_INV_LAPLACE = 5,     // scale / -(|k|^2), 0 where |k|^2 == 0      (src/operators.txx:318-327)
  OP_INV_BIHARMONIC = 6,  // scale / |k|^4 with the PLAIN wave numbers (the reference has the Nyquist rule commented out here,
                          // src/operators.txx:362-375) and the zero mode passed through (`if (wave == 0) wave = 1`, :379-381)
  OP_SYMBOL_LIN = 7,      // scale * (opA + opB |k|^2)                  (explicit heat step: 1 - dt |k|^2)
  OP_SYMBOL_INV = 8,      // scale / (opA + opB |k|^2), 0 where the denominator is 0   (Helmholtz / implicit heat / Poisson)
};

struct StageDesc {
  const void* in;
  long long inStrideA, inStrideB, inStrideI;   // in elements of the input type
  int nA, nB;                                  // columns are indexed (a, b), b fastest
  int nseg, blk;
  unsigned blkMagic;                           // ceil(2^32 / blk)
  int pad0;
  SegDst seg[kMaxSeg];
  const void* tw;                              // W_M[k], see fft_engine.cuh
  // spectral operator applied to the transform output (forward stages) -- see SpectralOp
  int op;
  int opAxis;            // OP_GRAD: which wave number (0,1,2) multiplies
  int n[3];              // global grid
  int kind[3];           // how to obtain the global frequency index along x,y,z for an output element:
                         //   0 = transform index k, 1 = a + off, 2 = b + off, 3 = (b / div) + off, 4 = (b % div) + off
  int off[3];
  int div;
  double scale;
  int opOnLoad;          // 1: apply `op` to the INPUT elements (inverse X stage of grad) instead of the outputs
  int accumulate;        // 1: outputs are added to the destination (divergence sums three spectra)
  int colShift;          // strided kernels: tiles start at column -colShift so that tile origins are line aligned
                         // in a destination whose rows this rank enters at an unaligned offset (inverse Y stage)
  int inv;               // 1: inverse transform (sign +1), done as swap(re,im) . forward . swap(re,im)
  double opA, opB;       // OP_SYMBOL_LIN / OP_SYMBOL_INV coefficients
  int opFullAxis;        // OP_GRAD on an axis that the 3-D transform treats as a full COMPLEX axis (the reference's "slow" 3-D scheme,
                         // src/operators_cuda.cu:34-37): only X == N/2 is zeroed.  0 = the per-axis real-transform scheme, where
                         // the mode pair +-floor(N/2) vanishes as a pair (differs for odd N only).
  int staged;            // contiguous kernels with nseg > 1: outputs are collected in shared memory per destination
                         // segment and leave the CTA as bulk (TMA) copies of whole runs -- full 128-byte lines on the
                         // NVLink wire instead of one partial-line packet per warp store.  Set by the launcher.
  // Fused two-pass kernel (fused.cuh): the buffer between the two passes is used as a RING of a few planes of the outer
  // index a (plane a lives at slot a % ring) that stays resident in L2, so the intermediate array never reaches HBM.
  // Only honoured by kernels whose context says so (ctx_ring); 0 = plain addressing.
  int inRing, outRing;   // ring - 1 (ring = a power of two planes); accesses that do NOT touch the ring are marked evict-first
};

// does this execution context run ring-addressed stages (the fused persistent kernel)?  Default: no.
template <class Ctx, class = void>
struct ctx_ring : std::false_type {};
template <class Ctx>
struct ctx_ring<Ctx, std::void_t<decltype(Ctx::kRing)>> : std::integral_constant<bool, Ctx::kRing> {};

// (inRing / outRing hold ring - 1 with ring a power of two, i.e. a mask; which side of a stage is the ring is a compile-time
// property of the context: Ctx::kRingIn / Ctx::kRingOut)
template <class Ctx> ACC_HD int ring_in(const StageDesc& d, int a) {
  if constexpr (ctx_ring<Ctx>::value) {
    if constexpr (Ctx::kRingIn) return a & d.inRing;
    else return a;
  } else {
    return a;
  }
}
template <class Ctx> ACC_HD int ring_out(const StageDesc& d, int a) {
  if constexpr (ctx_ring<Ctx>::value) {
    if constexpr (Ctx::kRingOut) return a & d.outRing;
    else return a;
  } else {
    return a;
  }
}

// Barrier among the G threads that work on ONE pencil of a contiguous stage (pencil `pcl` of the CTA's PC): the pencils of a
// CTA share nothing, so a CTA-wide barrier only makes every pencil wait for the slowest one.  Contexts that can do better
// (Ctx::kGroupSync: warp-level sync when a pencil's threads sit inside one warp, a named barrier per pencil when they are whole
// warps) provide sync_group; everything else (the CPU emulator) falls back to the CTA barrier, which is always sufficient.
template <class Ctx, class = void>
struct ctx_group_sync : std::false_type {};
template <class Ctx>
struct ctx_group_sync<Ctx, std::void_t<decltype(Ctx::kGroupSync)>> : std::integral_constant<bool, Ctx::kGroupSync> {};
template <int G, int PC, class Ctx>
ACC_HD void group_sync(Ctx& ctx, int pcl) {
  if constexpr (ctx_group_sync<Ctx>::value) ctx.template sync_group<G, PC>(pcl);
  else ctx.sync();
}

// global loads / stores of a stage.  In a ring context, ring data is read with ld.global.cg (L2 only: another SM wrote it
// during this kernel, an L1 line could be stale) and everything else may carry an evict-first hint.
template <class Ctx, class T>
ACC_HD cplx<T> ld_g(const StageDesc& d, const cplx<T>* p) {
#if defined(__CUDA_ARCH__)
  if constexpr (ctx_ring<Ctx>::value) {
    if constexpr (sizeof(T) == 8) {
      const double2 v = Ctx::kRingIn ? __ldcg(reinterpret_cast<const double2*>(p)) : __ldcs(reinterpret_cast<const double2*>(p));
      return {v.x, v.y};
    } else {
      const float2 v = Ctx::kRingIn ? __ldcg(reinterpret_cast<const float2*>(p)) : __ldcs(reinterpret_cast<const float2*>(p));
      return {v.x, v.y};
    }
  }
#endif
  return ld_stream(p);
}
template <class Ctx, class T>
ACC_HD void st_g(const StageDesc& d, cplx<T>* p, cplx<T> v) {
#if defined(__CUDA_ARCH__)
  if constexpr (ctx_ring<Ctx>::value) {
    if constexpr (!Ctx::kRingOut) {
      if constexpr (sizeof(T) == 8) __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
      else __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
      return;
    }
  }
#endif
  st_stream(p, v);
}

template <class T>
ACC_HD void store_out(const StageDesc& d, cplx<T>* dst, cplx<T> y) {
  if (d.accumulate) {
    const cplx<T> o = *dst;
    y = {o.x + y.x, o.y + y.y};
  }
  *dst = y;
}

ACC_HD int seg_of(const StageDesc& d, int k) {
  if (d.nseg == 1) return 0;
  if (d.blk == 1) return k;
  return (int)(((unsigned long long)(unsigned)k * d.blkMagic) >> 32);
}

// wave number with the reference's Nyquist filter:  k = X > N/2 ? X-N : X;  k = 0 at X == N/2
ACC_HD long long wave_number(int X, int N) {
  long long w = X;
  if (X > N / 2) w -= N;
  if (X == N / 2) w = 0;
  return w;
}

ACC_HD int freq_index(const StageDesc& d, int ax, int k, int a, int b) {
  switch (d.kind[ax]) {
    case 0: return k + d.off[ax];
    case 1: return a + d.off[ax];
    case 2: return b + d.off[ax];
    case 3: return b / d.div + d.off[ax];
    default: return b % d.div + d.off[ax];
  }
}

// The symbol multiply, split in two so that nothing column-invariant is recomputed per element:
//   op_prepare(d, a, b)        once per thread (out of line): what the two axes that are FIXED for column (a, b) contribute
//   apply_op(d, col, v, k)     per element: the wave number of the transform axis (index k) + the symbol itself
// (the transform axis is the one with d.kind[ax] == 0; for OP_GRAD the derivative axis may be the transform axis or a fixed one)
struct OpCol {
  double s_other;   // sum of the squared wave numbers of the fixed axes (laplace family)
  double g_fixed;   // OP_GRAD along a FIXED axis: scale * wave number of that axis for this column
  int kax;          // the transform axis
};

ACC_HD long long wave_number_plain(int X, int N) { return X > N / 2 ? (long long)X - N : (long long)X; }

ACC_HD_NOINLINE OpCol op_prepare(const StageDesc& d, int a, int b) {
  OpCol c;
  c.s_other = 0.0;
  c.g_fixed = 0.0;
  c.kax = d.kind[0] == 0 ? 0 : d.kind[1] == 0 ? 1 : 2;
  if (d.op == OP_GRAD) {
    const int ax = d.opAxis;
    if (d.kind[ax] == 0) {
      c.kax = ax;
    } else {
      const int X = freq_index(d, ax, 0, a, b), N = d.n[ax];
      c.g_fixed = (X == N - N / 2 && !d.opFullAxis) ? 0.0 : d.scale * (double)wave_number(X, N);
    }
    return c;
  }
  if (d.op == OP_SCALE || d.op == OP_NONE) return c;
  for (int ax = 0; ax < 3; ++ax) {
    if (ax == c.kax) continue;
    const int X = freq_index(d, ax, 0, a, b);
    // the reference's inv_biharmonic has the Nyquist rule commented out (src/operators.txx:362-375)
    const long long w = d.op == OP_INV_BIHARMONIC ? wave_number_plain(X, d.n[ax]) : wave_number(X, d.n[ax]);
    c.s_other += (double)(w * w);
  }
  return c;
}

// The symbol's (real) weight for transform index k of a prepared column -- out of line: ONE copy of the double-precision
// symbol arithmetic per translation unit instead of one per unrolled element of every general kernel (the general
// kernels shrank by a third; the call costs a handful of cycles per element of an HBM- or NVLink-bound stage).
// OP_GRAD multiplies by i * w, everything else by w.
ACC_HD_NOINLINE double op_weight(const StageDesc& d, const OpCol& c, int k) {
  if (d.op == OP_SCALE) return d.scale;
  if (d.op == OP_GRAD) {
    // The reference differentiates with a REAL transform along the axis (src/operators_gpu.txx:172-278), i.e. it
    // only ever sees X <= N/2 and c2r implies the mirrored half: the mode pair +-floor(N/2) is zeroed as a pair.
    // On a full (c2c) axis that means X == N/2 and X == N - N/2 (the same index when N is even); the reference's full
    // 3-D scheme (opFullAxis) zeroes X == N/2 only (src/operators_cuda.cu:34-37).
    if (d.kind[d.opAxis] != 0) return c.g_fixed;
    const int X = k + d.off[c.kax], N = d.n[c.kax];
    return (X == N - N / 2 && !d.opFullAxis) ? 0.0 : d.scale * (double)wave_number(X, N);
  }
  const int X = k + d.off[c.kax], N = d.n[c.kax];
  const long long wk = d.op == OP_INV_BIHARMONIC ? wave_number_plain(X, N) : wave_number(X, N);
  const double s2 = c.s_other + (double)(wk * wk);
  double w;
  switch (d.op) {
    case OP_LAPLACE: w = -s2; break;
    case OP_BIHARMONIC: w = s2 * s2; break;
    case OP_INV_LAPLACE: w = s2 == 0.0 ? 0.0 : -1.0 / s2; break;
    case OP_INV_BIHARMONIC: w = s2 == 0.0 ? 1.0 : 1.0 / (s2 * s2); break;   // zero mode passed through (:379-381)
    case OP_SYMBOL_LIN: w = d.opA + d.opB * s2; break;
    default: {
      const double den = d.opA + d.opB * s2;
      w = den == 0.0 ? 0.0 : 1.0 / den;
    }
  }
  return w * d.scale;
}

template <class T>
ACC_HD cplx<T> apply_op(const StageDesc& d, const OpCol& c, cplx<T> v, int k) {
  if (d.op == OP_NONE) return v;
  const double w = op_weight(d, c, k);
  if (d.op == OP_GRAD) return {T(-w * v.y), T(w * v.x)};
  return {T(w * v.x), T(w * v.y)};
}

// one-off form (copy stages, generic kernel): prepare + apply
template <class T>
ACC_HD cplx<T> apply_op(const StageDesc& d, cplx<T> v, int k, int a, int b) {
  if (d.op == OP_NONE) return v;
  const OpCol c = op_prepare(d, a, b);
  return apply_op(d, c, v, k);
}

// true when the stage needs none of the generic machinery: one destination segment, no spectral operator
// (every stage kernel exists in two instantiations: GEN = false serves exactly these stages -- the single-GPU transform --
// and carries no segment / operator / accumulate code at all; GEN = true is the general kernel)
ACC_HD bool stage_is_simple(const StageDesc& d) { return d.nseg == 1 && d.op == OP_NONE && d.accumulate == 0; }

template <class T>
ACC_HD cplx<T> cswap_if(cplx<T> a, bool sw) { return sw ? cplx<T>{a.y, a.x} : a; }

// generic (segmented / operator) store of output element k of column (a, b); col = op_prepare(d, a, b)
template <class T>
ACC_HD void store_generic(const StageDesc& d, const OpCol& col, cplx<T> y, int k, int a, int b) {
  const int s = seg_of(d, k);
  const SegDst& sd = d.seg[s];
  cplx<T>* dst = (cplx<T>*)sd.base + (long long)a * sd.strideA + (long long)b * sd.strideB +
                 (long long)(k - s * d.blk) * sd.strideI;
  store_out(d, dst, d.opOnLoad ? y : apply_op(d, col, y, k));
}
template <class T>
ACC_HD void store_generic(const StageDesc& d, cplx<T> y, int k, int a, int b) {
  OpCol col = {0.0, 0.0, 0};
  if (d.op != OP_NONE && !d.opOnLoad) col = op_prepare(d, a, b);
  store_generic(d, col, y, k, a, b);
}
ACC_HD OpCol op_prepare_if(const StageDesc& d, int a, int b) {
  OpCol col = {0.0, 0.0, 0};
  if (d.op != OP_NONE) col = op_prepare(d, a, b);
  return col;
}

// ================================================================================================
// strided c2c: transform index strided in memory, TC adjacent columns (unit stride) per CTA tile.
// Used for the Y and X passes (reference: fplan_1 / fplan_2, src/accfft_gpu.cpp:208-253).
// threads = S::G * TC;  tid -> (c = tid % TC, t = tid / TC);  smem layout [phys(i)][c]
// ================================================================================================
template <class T, class S, int TC, bool GEN, class Ctx>
ACC_HD void strided_c2c_body(const StageDesc& d, Ctx& ctx) {
  constexpr int G = S::G;
  constexpr int R0 = S::radix(0), NB0 = S::nb(0), CNT0 = S::cnt(0);
  constexpr int RL = S::radix(S::NP - 1), NBL = S::nb(S::NP - 1), CNTL = S::cnt(S::NP - 1);
  const int tid = ctx.tid();
  const int c = tid % TC, t = tid / TC;
  const int tpa = (d.nB + d.colShift + TC - 1) / TC;
  const long long tile = ctx.bid();
  const int a = (int)(tile / tpa);
  const int b = (int)(tile % tpa) * TC + c - d.colShift;
  const bool valid = b >= 0 && b < d.nB;
  const bool inv = d.inv != 0;
  OpCol col = {0.0, 0.0, 0};
  if constexpr (GEN) {
    if (valid) col = op_prepare_if(d, a, b);
  }
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t, (const cplx<T>*)d.tw);
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  if (valid) {
    const cplx<T>* in = (const cplx<T>*)d.in + (long long)ring_in<Ctx>(d, a) * d.inStrideA + b;
    const long long si = d.inStrideI;
    const long long step = (long long)NB0 * si;
    ACC_UNROLL
    for (int u = 0; u < CNT0; ++u) {
      const int j = t + u * G;
      if (CNT0 * G == NB0 || j < NB0) {
        const cplx<T>* p = in + (long long)j * si;
        ACC_UNROLL
        for (int r = 0; r < R0; ++r) v[u * R0 + r] = ld_g<Ctx>(d, p + r * step);
      }
    }
    if constexpr (ctx_ring<Ctx>::value) ctx.after_loads();
    if constexpr (GEN) {
      if (d.opOnLoad) pass_foreach<S, 0>(v, t, [&](int i, cplx<T>& r) { r = apply_op(d, col, r, i); });
    }
    if (inv) {
      ACC_UNROLL
      for (int i = 0; i < S::NREG; ++i) v[i] = cswap(v[i]);
    }
  }
  cplx<T>* sm = (cplx<T>*)ctx.smem();
  fft_passes<S>(v, t, w1, [&](int i) { return sm + S::phys(i) * TC + c; }, [&]() { ctx.sync(); });
  if (valid) {
    if (inv) {
      ACC_UNROLL
      for (int i = 0; i < S::NREG; ++i) v[i] = cswap(v[i]);
    }
    if constexpr (!GEN) {
      const SegDst& sd = d.seg[0];
      cplx<T>* out = (cplx<T>*)sd.base + (long long)ring_out<Ctx>(d, a) * sd.strideA + (long long)b * sd.strideB;
      const long long so = sd.strideI;
      const long long step = (long long)NBL * so;
      ACC_UNROLL
      for (int u = 0; u < CNTL; ++u) {
        const int j = t + u * G;
        if (CNTL * G == NBL || j < NBL) {
          cplx<T>* p = out + (long long)j * so;
          ACC_UNROLL
          for (int r = 0; r < RL; ++r) st_g<Ctx>(d, p + r * step, v[u * RL + r]);
        }
      }
    } else {
      const int ao = ring_out<Ctx>(d, a);
      pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { store_generic(d, col, r, k, ao, b); });
    }
  }
}


// ================================================================================================
// strided c2c, SINGLE precision, two adjacent columns per thread: the pair (b0, b0 + 1) travels as ONE 16-byte quantity
// cplx<f2> = {re0, re1, im0, im1} through registers and shared memory, so that the kernel issues 128-bit global and
// shared-memory accesses and half as many instructions per byte as the 8-byte-element form -- the same shape as the
// double-precision kernel (TCP pair-columns = 128-byte row segments with G * TCP threads).  Twiddles stay scalar floats and
// are shared by the two lanes.  A pair whose global address is not 16-byte aligned (odd row pitch on odd rows, odd column
// shift) or that straddles the edge of the column range moves as two 8-byte halves; the choice is uniform per tile row.
// ================================================================================================
struct alignas(16) f4raw {
  float x, y, z, w;
};

template <class Ctx>
ACC_HD cplx<f2> load_pair(const StageDesc& d, const cplx<float>* p, bool v0, bool v1) {
  if (v0 && v1 && (((unsigned long long)p) & 15ull) == 0) {
#if defined(__CUDA_ARCH__)
    if constexpr (ctx_ring<Ctx>::value) {
      const float4 q = Ctx::kRingIn ? __ldcg(reinterpret_cast<const float4*>(p)) : __ldcs(reinterpret_cast<const float4*>(p));
      return {f2(q.x, q.z), f2(q.y, q.w)};
    }
#endif
    const f4raw q = *reinterpret_cast<const f4raw*>(p);
    return {f2(q.x, q.z), f2(q.y, q.w)};
  }
  const cplx<float> e0 = v0 ? ld_g<Ctx>(d, p) : cplx<float>{0.f, 0.f};
  const cplx<float> e1 = v1 ? ld_g<Ctx>(d, p + 1) : cplx<float>{0.f, 0.f};
  return {f2(e0.x, e1.x), f2(e0.y, e1.y)};
}
template <class Ctx>
ACC_HD void store_pair(const StageDesc& d, cplx<float>* p, cplx<f2> v, bool v0, bool v1) {
  if (v0 && v1 && (((unsigned long long)p) & 15ull) == 0) {
#if defined(__CUDA_ARCH__)
    if constexpr (ctx_ring<Ctx>::value) {
      if constexpr (!Ctx::kRingOut) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v.x.a, v.y.a, v.x.b, v.y.b));
        return;
      }
    }
#endif
    *reinterpret_cast<f4raw*>(p) = f4raw{v.x.a, v.y.a, v.x.b, v.y.b};
    return;
  }
  if (v0) st_g<Ctx>(d, p, cplx<float>{v.x.a, v.y.a});
  if (v1) st_g<Ctx>(d, p + 1, cplx<float>{v.x.b, v.y.b});
}

template <class S, int TCP, bool GEN, class Ctx>
ACC_HD void strided_c2c_pair_body(const StageDesc& d, Ctx& ctx) {
  using T = f2;
  constexpr int G = S::G;
  constexpr int R0 = S::radix(0), NB0 = S::nb(0), CNT0 = S::cnt(0);
  constexpr int RL = S::radix(S::NP - 1), NBL = S::nb(S::NP - 1), CNTL = S::cnt(S::NP - 1);
  const int tid = ctx.tid();
  const int c = tid % TCP, t = tid / TCP;
  const int tpa = (d.nB + d.colShift + 2 * TCP - 1) / (2 * TCP);
  const long long tile = ctx.bid();
  const int a = (int)(tile / tpa);
  const int b0 = 2 * ((int)(tile % tpa) * TCP + c) - d.colShift;
  const bool v0 = b0 >= 0 && b0 < d.nB, v1 = b0 + 1 >= 0 && b0 + 1 < d.nB;
  const bool inv = d.inv != 0;
  OpCol col0 = {0.0, 0.0, 0}, col1 = {0.0, 0.0, 0};
  if constexpr (GEN) {
    if (v0) col0 = op_prepare_if(d, a, b0);
    if (v1) col1 = op_prepare_if(d, a, b0 + 1);
  }
  cplx<float> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t, (const cplx<float>*)d.tw);
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0.f), T(0.f)};
  if (v0 || v1) {
    const cplx<float>* in = (const cplx<float>*)d.in + (long long)ring_in<Ctx>(d, a) * d.inStrideA + b0;
    const long long si = d.inStrideI;
    const long long step = (long long)NB0 * si;
    ACC_UNROLL
    for (int u = 0; u < CNT0; ++u) {
      const int j = t + u * G;
      if (CNT0 * G == NB0 || j < NB0) {
        const cplx<float>* p = in + (long long)j * si;
        ACC_UNROLL
        for (int r = 0; r < R0; ++r) v[u * R0 + r] = load_pair<Ctx>(d, p + r * step, v0, v1);
      }
    }
    if constexpr (ctx_ring<Ctx>::value) ctx.after_loads();
    if constexpr (GEN) {
      if (d.opOnLoad)
        pass_foreach<S, 0>(v, t, [&](int i, cplx<T>& r) {
          const cplx<float> e0 = apply_op(d, col0, cplx<float>{r.x.a, r.y.a}, i);
          const cplx<float> e1 = apply_op(d, col1, cplx<float>{r.x.b, r.y.b}, i);
          r = {f2(e0.x, e1.x), f2(e0.y, e1.y)};
        });
    }
    if (inv) {
      ACC_UNROLL
      for (int i = 0; i < S::NREG; ++i) v[i] = cswap(v[i]);
    }
  }
  cplx<T>* sm = (cplx<T>*)ctx.smem();
  fft_passes<S>(v, t, w1, [&](int i) { return sm + S::phys(i) * TCP + c; }, [&]() { ctx.sync(); });
  if (v0 || v1) {
    if (inv) {
      ACC_UNROLL
      for (int i = 0; i < S::NREG; ++i) v[i] = cswap(v[i]);
    }
    if constexpr (!GEN) {
      const SegDst& sd = d.seg[0];
      cplx<float>* out = (cplx<float>*)sd.base + (long long)ring_out<Ctx>(d, a) * sd.strideA + (long long)b0 * sd.strideB;
      const long long so = sd.strideI;
      const long long step = (long long)NBL * so;
      const bool unit = sd.strideB == 1;
      ACC_UNROLL
      for (int u = 0; u < CNTL; ++u) {
        const int j = t + u * G;
        if (CNTL * G == NBL || j < NBL) {
          cplx<float>* p = out + (long long)j * so;
          ACC_UNROLL
          for (int r = 0; r < RL; ++r) {
            const cplx<T> y = v[u * RL + r];
            if (unit) {
              store_pair<Ctx>(d, p + r * step, y, v0, v1);
            } else {
              if (v0) st_g<Ctx>(d, p + r * step, cplx<float>{y.x.a, y.y.a});
              if (v1) st_g<Ctx>(d, p + r * step + sd.strideB, cplx<float>{y.x.b, y.y.b});
            }
          }
        }
      }
    } else {
      const int ao = ring_out<Ctx>(d, a);
      pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) {
        if (v0) store_generic(d, col0, cplx<float>{r.x.a, r.y.a}, k, ao, b0);
        if (v1) store_generic(d, col1, cplx<float>{r.x.b, r.y.b}, k, ao, b0 + 1);
      });
    }
  }
}

// ---- staged (shared memory -> bulk copy) stores of the contiguous kernels ------------------------------------
// Staging layout for a CTA of PC pencils and NOUT outputs per pencil:  [segment s][pencil][k - s*blk], i.e. element
// (pencil, k) lives at  PC*s*blk + pencil*size_s + (k - s*blk)  with size_s = min(blk, NOUT - s*blk).  Rows of one
// segment are adjacent, exactly like the destination buffer (dense rows of size_s elements, one per pencil), so the
// PC pencils of a CTA form ONE contiguous run per destination rank whenever they share the outer index a.
template <int PC>
ACC_HD int staged_index(const StageDesc& d, int nout, int pcl, int k) {
  const int s = seg_of(d, k);
  const int rest = nout - s * d.blk;
  const int size_s = rest < d.blk ? rest : d.blk;
  return PC * s * d.blk + pcl * size_s + (k - s * d.blk);
}

// all threads of the CTA call this after their staged writes; idx0 = first pencil of the CTA
template <class T, int PC, class Ctx>
ACC_HD void staged_flush(const StageDesc& d, Ctx& ctx, const cplx<T>* stage, int nout, long long idx0, int nthreads) {
  ctx.bulk_fence();
  ctx.sync();
  const long long total = (long long)d.nA * d.nB;
  const int npen = (int)((total - idx0) < PC ? (total - idx0) : PC);
  const int a0 = (int)(idx0 / d.nB), b0 = (int)(idx0 % d.nB);
  const int tid = ctx.tid();
  int issue = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const int rest = nout - s * d.blk;
    const int size_s = rest < d.blk ? rest : d.blk;
    if (size_s <= 0) continue;
    const SegDst& sd = d.seg[s];
    const bool merged = (b0 + npen <= d.nB) && sd.strideB == size_s && sd.strideI == 1;
    const int nruns = merged ? 1 : npen;
    for (int r = 0; r < nruns; ++r) {
      const long long idx = idx0 + r;
      const int a = merged ? a0 : (int)(idx / d.nB), b = merged ? b0 : (int)(idx % d.nB);
      cplx<T>* dst = (cplx<T>*)sd.base + (long long)a * sd.strideA + (long long)b * sd.strideB;
      const cplx<T>* src = stage + PC * s * d.blk + r * size_s;
      const int n = merged ? npen * size_s : size_s;
      const unsigned bytes = (unsigned)n * (unsigned)sizeof(cplx<T>);
      const bool aligned = sd.strideI == 1 && (((unsigned long long)dst | (unsigned long long)src | bytes) & 15ull) == 0;
      if (aligned) {
        if (tid == issue % nthreads) ctx.bulk_store(dst, src, bytes);
        ++issue;
      } else {
        for (int i = tid; i < n; i += nthreads) dst[(long long)i * sd.strideI] = src[i];
      }
    }
  }
  ctx.bulk_wait();
}

// ================================================================================================
// contiguous c2c: transform index has unit stride; PC pencils per CTA, G threads per pencil.
// Used for the Z pass of c2c plans (reference fplan_0 of the c2c plan, src/accfft_gpu.cpp:767-790).
// ================================================================================================
template <class T, class S, int PC, bool GEN, class Ctx>
ACC_HD void contig_c2c_body(const StageDesc& d, Ctx& ctx) {
  const int tid = ctx.tid();
  const int t = tid % S::G, pcl = tid / S::G;
  const long long idx = ctx.bid() * PC + pcl;
  const bool valid = idx < (long long)d.nA * d.nB;
  const int a = valid ? (int)(idx / d.nB) : 0;
  const int b = valid ? (int)(idx % d.nB) : 0;
  const bool inv = d.inv != 0;
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t, (const cplx<T>*)d.tw);
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  if (valid) {
    const cplx<T>* in = (const cplx<T>*)d.in + (long long)ring_in<Ctx>(d, a) * d.inStrideA + (long long)b * d.inStrideB;
    pass_foreach<S, 0>(v, t, [&](int i, cplx<T>& r) { r = cswap_if(ld_g<Ctx>(d, in + i), inv); });
    if constexpr (ctx_ring<Ctx>::value) ctx.after_loads();
  }
  cplx<T>* sm = (cplx<T>*)ctx.smem() + pcl * S::SMEM_ELEMS;
  cplx<T>* stage = (cplx<T>*)ctx.smem() + ((PC * S::SMEM_ELEMS + 1) & ~1);   // only touched when d.staged
  fft_passes<S>(v, t, w1, [&](int i) { return sm + S::phys(i); }, [&]() { group_sync<S::G, PC>(ctx, pcl); });
  if constexpr (!GEN) {
    if (valid) {
      const SegDst& sd = d.seg[0];
      cplx<T>* out = (cplx<T>*)sd.base + (long long)ring_out<Ctx>(d, a) * sd.strideA + (long long)b * sd.strideB;
      const long long so = sd.strideI;
      pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { st_g<Ctx>(d, out + (long long)k * so, cswap_if(r, inv)); });
    }
  } else {
    if (valid) {
      if (d.staged) {
        pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { stage[staged_index<PC>(d, S::N, pcl, k)] = cswap_if(r, inv); });
      } else {
        const OpCol col = op_prepare_if(d, a, b);
        pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { store_generic(d, col, cswap_if(r, inv), k, a, b); });
      }
    }
    if (d.staged) staged_flush<T, PC>(d, ctx, stage, S::N, ctx.bid() * PC, S::G * PC);
  }
}

// ================================================================================================
// contiguous r2c of even length N2 = 2*M: one M-point complex FFT on (x[2n], x[2n+1]) pairs plus the
// split step  X[k] = E[k] + W_N2^k O[k].   Reference: fplan_0 = D2Z (src/accfft_gpu.cpp:173-189).
// S is the schedule of the M-point transform; d.tw is the table of N2 = 2M entries.
// Input strides are in units of REAL elements and must be even.
// ================================================================================================
template <class T, class S, int PC, bool GEN, class Ctx>
ACC_HD void contig_r2c_body(const StageDesc& d, Ctx& ctx) {
  constexpr int M = S::N;
  const int tid = ctx.tid();
  const int t = tid % S::G, pcl = tid / S::G;
  const long long idx = ctx.bid() * PC + pcl;
  const bool valid = idx < (long long)d.nA * d.nB;
  const int a = valid ? (int)(idx / d.nB) : 0;
  const int b = valid ? (int)(idx % d.nB) : 0;
  const cplx<T>* tw = (const cplx<T>*)d.tw;
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 2>(w1, t, tw);
  cplx<T> ws[S::cnt(S::NP - 1)];
  ACC_UNROLL
  for (int u = 0; u < S::cnt(S::NP - 1); ++u) ws[u] = tw[(t + u * S::G) % M];
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  if (valid) {
    const cplx<T>* in = (const cplx<T>*)((const T*)d.in + (long long)a * d.inStrideA + (long long)b * d.inStrideB);
    pass_foreach<S, 0>(v, t, [&](int i, cplx<T>& r) { r = ld_g<Ctx>(d, in + i); });
    if constexpr (ctx_ring<Ctx>::value) ctx.after_loads();
  }
  cplx<T>* sm = (cplx<T>*)ctx.smem() + pcl * S::SMEM_ELEMS;
  fft_passes<S>(v, t, w1, [&](int i) { return sm + S::phys(i); }, [&]() { group_sync<S::G, PC>(ctx, pcl); });
  cplx<T>* stage = (cplx<T>*)ctx.smem() + ((PC * S::SMEM_ELEMS + 1) & ~1);   // only touched when d.staged
  // publish Z so every thread can read its mirror partner Z[M-k]
  if constexpr (S::NP > 1) group_sync<S::G, PC>(ctx, pcl);
  pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { sm[S::phys(k)] = r; });
  group_sync<S::G, PC>(ctx, pcl);
  if (valid) {
    constexpr int RL = S::radix(S::NP - 1), NBL = S::nb(S::NP - 1), CNTL = S::cnt(S::NP - 1);
    const SegDst& s0 = d.seg[0];
    cplx<T>* out0 = (cplx<T>*)s0.base + (long long)ring_out<Ctx>(d, a) * s0.strideA + (long long)b * s0.strideB;
    const long long so = s0.strideI;
    OpCol col = {0.0, 0.0, 0};
    if constexpr (GEN) col = op_prepare_if(d, a, b);
    auto put = [&](int kk, cplx<T> y) {
      if constexpr (!GEN) {
        st_g<Ctx>(d, out0 + (long long)kk * so, y);
      } else {
        if (d.staged) stage[staged_index<PC>(d, M + 1, pcl, kk)] = y;
        else store_generic(d, col, y, kk, a, b);
      }
    };
    // split step X[k] = (E + W_N2^k O)/2 for k = j + r*NBL:  W_N2^k = W_N2^j * W_(2 RL)^r, the second factor a
    // compile-time root, so only ws[u] = W_N2^j comes from the table (fetched up front)
    ACC_UNROLL
    for (int u = 0; u < CNTL; ++u) {
      const int j = t + u * S::G;
      if (CNTL * S::G == NBL || j < NBL) {
        const cplx<T> wj = ws[u];
        static_for<0, RL>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          const int k = j + r * NBL;
          const cplx<T> z = v[u * RL + r];
          if (r == 0 && k == 0) {
            put(0, cplx<T>{z.x + z.y, T(0)});
            put(M, cplx<T>{z.x - z.y, T(0)});
          } else {
            const cplx<T> p = sm[S::phys(M - k)];
            const cplx<T> e = {z.x + p.x, z.y - p.y};                       // Z[k] + conj(Z[M-k])
            const cplx<T> o = mul_root<2 * RL, r, T>(cplx<T>{z.y + p.y, p.x - z.x});   // -i (Z[k] - conj(Z[M-k])) * W_(2RL)^r
            const cplx<T> wo = cmul(wj, o);
            put(k, cplx<T>{T(0.5) * (e.x + wo.x), T(0.5) * (e.y + wo.y)});
          }
        });
      }
    }
  }
  if constexpr (GEN) {
    if (d.staged) staged_flush<T, PC>(d, ctx, stage, M + 1, ctx.bid() * PC, S::G * PC);
  }
}

// ================================================================================================
// contiguous c2r of even length N2 = 2*M (unnormalised, like cufftExecZ2D / fftw c2r):
//   Z'[k] = (X[k] + conj(X[M-k])) + i conj(W_N2^k) (X[k] - conj(X[M-k])),  z = IDFT_M(Z'),  x[2n] + i x[2n+1] = z[n]
// Imaginary parts of X[0] and X[M] are ignored, as FFTW / cuFFT do.   Reference: iplan_0 = Z2D
// (src/accfft_gpu.cpp:191-206).  Output strides (seg[0]) are in units of REAL elements and must be even.
// ================================================================================================
template <class T, class S, int PC, bool GEN, class Ctx>
ACC_HD void contig_c2r_body(const StageDesc& d, Ctx& ctx) {
  constexpr int M = S::N;
  const int tid = ctx.tid();
  const int t = tid % S::G, pcl = tid / S::G;
  const long long idx = ctx.bid() * PC + pcl;
  const bool valid = idx < (long long)d.nA * d.nB;
  const int a = valid ? (int)(idx / d.nB) : 0;
  const int b = valid ? (int)(idx % d.nB) : 0;
  cplx<T>* sm = (cplx<T>*)ctx.smem() + pcl * S::SMEM_ELEMS;
  const cplx<T>* tw = (const cplx<T>*)d.tw;
  constexpr int R0 = S::radix(0), NB0 = S::nb(0), CNT0 = S::cnt(0);
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 2>(w1, t, tw);
  // twiddles W_N2^j of the pre-processing step: fetched up front (the fetch overlaps the data loads) when that costs at most
  // eight registers, else where they are used (the table is L1-resident; measured: the 128- and 1024-point double kernels
  // spill otherwise, the 256-point one loses 5 % when it fetches late)
  constexpr bool WS_EARLY = CNT0 * (int)sizeof(cplx<T>) <= 32;
  cplx<T> ws[WS_EARLY ? CNT0 : 1];
  if constexpr (WS_EARLY) {
    ACC_UNROLL
    for (int u = 0; u < CNT0; ++u) ws[u] = tw[(t + u * S::G) % M];
  }
  // The M + 1 input coefficients: every thread loads the ones it needs as X[k] -- the operands of pass 0, k = j + r NB0, all
  // loads issued before the first use -- keeps them in registers and publishes them to shared memory, from which it then
  // only fetches the mirror partners X[M - k] (X[M], the partner of k = 0, is loaded by the first thread of the pencil).
  cplx<T> v[S::NREG];
  ACC_UNROLL
  for (int i = 0; i < S::NREG; ++i) v[i] = {T(0), T(0)};
  cplx<T> xM = {T(0), T(0)};
  if (valid) {
    const cplx<T>* in = (const cplx<T>*)d.in + (long long)ring_in<Ctx>(d, a) * d.inStrideA + (long long)b * d.inStrideB;
    pass_foreach<S, 0>(v, t, [&](int k, cplx<T>& r) { r = ld_g<Ctx>(d, in + k); });
    if (t == 0) xM = ld_g<Ctx>(d, in + M);
    if constexpr (ctx_ring<Ctx>::value) ctx.after_loads();
    if constexpr (GEN) {
      if (d.op != OP_NONE) {
        const OpCol col = op_prepare(d, a, b);
        pass_foreach<S, 0>(v, t, [&](int k, cplx<T>& r) { r = apply_op(d, col, r, k); });
        if (t == 0) xM = apply_op(d, col, xM, M);
      }
    }
  }
  pass_foreach<S, 0>(v, t, [&](int k, cplx<T>& r) { sm[S::phys(k)] = r; });
  if (t == 0) sm[S::phys(M)] = xM;
  group_sync<S::G, PC>(ctx, pcl);
  // Z'[k] = (X[k] + conj(X[M-k])) + i conj(W_N2^k) (X[k] - conj(X[M-k])) for k = j + r*NB0, with
  // W_N2^k = W_N2^j * W_(2 R0)^r (compile-time second factor)
  ACC_UNROLL
  for (int u = 0; u < CNT0; ++u) {
    const int j = t + u * S::G;
    if (CNT0 * S::G == NB0 || j < NB0) {
      cplx<T> wj;
      if constexpr (WS_EARLY) wj = ws[u];
      else wj = tw[j % M];
      static_for<0, R0>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int k = j + r * NB0;
        const cplx<T> A = v[u * R0 + r];
        const cplx<T> B = sm[S::phys(M - k)];
        cplx<T> z;
        if (r == 0 && k == 0) {
          z = {A.x + B.x, A.x - B.x};
        } else {
          const cplx<T> e = {A.x + B.x, A.y - B.y};                             // X[k] + conj(X[M-k])
          const cplx<T> D = mul_root<2 * R0, -r, T>(cplx<T>{A.x - B.x, A.y + B.y});   // (X[k] - conj(X[M-k])) conj(W_(2R0)^r)
          // i * conj(wj) * D
          const cplx<T> q = {-(wj.x * D.y - wj.y * D.x), wj.x * D.x + wj.y * D.y};
          z = cadd(e, q);
        }
        v[u * R0 + r] = cswap(z);   // inverse transform = swap . forward . swap
      });
    }
  }
  group_sync<S::G, PC>(ctx, pcl);
  fft_passes<S>(v, t, w1, [&](int i) { return sm + S::phys(i); }, [&]() { group_sync<S::G, PC>(ctx, pcl); });
  if (valid) {
    const SegDst& sd = d.seg[0];
    cplx<T>* out = (cplx<T>*)((T*)sd.base + (long long)a * sd.strideA + (long long)b * sd.strideB);
    if (GEN && d.accumulate) {   // divergence: this component is added to what the earlier ones left in the output
      pass_foreach<S, S::NP - 1>(v, t, [&](int nidx, cplx<T>& r) {
        const cplx<T> o = out[nidx];
        out[nidx] = {o.x + r.y, o.y + r.x};
      });
    } else {
      pass_foreach<S, S::NP - 1>(v, t, [&](int nidx, cplx<T>& r) { st_g<Ctx>(d, out + nidx, cswap(r)); });
    }
  }
}

}  // namespace accfft_b200
