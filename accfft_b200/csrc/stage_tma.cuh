// stage_tma.cuh -- TMA-pipelined, persistent variant of the strided c2c stage kernel (Y and X passes; reference:
// the per-slab cufftExecZ2Z loops of src/accfft_gpu.cpp:499-539, :541-579 plus the pack kernels around them).
//
// A CTA walks over column tiles (tile = blockIdx.x + i * gridDim.x).  One elected thread keeps NST tiles in flight with
// cp.async.bulk.tensor (a 3-D tensor map over the stage input: columns x transform index x outer index, box = one
// N x 128-byte tile split into <= 256-row boxes) that complete on mbarriers; the compute threads take pass-0 operands
// from the landed tile, run the same register-resident Stockham passes as strided_c2c_body IN PLACE in that tile buffer,
// and store the last pass straight from registers (fused pack: peer-mapped destinations, spectral operators).
// Out-of-range columns (ragged right edge, negative columns of a shifted tile grid) are zero-filled by the TMA unit.
// CUDA only; the CPU emulator keeps exercising strided_c2c_body, which shares every helper used here.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include "stage.cuh"

namespace accfft_b200 {

#ifndef ACCFFT_TMA_SMEM_BUDGET
#define ACCFFT_TMA_SMEM_BUDGET (204 * 1024)
#endif
#ifndef ACCFFT_TMA_MAX_STAGES
#define ACCFFT_TMA_MAX_STAGES 3
#endif
#ifndef ACCFFT_TMA_SPLIT
#define ACCFFT_TMA_SPLIT 0   // measured (profiles/variants_r05.txt): the two-round exchange costs more than the occupancy it buys
#endif
#ifndef ACCFFT_TMA_BOXR
#define ACCFFT_TMA_BOXR 256
#endif
#ifndef ACCFFT_TMA_TARGET_THREADS
#define ACCFFT_TMA_TARGET_THREADS 1024
#endif

// Shared memory of one CTA:  NST landing buffers (dense N x TC tiles, written by the TMA unit, read once by pass 0)
// + ONE exchange buffer for the Stockham passes.  With SPLIT the exchange moves real and imaginary parts in two rounds
// through a buffer of N x TC REALS (padded rows, see Sched::phys), which halves its size: a 512-point double tile then
// costs 64 KB + 37 KB, so two 512-thread CTAs stay resident per SM, each with its next tile already in flight.
template <class T, class S, int TC>
struct TmaCfg {
  static constexpr int ES = (int)sizeof(cplx<T>);
  static constexpr int ROW_BYTES = TC * ES;
  static constexpr int boxrows() {
    int r = S::N < ACCFFT_TMA_BOXR ? S::N : ACCFFT_TMA_BOXR;
    while (S::N % r) --r;
    return r;
  }
  static constexpr int BOXR = boxrows();
  static constexpr int NBOX = S::N / BOXR;
  static constexpr int TILE_BYTES = S::N * ROW_BYTES;
  static constexpr int THREADS = S::G * TC;
  static constexpr bool SPLIT = ACCFFT_TMA_SPLIT != 0 && S::NP > 1;
  static constexpr int XCH_BYTES = S::NP > 1 ? ((S::SMEM_ELEMS * TC * (SPLIT ? ES / 2 : ES) + 127) / 128) * 128 : 0;
  static constexpr int MISC_BYTES = 128 + 8 * ACCFFT_TMA_MAX_STAGES;
  // resident CTAs per SM: aim at ~1024 compute threads, every CTA needs one landing buffer + the exchange buffer
  static constexpr int ctas() {
    int want = (ACCFFT_TMA_TARGET_THREADS + THREADS - 1) / THREADS;
    want = want > 8 ? 8 : want;
    int fit = ACCFFT_TMA_SMEM_BUDGET / (TILE_BYTES + XCH_BYTES + MISC_BYTES);
    int c = want < fit ? want : fit;
    return c < 1 ? 1 : c;
  }
  static constexpr int CTAS = ctas();
  // landing buffers per CTA: whatever the budget still allows (deeper prefetch), at least 1
  static constexpr int nst() {
    int n = (ACCFFT_TMA_SMEM_BUDGET / CTAS - XCH_BYTES - MISC_BYTES) / TILE_BYTES;
    n = n > ACCFFT_TMA_MAX_STAGES ? ACCFFT_TMA_MAX_STAGES : n;
    return n < 1 ? 1 : n;
  }
  static constexpr int NST = nst();
  static constexpr int SMEM = NST * TILE_BYTES + XCH_BYTES + MISC_BYTES;
  // every box lands at  q * BOXR * ROW_BYTES  inside the tile: the TMA unit needs that offset 128-byte aligned
  // (lengths whose largest divisor <= 256 is odd or only twice-even with narrow rows, e.g. 1500 with 32-byte rows, decline)
  static constexpr bool OK = S::N >= 32 && ROW_BYTES % 16 == 0 && SMEM <= 220 * 1024 && (NBOX == 1 || (BOXR * ROW_BYTES) % 128 == 0) &&
                             TILE_BYTES % 128 == 0;
};

// Stockham passes 1..NP-1 with the inter-pass exchange done through a buffer of REALS in two rounds (re, then im).
// On entry the registers hold pass-0 operands (gather order) and the exchange buffer is free; xaddr(i) -> &X[phys(i)][c].
template <class S, class T, class XAddr>
__device__ __forceinline__ void fft_passes_split(cplx<T>* v, int t, const cplx<T>* w1, XAddr xaddr) {
  pass_compute<S, 0>(v, t, w1);
  static_for<1, S::NP>([&](auto pc) {
    constexpr int P = decltype(pc)::value;
    constexpr int RS = S::radix(P - 1), NSS = S::ns(P - 1), NBS = S::nb(P - 1), CNTS = S::cnt(P - 1);
    constexpr int RG = S::radix(P), NBG = S::nb(P), CNTG = S::cnt(P);
    if constexpr (P > 1) __syncthreads();   // previous round's gathers are done
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (half == 1) __syncthreads();
#pragma unroll
      for (int u = 0; u < CNTS; ++u) {
        const int j = t + u * S::G;
        if (CNTS * S::G == NBS || j < NBS) {
          const int j0 = (j / NSS) * NSS * RS + (j % NSS);
#pragma unroll
          for (int r = 0; r < RS; ++r) *xaddr(j0 + r * NSS) = half == 0 ? v[u * RS + r].x : v[u * RS + r].y;
        }
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < CNTG; ++u) {
        const int j = t + u * S::G;
        if (CNTG * S::G == NBG || j < NBG) {
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            const T val = *xaddr(j + r * NBG);
            if (half == 0) v[u * RG + r].x = val;
            else v[u * RG + r].y = val;
          }
        }
      }
    }
    pass_compute<S, P>(v, t, w1);
  });
}

__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void tma_mbar_expect(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* sdst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          (unsigned)__cvta_generic_to_shared(sdst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}

template <class T, class S, int TC, bool GEN>
__global__ void __launch_bounds__(TmaCfg<T, S, TC>::THREADS, TmaCfg<T, S, TC>::CTAS) k_strided_tma(const __grid_constant__ StageDesc d,
                                                                           const __grid_constant__ CUtensorMap tmap,
                                                                           long long ntiles) {
  using C = TmaCfg<T, S, TC>;
  constexpr int G = S::G, NST = C::NST;
  constexpr int R0 = S::radix(0), NB0 = S::nb(0), CNT0 = S::cnt(0);
  constexpr int RL = S::radix(S::NP - 1), NBL = S::nb(S::NP - 1), CNTL = S::cnt(S::NP - 1);
  extern __shared__ unsigned char accfft_tma_smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)accfft_tma_smem_raw + 127) & ~(uintptr_t)127);
  unsigned char* xch = base + (size_t)NST * C::TILE_BYTES;
  uint64_t* full = (uint64_t*)(xch + C::XCH_BYTES);
  const int tid = (int)threadIdx.x;
  const int c0 = tid % TC, t0 = tid / TC;
  const int tpa = (d.nB + d.colShift + TC - 1) / TC;
  const bool inv = d.inv != 0;
  constexpr int E = C::ES / 8;   // tensor-map elements (u64) per complex value

  auto issue = [&](int i) {   // elected thread: start the loads of this CTA's i-th tile
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int s = i % NST;
    const int a = tile / tpa;
    const int col0 = (tile % tpa) * TC - d.colShift;
    tma_mbar_expect(&full[s], (unsigned)C::TILE_BYTES);
#pragma unroll
    for (int q = 0; q < C::NBOX; ++q)
      tma_load_3d(base + (size_t)s * C::TILE_BYTES + (size_t)q * C::BOXR * C::ROW_BYTES, &tmap, col0 * E, q * C::BOXR, a, &full[s]);
  };
  const int nmine = (int)((ntiles - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x);
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) tma_mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < NST && i < nmine; ++i) issue(i);
  }
  cplx<T> w1[S::NTW];
  load_base_twiddles<S, 1>(w1, t0, (const cplx<T>*)d.tw);
  __syncthreads();

#pragma unroll 1
  for (int i = 0; i < nmine; ++i) {
    // opaque copies of the thread coordinates: keeps the compiler from hoisting the (many) loop-invariant shared-memory
    // addresses of the passes out of the tile loop, which would cost more registers than recomputing them
    int c = c0, t = t0;
    asm volatile("" : "+r"(c), "+r"(t));
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    const int s = i % NST;
    const int a = tile / tpa;
    const int b = (tile % tpa) * TC + c - d.colShift;
    const bool valid = b >= 0 && b < d.nB;
    cplx<T>* sm = (cplx<T>*)(base + (size_t)s * C::TILE_BYTES);
    tma_mbar_wait(&full[s], (unsigned)((i / NST) & 1));
    cplx<T> v[S::NREG];
#pragma unroll
    for (int k = 0; k < S::NREG; ++k) v[k] = {T(0), T(0)};
#pragma unroll
    for (int u = 0; u < CNT0; ++u) {
      const int j = t + u * G;
      if (CNT0 * G == NB0 || j < NB0) {
#pragma unroll
        for (int r = 0; r < R0; ++r) v[u * R0 + r] = sm[(j + r * NB0) * TC + c];
      }
    }
    OpCol col = {0.0, 0.0, 0};
    if constexpr (GEN) {
      if (valid) col = op_prepare_if(d, a, b);
      if (d.opOnLoad) pass_foreach<S, 0>(v, t, [&](int k, cplx<T>& r) { r = apply_op(d, col, r, k); });
    }
    if (inv) {
#pragma unroll
      for (int k = 0; k < S::NREG; ++k) v[k] = cswap(v[k]);
    }
    // every pass-0 operand is in registers: hand the landing buffer back to the TMA unit for tile i + NST
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0 && i + NST < nmine) issue(i + NST);
    if constexpr (C::SPLIT) {
      T* xs = (T*)xch;
      fft_passes_split<S>(v, t, w1, [&](int k) { return xs + S::phys(k) * TC + c; });
    } else {
      cplx<T>* xs = (cplx<T>*)xch;
      fft_passes<S>(v, t, w1, [&](int k) { return xs + S::phys(k) * TC + c; }, [&]() { __syncthreads(); });
    }
    if (valid) {
      if (inv) {
#pragma unroll
        for (int k = 0; k < S::NREG; ++k) v[k] = cswap(v[k]);
      }
      if constexpr (!GEN) {
        const SegDst& sd = d.seg[0];
        cplx<T>* out = (cplx<T>*)sd.base + (long long)a * sd.strideA + (long long)b * sd.strideB;
        const long long so = sd.strideI;
        const long long step = (long long)NBL * so;
#pragma unroll
        for (int u = 0; u < CNTL; ++u) {
          const int j = t + u * G;
          if (CNTL * G == NBL || j < NBL) {
            cplx<T>* p = out + (long long)j * so;
#pragma unroll
            for (int r = 0; r < RL; ++r) st_stream(p + r * step, v[u * RL + r]);
          }
        }
      } else {
        pass_foreach<S, S::NP - 1>(v, t, [&](int k, cplx<T>& r) { store_generic(d, col, r, k, a, b); });
      }
    }
  }
}

}  // namespace accfft_b200
