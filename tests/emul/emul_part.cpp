// emul_part.cpp -- TEST INFRASTRUCTURE ONLY.  The kernel-body instantiations of the CPU emulator, compiled EMU_NPARTS times
// (-DEMU_PART=k): part k instantiates the stage bodies for every EMU_NPARTS-th schedule of schedules.def, so that the
// emulator builds in parallel (one translation unit with all 91 lengths x 9 kernel bodies x 2 precisions takes ~5 minutes).
#include "emul_fiber.hpp"

#ifndef EMU_PART
#error "compile with -DEMU_PART=k -DEMU_NPARTS=n"
#endif

// ---------------------------------------------------------------------------------------------------
// dispatch (same schedule table as the CUDA backend)
// ---------------------------------------------------------------------------------------------------
template <class T, class S>
static void emu_launch_sched(const Launch& l) {
  using C = Cfg<T, S>;
  const StageDesc& d = l.desc;
  const bool gen = !stage_is_simple(d);   // the same choice the CUDA launchers make (launch_impl.cuh: launch_pair)
  switch (l.kind) {
    case K_STRIDED_C2C: {
      if constexpr (sizeof(T) == 4) {   // single precision: the pair kernel, as in launch_impl.cuh
        using PC = PairCfg<S>;
        if (d.inStrideB == 1) {
          const long long nbp = (long long)d.nA * ((d.nB + d.colShift + 2 * PC::TCP - 1) / (2 * PC::TCP));
          run_grid(nbp, PC::THREADS, PC::SMEM, [&](EmuCtx& c) {
            if (gen) strided_c2c_pair_body<S, PC::TCP, true>(d, c);
            else strided_c2c_pair_body<S, PC::TCP, false>(d, c);
          });
          break;
        }
      }
      const long long nb = (long long)d.nA * ((d.nB + d.colShift + C::TC - 1) / C::TC);
      run_grid(nb, C::STRIDED_THREADS, C::STRIDED_SMEM, [&](EmuCtx& c) {
        if (gen) strided_c2c_body<T, S, C::TC, true>(d, c);
        else strided_c2c_body<T, S, C::TC, false>(d, c);
      });
      break;
    }
    case K_CONTIG_C2C: {
      const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
      StageDesc ds = d;
      ds.staged = (d.nseg > 1 && d.op == OP_NONE && d.accumulate == 0) ? 1 : 0;
      run_grid(nb, C::CONTIG_THREADS, C::CONTIG_STAGED_SMEM, [&](EmuCtx& c) {
        if (gen) contig_c2c_body<T, S, C::PC, true>(ds, c);
        else contig_c2c_body<T, S, C::PC, false>(ds, c);
      });
      break;
    }
    case K_CONTIG_R2C: {
      const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
      StageDesc ds = d;
      ds.staged = (d.nseg > 1 && d.op == OP_NONE && d.accumulate == 0) ? 1 : 0;
      run_grid(nb, C::CONTIG_THREADS, C::CONTIG_STAGED_SMEM, [&](EmuCtx& c) {
        if (gen) contig_r2c_body<T, S, C::PC, true>(ds, c);
        else contig_r2c_body<T, S, C::PC, false>(ds, c);
      });
      break;
    }
    case K_CONTIG_C2R: {
      const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
      run_grid(nb, C::CONTIG_THREADS, C::CONTIG_SMEM, [&](EmuCtx& c) {
        if (gen) contig_c2r_body<T, S, C::PC, true>(d, c);
        else contig_c2r_body<T, S, C::PC, false>(d, c);
      });
      break;
    }
    case K_CONTIG_DERIV: {
      const long long npairs = ((long long)d.nA * d.nB + 1) / 2;
      run_grid((npairs + C::PC - 1) / C::PC, C::CONTIG_THREADS, C::CONTIG_SMEM, [&](EmuCtx& c) { contig_deriv_body<T, S, C::PC>(d, c); });
      break;
    }
    case K_STRIDED_R2C:
    case K_STRIDED_C2R:
    case K_STRIDED_DERIV: {
      const long long nb = (long long)d.nA * ((d.nB + 2 * C::TC - 1) / (2 * C::TC));
      run_grid(nb, C::STRIDED_THREADS, C::STRIDED_SMEM, [&](EmuCtx& c) {
        if (l.kind == K_STRIDED_R2C) strided_r2c_body<T, S, C::TC>(d, c);
        else if (l.kind == K_STRIDED_C2R) strided_c2r_body<T, S, C::TC>(d, c);
        else strided_deriv_body<T, S, C::TC>(d, c);
      });
      break;
    }
    default:
      fprintf(stderr, "emul: unsupported kind %d\n", (int)l.kind);
      abort();
  }
}


// position of a length in the schedule table
static constexpr int kLengths[] = {
#define X(N_, G_, R0, R1, R2, R3) N_,
#include ACCFFT_SCHEDULES_FILE
#undef X
};
static constexpr int index_of(int n) {
  int i = 0;
  for (int v : kLengths) {
    if (v == n) return i;
    ++i;
  }
  return -1;
}

template <int PART>
static bool launch_part(const Launch& l) {
#define X(N_, G_, R0, R1, R2, R3)                                                        \
  if constexpr (index_of(N_) % EMU_NPARTS == PART) {                                      \
    if (l.fftN == N_) {                                                                   \
      if (l.prec == PREC_DOUBLE) emu_launch_sched<double, Sched<N_, G_, R0, R1, R2, R3>>(l); \
      else emu_launch_sched<float, Sched<N_, G_, R0, R1, R2, R3>>(l);                        \
      return true;                                                                        \
    }                                                                                     \
  }
#include ACCFFT_SCHEDULES_FILE
#undef X
  return false;
}

#define EMU_CAT2(a, b) a##b
#define EMU_CAT(a, b) EMU_CAT2(a, b)
bool EMU_CAT(emu_launch_part_, EMU_PART)(const Launch& l) { return launch_part<EMU_PART>(l); }
