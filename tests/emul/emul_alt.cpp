// emul_alt.cpp -- TEST INFRASTRUCTURE ONLY.  The CUDA build gives some contiguous kernels (k_tu.cu kinds 1 = c2c, 2 = r2c,
// 3 = c2r) their own schedules per kind / precision (schedules.def: KEEP_C, KEEP_FC).  This unit, compiled once per
// (EMU_ALT_KIND, EMU_ALT_PREC), instantiates exactly those -- every line of schedules.def that differs from the default table
// under -DACCFFT_KIND / -DACCFFT_PREC -- so that the CPU emulator runs the same schedule the GPU runs for that stage.
#include "emul_fiber.hpp"

#if !defined(EMU_ALT_KIND) || !defined(EMU_ALT_PREC)
#error "compile with -DEMU_ALT_KIND=1|2|3 -DEMU_ALT_PREC=0|1"
#endif

namespace {
struct SchedRow {
  int n, g, r0, r1, r2, r3;
};
// the default table (what every other kernel kind and the emulator's main parts use)
constexpr SchedRow kDefault[] = {
#define X(N_, G_, R0, R1, R2, R3) {N_, G_, R0, R1, R2, R3},
#include ACCFFT_SCHEDULES_FILE
#undef X
};
constexpr bool differs(int n, int g, int r0, int r1, int r2, int r3) {
  for (const SchedRow& d : kDefault)
    if (d.n == n) return d.g != g || d.r0 != r0 || d.r1 != r1 || d.r2 != r2 || d.r3 != r3;
  return true;
}
}  // namespace

// from here on schedules.def is read the way translation unit k<EMU_ALT_KIND>_p<EMU_ALT_PREC> of the CUDA build reads it
#define ACCFFT_KIND EMU_ALT_KIND
#define ACCFFT_PREC EMU_ALT_PREC

#if EMU_ALT_PREC == 0
typedef double alt_real;
#else
typedef float alt_real;
#endif

template <class S>
static void alt_run(const Launch& l) {
  using T = alt_real;
  using C = Cfg<T, S>;
  const StageDesc& d = l.desc;
  const bool gen = !stage_is_simple(d);
  const long long nb = ((long long)d.nA * d.nB + C::PC - 1) / C::PC;
#if EMU_ALT_KIND == 1
  StageDesc ds = d;
  ds.staged = (d.nseg > 1 && d.op == OP_NONE && d.accumulate == 0) ? 1 : 0;
  run_grid(nb, C::CONTIG_THREADS, C::CONTIG_STAGED_SMEM, [&](EmuCtx& c) {
    if (gen) contig_c2c_body<T, S, C::PC, true>(ds, c);
    else contig_c2c_body<T, S, C::PC, false>(ds, c);
  });
#elif EMU_ALT_KIND == 2
  StageDesc ds = d;
  ds.staged = (d.nseg > 1 && d.op == OP_NONE && d.accumulate == 0) ? 1 : 0;
  run_grid(nb, C::CONTIG_THREADS, C::CONTIG_STAGED_SMEM, [&](EmuCtx& c) {
    if (gen) contig_r2c_body<T, S, C::PC, true>(ds, c);
    else contig_r2c_body<T, S, C::PC, false>(ds, c);
  });
#else
  run_grid(nb, C::CONTIG_THREADS, C::CONTIG_SMEM, [&](EmuCtx& c) {
    if (gen) contig_c2r_body<T, S, C::PC, true>(d, c);
    else contig_c2r_body<T, S, C::PC, false>(d, c);
  });
#endif
}

#define EMU_CAT4_(a, b, c, d) a##b##c##d
#define EMU_CAT4(a, b, c, d) EMU_CAT4_(a, b, c, d)
// true when this (kind, precision) has its own schedule for l.fftN and the stage was run with it
bool EMU_CAT4(emu_launch_alt_k, EMU_ALT_KIND, _p, EMU_ALT_PREC)(const Launch& l) {
#define X(N_, G_, R0, R1, R2, R3)                              \
  if constexpr (differs(N_, G_, R0, R1, R2, R3)) {             \
    if (l.fftN == N_) {                                        \
      alt_run<Sched<N_, G_, R0, R1, R2, R3>>(l);               \
      return true;                                             \
    }                                                          \
  }
#include ACCFFT_SCHEDULES_FILE
#undef X
  return false;
}

// how many lengths this unit overrides (the tests assert that the table is not silently empty)
int EMU_CAT4(emu_alt_count_k, EMU_ALT_KIND, _p, EMU_ALT_PREC)() {
  int n = 0;
#define X(N_, G_, R0, R1, R2, R3) \
  if (differs(N_, G_, R0, R1, R2, R3)) ++n;
#include ACCFFT_SCHEDULES_FILE
#undef X
  return n;
}
