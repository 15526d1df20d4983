// emul.cpp -- TEST INFRASTRUCTURE ONLY.  A CPU emulator for the stage kernels of accfft_b200.
//
// The kernel bodies in accfft_b200/csrc/stage.cuh are __host__ __device__ templates over an execution
// context.  Here each CUDA thread of a CTA runs as a ucontext fiber (__syncthreads == yield to a
// round-robin scheduler), each rank of the process grid runs as an OS thread, and the inter-rank
// barrier is a pthread barrier.  The plan / descriptor code (plan.cpp) is the very same code the CUDA
// backend uses, so index math, uneven splits and peer addressing are validated without a GPU.
// Nothing in the product links against this file; libaccfft_b200.so has no CPU execution path.
#include <pthread.h>
#include <string.h>
#include <ucontext.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <thread>
#include <vector>

#include "../../accfft_b200/csrc/kernel_cfg.hpp"
#include "../../accfft_b200/csrc/plan.hpp"

using namespace accfft_b200;

#include "emul_fiber.hpp"

// kernel-body instantiations live in emul_part.cpp (compiled EMU_NPARTS times, one slice of the schedule table each)
#ifndef EMU_NPARTS
#define EMU_NPARTS 8
#endif
bool emu_launch_part_0(const Launch&);
bool emu_launch_part_1(const Launch&);
bool emu_launch_part_2(const Launch&);
bool emu_launch_part_3(const Launch&);
bool emu_launch_part_4(const Launch&);
bool emu_launch_part_5(const Launch&);
bool emu_launch_part_6(const Launch&);
bool emu_launch_part_7(const Launch&);
// per-kind / per-precision schedules of the contiguous kernels (emul_alt.cpp; mirrors k_tu.cu kinds 1, 2, 3)
bool emu_launch_alt_k1_p0(const Launch&);
bool emu_launch_alt_k1_p1(const Launch&);
bool emu_launch_alt_k2_p0(const Launch&);
bool emu_launch_alt_k2_p1(const Launch&);
bool emu_launch_alt_k3_p0(const Launch&);
bool emu_launch_alt_k3_p1(const Launch&);
int emu_alt_count_k1_p0();
int emu_alt_count_k1_p1();
int emu_alt_count_k2_p0();
int emu_alt_count_k2_p1();
int emu_alt_count_k3_p0();
int emu_alt_count_k3_p1();
static long long g_alt_launches = 0;
static_assert(EMU_NPARTS == 8, "emul.cpp declares eight parts");

static bool emu_supported(int N) {
#define X(N_, G_, R0, R1, R2, R3) \
  if (N == N_) return true;
#include ACCFFT_SCHEDULES_FILE
#undef X
  return false;
}

// stage without a transform (partial-axis executes): same element walk as k_copy_stage in engine.cu
template <class T>
static void emu_copy_stage(const Launch& l) {
  const StageDesc& d = l.desc;
  const cplx<T>* in = (const cplx<T>*)d.in;
  for (int a = 0; a < d.nA; ++a)
    for (int b = 0; b < d.nB; ++b)
      for (int k = 0; k < l.fftN; ++k)
        store_generic(d, in[(long long)a * d.inStrideA + (long long)b * d.inStrideB + (long long)k * d.inStrideI], k, a, b);
}

static void emu_launch(const Launch& l) {
  if (l.kind == K_COPY_REAL) {   // same element walk as k_copy_real in engine.cu
    const long long total = (long long)l.desc.nA * l.desc.nB * l.copyN;
    for (long long idx = 0; idx < total; ++idx) {
      if (l.prec == PREC_DOUBLE) copy_real_element<double>(l.desc, idx, l.copyN);
      else copy_real_element<float>(l.desc, idx, l.copyN);
    }
    return;
  }
  if (l.kind == K_COPY) {
    if (l.prec == PREC_DOUBLE) emu_copy_stage<double>(l);
    else emu_copy_stage<float>(l);
    return;
  }
  {   // the schedule the CUDA build gives this kind / precision, where it has its own
    bool alt = false;
    const bool dbl = l.prec == PREC_DOUBLE;
    if (l.kind == K_CONTIG_C2C) alt = dbl ? emu_launch_alt_k1_p0(l) : emu_launch_alt_k1_p1(l);
    else if (l.kind == K_CONTIG_R2C) alt = dbl ? emu_launch_alt_k2_p0(l) : emu_launch_alt_k2_p1(l);
    else if (l.kind == K_CONTIG_C2R) alt = dbl ? emu_launch_alt_k3_p0(l) : emu_launch_alt_k3_p1(l);
    if (alt) {
      __atomic_add_fetch(&g_alt_launches, 1, __ATOMIC_RELAXED);
      return;
    }
  }
  if (emu_launch_part_0(l) || emu_launch_part_1(l) || emu_launch_part_2(l) || emu_launch_part_3(l) || emu_launch_part_4(l) ||
      emu_launch_part_5(l) || emu_launch_part_6(l) || emu_launch_part_7(l))
    return;
  fprintf(stderr, "emul: no schedule for N=%d\n", l.fftN);
  abort();
}

// ---------------------------------------------------------------------------------------------------
// backend + world
// ---------------------------------------------------------------------------------------------------
struct EmuWorld;
struct EmuBackend : Backend {
  EmuWorld* w;
  std::map<std::pair<int, int>, std::vector<char>> tables;
  bool supported(Kind, Prec, int fftN) override { return emu_supported(fftN); }
  const void* twiddle(Prec prec, int M) override {
    auto key = std::make_pair((int)prec, M);
    auto it = tables.find(key);
    if (it != tables.end()) return it->second.data();
    std::vector<char> buf((size_t)M * (prec == PREC_DOUBLE ? 16 : 8));
    for (int k = 0; k < M; ++k) {
      const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)M;
      if (prec == PREC_DOUBLE) {
        ((double*)buf.data())[2 * k] = (double)cosl(ang);
        ((double*)buf.data())[2 * k + 1] = (double)sinl(ang);
      } else {
        ((float*)buf.data())[2 * k] = (float)cosl(ang);
        ((float*)buf.data())[2 * k + 1] = (float)sinl(ang);
      }
    }
    auto& ref = tables[key] = std::move(buf);
    return ref.data();
  }
  int launch(const Launch& l) override {
    if (!l.dry) emu_launch(l);
    return 0;
  }
  int barrier() override;
  // software-pipelined stages: the emulator executes in program order, which satisfies every fork / hand / join
  // dependency, so only the chunked descriptors (Plan::forward_zy_pipelined) are exercised here
  int pipeline_chunks() override {
    const char* c = getenv("EMU_PIPELINE_CHUNKS");
    return c ? atoi(c) : 1;
  }
  int l2_planes(const Plan&) override {
    const char* c = getenv("EMU_L2_PLANES");
    return c ? atoi(c) : 0;
  }
  // copy-engine exchange (Plan::forward_slab_ce / inverse_slab_ce): copies are row memcpys executed in program order
  std::vector<char> stage_buf;
  int ce_chunks() override {
    const char* c = getenv("EMU_CE_CHUNKS");
    return c ? atoi(c) : 0;
  }
  int copy_lanes() override {
    const char* c = getenv("EMU_CE_LANES");
    return c ? atoi(c) : 2;
  }
  void* staging() override { return stage_buf.empty() ? nullptr : stage_buf.data(); }
  int copy2d(int lane, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) override {
    if (lane < 1 || lane > copy_lanes()) return -1;
    for (size_t r = 0; r < height; ++r) memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width);
    return 0;
  }
};

struct EmuWorld {
  int n[3], np[2], nprocs;
  bool r2c;
  Prec prec;
  std::vector<EmuBackend> be;
  std::vector<Plan> plans;
  std::vector<std::vector<char>> bufA, bufB, bufA2, bufB2, spec;
  std::vector<long long> strides;   // batched plans: bytes between the field sets of a rank's buffers
  pthread_barrier_t bar;
};
int EmuBackend::barrier() {
  pthread_barrier_wait(&w->bar);
  return 0;
}

extern "C" {

void* emu_world_create(const int* n, int P0, int P1, int r2c, int prec) {
  EmuWorld* w = new EmuWorld();
  for (int i = 0; i < 3; ++i) w->n[i] = n[i];
  w->np[0] = P0;
  w->np[1] = P1;
  w->nprocs = P0 * P1;
  w->r2c = r2c != 0;
  w->prec = (Prec)prec;
  w->be.resize(w->nprocs);
  w->plans.resize(w->nprocs);
  w->bufA.resize(w->nprocs);
  w->bufB.resize(w->nprocs);
  w->bufA2.resize(w->nprocs);
  w->bufB2.resize(w->nprocs);
  w->spec.resize(w->nprocs);
  pthread_barrier_init(&w->bar, nullptr, w->nprocs);
  const int sz = prec == PREC_DOUBLE ? 8 : 4;
  for (int r = 0; r < w->nprocs; ++r) {
    int coords[2] = {r / P1, r % P1};
    w->be[r].w = w;
    Plan& p = w->plans[r];
    p.g = make_geometry(w->n, w->np, coords, w->r2c, sz);
    p.prec = w->prec;
    p.inplace = false;
    p.be = &w->be[r];
    p.nprocs = w->nprocs;
    // +64 bytes so that out-of-range writes are more likely to be caught by the canary check below
    const size_t bb1 = (size_t)Plan::buffer_bytes(w->n, w->np, coords, w->r2c, w->prec, w->nprocs);
    // batched plans (EMU_BATCH=<fields>): every buffer holds `batch` field sets, one every `stride` bytes
    const char* benv = getenv("EMU_BATCH");
    const int batch = benv ? std::max(1, atoi(benv)) : 1;
    const size_t stride = (bb1 + 255) / 256 * 256;
    const size_t bb = batch > 1 ? stride * (size_t)batch : bb1;
    p.batch = batch;
    w->strides.resize(w->nprocs);
    w->strides[r] = (long long)stride;
    w->bufA[r].assign(bb + 64, (char)0xEE);
    w->bufB[r].assign(bb + 64, (char)0xEE);
    w->bufA2[r].assign(bb + 64, (char)0xEE);
    w->bufB2[r].assign(bb + 64, (char)0xEE);
    w->spec[r].assign(bb + 64, (char)0xEE);
    w->be[r].stage_buf.assign(bb + 64, (char)0xEE);
  }
  for (int r = 0; r < w->nprocs; ++r) {
    Plan& p = w->plans[r];
    const char* dbenv = getenv("EMU_DB");   // double-buffered exchange buffers (plan.hpp: flip): on unless EMU_DB=0
    const bool dbuf = w->nprocs > 1 && !(dbenv && dbenv[0] == '0');
    for (int q = 0; q < w->nprocs; ++q) {
      p.peerA.push_back(w->bufA[q].data());
      p.peerB.push_back(w->bufB[q].data());
      if (dbuf) {
        p.peerA_alt.push_back(w->bufA2[q].data());
        p.peerB_alt.push_back(w->bufB2[q].data());
      }
    }
    if (p.batch > 1) p.field_stride = w->strides;
    if (!p.check_supported()) {
      fprintf(stderr, "emul: %s\n", p.error.c_str());
      delete w;
      return nullptr;
    }
  }
  return w;
}

void emu_world_destroy(void* h) {
  EmuWorld* w = (EmuWorld*)h;
  pthread_barrier_destroy(&w->bar);
  delete w;
}

// lengths with a kind- / precision-specific schedule (sum over the six contiguous kernel kinds x precisions), and how many
// stage launches have used one so far
int emu_alt_schedule_count() {
  return emu_alt_count_k1_p0() + emu_alt_count_k1_p1() + emu_alt_count_k2_p0() + emu_alt_count_k2_p1() + emu_alt_count_k3_p0() +
         emu_alt_count_k3_p1();
}
long long emu_alt_launches() { return g_alt_launches; }

long long emu_alloc_max(void* h, int rank) { return ((EmuWorld*)h)->plans[rank].g.L.alloc_max; }

static int canaries_ok(EmuWorld* w) {
  for (int r = 0; r < w->nprocs; ++r)
    for (int i = 0; i < 64; ++i) {
      if ((unsigned char)w->bufA[r][w->bufA[r].size() - 64 + i] != 0xEE) return 0;
      if ((unsigned char)w->bufB[r][w->bufB[r].size() - 64 + i] != 0xEE) return 0;
      if ((unsigned char)w->bufA2[r][w->bufA2[r].size() - 64 + i] != 0xEE) return 0;
      if ((unsigned char)w->bufB2[r][w->bufB2[r].size() - 64 + i] != 0xEE) return 0;
    }
  return 1;
}

// mode: 0 forward, 1 inverse, 2 laplace, 3 biharmonic.  in/out: per-rank host pointers.
int emu_execute(void* h, int mode, void** in, void** out, int padded) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r) {
    th.emplace_back([&, r]() {
      Plan& p = w->plans[r];
      if (mode == 0) rc[r] = p.execute_forward(in[r], out[r], padded != 0);
      else if (mode == 1) rc[r] = p.execute_inverse(in[r], out[r], padded != 0);
      else if (mode == 2) rc[r] = p.execute_spectral(in[r], out[r], OP_LAPLACE, w->spec[r].data());
      else rc[r] = p.execute_spectral(in[r], out[r], OP_BIHARMONIC, w->spec[r].data());
    });
  }
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

// batched fields: in / out are [field][rank] pointer tables flattened field-major; mode 0 forward, 1 inverse
int emu_execute_many(void* h, int mode, int nf, void** in, void** out) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() {
      std::vector<const void*> ins(nf);
      std::vector<void*> outs(nf);
      for (int f = 0; f < nf; ++f) ins[f] = in[f * w->nprocs + r], outs[f] = out[f * w->nprocs + r];
      Plan& p = w->plans[r];
      rc[r] = mode == 0 ? p.execute_forward_many(nf, ins.data(), outs.data(), false) : p.execute_inverse_many(nf, ins.data(), outs.data(), false);
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

// partial-axis transforms: mode 0 forward, 1 inverse; xyz bit 2 = Z, bit 1 = Y, bit 0 = X
int emu_execute_xyz(void* h, int mode, void** in, void** out, int xyz) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() {
      Plan& p = w->plans[r];
      rc[r] = mode == 0 ? p.execute_forward(in[r], out[r], false, xyz) : p.execute_inverse(in[r], out[r], false, xyz);
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

// grad: outs = 3 arrays of per-rank pointers (x, y, z components); divergence: ins likewise
int emu_execute_grad(void* h, void** in, void** ox, void** oy, void** oz, int comps) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() {
      void* outs[3] = {ox[r], oy[r], oz[r]};
      rc[r] = w->plans[r].execute_grad(in[r], outs, comps, w->spec[r].data());
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

int emu_execute_div(void* h, void** ix, void** iy, void** iz, void** out) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() {
      const void* ins[3] = {ix[r], iy[r], iz[r]};
      rc[r] = w->plans[r].execute_divergence(ins, out[r], w->spec[r].data());
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

// solver-style symbols: op = SpectralOp value (OP_INV_LAPLACE, OP_INV_BIHARMONIC, OP_SYMBOL_LIN, OP_SYMBOL_INV)
int emu_execute_symbol(void* h, int op, double a, double b, void** in, void** out) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() { rc[r] = w->plans[r].execute_spectral(in[r], out[r], op, w->spec[r].data(), a, b); });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

// single-axis real paths: mode 0 = r2c, 1 = c2r, 2 = derivative, 3 = derivative accumulated into out
int emu_execute_axis(void* h, int mode, int axis, void** in, void** out) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() {
      Plan& p = w->plans[r];
      if (mode == 0) rc[r] = p.execute_r2c_axis(axis, in[r], out[r]);
      else if (mode == 1) rc[r] = p.execute_c2r_axis(axis, in[r], out[r]);
      else rc[r] = p.execute_deriv_axis(axis, in[r], out[r], mode == 3, w->spec[r].data());
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

// per-axis operator schemes (execute_grad_axes / execute_divergence_axes)
int emu_execute_grad_axes(void* h, void** in, void** ox, void** oy, void** oz, int comps) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() {
      void* outs[3] = {ox[r], oy[r], oz[r]};
      rc[r] = w->plans[r].execute_grad_axes(in[r], outs, comps, w->spec[r].data());
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

int emu_execute_div_axes(void* h, void** ix, void** iy, void** iz, void** out) {
  EmuWorld* w = (EmuWorld*)h;
  std::vector<std::thread> th;
  std::vector<int> rc(w->nprocs, 0);
  for (int r = 0; r < w->nprocs; ++r)
    th.emplace_back([&, r]() {
      const void* ins[3] = {ix[r], iy[r], iz[r]};
      rc[r] = w->plans[r].execute_divergence_axes(ins, out[r], w->spec[r].data());
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w->nprocs; ++r)
    if (rc[r]) return rc[r];
  return canaries_ok(w) ? 0 : -99;
}

}  // extern "C"
