// engine.cu -- CUDA backend of the plan (twiddle tables, stage launches, inter-GPU barrier) and the
// C ABI of libaccfft_b200.so (include/accfft_b200.h).
//
// Replaces, in the reference: Mem_Mgr_gpu / T_Plan_gpu (src/transpose_gpu.cpp:58-384), the cuFFT plan
// set-up of accfft_plan_dft_3d_{r2c,c2c}_gpu (src/accfft_gpu.cpp:77-442, :704-931), the execute drivers
// (:444-675, :943-1136) and the operator drivers (src/operators_gpu.txx:72-609).
// There is no CPU path in this library: every entry point that computes launches sm_100a kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/accfft_b200.h"
#include "plan.hpp"
#include "kernel_cfg.hpp"
#include "fused_launch.hpp"

namespace accfft_b200 {

// per-(kind, precision) launchers, one translation unit each (k_tu.cu)
int accfft_launch_k0_p0(const Launch&, cudaStream_t);
int accfft_launch_k1_p0(const Launch&, cudaStream_t);
int accfft_launch_k2_p0(const Launch&, cudaStream_t);
int accfft_launch_k3_p0(const Launch&, cudaStream_t);
int accfft_launch_k0_p1(const Launch&, cudaStream_t);
int accfft_launch_k1_p1(const Launch&, cudaStream_t);
int accfft_launch_k2_p1(const Launch&, cudaStream_t);
int accfft_launch_k3_p1(const Launch&, cudaStream_t);
int accfft_launch_k4_p0(const Launch&, cudaStream_t);   // TMA-pipelined strided stage; -2 = does not qualify
int accfft_launch_k4_p1(const Launch&, cudaStream_t);
int accfft_launch_k5_p0(const Launch&, cudaStream_t);   // strided r2c / c2r / derivative (stage_real.cuh)
int accfft_launch_k5_p1(const Launch&, cudaStream_t);
int accfft_launch_k6_p0(const Launch&, cudaStream_t);
int accfft_launch_k6_p1(const Launch&, cudaStream_t);
int accfft_launch_k7_p0(const Launch&, cudaStream_t);
int accfft_launch_k7_p1(const Launch&, cudaStream_t);
int accfft_launch_k8_p0(const Launch&, cudaStream_t);   // contiguous (Z-axis) derivative
int accfft_launch_k8_p1(const Launch&, cudaStream_t);
int accfft_launch_k9_p0(const Launch&, cudaStream_t);   // strided c2c, wide CTA shape; -2 = the length has none
int accfft_launch_k9_p1(const Launch&, cudaStream_t);
int accfft_launch_generic(const Launch&, int realN, cudaStream_t);   // generic_dft.cu: O(N^2) fallback, -1 = not served
int accfft_generic_max_length(Prec prec);
int accfft_has_wide_p0(int n);   // does the strided kernel have a wide CTA shape for this length? (k_tu.cu, kind 0)
int accfft_has_wide_p1(int n);
static bool has_wide(Prec p, int n) { return (p == PREC_DOUBLE ? accfft_has_wide_p0(n) : accfft_has_wide_p1(n)) != 0; }

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return (int)e;
}
#define CU(call)                                      \
  do {                                                \
    cudaError_t e__ = (call);                         \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

static bool length_supported(int N) {
#define X(N_, G_, R0, R1, R2, R3) \
  if (N == N_) return true;
#include ACCFFT_SCHEDULES_FILE
#undef X
  return false;
}

// ---------------------------------------------------------------------------------------------------
// inter-GPU barrier: every rank owns a flag array (one 64-bit epoch slot per peer) inside its exchange
// slab; rank r arrives by storing the epoch into slot r of every peer and waits until all of its own
// slots reached the epoch.  Stage kernels of the same stream have completed before the barrier kernel
// starts, so their NVLink stores are performed; the system-scope fences order them with the flags.
// ---------------------------------------------------------------------------------------------------
struct BarrierDesc {
  unsigned long long* flags[ACCFFTB_MAX_GROUP * ACCFFTB_MAX_GROUP];   // flag array of every rank (peer mapped)
  int rank, nprocs;
  unsigned long long epoch;
  long long timeout_cycles;
  int* status;   // host-mapped word, set to 1 on timeout: sticky, read by the host without synchronising (engine_check)
};

__global__ void __launch_bounds__(32) k_barrier(const __grid_constant__ BarrierDesc b) {
  // one warp: lane q signals ranks q, q+32, ... and then waits for them.  The stage kernel that precedes this launch
  // on the stream has completed, i.e. its (peer) stores are performed; the system-scope release orders them before
  // the flag, the acquire orders the peers' stores before everything that follows on this stream.
  for (int q = (int)threadIdx.x; q < b.nprocs; q += 32) {
    unsigned long long* remote = b.flags[q] + b.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(b.epoch) : "memory");
  }
  // A barrier that timed out leaves the plan dead: the status word is sticky, every later barrier of the plan returns at
  // once (this rank still signalled above, so that healthy peers are not dragged into the time-out as well) and every
  // later API call fails with ACCFFTB_ERR_TIMEOUT before it launches anything (engine_check).
  const bool dead = *(volatile int*)b.status != 0;
  for (int q = (int)threadIdx.x; q < b.nprocs && !dead; q += 32) {
    const unsigned long long* mine = b.flags[b.rank] + q;
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v >= b.epoch) break;
      if (++spins > 64) {   // peers normally arrive within a few microseconds: spin first, back off later
        if (clock64() - t0 > b.timeout_cycles) {
          *(volatile int*)b.status = 1;
          break;
        }
        __nanosleep(64);
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}


// ---------------------------------------------------------------------------------------------------
// TMA push kernel: the transfer engine of the slab exchange (Plan::forward_slab_ce / inverse_slab_ce).  Moves a 2-D block
// (`height` rows of `width` bytes) from the local staging buffer into a peer's exchange buffer over NVLink with a handful of
// CTAs: one elected thread per CTA streams tiles  global --cp.async.bulk--> shared --cp.async.bulk--> peer global  through a
// ring of shared-memory buffers (loads run NBUF-1 tiles ahead, completing on mbarriers; stores are bulk groups).  The TMA
// unit keeps hundreds of KB per SM in flight, so ~16 SMs fill the link while the other ~130 run the next chunk's FFT stages
// at HBM speed -- measured on this pool, copy engines (cudaMemcpyAsync) reach only ~500 GB/s per direction when both
// directions and a concurrent HBM-bound kernel are active, SM / TMA stores 650-680 (profiles/ce_vs_push_r2.txt).
// All addresses, pitches and the width are multiples of 16 bytes (the caller falls back to the copy engine otherwise).
// ---------------------------------------------------------------------------------------------------
struct PushDesc {
  const char* src;
  char* dst;
  unsigned long long spitch, dpitch, width, height;
  unsigned tile;   // bytes per tile (multiple of 16, <= kPushTile)
};
constexpr unsigned kPushTile = 32 * 1024;
constexpr int kPushBufs = 6;
constexpr int kPushSmem = kPushTile * kPushBufs + 128;

__device__ __forceinline__ void push_mbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)));
}
__device__ __forceinline__ void push_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void push_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "PW_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra PD_%=;\n"
      " bra PW_%=;\n"
      "PD_%=:\n"
      "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void push_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(32) k_push(const __grid_constant__ PushDesc d) {
  extern __shared__ unsigned char push_smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)push_smem_raw + 127) & ~(uintptr_t)127);
  unsigned long long* bars = (unsigned long long*)(base + (size_t)kPushTile * kPushBufs);
  if (threadIdx.x != 0) return;
  const unsigned long long tiles_per_row = (d.width + d.tile - 1) / d.tile;
  const unsigned long long ntiles = tiles_per_row * d.height;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const unsigned long long first = blockIdx.x, step = gridDim.x;
  if (first >= ntiles) return;
  const unsigned long long nmine = (ntiles - first + step - 1) / step;
  for (int b = 0; b < kPushBufs; ++b) push_mbar_init(&bars[b]);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  auto geom = [&](unsigned long long i, const char*& src, char*& dst, unsigned& bytes) {
    const unsigned long long t = first + i * step;
    const unsigned long long row = t / tiles_per_row, col = (t % tiles_per_row) * d.tile;
    const unsigned long long left = d.width - col;
    bytes = (unsigned)(left < d.tile ? left : d.tile);
    src = d.src + row * d.spitch + col;
    dst = d.dst + row * d.dpitch + col;
  };
  auto load = [&](unsigned long long i) {
    const char* src;
    char* dst;
    unsigned bytes;
    geom(i, src, dst, bytes);
    push_load(base + (size_t)(i % kPushBufs) * kPushTile, src, bytes, &bars[i % kPushBufs]);
  };
  for (unsigned long long j = 0; j + 1 < (unsigned long long)kPushBufs && j < nmine; ++j) load(j);
  for (unsigned long long i = 0; i < nmine; ++i) {
    if (i + kPushBufs - 1 < nmine) {
      // the buffer of tile i + NBUF - 1 last held tile i - 1: its store must have read it (at most store i-1's successor pending)
      if (i > 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      load(i + kPushBufs - 1);
    }
    const char* src;
    char* dst;
    unsigned bytes;
    geom(i, src, dst, bytes);
    push_wait(&bars[i % kPushBufs], (unsigned)((i / kPushBufs) & 1));
    push_store(dst, base + (size_t)(i % kPushBufs) * kPushTile, bytes);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// Stage without a transform (partial-axis executes, xyz != 111): the same descriptor as the FFT stage, but every
// element just moves to where the next layout wants it.  The reference does this with its transposes alone
// (src/accfft_gpu.cpp:488-495: "else data_out_d = data_d" followed by T_plan_1->execute_gpu).
template <class T>
__global__ void __launch_bounds__(256) k_copy_stage(const __grid_constant__ StageDesc d, int N, int kFast) {
  const long long total = (long long)d.nA * d.nB * N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const cplx<T>* in = (const cplx<T>*)d.in;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int a, b, k;
    if (kFast) {
      k = (int)(idx % N);
      b = (int)((idx / N) % d.nB);
      a = (int)(idx / ((long long)N * d.nB));
    } else {
      b = (int)(idx % d.nB);
      k = (int)((idx / d.nB) % N);
      a = (int)(idx / ((long long)N * d.nB));
    }
    const cplx<T> v = in[(long long)a * d.inStrideA + (long long)b * d.inStrideB + (long long)k * d.inStrideI];
    store_generic(d, v, k, a, b);
  }
}

// Real copy stage of the single-axis paths (stage_real.cuh: copy_real_element): the reference's real-data transposes
// T_plan_y / T_plan_x (+ inverses) and the cudaMemcpy D2D around them (src/accfft_gpu.cpp:1357-1362, :1404-1409,
// :1509-1511, :1541-1545) as ONE gather/scatter whose stores land in the owning rank's buffer.
template <class T>
__global__ void __launch_bounds__(256) k_copy_real(const __grid_constant__ StageDesc d, long long nK) {
  const long long total = (long long)d.nA * d.nB * nK;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
    copy_real_element<T>(d, idx, nK);
}

// ---------------------------------------------------------------------------------------------------
struct Engine;
struct CudaBackend : Backend {
  Engine* e = nullptr;
  bool supported(Kind, Prec, int fftN) override { return length_supported(fftN); }
  bool generic_ok(Prec prec, int N) override { return N >= 1 && N <= accfft_generic_max_length(prec); }
  const void* twiddle(Prec prec, int tableN) override;
  int launch(const Launch& l) override;
  int barrier() override;
  int pipeline_chunks() override;
  int fork() override;
  int hand(int from) override;
  int join() override;
  int barrier_on(int lane) override;
  int link_streams(cudaStream_t from, cudaStream_t to);
  // copy-engine exchange (plan.hpp)
  int ce_chunks() override;
  int copy_lanes() override;
  void* staging() override;
  int link(int from, int to) override;
  int copy2d(int lane, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) override;
  int local_variant(const Launch& l) override;
  int l2_planes(const Plan& p) override;
  int l2_window(const void* ptr, size_t bytes) override;
  int launch_fused(const Launch& lz, const Launch& ly, bool z_first) override;
};

constexpr size_t kFlagBytes = 4096;
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Engine {
  Plan plan;
  CudaBackend be;
  int device = 0;
  int rank = 0;
  cudaStream_t stream = 0;
  bool blocking = true;
  cudaStream_t aux = nullptr;  // lane 1 of software-pipelined stages (lowest priority: its kernels wait on NVLink anyway)
  int overlap_chunks = 1;      // ACCFFT_B200_OVERLAP: chunks of the pipelined Z/Y stages on slab grids (<= 1: off)
  // copy-engine exchange on slab grids (Plan::forward_slab_ce / inverse_slab_ce)
  static constexpr int kMaxCopyLanes = 8;
  int ce_chunks = 0;           // ACCFFT_B200_CE: chunks per exchanging stage (0: direct peer stores instead)
  int ce_lanes = 0;            // ACCFFT_B200_CE_LANES: copy streams (copies to different peers overlap)
  bool l2_persist = false;     // ACCFFT_B200_L2_PERSIST=1: give the hand-off buffer an L2-persisting access-policy window
  size_t l2_window_max = 0;
  long long l2_budget = 0;     // ACCFFT_B200_L2_MB: bytes of intermediate planes kept L2-resident between two local passes (0: off)
  // fused local Z / Y pair (fused.cuh).  ACCFFT_B200_FUSED: 0 = never, 1 = wherever a fused kernel exists, 2 = only where the
  // strided pass stores to peers (forward on a slab grid) and would run in the same CTA shape stand-alone; _CP / _LAG / _NS
  // tune it (_FLAGS: debugging).  Default: 2 on multi-rank slab grids (there the Y pass is NVLink-bound and the Z tiles fill the
  // SM time it leaves idle: 2 % on the step at 2 and at 8 GPUs, profiles/fused_n{2,8}_r2q.txt), 0 elsewhere (with both passes
  // local it is slower than the two stand-alone kernels although it halves their HBM traffic).
  int fused_mode = 0;
  FusedTune fused_tune = {0, 0, 0, 0};
  FusedInfo fused_info[2] = {};   // what the last forward / inverse fused launch used (accfftb_plan_fused_info)
  int* fused_ctr = nullptr;
  static constexpr int kFusedCtrInts = 16384;
  int push_ctas = 0;           // ACCFFT_B200_PUSH: CTAs of the TMA push kernel per transfer (0: copy engines / cudaMemcpyAsync)
  cudaStream_t copy_streams[kMaxCopyLanes] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  char* stage_buf = nullptr;   // local staging buffer of the exchanging stage (one exchange buffer long)
  int variant_local[ACCFFTB_PROFILE_SLOTS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // tuned kernel variant of a stage when it stores locally
  int* status_h = nullptr;     // host view of the barrier time-out word (mapped pinned memory)
  long long timeout_cycles = 20LL * 1000 * 1000 * 1000;
  cudaStream_t lane_stream(int lane) const { return lane <= 0 ? stream : (lane <= kMaxCopyLanes && copy_streams[lane - 1] ? copy_streams[lane - 1] : aux); }
  std::vector<cudaEvent_t> sync_events;   // ring of timing-less events for fork / hand / join
  size_t sync_next = 0;
  int tma_mode = 1;            // ACCFFT_B200_TMA: 0 = plain kernel only, 1 = tuned / heuristic, 2 = TMA kernel always, 3 = wide shape always
  bool wired = false;          // peers attached (always true for single-rank plans)
  char* slab = nullptr;        // [A | B | (A' | B' with double buffering) | flags] of this rank
  size_t slab_bytes = 0, bufsz = 0;
  int nbuf = 2;                // 4 = double buffered exchange buffers (multi-rank plans; ACCFFT_B200_DB=0: 2)
  int howmany = 1;             // field sets (accfftb_plan_create_many): slab = howmany x [A | B | (A' | B')] | flags
  void* spec = nullptr;        // lazily allocated spectrum scratch for grad / divergence
  int* status_d = nullptr;     // barrier timeout word
  std::vector<void*> opened;   // cudaIpcOpenMemHandle mappings to close
  std::vector<unsigned long long*> flag_ptrs;
  unsigned long long epoch = 0;
  long long launches = 0;
  long long copies = 0;        // copy-engine transfers enqueued (not kernels: counted apart from `launches`)
  int variant[ACCFFTB_PROFILE_SLOTS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // per stage slot: 0 heuristic, 1 plain, 2 TMA (set by autotune)
  std::map<std::pair<int, int>, void*> tables;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  // timing of the current call
  bool timing = false;
  double t_fft = 0, t_bar = 0;
  // per-launch profiling (event pairs on the plan's stream, read back by accfftb_plan_profile_read)
  bool profiling = false;
  struct Rec {
    int slot;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t take_event() {
    if (!pool.empty()) {
      cudaEvent_t ev = pool.back();
      pool.pop_back();
      return ev;
    }
    cudaEvent_t ev = nullptr;
    cudaEventCreate(&ev);
    return ev;
  }
};

struct ScopedProfile {
  Engine* e;
  Engine::Rec r;
  bool on;
  ScopedProfile(Engine* e_, int slot) : e(e_), on(e_->profiling && slot >= 0) {
    if (on) {
      r.slot = slot;
      r.a = e->take_event();
      r.b = e->take_event();
      cudaEventRecord(r.a, e->stream);
    }
  }
  ~ScopedProfile() {
    if (on) {
      cudaEventRecord(r.b, e->stream);
      e->recs.push_back(r);
    }
  }
};

const void* CudaBackend::twiddle(Prec prec, int M) {
  auto key = std::make_pair((int)prec, M);
  auto it = e->tables.find(key);
  if (it != e->tables.end()) return it->second;
  const size_t es = prec == PREC_DOUBLE ? 16 : 8;
  std::vector<char> buf((size_t)M * es);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < M; ++k) {
    // exact octant symmetry is not needed for the 1e-12 bound; long double cosl/sinl round correctly enough
    const long double ang = -two_pi * (long double)k / (long double)M;
    if (prec == PREC_DOUBLE) {
      ((double*)buf.data())[2 * k] = (double)cosl(ang);
      ((double*)buf.data())[2 * k + 1] = (double)sinl(ang);
    } else {
      ((float*)buf.data())[2 * k] = (float)cosl(ang);
      ((float*)buf.data())[2 * k + 1] = (float)sinl(ang);
    }
  }
  void* d = nullptr;
  if (cudaMalloc(&d, buf.size()) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, buf.data(), buf.size(), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  e->tables[key] = d;
  return d;
}

struct ScopedTimer {
  Engine* e;
  double* acc;
  ScopedTimer(Engine* e_, double* acc_) : e(e_), acc(acc_) {
    if (e->timing && acc) cudaEventRecord(e->ev[0], e->stream);
  }
  ~ScopedTimer() {
    if (e->timing && acc) {
      cudaEventRecord(e->ev[1], e->stream);
      cudaEventSynchronize(e->ev[1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
      *acc += 1e-3 * (double)ms;
    }
  }
};

// everything below enqueues on e->stream; a launch on lane 1 swaps the auxiliary stream in for its duration
struct LaneScope {
  Engine* e;
  cudaStream_t saved;
  LaneScope(Engine* e_, int lane) : e(e_), saved(e_->stream) {
    if (lane >= 1) {
      cudaStream_t s = e->lane_stream(lane);
      if (s) e->stream = s;
    }
  }
  ~LaneScope() { e->stream = saved; }
};

int CudaBackend::pipeline_chunks() { return (e->timing || !e->aux) ? 1 : e->overlap_chunks; }

int CudaBackend::link_streams(cudaStream_t from, cudaStream_t to) {
  if (from == to) return 0;
  if (e->sync_events.empty()) {
    e->sync_events.resize(256, nullptr);
    for (auto& ev : e->sync_events) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  cudaEvent_t ev = e->sync_events[e->sync_next++ % e->sync_events.size()];
  CU(cudaEventRecord(ev, from));
  CU(cudaStreamWaitEvent(to, ev, 0));
  return 0;
}
int CudaBackend::fork() { return link_streams(e->stream, e->aux); }
int CudaBackend::hand(int from) { return from == 0 ? link_streams(e->stream, e->aux) : link_streams(e->aux, e->stream); }
int CudaBackend::join() { return link_streams(e->aux, e->stream); }

// planes per chunk of the L2-resident Z <-> Y hand-off: as many whole i0 planes of the intermediate layout as fit the budget
int CudaBackend::l2_planes(const Plan& p) {
  if (e->timing || e->l2_budget <= 0) return 0;
  const long long cs = p.prec == PREC_DOUBLE ? 16 : 8;
  const long long plane = (long long)p.g.n[1] * p.g.k2.size * cs;   // bytes of one i0 plane of A ([N1][k2]); B's [i1][n2c] is the same size
  if (plane <= 0) return 0;
  const long long planes = e->l2_budget / plane;
  if (planes < 1 || planes >= p.g.i0.size) return 0;                 // a plane does not fit / the whole array fits: nothing to chunk
  return (int)planes;
}

// L2 access-policy window of the plan's stream: hits in the window persist, everything else streams through
int CudaBackend::l2_window(const void* ptr, size_t bytes) {
  if (!e->l2_persist) return 0;
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  v.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  v.accessPolicyWindow.num_bytes = std::min(bytes, e->l2_window_max);
  v.accessPolicyWindow.hitRatio = 1.0f;
  v.accessPolicyWindow.hitProp = bytes ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
  v.accessPolicyWindow.missProp = bytes ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
  CU(cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &v));
  return 0;
}

// ---- copy-engine exchange ---------------------------------------------------------------------------------------
int CudaBackend::ce_chunks() { return (e->timing || !e->stage_buf || e->ce_lanes < 1) ? 0 : e->ce_chunks; }
int CudaBackend::copy_lanes() { return e->ce_lanes; }
void* CudaBackend::staging() { return e->stage_buf; }
int CudaBackend::link(int from, int to) { return link_streams(e->lane_stream(from), e->lane_stream(to)); }
int CudaBackend::local_variant(const Launch& l) {
  const int v = e->variant_local[l.slot >= 0 && l.slot < ACCFFTB_PROFILE_SLOTS ? l.slot : 6];
  if (v) return v;
  // heuristic of CudaBackend::launch for a LOCAL stage (the launch itself sees nseg > 1 and would pick the peer-store shape)
  return (l.desc.nA > 1 || l.prec == PREC_FLOAT) ? 2 : 1;
}
int CudaBackend::copy2d(int lane, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height) {
  if (lane < 1 || lane > e->ce_lanes) return fail(ACCFFTB_ERR_STATE, "copy lane out of range");
  if (!width || !height) return 0;
  LaneScope lane_scope(e, lane);
  ScopedProfile pf(e, 9);
  const bool aligned = ((((uintptr_t)dst | (uintptr_t)src | width) & 15) == 0) && (height == 1 || ((dpitch | spitch) & 15) == 0);
  if (e->push_ctas > 0 && aligned) {
    // TMA push kernel on a few SMs (see k_push)
    PushDesc pd;
    pd.src = (const char*)src;
    pd.dst = (char*)dst;
    pd.spitch = spitch;
    pd.dpitch = dpitch;
    pd.width = width;
    pd.height = height;
    pd.tile = kPushTile;
    const unsigned long long ntiles = ((width + kPushTile - 1) / kPushTile) * height;
    const unsigned grid = (unsigned)std::min<unsigned long long>(ntiles, (unsigned long long)e->push_ctas);
    k_push<<<grid, 32, kPushSmem, e->stream>>>(pd);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, "push kernel launch");
    e->launches++;
    e->copies++;
    return 0;
  }
  if (height == 1) CU(cudaMemcpyAsync(dst, src, width, cudaMemcpyDeviceToDevice, e->stream));
  else CU(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToDevice, e->stream));
  e->copies++;
  return 0;
}
int CudaBackend::barrier_on(int lane) {
  LaneScope ls(e, lane);
  return barrier();
}

int CudaBackend::launch(const Launch& l) {
  if (!l.desc.tw && l.kind != K_COPY_REAL) return fail(ACCFFTB_ERR_STATE, "twiddle table allocation failed");
  LaneScope lane_scope(e, l.dry ? 0 : l.lane);
  ScopedTimer tm(e, l.dry ? nullptr : &e->t_fft);
  ScopedProfile pf(e, l.dry ? -1 : l.slot);
  int rc = -1;
  if (l.kind == K_COPY) {
    cudaFuncAttributes fa;
    if (l.dry) return (int)(l.prec == PREC_DOUBLE ? cudaFuncGetAttributes(&fa, k_copy_stage<double>)
                                                  : cudaFuncGetAttributes(&fa, k_copy_stage<float>));
    const long long total = (long long)l.desc.nA * l.desc.nB * l.fftN;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 16);
    const int kfast = l.desc.inStrideI == 1 ? 1 : 0;
    if (l.prec == PREC_DOUBLE) k_copy_stage<double><<<grid, 256, 0, e->stream>>>(l.desc, l.fftN, kfast);
    else k_copy_stage<float><<<grid, 256, 0, e->stream>>>(l.desc, l.fftN, kfast);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, "copy stage launch");
    e->launches++;
    return 0;
  }
  if (l.kind == K_COPY_REAL) {
    cudaFuncAttributes fa;
    if (l.dry) return (int)(l.prec == PREC_DOUBLE ? cudaFuncGetAttributes(&fa, k_copy_real<double>)
                                                  : cudaFuncGetAttributes(&fa, k_copy_real<float>));
    const long long total = (long long)l.desc.nA * l.desc.nB * l.copyN;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 16);
    if (l.prec == PREC_DOUBLE) k_copy_real<double><<<grid, 256, 0, e->stream>>>(l.desc, l.copyN);
    else k_copy_real<float><<<grid, 256, 0, e->stream>>>(l.desc, l.copyN);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(err, "real copy stage launch");
    e->launches++;
    return 0;
  }
  // lengths without a compiled schedule, and r2c / c2r with an odd real length: the generic O(N^2) stage kernel
  {
    const bool real_kind = l.kind == K_CONTIG_R2C || l.kind == K_CONTIG_C2R;
    const int realN = e->plan.g.n[2];
    const bool complex_kind = l.kind == K_STRIDED_C2C || l.kind == K_CONTIG_C2C;
    if ((real_kind && (realN % 2 != 0 || !length_supported(l.fftN))) || (complex_kind && !length_supported(l.fftN))) {
      const int grc = accfft_launch_generic(l, realN, e->stream);
      if (grc == -1) return fail(ACCFFTB_ERR_UNSUPPORTED, "no kernel for length " + std::to_string(real_kind ? realN : l.fftN));
      if (grc != 0) return cuda_fail((cudaError_t)grc, "generic stage kernel launch");
      if (!l.dry) e->launches++;
      return 0;
    }
  }
  const int sel = (int)l.kind * 2 + (int)l.prec;
  switch (sel) {
    case 0:
    case 1: {
      // Two kernels serve the strided stages: the plain one and the TMA-pipelined persistent one.  Choice: the launch's
      // own request, else what plan-time tuning (ACCFFT_MEASURE) found for this stage, else a heuristic from the
      // measurements in profiles/variants_r05.txt (TMA wins when the tile's rows are a few KB apart -- the Y passes --
      // and for single precision; the plain kernel wins on the plane-strided X passes in double).  Variant 3 = the plain
      // kernel in its wide CTA shape (kernel_cfg.hpp: CfgW).
      // Dry (plan-time) launches load BOTH; at run time the TMA kernel declines (-2) what it cannot take.
      // Stages that store to peers are NVLink-bound and want as many storing threads and as long row segments as
      // possible: never the one-CTA-per-SM TMA kernel (profiles/mgpu8_r05.log: it lost 5-10 % there), the wide shape where
      // the length has one (profiles/mgpu8_r14.log), else the plain kernel.
      int v = l.variant ? l.variant : e->variant[l.slot >= 0 && l.slot < ACCFFTB_PROFILE_SLOTS ? l.slot : 6];
      // heuristic (ACCFFT_ESTIMATE, and every stage that stores to peers): peer stores want the longest row segments
      // the length allows -> wide shape where it exists (else it declines and the plain kernel runs)
      if (v == 0) v = l.desc.nseg > 1 ? 3 : ((l.desc.nA > 1 || l.prec == PREC_FLOAT) ? 2 : 1);
      if (e->tma_mode == 0) v = 1;
      if (e->tma_mode == 2) v = 2;
      if (e->tma_mode == 3) v = 3;
      if (v == 3 && !has_wide(l.prec, l.fftN)) v = 1;
      rc = -2;
      if (v == 2 || l.dry) rc = sel == 0 ? accfft_launch_k4_p0(l, e->stream) : accfft_launch_k4_p1(l, e->stream);
      if (rc == -2 || (l.dry && rc == 0)) {
        rc = -2;
        if (v == 3 || l.dry) rc = sel == 0 ? accfft_launch_k9_p0(l, e->stream) : accfft_launch_k9_p1(l, e->stream);
        if (rc == -2 || (l.dry && rc == 0))   // default shape (also when the length has no wide shape)
          rc = sel == 0 ? accfft_launch_k0_p0(l, e->stream) : accfft_launch_k0_p1(l, e->stream);
      }
      break;
    }
    case 2: rc = accfft_launch_k1_p0(l, e->stream); break;
    case 3: rc = accfft_launch_k1_p1(l, e->stream); break;
    case 4: rc = accfft_launch_k2_p0(l, e->stream); break;
    case 5: rc = accfft_launch_k2_p1(l, e->stream); break;
    case 6: rc = accfft_launch_k3_p0(l, e->stream); break;
    case 7: rc = accfft_launch_k3_p1(l, e->stream); break;
    case 10: rc = accfft_launch_k5_p0(l, e->stream); break;
    case 11: rc = accfft_launch_k5_p1(l, e->stream); break;
    case 12: rc = accfft_launch_k6_p0(l, e->stream); break;
    case 13: rc = accfft_launch_k6_p1(l, e->stream); break;
    case 14: rc = accfft_launch_k7_p0(l, e->stream); break;
    case 15: rc = accfft_launch_k7_p1(l, e->stream); break;
    case 18: rc = accfft_launch_k8_p0(l, e->stream); break;   // K_CONTIG_DERIV = 9
    case 19: rc = accfft_launch_k8_p1(l, e->stream); break;
    default: break;
  }
  if (rc == -1) return fail(ACCFFTB_ERR_UNSUPPORTED, "no kernel schedule for length " + std::to_string(l.fftN));
  if (rc != 0) return cuda_fail((cudaError_t)rc, "stage kernel launch");
  if (!l.dry) e->launches++;
  return 0;
}

// the local Z / Y pair as one persistent kernel (fused.cuh); -2 = not available for this pair, run the stages separately
int CudaBackend::launch_fused(const Launch& lz, const Launch& ly, bool z_first) {
  if (!lz.dry) e->fused_info[z_first ? 0 : 1] = FusedInfo{};
  if (!e->fused_mode || !e->fused_ctr) return -2;
  // mode 2: only where the strided pass stores to peers (forward on a slab grid): there the pass is NVLink-bound and the
  // contiguous pass's tiles fill the SM time it leaves idle
  // (measured for r2c plans only -- c2c tiles are twice the size and lose more inside the persistent loop: not fused in mode 2)
  if (e->fused_mode == 2 && (stage_is_simple(ly.desc) || has_wide(ly.prec, ly.fftN) || lz.kind != K_CONTIG_R2C)) return -2;
  if (e->timing && !lz.dry) return -2;   // per-stage timers of the reference-style `timer` array want separate stages
  const bool r2c = lz.kind == K_CONTIG_R2C || lz.kind == K_CONTIG_C2R;
  if (!r2c && lz.kind != K_CONTIG_C2C) return -2;
  if (ly.kind != K_STRIDED_C2C || lz.prec != ly.prec) return -2;
  if (r2c && e->plan.g.n[2] % 2 != 0) return -2;
  if (!length_supported(lz.fftN) || !length_supported(ly.fftN)) return -2;
  if (!lz.desc.tw || !ly.desc.tw) return fail(ACCFFTB_ERR_STATE, "twiddle table allocation failed");
  const bool dry = lz.dry;
  const int mode = (r2c ? 0 : 2) + (z_first ? 0 : 1);
  ScopedProfile pf(e, dry ? -1 : (z_first ? lz.slot : ly.slot));
  FusedInfo* info = dry ? nullptr : &e->fused_info[z_first ? 0 : 1];
  if (info) *info = FusedInfo{};
  int rc = -2;
  const int p = (int)lz.prec;
  switch (mode * 2 + p) {
    case 0: rc = accfft_launch_fused_m0_p0(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
    case 1: rc = accfft_launch_fused_m0_p1(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
    case 2: rc = accfft_launch_fused_m1_p0(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
    case 3: rc = accfft_launch_fused_m1_p1(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
    case 4: rc = accfft_launch_fused_m2_p0(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
    case 5: rc = accfft_launch_fused_m2_p1(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
    case 6: rc = accfft_launch_fused_m3_p0(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
    default: rc = accfft_launch_fused_m3_p1(lz, ly, e->fused_tune, e->fused_ctr, Engine::kFusedCtrInts, e->stream, dry, info); break;
  }
  if (rc == -2) {
    if (pf.on) e->pool.push_back(pf.r.a), e->pool.push_back(pf.r.b), pf.on = false;
    return -2;
  }
  if (rc != 0) return cuda_fail((cudaError_t)rc, "fused stage kernel launch");
  if (!dry) e->launches++;
  return 0;
}

int CudaBackend::barrier() {
  if (!e->wired) return fail(ACCFFTB_ERR_STATE, "multi-rank plan used before its peers were attached");
  ScopedTimer tm(e, &e->t_bar);
  ScopedProfile pf(e, 6);
  BarrierDesc b;
  memset(&b, 0, sizeof(b));
  for (int q = 0; q < e->plan.nprocs; ++q) b.flags[q] = e->flag_ptrs[q];
  b.rank = e->rank;
  b.nprocs = e->plan.nprocs;
  b.epoch = ++e->epoch;
  b.timeout_cycles = e->timeout_cycles;   // 10 s of SM clock (cudaDevAttrClockRate): a peer died or diverged
  b.status = e->status_d;
  k_barrier<<<1, 32, 0, e->stream>>>(b);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return cuda_fail(err, "barrier kernel launch");
  e->launches++;
  return 0;
}

// The barrier time-out word lives in mapped pinned host memory: the host reads it without synchronising anything.  It is
// sticky -- once a barrier of this plan has timed out the stages after it ran on incomplete peer data, so every later call on
// the plan fails before launching (blocking or not); non-blocking callers can also poll accfftb_plan_check().
static int engine_check(Engine* e) {
  if (e->status_h && *(volatile int*)e->status_h)
    return fail(ACCFFTB_ERR_TIMEOUT, "an inter-GPU barrier of this plan timed out (a peer died or diverged): results since then are "
                                     "invalid and the plan is unusable; destroy it");
  return 0;
}

static int engine_finish(Engine* e, int rc) {
  if (rc) return rc;
  if (e->blocking) {
    CU(cudaStreamSynchronize(e->stream));
    return engine_check(e);
  }
  return 0;
}

static void engine_free(Engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->aux) cudaStreamSynchronize(e->aux), cudaStreamDestroy(e->aux);
  for (auto ev : e->sync_events)
    if (ev) cudaEventDestroy(ev);
  for (void* p : e->opened) cudaIpcCloseMemHandle(p);
  for (auto& kv : e->tables) cudaFree(kv.second);
  if (e->slab) cudaFree(e->slab);
  if (e->spec) cudaFree(e->spec);
  if (e->status_h) cudaFreeHost(e->status_h);
  if (e->stage_buf) cudaFree(e->stage_buf);
  if (e->fused_ctr) cudaFree(e->fused_ctr);
  for (auto& cs : e->copy_streams)
    if (cs) cudaStreamSynchronize(cs), cudaStreamDestroy(cs);
  if (e->ev[0]) cudaEventDestroy(e->ev[0]);
  if (e->ev[1]) cudaEventDestroy(e->ev[1]);
  for (auto& r : e->recs) cudaEventDestroy(r.a), cudaEventDestroy(r.b);
  for (auto ev : e->pool) cudaEventDestroy(ev);
}

static void wire_self(Engine* e) {
  const int np = e->plan.nprocs;
  e->plan.peerA.assign(np, nullptr);
  e->plan.peerB.assign(np, nullptr);
  e->flag_ptrs.assign(np, nullptr);
  e->plan.peerA[e->rank] = e->slab;
  e->plan.peerB[e->rank] = e->slab + e->bufsz;
  if (e->nbuf == 4) {
    e->plan.peerA_alt.assign(np, nullptr);
    e->plan.peerB_alt.assign(np, nullptr);
    e->plan.peerA_alt[e->rank] = e->slab + 2 * e->bufsz;
    e->plan.peerB_alt[e->rank] = e->slab + 3 * e->bufsz;
  }
  e->flag_ptrs[e->rank] = (unsigned long long*)(e->slab + (size_t)e->howmany * e->nbuf * e->bufsz);
}

// size of rank q's exchange buffers (256-byte aligned) -- every rank can compute it for every peer
static size_t bufsz_of(const Engine* e, int q) {
  int coords[2] = {q / e->plan.g.np[1], q % e->plan.g.np[1]};
  const long long b = Plan::buffer_bytes(e->plan.g.n, e->plan.g.np, coords, e->plan.g.r2c, e->plan.prec, e->plan.nprocs);
  return align_up((size_t)b, 256);
}

struct IpcBlob {
  cudaIpcMemHandle_t handle;   // 64 bytes
  unsigned long long slab_bytes;
  int rank, nprocs;
  unsigned magic;
  char pad[ACCFFTB_IPC_BLOB_BYTES - 64 - 8 - 12];
};
static_assert(sizeof(IpcBlob) == ACCFFTB_IPC_BLOB_BYTES, "blob size");

struct TimerScope {
  Engine* e;
  double* timer;
  int nslots;
  TimerScope(Engine* e_, double* timer_, int nslots_) : e(e_), timer(timer_), nslots(nslots_) {
    e->timing = timer != nullptr;
    e->t_fft = e->t_bar = 0;
  }
  ~TimerScope() {
    if (timer) {
      timer[0] += e->t_bar;
      timer[2] += e->t_bar;
      timer[4] += e->t_fft;
    }
    e->timing = false;
  }
};

// Plan-time tuning (ACCFFT_MEASURE / ACCFFT_PATIENT; the reference tunes its transpose method the same way,
// src/transpose_gpu.cpp:632-743, but on the caller's buffers -- here only plan-owned scratch is touched).
// Each strided stage is timed with both kernels on this rank's own input buffer, storing to a scratch array.
static int autotune(Engine* e) {
  Plan& pl = e->plan;
  void* scratch = nullptr;
  CU(cudaMalloc(&scratch, e->bufsz));
  CU(cudaMemsetAsync(e->slab, 0, 2 * e->bufsz, e->stream));
  Launch ls[4] = {pl.fwd_y(), pl.fwd_x(scratch), pl.inv_x(e->slab), pl.inv_y()};
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  int rc = 0;
  for (Launch& l : ls) {
    StageDesc& d = l.desc;
    if (d.nA <= 0 || d.nB <= 0 || !length_supported(l.fftN)) continue;
    const bool to_peers = d.nseg > 1;
    if (to_peers) {   // stores to peers: NVLink-bound, not tunable on this rank alone -- see CudaBackend::launch
      e->variant[l.slot] = has_wide(l.prec, l.fftN) ? 3 : 1;   // wide shape where the length has one, else the plain kernel
      // copy-engine exchange: the same stage packs into LOCAL staging instead -- tune it like a local stage (variant_local)
      if (!(e->ce_chunks > 0 && e->stage_buf)) continue;
    }
    // single local destination [a][k][b]: same bytes as the stage input, which fits one exchange buffer
    d.nseg = 1;
    d.blk = l.fftN;
    d.seg[0].base = scratch;
    d.seg[0].strideA = (long long)l.fftN * d.nB;
    d.seg[0].strideB = 1;
    d.seg[0].strideI = d.nB;
    d.colShift = 0;
    float best[4] = {0, 1e30f, 1e30f, 1e30f};
    for (int v = 1; v <= 3 && !rc; ++v) {
      l.variant = v;
      for (int it = 0; it < 4 && !rc; ++it) {   // first iteration = warm-up
        cudaEventRecord(a, e->stream);
        rc = e->be.launch(l);
        cudaEventRecord(b, e->stream);
        if (rc) break;
        cudaError_t err = cudaEventSynchronize(b);
        if (err != cudaSuccess) rc = cuda_fail(err, "autotune");
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best[v]) best[v] = ms;
      }
    }
    if (rc) break;
    if (!has_wide(l.prec, l.fftN)) best[3] = 1e30f;   // that run fell back to the plain kernel
    const int pick = best[2] < best[1] ? (best[3] < best[2] ? 3 : 2) : (best[3] < best[1] ? 3 : 1);
    if (to_peers) e->variant_local[l.slot] = pick;
    else e->variant[l.slot] = pick;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaMemsetAsync(e->slab, 0, 2 * e->bufsz, e->stream);
  cudaStreamSynchronize(e->stream);
  cudaFree(scratch);
  e->launches = 0;
  return rc;
}

}  // namespace accfft_b200

using namespace accfft_b200;

struct accfftb_plan {
  Engine e;
};

static bool overflow_guard(const LocalSizes& L, int) {
  // the reference refuses local input blocks of more than 2^31-1 elements (src/accfft_common.cpp:179-184)
  return (long long)L.isize[0] * L.isize[1] * L.isize[2] > 2147483647LL;
}

extern "C" {

const char* accfftb_last_error(void) { return g_err.c_str(); }
const char* accfftb_version(void) { return "accfft_b200 0.1 (sm_100a)"; }

int accfftb_create_comm_dims(int nprocs, int c_dims[2]) {
  if (nprocs > 0 && c_dims[0] * c_dims[1] == nprocs && c_dims[0] > 0) return 0;
  dims_create(nprocs, c_dims);
  return 1;
}

void accfftb_rank_to_coords(int rank, const int c_dims[2], int coords[2]) {
  coords[0] = rank / c_dims[1];
  coords[1] = rank % c_dims[1];
}

static void copy3(int* dst, const int* src) {
  if (dst)
    for (int i = 0; i < 3; ++i) dst[i] = src[i];
}

long long accfftb_local_size_dft_r2c(const int n[3], const int c_dims[2], const int coords[2], int sizeof_real,
                                     int isize[3], int istart[3], int osize[3], int ostart[3]) {
  LocalSizes L = local_size_r2c(n, c_dims, coords, sizeof_real);
  copy3(isize, L.isize);
  copy3(istart, L.istart);
  copy3(osize, L.osize);
  copy3(ostart, L.ostart);
  if (overflow_guard(L, sizeof_real)) return fail(ACCFFTB_ERR_OVERFLOW, "local block exceeds 2^31 elements");
  return L.alloc_max;
}

long long accfftb_local_size_dft_c2c(const int n[3], const int c_dims[2], const int coords[2], int sizeof_real,
                                     int isize[3], int istart[3], int osize[3], int ostart[3]) {
  LocalSizes L = local_size_c2c(n, c_dims, coords, sizeof_real);
  copy3(isize, L.isize);
  copy3(istart, L.istart);
  copy3(osize, L.osize);
  copy3(ostart, L.ostart);
  // (the reference's c2c query has no overflow guard, src/accfft_common.cpp:203-256; sizes are 64-bit here)
  return L.alloc_max;
}

int accfftb_stage_size(const int n[3], const int c_dims[2], const int coords[2], int kind, int sizeof_real, int which,
                       int size[3], int start[3]) {
  LocalSizes L = kind == ACCFFTB_R2C ? local_size_r2c(n, c_dims, coords, sizeof_real)
                                     : local_size_c2c(n, c_dims, coords, sizeof_real);
  const int zero[3] = {0, 0, 0};
  switch (which) {
    case 0: copy3(size, L.osize_0); copy3(start, L.ostart_0); return 0;
    case 1: copy3(size, L.osize_1); copy3(start, L.ostart_1); return 0;
    case 2: copy3(size, L.osize_2); copy3(start, L.ostart_2); return 0;
    case 3: copy3(size, L.osize_y); copy3(start, L.ostart_y); return 0;
    case 4: copy3(size, L.osize_yi); copy3(start, zero); return 0;
    case 5: copy3(size, L.osize_x); copy3(start, zero); return 0;
    case 6: copy3(size, L.osize_xi); copy3(start, zero); return 0;
    default: return fail(ACCFFTB_ERR_ARG, "accfftb_stage_size: which must be 0..6");
  }
}

int accfftb_length_supported(int n) { return (length_supported(n) || (n >= 1 && n <= accfft_generic_max_length(PREC_DOUBLE))) ? 1 : 0; }
int accfftb_length_fast(int n) { return length_supported(n) ? 1 : 0; }

accfftb_plan* accfftb_plan_create(const int n[3], const int c_dims[2], int rank, int kind, int prec, unsigned flags) {
  return accfftb_plan_create_many(n, c_dims, rank, kind, prec, flags, 1);
}

accfftb_plan* accfftb_plan_create_many(const int n[3], const int c_dims[2], int rank, int kind, int prec, unsigned flags, int howmany) {
  if (howmany < 1 || howmany > 64) {
    fail(ACCFFTB_ERR_ARG, "accfftb_plan_create_many: howmany must be 1..64");
    return nullptr;
  }
  if (!n || !c_dims || n[0] <= 0 || n[1] <= 0 || n[2] <= 0 || c_dims[0] <= 0 || c_dims[1] <= 0 || rank < 0 ||
      rank >= c_dims[0] * c_dims[1] || (kind != ACCFFTB_R2C && kind != ACCFFTB_C2C) ||
      (prec != ACCFFTB_DOUBLE && prec != ACCFFTB_FLOAT)) {
    fail(ACCFFTB_ERR_ARG, "accfftb_plan_create: bad argument");
    return nullptr;
  }
  accfftb_plan* p = new accfftb_plan();
  Engine* e = &p->e;
  e->be.e = e;
  e->rank = rank;
  int coords[2] = {rank / c_dims[1], rank % c_dims[1]};
  Plan& pl = e->plan;
  pl.g = make_geometry(n, c_dims, coords, kind == ACCFFTB_R2C, prec == ACCFFTB_DOUBLE ? 8 : 4);
  pl.prec = (Prec)prec;
  pl.inplace = false;
  pl.be = &e->be;
  pl.nprocs = c_dims[0] * c_dims[1];
  e->howmany = howmany;
  pl.batch = howmany;
  if (!pl.check_supported()) {
    fail(ACCFFTB_ERR_UNSUPPORTED, pl.error);
    delete p;
    return nullptr;
  }
  if ((long long)pl.g.n[0] * pl.g.k1.size * pl.g.k2.size > 2147483647LL ||
      (long long)pl.g.i0.size * pl.g.n[1] * pl.g.k2.size > 2147483647LL || overflow_guard(pl.g.L, 0)) {
    fail(ACCFFTB_ERR_OVERFLOW, "local block exceeds 2^31 elements; use more ranks");
    delete p;
    return nullptr;
  }
  cudaError_t err = cudaGetDevice(&e->device);
  if (err == cudaSuccess) {
    e->bufsz = bufsz_of(e, rank);
    e->nbuf = pl.nprocs > 1 ? 4 : 2;
    if (const char* m = getenv("ACCFFT_B200_DB"))
      if (m[0] == '0') e->nbuf = 2;
    e->slab_bytes = (size_t)e->howmany * e->nbuf * e->bufsz + kFlagBytes;
    err = cudaMalloc((void**)&e->slab, e->slab_bytes);
  }
  if (err == cudaSuccess) err = cudaMemset(e->slab + (size_t)e->howmany * e->nbuf * e->bufsz, 0, kFlagBytes);
  // spectrum scratch of the operators: allocated now, never inside a collective call (cudaMalloc synchronises
  // the device, which must not happen while a peer's barrier kernel waits for this rank)
  if (err == cudaSuccess && kind == ACCFFTB_R2C) err = cudaMalloc(&e->spec, e->bufsz);
  if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->status_h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable);
  if (err == cudaSuccess) {
    *e->status_h = 0;
    err = cudaHostGetDevicePointer((void**)&e->status_d, e->status_h, 0);
  }
  if (err == cudaSuccess) {
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device) == cudaSuccess && khz > 0)
      e->timeout_cycles = 10LL * 1000 * (long long)khz;   // 10 seconds of SM clock
    if (const char* m = getenv("ACCFFT_B200_BARRIER_TIMEOUT_MS")) e->timeout_cycles = (long long)atoll(m) * (long long)(khz > 0 ? khz : 2000000);
  }
  e->fused_mode = (pl.nprocs > 1 && c_dims[1] == 1) ? 2 : 0;
  if (const char* m = getenv("ACCFFT_B200_FUSED")) e->fused_mode = atoi(m);
  if (const char* m = getenv("ACCFFT_B200_FUSED_CP")) e->fused_tune.cp = atoi(m);
  if (const char* m = getenv("ACCFFT_B200_FUSED_LAG")) e->fused_tune.lag = atoi(m);
  if (const char* m = getenv("ACCFFT_B200_FUSED_NS")) e->fused_tune.ns = atoi(m);
  if (const char* m = getenv("ACCFFT_B200_FUSED_FLAGS")) e->fused_tune.flags = atoi(m);
  if (err == cudaSuccess && e->fused_mode && c_dims[1] == 1) err = cudaMalloc((void**)&e->fused_ctr, sizeof(int) * Engine::kFusedCtrInts);
  if (err == cudaSuccess) err = cudaEventCreate(&e->ev[0]);
  if (err == cudaSuccess) err = cudaEventCreate(&e->ev[1]);
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    cuda_fail(err, "accfftb_plan_create");
    engine_free(e);
    delete p;
    return nullptr;
  }
  wire_self(e);
  if (e->howmany > 1) {
    pl.field_stride.resize(pl.nprocs);
    for (int q = 0; q < pl.nprocs; ++q) pl.field_stride[q] = (long long)e->nbuf * (long long)bufsz_of(e, q);
  }
  e->wired = pl.nprocs == 1;
  {
    // load every kernel now (see Plan::prepare)
    cudaFuncAttributes fa;
    err = cudaFuncGetAttributes(&fa, k_barrier);
    int rc = err == cudaSuccess ? pl.prepare() : (int)err;
    if (const char* m = getenv("ACCFFT_B200_OVERLAP")) e->overlap_chunks = atoi(m);
    if (pl.nprocs > 1 && e->overlap_chunks > 1) {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least priority (numerically greatest)
      if (cudaStreamCreateWithPriority(&e->aux, cudaStreamNonBlocking, lo) != cudaSuccess) e->aux = nullptr, cudaGetLastError();
    }
    // copy-engine exchange: slab grids only (one exchange per transform, both of its ends local)
    if (pl.nprocs > 1 && c_dims[1] == 1) {
      e->ce_chunks = 0;   // default: direct peer stores (see DESIGN.md section 5 for the measurements behind this choice)
      e->ce_lanes = 4;
      if (const char* m = getenv("ACCFFT_B200_CE")) e->ce_chunks = atoi(m);
      if (const char* m = getenv("ACCFFT_B200_CE_LANES")) e->ce_lanes = std::max(1, std::min(atoi(m), (int)Engine::kMaxCopyLanes));
      e->push_ctas = 16;
      if (const char* m = getenv("ACCFFT_B200_PUSH")) e->push_ctas = std::max(0, atoi(m));
      if (e->push_ctas > 0) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, k_push) != cudaSuccess ||
            cudaFuncSetAttribute(k_push, cudaFuncAttributeMaxDynamicSharedMemorySize, kPushSmem) != cudaSuccess) {
          cudaGetLastError();
          e->push_ctas = 0;
        }
      }
      if (e->ce_chunks > 0) {
        cudaError_t ce = cudaMalloc((void**)&e->stage_buf, e->bufsz);
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // hi = greatest priority (numerically least)
        for (int l = 0; l < e->ce_lanes && ce == cudaSuccess; ++l)   // highest priority: push CTAs take SMs as compute CTAs retire
          ce = cudaStreamCreateWithPriority(&e->copy_streams[l], cudaStreamNonBlocking, prio_hi);
        if (ce != cudaSuccess) {   // no room / no streams: fall back to the direct peer-store path
          cudaGetLastError();
          if (e->stage_buf) cudaFree(e->stage_buf), e->stage_buf = nullptr;
          e->ce_chunks = 0;
        }
      }
    }
    if (const char* m = getenv("ACCFFT_B200_L2_MB")) e->l2_budget = (long long)atoi(m) << 20;
    if (const char* m = getenv("ACCFFT_B200_L2_PERSIST")) e->l2_persist = atoi(m) != 0;
    if (e->l2_budget > 0 && e->l2_persist) {
      int max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
      e->l2_window_max = (size_t)max_window;
      if (max_persist <= 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)e->l2_budget, (size_t)max_persist)) != cudaSuccess) {
        cudaGetLastError();
        e->l2_persist = false;
      }
    }
    if (const char* m = getenv("ACCFFT_B200_TMA")) e->tma_mode = m[0] == '0' ? 0 : m[0] == '2' ? 2 : m[0] == '3' ? 3 : 1;
    // ACCFFT_MEASURE / ACCFFT_PATIENT: pick the faster kernel per strided stage by timing both (plan-owned buffers only)
    if (!rc && (flags & (ACCFFTB_MEASURE | ACCFFTB_PATIENT)) && e->tma_mode == 1) rc = autotune(e);
    if (rc) {
      if (err != cudaSuccess) cuda_fail(err, "accfftb_plan_create (kernel load)");
      engine_free(e);
      delete p;
      return nullptr;
    }
  }
  return p;
}

int accfftb_plan_ipc_export(accfftb_plan* plan, void* blob) {
  if (!plan || !blob) return fail(ACCFFTB_ERR_ARG, "accfftb_plan_ipc_export: null argument");
  Engine* e = &plan->e;
  IpcBlob b;
  memset(&b, 0, sizeof(b));
  CU(cudaIpcGetMemHandle(&b.handle, e->slab));
  b.slab_bytes = e->slab_bytes;
  b.rank = e->rank;
  b.nprocs = e->plan.nprocs;
  b.magic = 0xACCFF7B2u;
  memcpy(blob, &b, sizeof(b));
  return 0;
}

int accfftb_plan_ipc_import(accfftb_plan* plan, const void* blobs, int nblobs) {
  if (!plan || !blobs) return fail(ACCFFTB_ERR_ARG, "accfftb_plan_ipc_import: null argument");
  Engine* e = &plan->e;
  if (nblobs != e->plan.nprocs) return fail(ACCFFTB_ERR_ARG, "accfftb_plan_ipc_import: need one blob per rank");
  if (e->wired && e->plan.nprocs > 1) return fail(ACCFFTB_ERR_STATE, "peers already attached");
  for (int q = 0; q < nblobs; ++q) {
    IpcBlob b;
    memcpy(&b, (const char*)blobs + (size_t)q * sizeof(IpcBlob), sizeof(b));
    if (b.magic != 0xACCFF7B2u || b.rank != q || b.nprocs != nblobs)
      return fail(ACCFFTB_ERR_ARG, "accfftb_plan_ipc_import: blob " + std::to_string(q) + " is not rank-ordered / valid");
    const size_t bs = bufsz_of(e, q);
    if (b.slab_bytes != (size_t)e->howmany * e->nbuf * bs + kFlagBytes)
      return fail(ACCFFTB_ERR_ARG, "accfftb_plan_ipc_import: peer plan geometry (or its buffering mode / field count) differs");
    if (q == e->rank) continue;
    void* base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess));
    e->opened.push_back(base);
    e->plan.peerA[q] = base;
    e->plan.peerB[q] = (char*)base + bs;
    if (e->nbuf == 4) {
      e->plan.peerA_alt[q] = (char*)base + 2 * bs;
      e->plan.peerB_alt[q] = (char*)base + 3 * bs;
    }
    e->flag_ptrs[q] = (unsigned long long*)((char*)base + (size_t)e->howmany * e->nbuf * bs);
  }
  e->wired = true;
  return 0;
}

int accfftb_plan_attach_local_peers(accfftb_plan* plan, accfftb_plan* const* plans, int nplans) {
  if (!plan || !plans) return fail(ACCFFTB_ERR_ARG, "accfftb_plan_attach_local_peers: null argument");
  Engine* e = &plan->e;
  if (nplans != e->plan.nprocs) return fail(ACCFFTB_ERR_ARG, "need one plan per rank");
  CU(cudaSetDevice(e->device));
  for (int q = 0; q < nplans; ++q) {
    if (q == e->rank) continue;
    Engine* o = &plans[q]->e;
    if (o->rank != q) return fail(ACCFFTB_ERR_ARG, "plans are not rank-ordered");
    if (o->device != e->device) {
      cudaError_t err = cudaDeviceEnablePeerAccess(o->device, 0);
      if (err == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (err != cudaSuccess) return cuda_fail(err, "cudaDeviceEnablePeerAccess");
    }
    if (o->nbuf != e->nbuf || o->howmany != e->howmany) return fail(ACCFFTB_ERR_ARG, "plans differ in their buffering mode / field count");
    e->plan.peerA[q] = o->slab;
    e->plan.peerB[q] = o->slab + o->bufsz;
    if (e->nbuf == 4) {
      e->plan.peerA_alt[q] = o->slab + 2 * o->bufsz;
      e->plan.peerB_alt[q] = o->slab + 3 * o->bufsz;
    }
    e->flag_ptrs[q] = (unsigned long long*)(o->slab + (size_t)o->howmany * o->nbuf * o->bufsz);
  }
  e->wired = true;
  return 0;
}

int accfftb_plan_set_stream(accfftb_plan* plan, void* cuda_stream) {
  if (!plan) return fail(ACCFFTB_ERR_ARG, "null plan");
  plan->e.stream = (cudaStream_t)cuda_stream;
  return 0;
}

int accfftb_plan_set_blocking(accfftb_plan* plan, int blocking) {
  if (!plan) return fail(ACCFFTB_ERR_ARG, "null plan");
  plan->e.blocking = blocking != 0;
  return 0;
}

int accfftb_plan_set_profiling(accfftb_plan* plan, int on) {
  if (!plan) return fail(ACCFFTB_ERR_ARG, "null plan");
  plan->e.profiling = on != 0;
  return 0;
}

int accfftb_plan_profile_read(accfftb_plan* plan, double ms[ACCFFTB_PROFILE_SLOTS], long long count[ACCFFTB_PROFILE_SLOTS]) {
  if (!plan || !ms || !count) return fail(ACCFFTB_ERR_ARG, "null argument");
  Engine* e = &plan->e;
  CU(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < ACCFFTB_PROFILE_SLOTS; ++i) ms[i] = 0, count[i] = 0;
  // slot 9 (copy-engine transfers): transfers on different lanes overlap, so report the UNION of their intervals -- the
  // time during which at least one transfer was in flight -- not the sum
  std::vector<std::pair<float, float>> spans;
  cudaEvent_t ref = nullptr;
  for (auto& r : e->recs) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.slot == 9) {
      if (!ref) ref = r.a;
      float t0 = 0;
      CU(cudaEventElapsedTime(&t0, ref, r.a));
      spans.push_back({t0, t0 + t});
      count[9]++;
    } else if (r.slot >= 0 && r.slot < ACCFFTB_PROFILE_SLOTS) {
      ms[r.slot] += t, count[r.slot]++;
    }
  }
  std::sort(spans.begin(), spans.end());
  float hi = -1e30f;
  for (auto& sp : spans) {
    if (sp.first > hi) ms[9] += sp.second - sp.first, hi = sp.second;
    else if (sp.second > hi) ms[9] += sp.second - hi, hi = sp.second;
  }
  for (auto& r : e->recs) {
    e->pool.push_back(r.a);
    e->pool.push_back(r.b);
  }
  e->recs.clear();
  return 0;
}

int accfftb_plan_stage_variant(const accfftb_plan* plan, int slot) {
  if (!plan || slot < 0 || slot >= ACCFFTB_PROFILE_SLOTS) return -1;
  return plan->e.variant[slot];
}

int accfftb_plan_fused_info(const accfftb_plan* plan, int dir, int info[8]) {
  if (!plan || dir < 0 || dir > 1) return -1;
  const FusedInfo& f = plan->e.fused_info[dir];
  if (info) {
    const int v[8] = {f.cp, f.lag, f.ns, f.nch, f.ring_planes, f.grid, f.threads, f.smem};
    for (int i = 0; i < 8; ++i) info[i] = v[i];
  }
  return f.grid > 0 ? 1 : 0;
}

long long accfftb_plan_alloc_max(const accfftb_plan* plan) { return plan ? plan->e.plan.g.L.alloc_max : -1; }
long long accfftb_plan_launch_count(const accfftb_plan* plan) { return plan ? plan->e.launches : -1; }
long long accfftb_plan_copy_count(const accfftb_plan* plan) { return plan ? plan->e.copies : -1; }

int accfftb_plan_check(accfftb_plan* plan) {
  if (!plan) return fail(ACCFFTB_ERR_ARG, "null plan");
  return engine_check(&plan->e);
}

int accfftb_plan_exchange_mode(const accfftb_plan* plan, int* chunks, int* lanes) {
  if (!plan) return -1;
  const Engine* e = &plan->e;
  const bool ce = e->plan.nprocs > 1 && e->stage_buf && e->ce_chunks > 0 && e->ce_lanes > 0;
  if (chunks) *chunks = ce ? e->ce_chunks : 0;
  if (lanes) *lanes = ce ? e->ce_lanes : 0;
  return ce ? 1 : 0;
}

int accfftb_plan_destroy(accfftb_plan* plan) {
  if (!plan) return 0;
  engine_free(&plan->e);
  delete plan;
  return 0;
}

#define CHECK_PLAN(plan, want_r2c)                                                                  \
  if (!(plan)) return fail(ACCFFTB_ERR_ARG, "null plan");                                          \
  if ((plan)->e.plan.g.r2c != (want_r2c))                                                           \
    return fail(ACCFFTB_ERR_ARG, (want_r2c) ? "this call needs an r2c plan" : "this call needs a c2c plan"); \
  if (xyz < 0 || xyz > 7) return fail(ACCFFTB_ERR_ARG, "xyz must be a 3-bit mask (bit 2 = Z, bit 1 = Y, bit 0 = X)"); \
  if ((want_r2c) && !(xyz & 4))                                                                      \
    return fail(ACCFFTB_ERR_UNSUPPORTED, "r2c / c2r need the Z axis selected: the real <-> complex step lives there " \
                                         "(the reference reinterprets the real rows as complex in that case)"); \
  if (!(plan)->e.wired) return fail(ACCFFTB_ERR_STATE, "multi-rank plan used before its peers were attached"); \
  if (int dead__ = engine_check(&(plan)->e)) return dead__;

int accfftb_execute_r2c(accfftb_plan* plan, const void* data, void* data_out, double* timer, int xyz) {
  CHECK_PLAN(plan, true);
  if (!data || !data_out) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  return engine_finish(e, e->plan.execute_forward(data, data_out, data == data_out, xyz));
}

int accfftb_execute_c2r(accfftb_plan* plan, const void* data, void* data_out, double* timer, int xyz) {
  CHECK_PLAN(plan, true);
  if (!data || !data_out) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  return engine_finish(e, e->plan.execute_inverse(data, data_out, data == data_out, xyz));
}

int accfftb_execute_c2c(accfftb_plan* plan, int direction, const void* data, void* data_out, double* timer, int xyz) {
  CHECK_PLAN(plan, false);
  if (!data || !data_out) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  if (direction != ACCFFTB_FORWARD && direction != ACCFFTB_BACKWARD)
    return fail(ACCFFTB_ERR_ARG, "direction must be ACCFFTB_FORWARD or ACCFFTB_BACKWARD");
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  if (direction == ACCFFTB_FORWARD) return engine_finish(e, e->plan.execute_forward(data, data_out, false, xyz));
  return engine_finish(e, e->plan.execute_inverse(data, data_out, false, xyz));
}

// ---- batched fields ------------------------------------------------------------------------------------------------
static int many_args(accfftb_plan* plan, int howmany, const void* const* in, void* const* out) {
  if (howmany < 1 || howmany > plan->e.howmany)
    return fail(ACCFFTB_ERR_ARG, "howmany exceeds the field count the plan was made for (accfftb_plan_create_many)");
  if (!in || !out) return fail(ACCFFTB_ERR_ARG, "null pointer array");
  for (int f = 0; f < howmany; ++f)
    if (!in[f] || !out[f]) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  return 0;
}

int accfftb_plan_howmany(const accfftb_plan* plan) { return plan ? plan->e.howmany : -1; }

int accfftb_execute_r2c_many(accfftb_plan* plan, int howmany, const void* const* data, void* const* data_out, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (int rc = many_args(plan, howmany, data, data_out)) return rc;
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  return engine_finish(e, e->plan.execute_forward_many(howmany, data, data_out, false));
}

int accfftb_execute_c2r_many(accfftb_plan* plan, int howmany, const void* const* data, void* const* data_out, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (int rc = many_args(plan, howmany, data, data_out)) return rc;
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  return engine_finish(e, e->plan.execute_inverse_many(howmany, data, data_out, false));
}

int accfftb_execute_c2c_many(accfftb_plan* plan, int direction, int howmany, const void* const* data, void* const* data_out, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, false);
  if (int rc = many_args(plan, howmany, data, data_out)) return rc;
  if (direction != ACCFFTB_FORWARD && direction != ACCFFTB_BACKWARD)
    return fail(ACCFFTB_ERR_ARG, "direction must be ACCFFTB_FORWARD or ACCFFTB_BACKWARD");
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  if (direction == ACCFFTB_FORWARD) return engine_finish(e, e->plan.execute_forward_many(howmany, data, data_out, false));
  return engine_finish(e, e->plan.execute_inverse_many(howmany, data, data_out, false));
}

static int ensure_spec(Engine* e) {
  if (e->spec) return 0;
  return fail(ACCFFTB_ERR_STATE, "operator scratch buffer missing");
}

int accfftb_laplace(accfftb_plan* plan, const void* A, void* LA, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (!A || !LA) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  Engine* e = &plan->e;
  int rc = ensure_spec(e);
  if (rc) return rc;
  TimerScope ts(e, timer, 7);
  return engine_finish(e, e->plan.execute_spectral(A, LA, OP_LAPLACE, e->spec));
}

int accfftb_biharmonic(accfftb_plan* plan, const void* A, void* BA, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (!A || !BA) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  Engine* e = &plan->e;
  int rc = ensure_spec(e);
  if (rc) return rc;
  TimerScope ts(e, timer, 7);
  return engine_finish(e, e->plan.execute_spectral(A, BA, OP_BIHARMONIC, e->spec));
}

static int spectral_impl(accfftb_plan* plan, const void* A, void* out, int op, double a, double b, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (!A || !out) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  Engine* e = &plan->e;
  int rc = ensure_spec(e);
  if (rc) return rc;
  TimerScope ts(e, timer, 7);
  return engine_finish(e, e->plan.execute_spectral(A, out, op, e->spec, a, b));
}

int accfftb_inv_laplace(accfftb_plan* plan, const void* A, void* invLA, double* timer) {
  return spectral_impl(plan, A, invLA, OP_INV_LAPLACE, 0, 0, timer);
}
int accfftb_inv_biharmonic(accfftb_plan* plan, const void* A, void* invBA, double* timer) {
  return spectral_impl(plan, A, invBA, OP_INV_BIHARMONIC, 0, 0, timer);
}
int accfftb_symbol_solve(accfftb_plan* plan, const void* A, void* out, double alpha, double beta, int inverse, double* timer) {
  return spectral_impl(plan, A, out, inverse ? OP_SYMBOL_INV : OP_SYMBOL_LIN, alpha, beta, timer);
}

static int grad_impl(accfftb_plan* plan, const void* A, void* Ax, void* Ay, void* Az, int comps, double* timer, bool slow) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (!A) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  if (comps < 0 || comps > 7) return fail(ACCFFTB_ERR_ARG, "component mask must be 0..7");
  if (((comps & 1) && !Ax) || ((comps & 2) && !Ay) || ((comps & 4) && !Az))
    return fail(ACCFFTB_ERR_ARG, "selected gradient component has a null output pointer");
  Engine* e = &plan->e;
  int rc = ensure_spec(e);
  if (rc) return rc;
  TimerScope ts(e, timer, 7);
  void* outs[3] = {Ax, Ay, Az};
  return engine_finish(e, slow ? e->plan.execute_grad(A, outs, comps, e->spec) : e->plan.execute_grad_axes(A, outs, comps, e->spec));
}

static int div_impl(accfftb_plan* plan, const void* Ax, const void* Ay, const void* Az, void* div, double* timer, bool slow) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (!Ax || !Ay || !Az || !div) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  Engine* e = &plan->e;
  int rc = ensure_spec(e);
  if (rc) return rc;
  TimerScope ts(e, timer, 7);
  const void* ins[3] = {Ax, Ay, Az};
  // The per-axis scheme accumulates into the output while it still reads inputs: an output that aliases a component (the
  // reference's fast divergence tolerates divA == A_x, src/operators_gpu.txx:448-553) goes through the 3-D scheme instead,
  // with the per-axis scheme's Nyquist rule -- the same operator; its output is written only after all inputs are consumed.
  const bool aliased = div == Ax || div == Ay || div == Az;
  if (!slow && aliased) return engine_finish(e, e->plan.execute_divergence(ins, div, e->spec, true));
  return engine_finish(e, slow ? e->plan.execute_divergence(ins, div, e->spec) : e->plan.execute_divergence_axes(ins, div, e->spec));
}

int accfftb_grad(accfftb_plan* plan, const void* A, void* Ax, void* Ay, void* Az, int comps, double* timer) {
  return grad_impl(plan, A, Ax, Ay, Az, comps, timer, false);
}
int accfftb_grad_slow(accfftb_plan* plan, const void* A, void* Ax, void* Ay, void* Az, int comps, double* timer) {
  return grad_impl(plan, A, Ax, Ay, Az, comps, timer, true);
}
int accfftb_divergence(accfftb_plan* plan, const void* Ax, const void* Ay, const void* Az, void* div, double* timer) {
  return div_impl(plan, Ax, Ay, Az, div, timer, false);
}
int accfftb_divergence_slow(accfftb_plan* plan, const void* Ax, const void* Ay, const void* Az, void* div, double* timer) {
  return div_impl(plan, Ax, Ay, Az, div, timer, true);
}

int accfftb_execute_r2c_axis(accfftb_plan* plan, int axis, const void* data, void* data_out, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (axis < 0 || axis > 2) return fail(ACCFFTB_ERR_ARG, "axis must be 0 (x), 1 (y) or 2 (z)");
  if (!data || !data_out) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  if (data == data_out) return fail(ACCFFTB_ERR_UNSUPPORTED, "single-axis transforms are out of place");
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  return engine_finish(e, e->plan.execute_r2c_axis(axis, data, data_out));
}

int accfftb_execute_c2r_axis(accfftb_plan* plan, int axis, const void* data, void* data_out, double* timer) {
  const int xyz = 7;
  CHECK_PLAN(plan, true);
  if (axis < 0 || axis > 2) return fail(ACCFFTB_ERR_ARG, "axis must be 0 (x), 1 (y) or 2 (z)");
  if (!data || !data_out) return fail(ACCFFTB_ERR_ARG, "null data pointer");
  if (data == data_out) return fail(ACCFFTB_ERR_UNSUPPORTED, "single-axis transforms are out of place");
  Engine* e = &plan->e;
  TimerScope ts(e, timer, 5);
  return engine_finish(e, e->plan.execute_c2r_axis(axis, data, data_out));
}

}  // extern "C"
