// nvlink_push.cu -- microbenchmark (measurement tooling, not product): achieved NVLink bandwidth of SM stores /
// loads to peer memory for the access patterns of the fused FFT+exchange stage kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvlink_push nvlink_push.cu -lpthread
//   ./nvlink_push [ngpus]
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Ptrs { uint4* p[8]; };

// mode 0: push (read local, write dst[peer]); mode 1: pull (read dst[peer], write local)
// element idx (16 B): run = idx / run16, within = idx % run16, peer = run % npeers, slot = run / npeers
__global__ void k_xfer(Ptrs peers, uint4* local, size_t n16, int npeers, int run16, int pitch16, int mode) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n16; idx += stride) {
    const size_t run = idx / run16;
    const int within = (int)(idx % run16);
    const int peer = (int)(run % npeers);
    const size_t slot = run / npeers;
    uint4* rp = peers.p[peer] + slot * pitch16 + within;
    if (mode == 0) *rp = local[idx];
    else local[idx] = *rp;
  }
}

// row-segment pattern of the strided (Y/X) stages: a CTA owns a tile of 8 columns (128 B) x NR rows; quarter-warps
// write one 128-byte row segment each; consecutive rows are rowpitch16 apart; rows [s*blk, (s+1)*blk) go to peer s.
__global__ void k_rows(Ptrs peers, const uint4* local, int ntiles, int nrows, int npeers, int rowpitch16) {
  const int c = threadIdx.x % 8, t = threadIdx.x / 8;   // 512 threads: t in [0,64)
  const int blk = nrows / npeers;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    uint4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = local[((size_t)tile * nrows + t + r * 64) * 8 + c];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = t + r * 64;
      const int s = row / blk;
      peers.p[s][((size_t)(row - s * blk)) * rowpitch16 + (size_t)tile * 8 + c] = v[r];
    }
  }
}

static int g_n = 8;
static pthread_barrier_t g_bar;
static uint4* g_buf[8];       // receive buffer of each GPU
static uint4* g_src[8];
static const size_t BYTES = 256u << 20;   // per GPU payload per iteration

struct Case { const char* name; int kind; int npeers_incl_self; int run16, pitch16; int grid; int mode; };

static double g_res[8];

static void* worker(void* arg) {
  const int g = (int)(size_t)arg;
  CK(cudaSetDevice(g));
  for (int q = 0; q < g_n; ++q) if (q != g) { cudaError_t e = cudaDeviceEnablePeerAccess(q, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("peer access %d->%d failed\n", g, q); exit(1);} cudaGetLastError(); }
  CK(cudaMalloc(&g_buf[g], BYTES * 2));
  CK(cudaMalloc(&g_src[g], BYTES));
  CK(cudaMemset(g_src[g], 1, BYTES));
  CK(cudaMemset(g_buf[g], 0, BYTES * 2));
  CK(cudaDeviceSynchronize());
  pthread_barrier_wait(&g_bar);
  std::vector<Case> cases;
  const int G148 = 148 * 8;
  for (int np : {2, 4, 8}) {
    if (np > g_n) continue;
    cases.push_back({"push contiguous 1MB runs", 0, np, 65536, 65536, G148, 0});
    cases.push_back({"push run 64x16B aligned", 0, np, 64, 64, G148, 0});
    cases.push_back({"push run 65x16B dense (unaligned)", 0, np, 65, 65, G148, 0});
    cases.push_back({"push run 65x16B pitch 72 (aligned)", 0, np, 65, 72, G148, 0});
    cases.push_back({"push run 8x16B (128B lines)", 0, np, 8, 8, G148, 0});
    cases.push_back({"push run 33x16B dense", 0, np, 33, 33, G148, 0});
    cases.push_back({"pull contiguous 1MB runs", 0, np, 65536, 65536, G148, 1});
    cases.push_back({"pull run 65x16B dense", 0, np, 65, 65, G148, 1});
    cases.push_back({"rows 128B segs pitch 72", 1, np, 0, 72, 148 * 2, 0});
    cases.push_back({"rows 128B segs pitch 4104", 1, np, 0, 4104, 148 * 2, 0});
  }
  for (int grid : {16, 32, 64, 148, 296, 592}) cases.push_back({"push contiguous, grid sweep", 0, g_n, 65536, 65536, grid, 0});
  for (int grid : {16, 32, 64, 148}) cases.push_back({"pull contiguous, grid sweep", 0, g_n, 65536, 65536, grid, 1});
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (const Case& c : cases) {
    const int np = c.npeers_incl_self;
    // destinations: self first, then the next np-1 GPUs in the same aligned group (like a row / column group)
    Ptrs P;
    const int base = g / np * np;
    for (int s = 0; s < np; ++s) {
      const int q = base + (g - base + s) % np;
      // every sender writes into its own region of the receiver: region index = sender's position in the group
      P.p[s] = g_buf[q] + (size_t)(g - base) * (BYTES * 2 / 16 / np);
    }
    const size_t n16 = BYTES / 16;
    auto run = [&]() {
      if (c.kind == 0) {
        // cap so that (slot * pitch) stays inside the per-sender region
        size_t n = n16;
        if (c.pitch16 > c.run16) n = (size_t)((double)n16 * c.run16 / c.pitch16);
        k_xfer<<<c.grid, 256>>>(P, g_src[g], n, np, c.run16, c.pitch16, c.mode);
      } else {
        const int nrows = 512;
        const int ntiles = std::min<size_t>(n16 / (nrows * 8), (size_t)c.pitch16 / 8);
        // repeat to move enough data
        const int reps = (int)(n16 / ((size_t)nrows * 8 * ntiles));
        for (int r = 0; r < std::min(reps, 64); ++r) k_rows<<<c.grid, 512>>>(P, g_src[g], ntiles, nrows, np, c.pitch16);
      }
    };
    run();
    CK(cudaDeviceSynchronize());
    pthread_barrier_wait(&g_bar);
    CK(cudaEventRecord(e0));
    const int iters = 5;
    for (int i = 0; i < iters; ++i) run();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double moved = (double)BYTES;
    if (c.kind == 0 && c.pitch16 > c.run16) moved *= (double)c.run16 / c.pitch16;
    if (c.kind == 1) {
      const int nrows = 512;
      const int ntiles = std::min<size_t>(BYTES / 16 / (nrows * 8), (size_t)c.pitch16 / 8);
      const int reps = std::min((int)(BYTES / 16 / ((size_t)nrows * 8 * ntiles)), 64);
      moved = (double)reps * nrows * ntiles * 128.0;
    }
    const double remote = moved * (np - 1) / np;
    g_res[g] = remote * iters / (ms * 1e-3) / 1e9;
    pthread_barrier_wait(&g_bar);
    if (g == 0) {
      double mn = 1e30, mx = 0;
      for (int q = 0; q < g_n; ++q) mn = std::min(mn, g_res[q]), mx = std::max(mx, g_res[q]);
      printf("%-40s group %d grid %5d : NVLink out per GPU min %7.1f max %7.1f GB/s  (total moved/GPU %.0f MB, %.3f ms)\n", c.name, np,
             c.grid, mn, mx, moved / 1e6, ms / iters);
      fflush(stdout);
    }
    pthread_barrier_wait(&g_bar);
  }
  return nullptr;
}

int main(int argc, char** argv) {
  int have = 0;
  CK(cudaGetDeviceCount(&have));
  g_n = argc > 1 ? atoi(argv[1]) : have;
  if (g_n > have) g_n = have;
  printf("nvlink_push: %d GPUs\n", g_n);
  pthread_barrier_init(&g_bar, nullptr, g_n);
  pthread_t th[8];
  for (int g = 0; g < g_n; ++g) pthread_create(&th[g], nullptr, worker, (void*)(size_t)g);
  for (int g = 0; g < g_n; ++g) pthread_join(th[g], nullptr);
  return 0;
}
