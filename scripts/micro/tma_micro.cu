// tma_micro.cu -- microbenchmark (measurement tooling, not product): throughput of 1-D bulk (TMA) copies
// global -> shared -> global through an mbarrier pipeline, as a function of chunk size, pipeline depth and CTA count;
// destination either local HBM or a peer GPU (NVLink).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_micro tma_micro.cu
//   ./tma_micro            (1 GPU: local copies;  >= 2 GPUs: also GPU0 -> GPU1 pushes)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          (unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(sdst)),
               "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

// One elected thread per CTA runs the whole pipeline.  A "stage" holds `ops` chunks of `chunk` bytes (stage bytes =
// ops*chunk); each chunk is its own bulk op on both sides (models 128-byte row segments when chunk = 128).
// Source chunk i of stage-tile T is at src + (T*ops + i)*chunk*srcStrideMul; dst likewise.
template <int NST>
__global__ void k_pipe(const char* src, char* dst, long long ntiles, int ops, int chunk, long long srcStride, long long dstStride) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[NST];
  const int stageBytes = ops * chunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x;   // one warp per CTA; lanes share the bulk-op issue work
  auto tile_of = [&](long long i) { return (long long)blockIdx.x + i * gridDim.x; };
  long long nmine = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (nmine < 0) nmine = 0;
  auto issue_load = [&](long long i) {
    const int s = (int)(i % NST);
    const long long T = tile_of(i);
    if (lane == 0) mbar_expect(&full[s], (unsigned)stageBytes);
    __syncwarp();
    for (int o = lane; o < ops; o += 32) bulk_load(smem + (size_t)s * stageBytes + (size_t)o * chunk, src + (T * ops + o) * srcStride, chunk, &full[s]);
  };
  for (long long i = 0; i < NST && i < nmine; ++i) issue_load(i);
  for (long long i = 0; i < nmine; ++i) {
    const int s = (int)(i % NST);
    mbar_wait(&full[s], (unsigned)((i / NST) & 1));
    const long long T = tile_of(i);
    for (int o = lane; o < ops; o += 32) bulk_store(dst + (T * ops + o) * dstStride, smem + (size_t)s * stageBytes + (size_t)o * chunk, chunk);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // the stage written out one iteration ago may be refilled once its stores have read shared memory
    if (i >= 1) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      if (i - 1 + NST < nmine) issue_load(i - 1 + NST);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static float run(int nst, int grid, const char* src, char* dst, size_t bytes, int ops, int chunk, long long srcStride, long long dstStride, int iters) {
  const long long ntiles = (long long)(bytes / ((size_t)ops * chunk));
  const int smem = nst * ops * chunk;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto launch = [&]() {
    if (nst == 2) { CK(cudaFuncSetAttribute(k_pipe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_pipe<2><<<grid, 32, smem>>>(src, dst, ntiles, ops, chunk, srcStride, dstStride); }
    else if (nst == 4) { CK(cudaFuncSetAttribute(k_pipe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_pipe<4><<<grid, 32, smem>>>(src, dst, ntiles, ops, chunk, srcStride, dstStride); }
    else { CK(cudaFuncSetAttribute(k_pipe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); k_pipe<8><<<grid, 32, smem>>>(src, dst, ntiles, ops, chunk, srcStride, dstStride); }
  };
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / iters;
}

int main() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  CK(cudaSetDevice(0));
  const size_t BYTES = 1ull << 30;
  char *src, *dst, *peer = nullptr;
  CK(cudaMalloc(&src, BYTES));
  CK(cudaMalloc(&dst, BYTES));
  CK(cudaMemset(src, 1, BYTES));
  CK(cudaMemset(dst, 0, BYTES));
  if (ndev > 1) {
    CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&peer, BYTES));
    CK(cudaMemset(peer, 0, BYTES));
    CK(cudaDeviceSynchronize());
    CK(cudaSetDevice(0));
  }
  printf("tma_micro: %d device(s); 1 GiB per launch; GB/s = payload bytes / time (HBM traffic is 2x for local copies)\n", ndev);
  struct Cfg { int nst, ctas_per_sm, ops, chunk; };
  // stage sizes: 64 KB tiles as 1 x 64 KB, 8 x 8 KB, 64 x 1 KB, 512 x 128 B (row segments); 16 KB stages too
  const Cfg cfgs[] = {{2, 1, 1, 65536}, {2, 1, 8, 8192}, {2, 1, 64, 1024}, {2, 1, 512, 128}, {2, 1, 256, 256},
                      {4, 1, 1, 32768}, {4, 1, 256, 128}, {4, 2, 1, 16384}, {4, 2, 128, 128}, {8, 2, 1, 8192}, {8, 2, 64, 128},
                      {4, 3, 1, 16384}, {4, 3, 16, 1024}, {4, 3, 128, 128}};
  for (const Cfg& c : cfgs) {
    const float ms = run(c.nst, 148 * c.ctas_per_sm, src, dst, BYTES, c.ops, c.chunk, c.chunk, c.chunk, 5);
    printf("local  nst %d ctas/sm %d ops %4d x %6d B : %.3f ms  payload %.0f GB/s (HBM r+w %.0f GB/s)\n", c.nst, c.ctas_per_sm, c.ops, c.chunk, ms,
           BYTES / ms / 1e6, 2.0 * BYTES / ms / 1e6);
    fflush(stdout);
  }
  // strided source (column tile of a [512][stride] array: 128-byte segments 8 KB apart), dense destination
  {
    const float ms = run(4, 148, src, dst, BYTES / 2, 512, 128, 8192, 128, 5);
    printf("local  strided src (128 B every 8 KB) nst 4 ctas/sm 1 : %.3f ms payload %.0f GB/s\n", ms, BYTES / 2 / ms / 1e6);
  }
  if (peer) {
    for (int grid : {8, 16, 32, 64, 148, 296}) {
      for (const Cfg& c : {Cfg{4, 1, 1, 16384}, Cfg{4, 1, 16, 1024}, Cfg{4, 1, 128, 128}}) {
        const float ms = run(c.nst, grid, src, peer, BYTES / 4, c.ops, c.chunk, c.chunk, c.chunk, 5);
        printf("peer   grid %4d nst %d ops %4d x %6d B : %.3f ms  NVLink %.0f GB/s\n", grid, c.nst, c.ops, c.chunk, ms, BYTES / 4 / ms / 1e6);
        fflush(stdout);
      }
    }
  }
  return 0;
}
