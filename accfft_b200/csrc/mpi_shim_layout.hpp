// mpi_shim_layout.hpp -- shared-memory layout of the single-node MPI shim (mpi_shim.cpp) and its launcher
// (accfft_mpirun.cpp).  The launcher creates a segment of sizeof(ShimShm) bytes (sparse: tmpfs pages are committed on
// first touch, so the large outboxes cost nothing until a rank sends that much) and zero-fills it by ftruncate.
#pragma once
#include <stddef.h>
#include <atomic>

namespace accfft_mpi_shim {

constexpr int kMaxRanks = 64;
constexpr int kMaxComms = 256;
constexpr size_t kSlot = 64 * 1024;            // collective staging slot per world rank
constexpr size_t kOutbox = 48ull << 20;        // eager-send payload area per world rank
constexpr int kQueue = 128;                    // in-flight messages per (source, destination) pair

struct ShmBarrier {
  std::atomic<int> count;
  std::atomic<int> sense;
  std::atomic<int> leavers;   // Comm_free: ranks that have left the last barrier (the last one recycles the id)
};
// one eager message: payload lives at `off` in the SOURCE rank's outbox
struct ShmMsg {
  std::atomic<int> state;     // 0 free, 1 ready (set by the sender), 2 consumed (set by the receiver; the sender frees it)
  int tag, comm;
  unsigned long long bytes, off;
};
struct ShmChan {              // FIFO from one source to one destination
  std::atomic<unsigned> head; // next slot the source fills
  ShmMsg q[kQueue];
};
struct ShimShm {
  std::atomic<int> comm_used[kMaxComms];       // barrier / communicator ids in use (0 is MPI_COMM_WORLD)
  ShmBarrier bar[kMaxComms];
  ShmChan chan[kMaxRanks][kMaxRanks];
  alignas(64) unsigned char slot[kMaxRanks][kSlot];
  alignas(64) unsigned char outbox[kMaxRanks][kOutbox];
};

}  // namespace accfft_mpi_shim
